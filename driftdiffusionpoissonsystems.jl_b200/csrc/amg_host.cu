// Host-side setup of the smoothed-aggregation multigrid hierarchy (a better
// preconditioner behind the same PCG interface, replacing the reference's `\` at src:394,588,1063).
//
// This file contains NO device code: it is the one-time, per-mesh symbolic/numeric setup that freezes the
// transfer operators.  Per level it
//   1. marks strong couplings of the BASE operator (the constant stiffness MV / semlin M, src:284,1028):
//      -a_ij >= theta * max_k(-a_ik);
//   2. aggregates the nodes greedily (root + all strong neighbours, leftovers join the strongest neighbouring
//      aggregate) -- sequential, hence deterministic; Dirichlet rows ("fixed") are left out: their rows are
//      identity rows (src:277-284) and are handled by the smoother alone;
//   3. smooths the piecewise-constant tentative prolongator with one damped-Jacobi step of the base operator,
//      P = (I - omega D^-1 A) T;
//   4. stores R = P^T row-wise (so that restriction and the Galerkin product are deterministic gathers) and
//      forms the coarse base operator R A P with a row-parallel marker-array product.
// The per-operator work (Galerkin products of the CURRENT operator M - J0 on the frozen pattern, smoothing,
// transfers, coarse solve) runs on the device (amg.cu).  Everything here is reachable WITHOUT a GPU through the
// ddps_amg_host_* entry points so that the CPU test-suite can check it against scipy.
#include <stdint.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <thread>
#include <vector>

#include "amg.h"

namespace ddps {

namespace {

template <typename F>
void parallel_ranges(int64_t n, int nthreads, F&& fn) {
  nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, (n + 4095) / 4096));
  if (nthreads <= 1) { fn(0, (int64_t)0, n); return; }
  std::vector<std::thread> th;
  const int64_t chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    const int64_t lo = std::min<int64_t>(n, t * chunk), hi = std::min<int64_t>(n, lo + chunk);
    th.emplace_back([&fn, t, lo, hi]() { fn(t, lo, hi); });
  }
  for (auto& x : th) x.join();
}

// Strong-neighbour lists: for every row the columns j != i with -a_ij >= theta * max_k(-a_ik) (fixed rows/columns
// excluded), sorted by strength (strongest first, ties by column).  The device setup (amg_setup.cu) produces the
// same lists with a kernel; everything downstream (aggregation) only sees the lists.
void strong_lists(const AmgHostLevel& L, const unsigned char* fixed, double theta, int nthreads, std::vector<int>& sptr,
                  std::vector<int>& scol) {
  const int n = L.n;
  sptr.assign(n + 1, 0);
  parallel_ranges(n, nthreads, [&](int, int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      if (fixed && fixed[i]) continue;
      double m = 0.0;
      for (int k = L.rowptr[i]; k < L.rowptr[i + 1]; ++k) {
        const int j = L.col[k];
        if (j != i && !(fixed && fixed[j])) m = std::max(m, -L.val[k]);
      }
      if (!(m > 0.0)) continue;
      const double cut = theta * m;
      int c = 0;
      for (int k = L.rowptr[i]; k < L.rowptr[i + 1]; ++k) {
        const int j = L.col[k];
        if (j != i && !(fixed && fixed[j]) && -L.val[k] >= cut) ++c;
      }
      sptr[i + 1] = c;
    }
  });
  for (int i = 0; i < n; ++i) sptr[i + 1] += sptr[i];
  scol.resize(sptr[n]);
  parallel_ranges(n, nthreads, [&](int, int64_t lo, int64_t hi) {
    std::vector<std::pair<double, int>> nb;
    for (int64_t i = lo; i < hi; ++i) {
      if (sptr[i + 1] == sptr[i]) continue;
      double m = 0.0;
      for (int k = L.rowptr[i]; k < L.rowptr[i + 1]; ++k) {
        const int j = L.col[k];
        if (j != i && !(fixed && fixed[j])) m = std::max(m, -L.val[k]);
      }
      const double cut = theta * m;
      nb.clear();
      for (int k = L.rowptr[i]; k < L.rowptr[i + 1]; ++k) {
        const int j = L.col[k];
        if (j != i && !(fixed && fixed[j]) && -L.val[k] >= cut) nb.emplace_back(L.val[k], j);   // most negative first
      }
      std::sort(nb.begin(), nb.end());
      for (size_t q = 0; q < nb.size(); ++q) scol[sptr[i] + q] = nb[q].second;
    }
  });
}

}  // namespace

// Greedy aggregation on the strong-neighbour lists (sequential: the result is a pure function of the lists).
// cap == 0: a node whose strong neighbourhood is still free becomes the root of a new aggregate with all of it;
// leftovers join the aggregate of their strongest aggregated neighbour; what is still free forms new aggregates.
// cap > 0 (gentle coarsening for the last, tiny levels: the global smooth modes live there and a coarsening ratio
// of ~7 from a few hundred rows costs CG iterations): a free node takes at most `cap` of its strongest free strong
// neighbours; leftovers join their strongest aggregated neighbour or stay alone.
int amg_aggregate_lists(int n, const int* sptr, const int* scol, int cap, std::vector<int>& agg,
                        const std::atomic<long long>* rows_ready) {
  const bool timing = getenv("DDPS_AMG_TIMING") != nullptr && n > 1000000;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double ta = now();
  agg.assign(n, -1);
  int nc = 0;
  auto wait_all = [&]() {
    if (rows_ready)
      while (rows_ready->load(std::memory_order_acquire) < n) std::this_thread::yield();
  };
  if (cap > 0) {
    wait_all();
    for (int i = 0; i < n; ++i) {
      if (agg[i] >= 0 || sptr[i + 1] == sptr[i]) continue;
      int taken = 0;
      for (int k = sptr[i]; k < sptr[i + 1] && taken < cap; ++k)
        if (agg[scol[k]] < 0) { agg[scol[k]] = nc; ++taken; }
      if (!taken) continue;
      agg[i] = nc++;
    }
    for (int i = 0; i < n; ++i) {
      if (agg[i] >= 0 || sptr[i + 1] == sptr[i]) continue;
      int pick = -1;
      for (int k = sptr[i]; k < sptr[i + 1]; ++k)
        if (agg[scol[k]] >= 0) { pick = agg[scol[k]]; break; }
      agg[i] = pick >= 0 ? pick : nc++;
    }
    return nc;
  }
  // pass 1.  Large levels: the rows are cut into kAggChunks contiguous chunks that run concurrently; a root and its
  // whole strong neighbourhood must lie inside one chunk (the few rows along the seams become leftovers of pass 2/3).
  // The chunk count is a constant, so the result does not depend on the number of threads.
  constexpr int kAggChunks = 16;
  // (measured: chunking a 8.4M-row level saves 30 ms of 100 but the seams' leftovers densify the coarser levels and cost
  // +12 % CG iterations -> the chunked path is kept for experiments only, DDPS_AMG_AGG_CHUNKS=1)
  const int nchunks = (n >= 200000 && getenv("DDPS_AMG_AGG_CHUNKS")) ? kAggChunks : 1;
  if (nchunks > 1) wait_all();
  const int64_t csize = ((int64_t)n + nchunks - 1) / nchunks;
  std::vector<int> cbase(nchunks + 1, 0);
  auto pass1 = [&](int c) {
    const int lo = (int)std::min<int64_t>(n, c * csize), hi = (int)std::min<int64_t>(n, lo + csize);
    int local = 0;
    long long avail = rows_ready ? 0 : n;
    for (int i = lo; i < hi; ++i) {
      if (i >= avail) {   // streamed input: wait until the producer has delivered this row's list
        while ((avail = rows_ready->load(std::memory_order_acquire)) <= i) std::this_thread::yield();
      }
      if (agg[i] >= 0 || sptr[i + 1] == sptr[i]) continue;
      bool ok = true;
      for (int k = sptr[i]; k < sptr[i + 1]; ++k) {
        const int j = scol[k];
        if (j < lo || j >= hi || agg[j] >= 0) { ok = false; break; }
      }
      if (!ok) continue;
      agg[i] = local;
      for (int k = sptr[i]; k < sptr[i + 1]; ++k) agg[scol[k]] = local;
      ++local;
    }
    cbase[c + 1] = local;
  };
  if (nchunks == 1) pass1(0);
  else {
    const int nt = (int)std::max(1u, std::min((unsigned)nchunks, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
      th.emplace_back([&, t]() { for (int c = t; c < nchunks; c += nt) pass1(c); });
    for (auto& x : th) x.join();
    for (int c = 0; c < nchunks; ++c) cbase[c + 1] += cbase[c];
    parallel_ranges(n, nt, [&](int, int64_t lo, int64_t hi) {
      for (int64_t i = lo; i < hi; ++i)
        if (agg[i] >= 0) agg[i] += cbase[i / csize];
    });
  }
  nc = cbase[nchunks];
  const double tb = now();
  // pass 2 (no chaining: decisions use the pass-1 state only).  In place: a leftover row stores its choice as -2 - aggregate,
  // which readers still see as "not aggregated in pass 1" (negative); a second sweep decodes.  Reads the pass-1 state only,
  // so the result does not depend on the number of threads.
  {
    const int nt = n < 200000 ? 1 : (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    parallel_ranges(n, nt, [&](int, int64_t lo, int64_t hi) {
      for (int64_t i = lo; i < hi; ++i) {
        if (agg[i] >= 0) continue;
        for (int k = sptr[i]; k < sptr[i + 1]; ++k) {
          const int a = agg[scol[k]];
          if (a >= 0) { agg[i] = -2 - a; break; }
        }
      }
    });
    parallel_ranges(n, nt, [&](int, int64_t lo, int64_t hi) {
      for (int64_t i = lo; i < hi; ++i)
        if (agg[i] < -1) agg[i] = -2 - agg[i];
    });
  }
  const double tc = now();
  // pass 3
  for (int i = 0; i < n; ++i) {
    if (agg[i] >= 0 || sptr[i + 1] == sptr[i]) continue;   // isolated row (diagonally dominant): smoother only
    agg[i] = nc;
    for (int k = sptr[i]; k < sptr[i + 1]; ++k)
      if (agg[scol[k]] < 0) agg[scol[k]] = nc;
    ++nc;
  }
  if (timing) fprintf(stderr, "[ddps amg]     aggregation n=%d: pass 1 %.1f ms, pass 2 %.1f, pass 3 %.1f\n", n, tb - ta, tc - tb, now() - tc);
  return nc;
}

namespace {

// One row of P = (I - omega D^-1 A) T: distinct aggregates of the row's neighbourhood, sorted by aggregate id; the
// contributions to one aggregate are added in entry order (own term first).  Returns the number of entries.
inline int prolongator_row(const AmgHostLevel& L, const unsigned char* fixed, double omega, int i, int* ag, double* v) {
  if (fixed && fixed[i]) return 0;
  int cnt = 0;
  double d = 0.0;
  for (int k = L.rowptr[i]; k < L.rowptr[i + 1]; ++k)
    if (L.col[k] == i) d = L.val[k];
  if (L.agg[i] >= 0) { ag[0] = L.agg[i]; v[0] = 1.0 - omega; cnt = 1; }
  if (d != 0.0) {
    for (int k = L.rowptr[i]; k < L.rowptr[i + 1]; ++k) {
      const int j = L.col[k];
      if (j == i || L.val[k] == 0.0 || L.agg[j] < 0) continue;
      const int a = L.agg[j];
      const double c = -omega * L.val[k] / d;
      int pos = 0;
      while (pos < cnt && ag[pos] != a) ++pos;
      if (pos == cnt) { ag[cnt] = a; v[cnt] = c; ++cnt; }
      else v[pos] += c;
    }
  }
  for (int p = 1; p < cnt; ++p) {   // insertion sort by aggregate id (a handful of entries)
    const int a = ag[p];
    const double c = v[p];
    int q = p - 1;
    while (q >= 0 && ag[q] > a) { ag[q + 1] = ag[q]; v[q + 1] = v[q]; --q; }
    ag[q + 1] = a; v[q + 1] = c;
  }
  return cnt;
}

// P = (I - omega D^-1 A) T with T the piecewise-constant tentative prolongator; rows of fixed nodes stay empty
void smooth_prolongator(AmgHostLevel& L, const unsigned char* fixed, double omega, int nthreads) {
  const int n = L.n;
  int maxlen = 1;
  for (int i = 0; i < n; ++i) maxlen = std::max(maxlen, L.rowptr[i + 1] - L.rowptr[i] + 1);
  L.Pptr.assign(n + 1, 0);
  parallel_ranges(n, nthreads, [&](int, int64_t lo, int64_t hi) {
    std::vector<int> ag(maxlen);
    std::vector<double> v(maxlen);
    for (int64_t i = lo; i < hi; ++i) L.Pptr[i + 1] = prolongator_row(L, fixed, omega, (int)i, ag.data(), v.data());
  });
  for (int i = 0; i < n; ++i) L.Pptr[i + 1] += L.Pptr[i];
  L.Pcol.resize(L.Pptr[n]);
  L.Pval.resize(L.Pptr[n]);
  parallel_ranges(n, nthreads, [&](int, int64_t lo, int64_t hi) {
    std::vector<int> ag(maxlen);
    std::vector<double> v(maxlen);
    for (int64_t i = lo; i < hi; ++i) {
      const int cnt = prolongator_row(L, fixed, omega, (int)i, ag.data(), v.data());
      std::copy(ag.begin(), ag.begin() + cnt, L.Pcol.begin() + L.Pptr[i]);
      // the transfer operators are stored in fp32 on the device: round here so that P, R = P^T and every Galerkin
      // product (host and device) use the same numbers
      for (int c = 0; c < cnt; ++c) L.Pval[L.Pptr[i] + c] = (double)(float)v[c];
    }
  });
}

// R = P^T row-wise, fine indices increasing inside every coarse row.  Parallel counting sort: thread t owns a
// contiguous chunk of fine rows; per-thread column histograms give every (column, thread) its write offset.
void transpose_prolongator(AmgHostLevel& L, int nthreads) {
  const int n = L.n, nc = L.nc;
  const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, ((int64_t)n + 65535) / 65536));
  const int64_t chunk = ((int64_t)n + nt - 1) / nt;
  std::vector<std::vector<int>> cnt(nt);
  auto run = [&](auto&& fn) {
    if (nt == 1) { fn(0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([&fn, t]() { fn(t); });
    for (auto& x : th) x.join();
  };
  run([&](int t) {
    cnt[t].assign(nc, 0);
    const int64_t lo = std::min<int64_t>(n, t * chunk), hi = std::min<int64_t>(n, lo + chunk);
    for (int k = L.Pptr[lo]; k < L.Pptr[hi]; ++k) cnt[t][L.Pcol[k]]++;
  });
  L.Rptr.assign(nc + 1, 0);
  for (int I = 0; I < nc; ++I) {
    int s = L.Rptr[I];
    for (int t = 0; t < nt; ++t) { const int c = cnt[t][I]; cnt[t][I] = s; s += c; }   // cnt becomes the write offset
    L.Rptr[I + 1] = s;
  }
  L.Rrow.resize(L.Pcol.size());
  L.Rval.resize(L.Pcol.size());
  run([&](int t) {
    const int64_t lo = std::min<int64_t>(n, t * chunk), hi = std::min<int64_t>(n, lo + chunk);
    for (int64_t i = lo; i < hi; ++i)
      for (int k = L.Pptr[i]; k < L.Pptr[i + 1]; ++k) {
        const int pos = cnt[t][L.Pcol[k]]++;
        L.Rrow[pos] = (int)i;
        L.Rval[pos] = L.Pval[k];
      }
  });
}

// C = R A P, row-parallel, columns sorted.  Structural zeros of A are kept (the current operator M - J0 has
// nonzeros where the base stiffness of a right-angle mesh has exact zeros).
void galerkin(const AmgHostLevel& L, AmgHostLevel& C, int nthreads) {
  const int nc = L.nc;
  C.n = nc;
  std::vector<int> len(nc, 0);
  const int nt = std::max(1, nthreads);
  std::vector<std::vector<int>> tcol(nt);
  std::vector<std::vector<double>> tval(nt);
  std::vector<int64_t> tlo(nt, 0), thi(nt, 0);
  parallel_ranges(nc, nt, [&](int t, int64_t lo, int64_t hi) {
    tlo[t] = lo; thi[t] = hi;
    std::vector<int> marker(nc, -1), list;
    std::vector<double> acc(nc, 0.0);
    for (int64_t I = lo; I < hi; ++I) {
      list.clear();
      for (int q = L.Rptr[I]; q < L.Rptr[I + 1]; ++q) {
        const int i = L.Rrow[q];
        const double r = L.Rval[q];
        for (int k = L.rowptr[i]; k < L.rowptr[i + 1]; ++k) {
          const int j = L.col[k];
          const double ra = r * L.val[k];
          for (int m = L.Pptr[j]; m < L.Pptr[j + 1]; ++m) {
            const int J = L.Pcol[m];
            if (marker[J] != (int)I) { marker[J] = (int)I; acc[J] = 0.0; list.push_back(J); }
            acc[J] += ra * L.Pval[m];
          }
        }
      }
      std::sort(list.begin(), list.end());
      for (int J : list) { tcol[t].push_back(J); tval[t].push_back(acc[J]); }
      len[I] = (int)list.size();
    }
  });
  C.rowptr.assign(nc + 1, 0);
  for (int I = 0; I < nc; ++I) C.rowptr[I + 1] = C.rowptr[I] + len[I];
  C.col.resize(C.rowptr[nc]);
  C.val.resize(C.rowptr[nc]);
  for (int t = 0; t < nt; ++t) {
    if (thi[t] <= tlo[t]) continue;
    std::copy(tcol[t].begin(), tcol[t].end(), C.col.begin() + C.rowptr[tlo[t]]);
    std::copy(tval[t].begin(), tval[t].end(), C.val.begin() + C.rowptr[tlo[t]]);
  }
}

}  // namespace

int amg_host_build(AmgHost& H, int n, std::vector<int>&& rowptr, std::vector<int>&& col, std::vector<double>&& val,
                   const unsigned char* fixed, const AmgParams& prm) {
  H.lv.clear();
  H.lv.emplace_back();
  H.lv[0].n = n;
  H.lv[0].rowptr = std::move(rowptr);
  H.lv[0].col = std::move(col);
  H.lv[0].val = std::move(val);
  int nthreads = prm.nthreads > 0 ? prm.nthreads : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  const bool timing = getenv("DDPS_AMG_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  while ((int)H.lv.size() < prm.max_levels) {
    AmgHostLevel& L = H.lv.back();
    if (L.n <= prm.coarse_max) break;
    const unsigned char* fx = H.lv.size() == 1 ? fixed : nullptr;
    std::vector<int> sptr, scol;
    const double t0 = now();
    strong_lists(L, fx, prm.theta, nthreads, sptr, scol);
    const double t1 = now();
    const int nc = amg_aggregate_lists(L.n, sptr.data(), scol.data(), amg_small_cap(prm, L.n), L.agg);
    const double t2 = now();
    if (nc <= 0 || nc > 0.7 * L.n) { L.agg.clear(); break; }   // coarsening stalled
    L.nc = nc;
    smooth_prolongator(L, fx, prm.omega, nthreads);
    const double t3 = now();
    transpose_prolongator(L, nthreads);
    const double t4 = now();
    AmgHostLevel C;
    galerkin(L, C, nthreads);
    const double t5 = now();
    if (timing)
      fprintf(stderr, "[ddps amg] level %d n=%d -> %d (%d threads): strong %.1f ms, aggregate %.1f, prolongator %.1f, transpose %.1f, galerkin %.1f\n",
              (int)H.lv.size() - 1, L.n, nc, nthreads, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4);
    H.lv.push_back(std::move(C));
  }
  return (int)H.lv.size();
}

}  // namespace ddps

// ------------------------------------------------------------------------------------------------------------
// host-only C ABI (no CUDA context needed): used by the CPU tests to check the hierarchy against scipy
// ------------------------------------------------------------------------------------------------------------
struct ddps_amg_host : public ddps::AmgHost {};

extern "C" {

int ddps_amg_host_create(ddps_amg_host** out, int64_t n, const int64_t* rowptr, const int64_t* col, const double* val,
                         const uint8_t* fixed, int nthreads, int coarse_max) {
  if (!out || !rowptr || !col || !val || n <= 0 || n > 2000000000LL) return DDPS_ERR_ARG;
  if (rowptr[n] > 2000000000LL) return DDPS_ERR_ARG;
  ddps_amg_host* h = new ddps_amg_host();
  std::vector<int> rp(n + 1), cc(rowptr[n]);
  std::vector<double> vv(val, val + rowptr[n]);
  for (int64_t i = 0; i <= n; ++i) rp[i] = (int)rowptr[i];
  for (int64_t k = 0; k < rowptr[n]; ++k) cc[k] = (int)col[k];
  ddps::AmgParams prm;
  prm.nthreads = nthreads;
  if (coarse_max > 0) prm.coarse_max = coarse_max;
  if (const char* e = getenv("DDPS_AMG_SMALL_LEVEL")) prm.small_level = atoi(e);   // experiments
  if (const char* e = getenv("DDPS_AMG_SMALL_CAP")) prm.small_cap = atoi(e);
  ddps::amg_host_build(*h, (int)n, std::move(rp), std::move(cc), std::move(vv), fixed, prm);
  *out = h;
  return DDPS_OK;
}

int ddps_amg_host_nlevels(const ddps_amg_host* h) { return h ? (int)h->lv.size() : DDPS_ERR_ARG; }

/* sizes = {n, nnz, nc (0 on the coarsest level), nnz(P)} */
int ddps_amg_host_level_sizes(const ddps_amg_host* h, int level, int64_t sizes[4]) {
  if (!h || level < 0 || level >= (int)h->lv.size() || !sizes) return DDPS_ERR_ARG;
  const ddps::AmgHostLevel& L = h->lv[level];
  sizes[0] = L.n;
  sizes[1] = (int64_t)L.col.size();
  sizes[2] = L.Pptr.empty() ? 0 : L.nc;
  sizes[3] = (int64_t)L.Pcol.size();
  return DDPS_OK;
}

int ddps_amg_host_level_get(const ddps_amg_host* h, int level, int64_t* rowptr, int64_t* col, double* val, int64_t* Pptr,
                            int64_t* Pcol, double* Pval, int64_t* agg) {
  if (!h || level < 0 || level >= (int)h->lv.size()) return DDPS_ERR_ARG;
  const ddps::AmgHostLevel& L = h->lv[level];
  if (rowptr) for (size_t i = 0; i < L.rowptr.size(); ++i) rowptr[i] = L.rowptr[i];
  if (col) for (size_t i = 0; i < L.col.size(); ++i) col[i] = L.col[i];
  if (val) for (size_t i = 0; i < L.val.size(); ++i) val[i] = L.val[i];
  if (Pptr) for (size_t i = 0; i < L.Pptr.size(); ++i) Pptr[i] = L.Pptr[i];
  if (Pcol) for (size_t i = 0; i < L.Pcol.size(); ++i) Pcol[i] = L.Pcol[i];
  if (Pval) for (size_t i = 0; i < L.Pval.size(); ++i) Pval[i] = L.Pval[i];
  if (agg) for (size_t i = 0; i < L.agg.size(); ++i) agg[i] = L.agg[i];
  return DDPS_OK;
}

void ddps_amg_host_free(ddps_amg_host* h) { delete h; }

}  // extern "C"
