// Device-side construction of one multigrid level.  The host reference of the same
// algorithm is amg_host.cu (tested against scipy on the CPU); this file does the work for the LARGE levels on the GPU
// so that only the sorted strong-neighbour lists go to the host (for the sequential greedy aggregation) and only
// the aggregate ids come back:
//   1. k_strong_*     strong-neighbour lists of the base operator, strongest first (count, scan, fill);
//   2. host           amg_aggregate_lists (shared with the host path) -> aggregate id per row;
//   3. k_prolong_*    P = (I - omega D^-1 A) T row by row, entries sorted by aggregate, values rounded to fp32,
//                     packed {aggregate, float} (count, scan, fill) -- same arithmetic, same order as the host;
//   4. cub sort       R = P^T row-wise: stable radix sort of the entries by aggregate keeps the fine indices of
//                     every coarse row increasing;
//   5. symbolic       coarse pattern of R A P: keys (I << 32 | J) of every structural product, radix sort, unique,
//                     lower_bound -> CSR with sorted columns (the same recipe as the fine pattern, setup.cu);
//   6. SELL-32 layout of the coarse pattern and the Galerkin product of the BASE values (k_amg_rap2 of amg.cu), which
//      the next level's strength / prolongator smoothing need.
// The arrays built here are installed in the hierarchy as they are (no host round trip, no upload).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <thread>

#include <cub/cub.cuh>

#include "amg_dev.cuh"

namespace ddps {

namespace {

constexpr int kMaxLocalRow = 64;   // longest matrix row the per-thread kernels of this file handle

// Temporaries come from the device's default stream-ordered memory pool (cudaMallocAsync on the library stream; every
// use is on that stream): a plain cudaFree of several GB is synchronous and was measured at up to 100 ms per level.
struct Tmp {
  cudaStream_t st;
  std::vector<void*> ptrs;
  explicit Tmp(cudaStream_t s) : st(s) {}
  ~Tmp() { for (void* p : ptrs) cudaFreeAsync(p, st); }
  template <typename T> cudaError_t alloc(T** p, size_t n) {
    cudaError_t e = cudaMallocAsync((void**)p, std::max<size_t>(n, 1) * sizeof(T), st);
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
};

inline int bits_for(uint64_t n) { int b = 1; while (b < 63 && (1ull << b) < n) ++b; return b; }

// ---- 1. strong-neighbour lists ------------------------------------------------------------------------------
// FILL == false: cnt[i] = number of strong neighbours.  FILL == true: scol[sptr[i] ...] = their columns sorted by
// (value ascending = strongest first, column ascending).
template <bool FILL>
__global__ void k_strong(int n, const int* __restrict__ rowptr, const int* __restrict__ slice_ptr, const int* __restrict__ col,
                         const double* __restrict__ val, const unsigned char* __restrict__ fixed, double theta,
                         int* __restrict__ cnt, const int* __restrict__ sptr, int* __restrict__ scol) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = 0;
  double sv[kMaxLocalRow];
  int sj[kMaxLocalRow];
  if (!(fixed && fixed[i])) {
    const int len = rowptr[i + 1] - rowptr[i];
    const int base = slice_ptr[i >> 5] + (i & 31);
    double m = 0.0;
    for (int k = 0; k < len; ++k) {
      const int j = col[base + 32 * k];
      if (j != i && !(fixed && fixed[j])) m = fmax(m, -val[base + 32 * k]);
    }
    if (m > 0.0) {
      const double cut = theta * m;
      for (int k = 0; k < len; ++k) {
        const int j = col[base + 32 * k];
        const double a = val[base + 32 * k];
        if (j != i && !(fixed && fixed[j]) && -a >= cut) {
          if (FILL) {   // insertion by (value, column); entries arrive in increasing column order
            int p = c;
            while (p > 0 && sv[p - 1] > a) { sv[p] = sv[p - 1]; sj[p] = sj[p - 1]; --p; }
            sv[p] = a; sj[p] = j;
          }
          ++c;
        }
      }
    }
  }
  if (!FILL) cnt[i] = c;
  else {
    const int o = sptr[i];
    for (int q = 0; q < c; ++q) scol[o + q] = sj[q];
  }
}

// ---- 3. prolongator -----------------------------------------------------------------------------------------
// One row of P = (I - omega D^-1 A) T (the arithmetic and its order are those of prolongator_row in amg_host.cu).
template <bool FILL>
__global__ void k_prolongator(int n, const int* __restrict__ rowptr, const int* __restrict__ slice_ptr,
                              const int* __restrict__ col, const double* __restrict__ val,
                              const unsigned char* __restrict__ fixed, const int* __restrict__ agg, double omega,
                              int* __restrict__ cnt, const int* __restrict__ Pptr, int2* __restrict__ Pent,
                              int* __restrict__ Pkey) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = 0;
  int ag[kMaxLocalRow + 1];
  double v[kMaxLocalRow + 1];
  if (!(fixed && fixed[i])) {
    const int len = rowptr[i + 1] - rowptr[i];
    const int base = slice_ptr[i >> 5] + (i & 31);
    double d = 0.0;
    for (int k = 0; k < len; ++k)
      if (col[base + 32 * k] == i) d = val[base + 32 * k];
    const int ai = agg[i];
    if (ai >= 0) { ag[0] = ai; v[0] = 1.0 - omega; c = 1; }
    if (d != 0.0) {
      for (int k = 0; k < len; ++k) {
        const int j = col[base + 32 * k];
        const double a = val[base + 32 * k];
        if (j == i || a == 0.0) continue;
        const int aj = agg[j];
        if (aj < 0) continue;
        const double w = -omega * a / d;
        int p = 0;
        while (p < c && ag[p] != aj) ++p;
        if (p == c) { ag[c] = aj; v[c] = w; ++c; }
        else v[p] += w;
      }
    }
    if (FILL) {
      for (int p = 1; p < c; ++p) {   // insertion sort by aggregate id
        const int a = ag[p];
        const double w = v[p];
        int q = p - 1;
        while (q >= 0 && ag[q] > a) { ag[q + 1] = ag[q]; v[q + 1] = v[q]; --q; }
        ag[q + 1] = a; v[q + 1] = w;
      }
    }
  }
  if (!FILL) cnt[i] = c;
  else {
    const int o = Pptr[i];
    for (int q = 0; q < c; ++q) {
      Pent[o + q] = make_int2(ag[q], __float_as_int(__double2float_rn(v[q])));
      Pkey[o + q] = ag[q];
    }
  }
}

// Pattern of row i of A*P: the sorted distinct coarse columns reachable through the (structural) entries of row i.
// Rows whose prolongator row is empty belong to no restriction row, so their A*P row is never read: left empty.
template <bool FILL>
__global__ void k_ap_pattern(int n, const int* __restrict__ rowptr, const int* __restrict__ slice_ptr, const int* __restrict__ col,
                             const int* __restrict__ Pptr, const int2* __restrict__ Pent, int* __restrict__ cnt,
                             const int* __restrict__ APptr, int* __restrict__ APcol, int* __restrict__ overflow) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int seen[kMaxLocalRow];
  int c = 0;
  if (Pptr[i + 1] > Pptr[i]) {
    const int len = rowptr[i + 1] - rowptr[i];
    const int base = slice_ptr[i >> 5] + (i & 31);
    for (int k = 0; k < len; ++k) {
      const int j = col[base + 32 * k];
      for (int m = Pptr[j]; m < Pptr[j + 1]; ++m) {
        const int J = Pent[m].x;
        int p = 0;
        while (p < c && seen[p] != J) ++p;
        if (p == c) {
          if (c == kMaxLocalRow) { *overflow = 1; break; }
          seen[c++] = J;
        }
      }
    }
  }
  if (!FILL) cnt[i] = c;
  else {
    for (int p = 1; p < c; ++p) {   // ascending columns
      const int J = seen[p];
      int q = p - 1;
      while (q >= 0 && seen[q] > J) { seen[q + 1] = seen[q]; --q; }
      seen[q + 1] = J;
    }
    const int o = APptr[i];
    for (int p = 0; p < c; ++p) APcol[o + p] = seen[p];
  }
}

// entry (fine row i, slot m of P) -> 64-bit value {row, float bits} laid out like an int2 {x = row, y = bits}
__global__ void k_transpose_values(int n, const int* __restrict__ Pptr, const int2* __restrict__ Pent,
                                   unsigned long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int m = Pptr[i]; m < Pptr[i + 1]; ++m)
    out[m] = ((unsigned long long)(unsigned)Pent[m].y << 32) | (unsigned long long)(unsigned)i;
}

template <typename K>
__global__ void k_lower_bound_sorted(const K* __restrict__ sorted, long long n, int nq, int shift, int* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)(sorted[mid] >> shift) < (long long)q) lo = mid + 1; else hi = mid;
  }
  out[q] = (int)lo;
}

// ---- 5. symbolic Galerkin product ---------------------------------------------------------------------------
// per entry q of R (coarse row I = rkey[q], fine row i): the DISTINCT coarse columns J reachable through the entries of
// fine row i (= the pattern of row i of A P), counted (FILL == false) or written as keys (I << 32 | J)
template <bool FILL>
__global__ void k_sym_keys(int nR, const int2* __restrict__ Rent, const int* __restrict__ rkey, const int* __restrict__ rowptr,
                           const int* __restrict__ slice_ptr, const int* __restrict__ col, const int* __restrict__ Pptr,
                           const int2* __restrict__ Pent, long long* __restrict__ cnt, const long long* __restrict__ off,
                           unsigned long long* __restrict__ keys, int* __restrict__ overflow) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nR) return;
  const int i = Rent[q].x;
  const int len = rowptr[i + 1] - rowptr[i];
  const int base = slice_ptr[i >> 5] + (i & 31);
  int seen[kMaxLocalRow];
  int c = 0;
  for (int k = 0; k < len; ++k) {
    const int j = col[base + 32 * k];
    for (int m = Pptr[j]; m < Pptr[j + 1]; ++m) {
      const int J = Pent[m].x;
      int p = 0;
      while (p < c && seen[p] != J) ++p;
      if (p == c) {
        if (c == kMaxLocalRow) { *overflow = 1; break; }
        seen[c++] = J;
      }
    }
  }
  if (!FILL) cnt[q] = c;
  else {
    const unsigned long long hi = (unsigned long long)(unsigned)rkey[q] << 32;
    long long o = off[q];
    for (int p = 0; p < c; ++p) keys[o++] = hi | (unsigned long long)(unsigned)seen[p];
  }
}

__global__ void k_low32_keys(const unsigned long long* __restrict__ keys, long long n, int* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)(keys[i] & 0xffffffffull);
}

// ---- 6. SELL-32 layout ----------------------------------------------------------------------------------------
__global__ void k_sell_width(const int* __restrict__ rowptr, int nrows, int nslices, long long* __restrict__ width32,
                             int* __restrict__ maxlen) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int slice = row >> 5;
  if (slice >= nslices) return;
  int m = (row < nrows) ? rowptr[row + 1] - rowptr[row] : 0;
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((row & 31) == 0) {
    width32[slice] = 32LL * m;
    atomicMax(maxlen, m);
  }
}
__global__ void k_narrow_ll(const long long* __restrict__ in, int n, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)in[i];
}
__global__ void k_sell_fill(const int* __restrict__ rowptr, const int* __restrict__ csrcol, int nrows, int nslices,
                            const int* __restrict__ slice_ptr, int* __restrict__ col) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int slice = row >> 5, lane = row & 31;
  if (slice >= nslices) return;
  const int base = slice_ptr[slice];
  const int w = (slice_ptr[slice + 1] - base) >> 5;
  int beg = 0, len = 0;
  if (row < nrows) { beg = rowptr[row]; len = rowptr[row + 1] - beg; }
  const int self = row < nrows ? row : 0;
  for (int k = 0; k < w; ++k) col[base + k * 32 + lane] = (k < len) ? csrcol[beg + k] : self;
}

void amg_release_transfers(AmgLevelDev& F) {
  cudaFree(F.Pptr); cudaFree(F.Pent); cudaFree(F.Rptr); cudaFree(F.Rent); cudaFree(F.APptr); cudaFree(F.APcol); cudaFree(F.APval);
  F.Pptr = F.Rptr = F.APptr = F.APcol = nullptr;
  F.APval = nullptr;
  F.Pent = F.Rent = nullptr;
  F.nc = 0;
}

template <typename T>
int scan_exclusive(Context* ctx, Tmp& tmp, const T* in, T* out, int n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, ctx->stream);
  void* ws;
  DDPS_CUDA(tmp.alloc((char**)&ws, bytes));
  DDPS_CUDA(cub::DeviceScan::ExclusiveSum(ws, bytes, in, out, n, ctx->stream));
  return 0;
}

}  // namespace

int amg_dev_coarsen(Context* ctx, Amg* A, int l, const unsigned char* d_fixed, const AmgParams& prm, int* nc_out) {
  *nc_out = 0;
  AmgLevelDev& F = A->lv[l];
  const int n = F.n;
  if (F.wmax > kMaxLocalRow) return 1;   // rows too long for the per-thread kernels: host path
  cudaStream_t st = ctx->stream;
  const int T = 256;
  const double* fval = F.cur_val;        // BASE values of this level (SELL order)
  {   // keep freed temporaries cached in the pool (the next level / the next solve reuses them)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  Tmp tmp(st);
  int rc;
  const bool timing = getenv("DDPS_AMG_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();

  // ---- 1. strong-neighbour lists -> host ----
  int *d_cnt, *d_sptr, *d_scol;
  DDPS_CUDA(tmp.alloc(&d_cnt, (size_t)n + 1));
  DDPS_CUDA(tmp.alloc(&d_sptr, (size_t)n + 1));
  DDPS_CUDA(cudaMemsetAsync(d_cnt, 0, ((size_t)n + 1) * sizeof(int), st));
  k_strong<false><<<(n + T - 1) / T, T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, fval, d_fixed, prm.theta, d_cnt, nullptr, nullptr);
  if ((rc = scan_exclusive(ctx, tmp, d_cnt, d_sptr, n + 1))) return rc;
  int ns = 0;
  DDPS_CUDA(cudaMemcpyAsync(&ns, d_sptr + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));
  DDPS_CUDA(tmp.alloc(&d_scol, (size_t)ns));
  k_strong<true><<<(n + T - 1) / T, T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, fval, d_fixed, prm.theta, nullptr, d_sptr, d_scol);
  std::vector<int> sptr((size_t)n + 1);
  DDPS_CUDA(cudaMemcpyAsync(sptr.data(), d_sptr, sptr.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));   // the lists are complete on the device, the row pointers on the host

  const double t1 = now();
  // ---- 2. aggregation (host, sequential greedy: the quality of the coarsening lives here).  The neighbour lists are
  //         copied by a producer thread in row chunks while the first, sequential pass already consumes them ----
  std::unique_ptr<int[]> scol(new int[(size_t)std::max(ns, 1)]);   // not value-initialised: the producer touches the pages
  std::atomic<long long> rows_ready(0);
  std::atomic<int> copy_err(0);
  const int dev = ctx->device;
  std::thread producer([&]() {
    cudaSetDevice(dev);
    const int nchunk = 32;
    for (int c = 0; c < nchunk; ++c) {
      const long long r0 = (long long)n * c / nchunk, r1 = (long long)n * (c + 1) / nchunk;
      const size_t o0 = (size_t)sptr[r0], o1 = (size_t)sptr[r1];
      if (o1 > o0 && cudaMemcpy(scol.get() + o0, d_scol + o0, (o1 - o0) * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
        copy_err.store(1);
      rows_ready.store(r1, std::memory_order_release);
    }
  });
  std::vector<int> agg;
  const int nc = amg_aggregate_lists(n, sptr.data(), scol.get(), amg_small_cap(prm, n), agg, &rows_ready);
  producer.join();
  if (copy_err.load()) { ctx->fail(DDPS_ERR_CUDA, "copy of the strong-neighbour lists failed"); return DDPS_ERR_CUDA; }
  if (nc <= 0 || nc > 0.7 * n) return 0;   // coarsening stalled
  std::vector<int>().swap(sptr);
  scol.reset();
  const double t2 = now();
  int* d_agg;
  DDPS_CUDA(tmp.alloc(&d_agg, (size_t)n));
  DDPS_CUDA(cudaMemcpyAsync(d_agg, agg.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));

  // ---- 3. prolongator ----
  int* Pptr = nullptr;
  int2* Pent = nullptr;
  DDPS_CUDA(cudaMalloc(&Pptr, ((size_t)n + 1) * sizeof(int)));
  F.Pptr = Pptr;   // owned by the level from here on (amg_free releases it)
  DDPS_CUDA(cudaMemsetAsync(d_cnt, 0, ((size_t)n + 1) * sizeof(int), st));
  k_prolongator<false><<<(n + T - 1) / T, T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, fval, d_fixed, d_agg, prm.omega, d_cnt,
                                                     nullptr, nullptr, nullptr);
  if ((rc = scan_exclusive(ctx, tmp, d_cnt, Pptr, n + 1))) return rc;
  int nP = 0;
  DDPS_CUDA(cudaMemcpyAsync(&nP, Pptr + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));   // also: agg (host vector) has been consumed
  DDPS_CUDA(cudaMalloc(&Pent, std::max<size_t>(nP, 1) * sizeof(int2)));
  F.Pent = Pent;
  int* d_pkey;
  DDPS_CUDA(tmp.alloc(&d_pkey, (size_t)nP));
  k_prolongator<true><<<(n + T - 1) / T, T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, fval, d_fixed, d_agg, prm.omega, nullptr,
                                                    Pptr, Pent, d_pkey);

  // ---- 4. R = P^T by a stable sort on the aggregate id ----
  unsigned long long* d_pv;
  int* d_rkey;
  int2* Rent = nullptr;
  int* Rptr = nullptr;
  DDPS_CUDA(tmp.alloc(&d_pv, (size_t)nP));
  DDPS_CUDA(tmp.alloc(&d_rkey, (size_t)nP));
  DDPS_CUDA(cudaMalloc(&Rent, std::max<size_t>(nP, 1) * sizeof(int2)));
  F.Rent = Rent;
  DDPS_CUDA(cudaMalloc(&Rptr, ((size_t)nc + 1) * sizeof(int)));
  F.Rptr = Rptr;
  k_transpose_values<<<(n + T - 1) / T, T, 0, st>>>(n, Pptr, Pent, d_pv);
  {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_pkey, d_rkey, d_pv, (unsigned long long*)Rent, nP, 0, bits_for(nc), st);
    void* ws;
    DDPS_CUDA(tmp.alloc((char**)&ws, bytes));
    DDPS_CUDA(cub::DeviceRadixSort::SortPairs(ws, bytes, d_pkey, d_rkey, d_pv, (unsigned long long*)Rent, nP, 0, bits_for(nc), st));
  }
  k_lower_bound_sorted<int><<<(nc + 1 + T - 1) / T, T, 0, st>>>(d_rkey, nP, nc + 1, 0, Rptr);

  // ---- 4b. pattern of A*P for the two-stage Galerkin product (skipped when a row is too long: one-stage kernel) ----
  {
    int* d_ovf2;
    DDPS_CUDA(tmp.alloc(&d_ovf2, 4));
    DDPS_CUDA(cudaMemsetAsync(d_ovf2, 0, sizeof(int), st));
    DDPS_CUDA(cudaMemsetAsync(d_cnt, 0, ((size_t)n + 1) * sizeof(int), st));
    k_ap_pattern<false><<<(n + T - 1) / T, T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, Pptr, Pent, d_cnt, nullptr, nullptr, d_ovf2);
    DDPS_CUDA(cudaMalloc(&F.APptr, ((size_t)n + 1) * sizeof(int)));
    if ((rc = scan_exclusive(ctx, tmp, d_cnt, F.APptr, n + 1))) return rc;
    int nAP = 0, ovf2 = 0;
    DDPS_CUDA(cudaMemcpyAsync(&nAP, F.APptr + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaMemcpyAsync(&ovf2, d_ovf2, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    if (ovf2 || getenv("DDPS_AMG_ONE_STAGE")) {
      cudaFree(F.APptr);
      F.APptr = nullptr;
    } else {
      F.nAP = nAP;
      DDPS_CUDA(cudaMalloc(&F.APcol, std::max<size_t>(nAP, 1) * sizeof(int)));
      DDPS_CUDA(cudaMalloc(&F.APval, std::max<size_t>(nAP, 1) * sizeof(double)));
      k_ap_pattern<true><<<(n + T - 1) / T, T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, Pptr, Pent, nullptr, F.APptr, F.APcol, d_ovf2);
    }
  }

  if (timing) DDPS_CUDA(cudaStreamSynchronize(st));
  const double t3 = now();
  // ---- 5. symbolic Galerkin product: coarse CSR pattern with sorted columns ----
  long long *d_qcnt, *d_qoff;
  DDPS_CUDA(tmp.alloc(&d_qcnt, (size_t)nP + 1));
  DDPS_CUDA(tmp.alloc(&d_qoff, (size_t)nP + 1));
  DDPS_CUDA(cudaMemsetAsync(d_qcnt, 0, ((size_t)nP + 1) * sizeof(long long), st));
  int* d_ovf;
  DDPS_CUDA(tmp.alloc(&d_ovf, 4));
  DDPS_CUDA(cudaMemsetAsync(d_ovf, 0, sizeof(int), st));
  k_sym_keys<false><<<(nP + T - 1) / T, T, 0, st>>>(nP, Rent, d_rkey, F.rowptr, F.slice_ptr, F.col, Pptr, Pent, d_qcnt, nullptr, nullptr, d_ovf);
  if ((rc = scan_exclusive(ctx, tmp, d_qcnt, d_qoff, nP + 1))) return rc;
  long long nk = 0;
  int ovf = 0;
  DDPS_CUDA(cudaMemcpyAsync(&nk, d_qoff + nP, sizeof(long long), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaMemcpyAsync(&ovf, d_ovf, sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));
  if (ovf || nk >= (1LL << 31) - 1) {   // a row of A P with more than kMaxLocalRow entries, or too many keys: host path
    amg_release_transfers(F);
    return 1;
  }
  unsigned long long *d_k0, *d_k1;
  int* d_num;
  DDPS_CUDA(tmp.alloc(&d_k0, (size_t)nk));
  DDPS_CUDA(tmp.alloc(&d_k1, (size_t)nk));
  DDPS_CUDA(tmp.alloc(&d_num, 4));
  k_sym_keys<true><<<(nP + T - 1) / T, T, 0, st>>>(nP, Rent, d_rkey, F.rowptr, F.slice_ptr, F.col, Pptr, Pent, nullptr, d_qoff, d_k0, d_ovf);
  {
    const int endbit = 32 + bits_for(nc);
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, d_k0, d_k1, (int)nk, 0, endbit, st);
    void* ws;
    DDPS_CUDA(tmp.alloc((char**)&ws, bytes));
    DDPS_CUDA(cub::DeviceRadixSort::SortKeys(ws, bytes, d_k0, d_k1, (int)nk, 0, endbit, st));
    size_t b2 = 0;
    cub::DeviceSelect::Unique(nullptr, b2, d_k1, d_k0, d_num, (int)nk, st);
    void* ws2;
    DDPS_CUDA(tmp.alloc((char**)&ws2, b2));
    DDPS_CUDA(cub::DeviceSelect::Unique(ws2, b2, d_k1, d_k0, d_num, (int)nk, st));
  }
  int cnnz = 0;
  DDPS_CUDA(cudaMemcpyAsync(&cnnz, d_num, sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));

  AmgLevelDev& C = A->lv[l + 1];
  C = AmgLevelDev();
  C.n = nc;
  C.nnz = cnnz;
  C.owns = true;
  C.nslices = (nc + 31) / 32;
  DDPS_CUDA(cudaMalloc(&C.rowptr, ((size_t)nc + 1) * sizeof(int)));
  k_lower_bound_sorted<unsigned long long><<<(nc + 1 + T - 1) / T, T, 0, st>>>(d_k0, cnnz, nc + 1, 32, C.rowptr);
  int* d_ccol;
  DDPS_CUDA(tmp.alloc(&d_ccol, (size_t)cnnz));
  k_low32_keys<<<(unsigned)((cnnz + T - 1) / T), T, 0, st>>>(d_k0, cnnz, d_ccol);

  if (timing) DDPS_CUDA(cudaStreamSynchronize(st));
  const double t4 = now();
  // ---- 6. SELL-32 layout + Galerkin product of the base values ----
  {
    long long *w32, *w32scan;
    int* d_max;
    DDPS_CUDA(tmp.alloc(&w32, (size_t)C.nslices + 1));
    DDPS_CUDA(tmp.alloc(&w32scan, (size_t)C.nslices + 1));
    DDPS_CUDA(tmp.alloc(&d_max, 4));
    DDPS_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), st));
    DDPS_CUDA(cudaMemsetAsync(w32, 0, ((size_t)C.nslices + 1) * sizeof(long long), st));
    k_sell_width<<<(C.nslices * 32 + T - 1) / T, T, 0, st>>>(C.rowptr, nc, C.nslices, w32, d_max);
    if ((rc = scan_exclusive(ctx, tmp, w32, w32scan, C.nslices + 1))) return rc;
    long long total = 0;
    DDPS_CUDA(cudaMemcpyAsync(&total, w32scan + C.nslices, sizeof(long long), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaMemcpyAsync(&C.wmax, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    if (total >= (1LL << 31)) { ctx->fail(DDPS_ERR_ARG, "multigrid level too large for 32-bit entry offsets"); return DDPS_ERR_ARG; }
    C.nentries = total;
    DDPS_CUDA(cudaMalloc(&C.slice_ptr, ((size_t)C.nslices + 1) * sizeof(int)));
    k_narrow_ll<<<(C.nslices + 1 + T - 1) / T, T, 0, st>>>(w32scan, C.nslices + 1, C.slice_ptr);
    DDPS_CUDA(cudaMalloc(&C.col, std::max<size_t>((size_t)C.nentries, 1) * sizeof(int)));
    k_sell_fill<<<(C.nslices * 32 + T - 1) / T, T, 0, st>>>(C.rowptr, d_ccol, nc, C.nslices, C.slice_ptr, C.col);
  }
  DDPS_CUDA(cudaMalloc(&C.val, std::max<size_t>((size_t)C.nentries, 1) * sizeof(double)));
  DDPS_CUDA(cudaMalloc(&C.dinv, (size_t)nc * sizeof(double)));
  DDPS_CUDA(cudaMemsetAsync(C.val, 0, std::max<size_t>((size_t)C.nentries, 1) * sizeof(double), st));
  DDPS_CUDA(cudaMemsetAsync(C.dinv, 0, (size_t)nc * sizeof(double), st));
  F.nc = nc;
  amg_launch_rap(ctx, F, fval, C, C.val, C.dinv);
  C.cur_val = C.val;     // base values of the next level (overwritten by the first per-operator numeric phase)
  C.cur_dinv = C.dinv;
  DDPS_CUDA(cudaStreamSynchronize(st));
  DDPS_CUDA(cudaGetLastError());
  *nc_out = nc;
  if (timing)
    fprintf(stderr, "[ddps amg]   device level n=%d: strong lists + D2H %.1f ms, host aggregation %.1f, P + R %.1f, symbolic (%lld keys) %.1f, "
                    "SELL + base Galerkin %.1f\n", n, t1 - t0, t2 - t1, t3 - t2, nk, t4 - t3, now() - t4);
  return 0;
}

}  // namespace ddps
