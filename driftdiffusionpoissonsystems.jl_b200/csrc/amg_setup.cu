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

// Coarse level from the sorted distinct keys (row << 32 | column) of its pattern: CSR row pointers, SELL-32 layout (the
// recipe of the fine pattern, setup.cu), value / inverse-diagonal arrays (zeroed: SELL padding stays 0).
int finish_coarse_level(Context* ctx, Tmp& tmp, const unsigned long long* d_keys, int cnnz, int nc, AmgLevelDev& C) {
  cudaStream_t st = ctx->stream;
  const int T = 256;
  int rc;
  C = AmgLevelDev();
  C.n = nc;
  C.n_loc = nc;
  C.nnz = cnnz;
  C.owns = true;
  C.nslices = (nc + 31) / 32;
  DDPS_CUDA(cudaMalloc(&C.rowptr, ((size_t)nc + 1) * sizeof(int)));
  k_lower_bound_sorted<unsigned long long><<<(nc + 1 + T - 1) / T, T, 0, st>>>(d_keys, cnnz, nc + 1, 32, C.rowptr);
  int* d_ccol;
  DDPS_CUDA(tmp.alloc(&d_ccol, (size_t)cnnz));
  if (cnnz) k_low32_keys<<<(unsigned)((cnnz + T - 1) / T), T, 0, st>>>(d_keys, cnnz, d_ccol);
  long long *w32, *w32scan;
  int* d_max;
  DDPS_CUDA(tmp.alloc(&w32, (size_t)C.nslices + 1));
  DDPS_CUDA(tmp.alloc(&w32scan, (size_t)C.nslices + 1));
  DDPS_CUDA(tmp.alloc(&d_max, 4));
  DDPS_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), st));
  DDPS_CUDA(cudaMemsetAsync(w32, 0, ((size_t)C.nslices + 1) * sizeof(long long), st));
  if (C.nslices) k_sell_width<<<(C.nslices * 32 + T - 1) / T, T, 0, st>>>(C.rowptr, nc, C.nslices, w32, d_max);
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
  if (C.nslices) k_sell_fill<<<(C.nslices * 32 + T - 1) / T, T, 0, st>>>(C.rowptr, d_ccol, nc, C.nslices, C.slice_ptr, C.col);
  DDPS_CUDA(cudaMalloc(&C.val, std::max<size_t>((size_t)C.nentries, 1) * sizeof(double)));
  DDPS_CUDA(cudaMalloc(&C.dinv, std::max<size_t>((size_t)nc, 1) * sizeof(double)));
  DDPS_CUDA(cudaMemsetAsync(C.val, 0, std::max<size_t>((size_t)C.nentries, 1) * sizeof(double), st));
  DDPS_CUDA(cudaMemsetAsync(C.dinv, 0, std::max<size_t>((size_t)nc, 1) * sizeof(double), st));
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
  if (timing) DDPS_CUDA(cudaStreamSynchronize(st));
  const double t4 = now();
  if ((rc = finish_coarse_level(ctx, tmp, d_k0, cnnz, nc, C))) return rc;
  F.nc = nc;
  if ((rc = amg_launch_rap(ctx, F, fval, C, C.val, C.dinv))) return rc;
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


// ================================================================================================================
// Distributed hierarchy (multi-rank contexts; VERDICT round 1 item 2, blueprint: tools/amg_dist_prototype.py)
// ================================================================================================================
namespace {

__global__ void k_ghostfix(int nl, int n, const unsigned char* __restrict__ isD, unsigned char* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nl) out[i] = i >= n ? 1 : (isD ? isD[i] : 0);
}
__global__ void k_global_ids(int nl, int n, const int* __restrict__ loc, int off, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nl) return;
  if (i >= n) { out[i] = -1; return; }
  const int a = loc ? loc[i] : i;
  out[i] = a >= 0 ? off + a : -1;
}
__global__ void k_row_len(int n, const int* __restrict__ ptr, int* __restrict__ len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) len[i] = ptr[i + 1] - ptr[i];
}
__global__ void k_add_base(int m, const int* __restrict__ incl, const int* __restrict__ base, int* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) out[j] = *base + incl[j];
}
__global__ void k_gather_int(int m, const int* __restrict__ idx, const int* __restrict__ src, int* __restrict__ dst) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) dst[j] = src[idx[j]];
}
// send row k = local row send_idx[k]: its entries go to buf[send_ptr[k] ...]; pos remembers where they came from
__global__ void k_pack_rows(int nsend, const int* __restrict__ send_idx, const int* __restrict__ ptr, const int* __restrict__ send_ptr,
                            const char* __restrict__ ent, char* __restrict__ buf, int* __restrict__ pos, int elem) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nsend) return;
  const int i = send_idx[k];
  int o = send_ptr[k];
  for (int e = ptr[i]; e < ptr[i + 1]; ++e, ++o) {
    for (int b = 0; b < elem; ++b) buf[(size_t)o * elem + b] = ent[(size_t)e * elem + b];
    if (pos) pos[o] = e;
  }
}

struct RowXchg {
  std::vector<int> send_ent_off, recv_ent_off;   // [nneigh+1] entry offsets: packed send buffer / positions in the row array
  int* send_pos = nullptr;                       // device, cudaMalloc'ed (kept by the caller)
  int nsend_ent = 0;
};

// Variable-length rows of the neighbours' boundary nodes (= my ghost nodes), in ghost order: row lengths by one halo of
// ints, then the packed entries.  ptr [nl+1]: entries 0..n valid on entry, all valid on return; *ent grows from ptr[n] to
// ptr[nl] elements of `elem` bytes (reallocated with cudaMalloc).  Collective.
int row_exchange(Context* ctx, Tmp& tmp, const HaloPlan& H, int n, int nl, int* ptr, void** ent, size_t elem, RowXchg* keep) {
  cudaStream_t st = ctx->stream;
  const int T = 256;
  const int ng = nl - n;
  int rc;
  int* d_len;
  DDPS_CUDA(tmp.alloc(&d_len, (size_t)nl + 1));
  DDPS_CUDA(cudaMemsetAsync(d_len, 0, ((size_t)nl + 1) * sizeof(int), st));
  if (n) k_row_len<<<(n + T - 1) / T, T, 0, st>>>(n, ptr, d_len);
  if ((rc = halo_exchange_bytes(ctx, H, d_len, sizeof(int)))) return rc;
  if (ng > 0) {
    int* d_incl;
    DDPS_CUDA(tmp.alloc(&d_incl, (size_t)ng));
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, bytes, d_len + n, d_incl, ng, st);
    void* ws;
    DDPS_CUDA(tmp.alloc((char**)&ws, bytes));
    DDPS_CUDA(cub::DeviceScan::InclusiveSum(ws, bytes, d_len + n, d_incl, ng, st));
    k_add_base<<<(ng + T - 1) / T, T, 0, st>>>(ng, d_incl, ptr + n, ptr + n + 1);
  }
  const int ns = H.nsend;
  int *d_slen, *d_sptr;
  DDPS_CUDA(tmp.alloc(&d_slen, (size_t)ns + 1));
  DDPS_CUDA(tmp.alloc(&d_sptr, (size_t)ns + 1));
  DDPS_CUDA(cudaMemsetAsync(d_slen, 0, ((size_t)ns + 1) * sizeof(int), st));
  if (ns) k_gather_int<<<(ns + T - 1) / T, T, 0, st>>>(ns, H.send_idx, d_len, d_slen);
  if ((rc = scan_exclusive(ctx, tmp, d_slen, d_sptr, ns + 1))) return rc;
  std::vector<int> h_sptr((size_t)ns + 1), h_gptr((size_t)ng + 1);
  DDPS_CUDA(cudaMemcpyAsync(h_sptr.data(), d_sptr, h_sptr.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaMemcpyAsync(h_gptr.data(), ptr + n, h_gptr.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));
  const int own_cnt = h_gptr[0], total = h_gptr[ng], total_send = h_sptr[ns];
  char* nent = nullptr;
  DDPS_CUDA(cudaMalloc(&nent, std::max<size_t>((size_t)total, 1) * elem));
  if (own_cnt) DDPS_CUDA(cudaMemcpyAsync(nent, *ent, (size_t)own_cnt * elem, cudaMemcpyDeviceToDevice, st));
  char* d_buf;
  DDPS_CUDA(tmp.alloc(&d_buf, std::max<size_t>((size_t)total_send, 1) * elem));
  int* d_pos = nullptr;
  if (keep) DDPS_CUDA(cudaMalloc(&d_pos, std::max<size_t>((size_t)total_send, 1) * sizeof(int)));
  if (ns) k_pack_rows<<<(ns + T - 1) / T, T, 0, st>>>(ns, H.send_idx, ptr, d_sptr, (const char*)*ent, d_buf, d_pos, (int)elem);
  std::vector<CommMsg> msgs;
  for (int j = 0; j < H.nneigh; ++j) {
    CommMsg m;
    m.peer = H.peer[j];
    m.send = d_buf + (size_t)h_sptr[H.send_off[j]] * elem;
    m.send_bytes = (size_t)(h_sptr[H.send_off[j + 1]] - h_sptr[H.send_off[j]]) * elem;
    m.recv = nent + (size_t)h_gptr[H.recv_off[j]] * elem;
    m.recv_bytes = (size_t)(h_gptr[H.recv_off[j + 1]] - h_gptr[H.recv_off[j]]) * elem;
    msgs.push_back(m);
  }
  if ((rc = comm_exchange(ctx, msgs.data(), (int)msgs.size()))) return rc;
  DDPS_CUDA(cudaStreamSynchronize(st));
  cudaFree(*ent);
  *ent = nent;
  if (keep) {
    keep->send_pos = d_pos;
    keep->nsend_ent = total_send;
    keep->send_ent_off.assign(H.nneigh + 1, 0);
    keep->recv_ent_off.assign(H.nneigh + 1, 0);
    for (int j = 0; j <= H.nneigh; ++j) { keep->send_ent_off[j] = h_sptr[H.send_off[j]]; keep->recv_ent_off[j] = h_gptr[H.recv_off[j]]; }
  }
  return 0;
}

// coarse pattern keys from the A*P rows: entry q of R (own coarse row I = rkey[q], fine row i, own or ghost) contributes
// (I << 32 | J) for every J of row i of A*P
template <bool FILL>
__global__ void k_sym_keys_ap(int nR, const int2* __restrict__ Rent, const int* __restrict__ rkey, const int* __restrict__ APptr,
                              const int* __restrict__ APcol, long long* __restrict__ cnt, const long long* __restrict__ off,
                              unsigned long long* __restrict__ keys) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nR) return;
  const int i = Rent[q].x;
  const int e0 = APptr[i], e1 = APptr[i + 1];
  if (!FILL) { cnt[q] = e1 - e0; return; }
  const unsigned long long hi = (unsigned long long)(unsigned)rkey[q] << 32;
  long long o = off[q];
  for (int e = e0; e < e1; ++e) keys[o++] = hi | (unsigned long long)(unsigned)APcol[e];
}

// restriction keys: P entry (fine row i, global coarse column J) belongs to my coarse row J - off when I own J, else to
// the sentinel row nc (cut off after the sort)
__global__ void k_restrict_keys(int nP, const int2* __restrict__ Pent, int off, int nc, int* __restrict__ key) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nP) return;
  const int J = Pent[m].x;
  key[m] = (J >= off && J < off + nc) ? J - off : nc;
}

// candidates for the ghost coarse set: every non-owned global coarse id an owned row needs (0x7fffffff = no candidate)
__global__ void k_cand_int(long long cnt, const int* __restrict__ ids, int off, int nc, int* __restrict__ out) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const int J = ids[k];
  out[k] = (J < 0 || (J >= off && J < off + nc)) ? 0x7fffffff : J;
}
__global__ void k_cand_int2(long long cnt, const int2* __restrict__ ent, int off, int nc, int* __restrict__ out) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const int J = ent[k].x;
  out[k] = (J < 0 || (J >= off && J < off + nc)) ? 0x7fffffff : J;
}
__global__ void k_cand_keys(long long cnt, const unsigned long long* __restrict__ keys, int off, int nc, int* __restrict__ out) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const int J = (int)(keys[k] & 0xffffffffull);
  out[k] = (J >= off && J < off + nc) ? 0x7fffffff : J;
}

// global coarse id -> local id: owned first (J - off), then the sorted ghost list; ids this rank has no use for -> -2
__device__ __forceinline__ int map_coarse(int J, int off, int nc, const int* __restrict__ gl, int ng) {
  if (J < 0) return J;
  if (J >= off && J < off + nc) return J - off;
  int lo = 0, hi = ng;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (gl[mid] < J) lo = mid + 1; else hi = mid;
  }
  return (lo < ng && gl[lo] == J) ? nc + lo : -2;
}
__global__ void k_relabel_int(long long cnt, int* __restrict__ ids, int off, int nc, const int* __restrict__ gl, int ng) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < cnt) ids[k] = map_coarse(ids[k], off, nc, gl, ng);
}
__global__ void k_relabel_int2(long long cnt, int2* __restrict__ ent, int off, int nc, const int* __restrict__ gl, int ng) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < cnt) ent[k].x = map_coarse(ent[k].x, off, nc, gl, ng);
}
__global__ void k_relabel_keys(long long cnt, unsigned long long* __restrict__ keys, int off, int nc, const int* __restrict__ gl, int ng) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const unsigned long long key = keys[k];
  const int J = map_coarse((int)(key & 0xffffffffull), off, nc, gl, ng);
  keys[k] = (key & 0xffffffff00000000ull) | (unsigned long long)(unsigned)J;
}

__global__ void k_gather_double(int n, const int* __restrict__ idx, const double* __restrict__ src, double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

inline unsigned nblk(long long n, int T) { return (unsigned)std::max<long long>(1, (n + T - 1) / T); }

}  // namespace

// The A*P values of my boundary rows go to the neighbours that hold them as ghost rows (fixed pattern: a halo of values)
int amg_dist_ap_halo(Context* ctx, AmgLevelDev& F) {
  if (F.dist != 1) return 0;
  const HaloPlan& H = *F.plan;
  if (F.ap_nsend > 0) {
    k_gather_double<<<(F.ap_nsend + 255) / 256, 256, 0, ctx->stream>>>(F.ap_nsend, F.ap_send_pos, F.APval, F.ap_send_buf);
    ctx->stats.kernel_launches++;
  }
  std::vector<CommMsg> msgs;
  for (int j = 0; j < H.nneigh; ++j) {
    CommMsg m;
    m.peer = H.peer[j];
    m.send = F.ap_send_buf + F.ap_send_off[j];
    m.send_bytes = (size_t)(F.ap_send_off[j + 1] - F.ap_send_off[j]) * sizeof(double);
    m.recv = F.APval + F.ap_recv_off[j];
    m.recv_bytes = (size_t)(F.ap_recv_off[j + 1] - F.ap_recv_off[j]) * sizeof(double);
    msgs.push_back(m);
  }
  return comm_exchange(ctx, msgs.data(), (int)msgs.size());
}

// One distributed coarsening step of level l (collective).  Aggregates never cross a rank boundary (the strong-neighbour
// lists ignore ghost columns), everything else is the single-GPU algorithm on the owned rows plus three exchanges on the
// level's halo plan: aggregate ids of the ghost nodes, their prolongator rows, the pattern of their A*P rows.
int amg_dist_coarsen(Context* ctx, Amg* A, int l, const AmgParams& prm, long long* ncg_out) {
  *ncg_out = 0;
  AmgLevelDev& F = A->lv[l];
  const int n = F.n, nl = F.n_loc;
  const HaloPlan& H = *F.plan;
  const int me = ctx->rank, nr = ctx->nranks;
  cudaStream_t st = ctx->stream;
  const int T = 256;
  const double* fval = F.cur_val;
  Tmp tmp(st);
  int rc;
  // every rank must take the device path or none does
  {
    int ok_me = F.wmax <= kMaxLocalRow ? 1 : 0;
    std::vector<int> all(nr);
    if ((rc = comm_allgather_host(ctx, &ok_me, sizeof(int), all.data()))) return rc;
    for (int r = 0; r < nr; ++r) if (!all[r]) return 1;
  }
  // ---- fixed flags: Dirichlet rows (finest level) and every ghost column ----
  unsigned char* d_fix;
  DDPS_CUDA(tmp.alloc(&d_fix, (size_t)nl));
  if (nl) k_ghostfix<<<nblk(nl, T), T, 0, st>>>(nl, n, l == 0 ? ctx->isD : nullptr, d_fix);
  // ---- 1. strong-neighbour lists of the owned rows -> host ----
  int *d_cnt, *d_sptr, *d_scol;
  DDPS_CUDA(tmp.alloc(&d_cnt, (size_t)nl + 2));
  DDPS_CUDA(tmp.alloc(&d_sptr, (size_t)n + 1));
  DDPS_CUDA(cudaMemsetAsync(d_cnt, 0, ((size_t)nl + 2) * sizeof(int), st));
  if (n) k_strong<false><<<nblk(n, T), T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, fval, d_fix, prm.theta, d_cnt, nullptr, nullptr);
  if ((rc = scan_exclusive(ctx, tmp, d_cnt, d_sptr, n + 1))) return rc;
  std::vector<int> sptr((size_t)n + 1);
  DDPS_CUDA(cudaMemcpyAsync(sptr.data(), d_sptr, sptr.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));
  const int ns = sptr[n];
  DDPS_CUDA(tmp.alloc(&d_scol, (size_t)ns));
  if (n) k_strong<true><<<nblk(n, T), T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, fval, d_fix, prm.theta, nullptr, d_sptr, d_scol);
  std::vector<int> scol((size_t)std::max(ns, 1));
  DDPS_CUDA(cudaMemcpyAsync(scol.data(), d_scol, (size_t)ns * sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));
  // ---- 2. rank-local greedy aggregation (host) ----
  std::vector<int> agg;
  int nc = amg_aggregate_lists(n, sptr.data(), scol.data(), 0, agg);
  if (nc < 0) nc = 0;
  std::vector<int>().swap(sptr);
  std::vector<int>().swap(scol);
  // ---- 3. global numbering of the aggregates ----
  long long cnt_me[2] = {nc, n};
  std::vector<long long> cnt_all(2 * (size_t)nr);
  if ((rc = comm_allgather_host(ctx, cnt_me, sizeof(cnt_me), cnt_all.data()))) return rc;
  std::vector<long long> coff((size_t)nr + 1, 0);
  long long nfg = 0;
  for (int r = 0; r < nr; ++r) { coff[r + 1] = coff[r] + cnt_all[2 * r]; nfg += cnt_all[2 * r + 1]; }
  const long long ncg = coff[nr];
  if (getenv("DDPS_AMG_TIMING")) fprintf(stderr, "[ddps amg] rank %d level %d: %d owned rows (%d local columns), %d strong couplings, %d aggregates; %lld of %lld in total\n",
                                         me, l, n, nl, ns, nc, ncg, nfg);
  if (ncg <= 0 || ncg > 0.7 * (double)nfg) return 0;   // coarsening stalled (same decision on every rank)
  if (ncg >= (1LL << 31) - 2) { ctx->fail(DDPS_ERR_ARG, "multigrid level too large for 32-bit global ids"); return DDPS_ERR_ARG; }
  const int off = (int)coff[me];
  int *d_agg, *d_gagg;
  DDPS_CUDA(tmp.alloc(&d_agg, (size_t)std::max(n, 1)));
  DDPS_CUDA(tmp.alloc(&d_gagg, (size_t)std::max(nl, 1)));
  if (n) DDPS_CUDA(cudaMemcpyAsync(d_agg, agg.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  if (nl) k_global_ids<<<nblk(nl, T), T, 0, st>>>(nl, n, d_agg, off, d_gagg);
  if ((rc = halo_exchange_bytes(ctx, H, d_gagg, sizeof(int)))) return rc;   // aggregate ids of the ghost nodes
  DDPS_CUDA(cudaStreamSynchronize(st));   // agg (host vector) consumed

  // ---- 4. prolongator rows of the owned nodes (GLOBAL coarse ids), then the ghost nodes' rows from their owners ----
  DDPS_CUDA(cudaMalloc(&F.Pptr, ((size_t)nl + 1) * sizeof(int)));
  DDPS_CUDA(cudaMemsetAsync(F.Pptr, 0, ((size_t)nl + 1) * sizeof(int), st));
  DDPS_CUDA(cudaMemsetAsync(d_cnt, 0, ((size_t)nl + 2) * sizeof(int), st));
  if (n) k_prolongator<false><<<nblk(n, T), T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, fval, d_fix, d_gagg, prm.omega, d_cnt, nullptr, nullptr, nullptr);
  if ((rc = scan_exclusive(ctx, tmp, d_cnt, F.Pptr, n + 1))) return rc;
  int nP_own = 0;
  DDPS_CUDA(cudaMemcpyAsync(&nP_own, F.Pptr + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));
  DDPS_CUDA(cudaMalloc(&F.Pent, std::max<size_t>(nP_own, 1) * sizeof(int2)));
  {
    int* d_pkey;
    DDPS_CUDA(tmp.alloc(&d_pkey, (size_t)std::max(nP_own, 1)));
    if (n) k_prolongator<true><<<nblk(n, T), T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, fval, d_fix, d_gagg, prm.omega, nullptr, F.Pptr, F.Pent, d_pkey);
  }
  if ((rc = row_exchange(ctx, tmp, H, n, nl, F.Pptr, (void**)&F.Pent, sizeof(int2), nullptr))) return rc;
  int nP = 0;
  DDPS_CUDA(cudaMemcpyAsync(&nP, F.Pptr + nl, sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));

  // ---- 5. R = P^T restricted to my aggregates: stable sort of (own-or-sentinel key, {fine row, value}) ----
  int* d_rkey;
  {
    unsigned long long *d_pv, *d_pvs;
    int* d_key;
    DDPS_CUDA(tmp.alloc(&d_pv, (size_t)std::max(nP, 1)));
    DDPS_CUDA(tmp.alloc(&d_pvs, (size_t)std::max(nP, 1)));
    DDPS_CUDA(tmp.alloc(&d_key, (size_t)std::max(nP, 1)));
    DDPS_CUDA(tmp.alloc(&d_rkey, (size_t)std::max(nP, 1)));
    if (nl) k_transpose_values<<<nblk(nl, T), T, 0, st>>>(nl, F.Pptr, F.Pent, d_pv);
    if (nP) k_restrict_keys<<<nblk(nP, T), T, 0, st>>>(nP, F.Pent, off, nc, d_key);
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_key, d_rkey, d_pv, d_pvs, nP, 0, bits_for((uint64_t)nc + 1), st);
    void* ws;
    DDPS_CUDA(tmp.alloc((char**)&ws, bytes));
    DDPS_CUDA(cub::DeviceRadixSort::SortPairs(ws, bytes, d_key, d_rkey, d_pv, d_pvs, nP, 0, bits_for((uint64_t)nc + 1), st));
    DDPS_CUDA(cudaMalloc(&F.Rptr, ((size_t)nc + 1) * sizeof(int)));
    k_lower_bound_sorted<int><<<nblk(nc + 1, T), T, 0, st>>>(d_rkey, nP, nc + 1, 0, F.Rptr);
    int nR = 0;
    DDPS_CUDA(cudaMemcpyAsync(&nR, F.Rptr + nc, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    DDPS_CUDA(cudaMalloc(&F.Rent, std::max<size_t>(nR, 1) * sizeof(int2)));
    if (nR) DDPS_CUDA(cudaMemcpyAsync(F.Rent, d_pvs, (size_t)nR * sizeof(int2), cudaMemcpyDeviceToDevice, st));
  }
  int nR = 0;
  DDPS_CUDA(cudaMemcpyAsync(&nR, F.Rptr + nc, sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));

  // ---- 6. pattern of A*P: owned rows here, the neighbours' boundary rows by exchange (GLOBAL coarse ids) ----
  RowXchg apx;
  {
    int* d_ovf;
    DDPS_CUDA(tmp.alloc(&d_ovf, 4));
    DDPS_CUDA(cudaMemsetAsync(d_ovf, 0, sizeof(int), st));
    DDPS_CUDA(cudaMemsetAsync(d_cnt, 0, ((size_t)nl + 2) * sizeof(int), st));
    if (n) k_ap_pattern<false><<<nblk(n, T), T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, F.Pptr, F.Pent, d_cnt, nullptr, nullptr, d_ovf);
    DDPS_CUDA(cudaMalloc(&F.APptr, ((size_t)nl + 1) * sizeof(int)));
    DDPS_CUDA(cudaMemsetAsync(F.APptr, 0, ((size_t)nl + 1) * sizeof(int), st));
    if ((rc = scan_exclusive(ctx, tmp, d_cnt, F.APptr, n + 1))) return rc;
    int nAP_own = 0, ovf = 0;
    DDPS_CUDA(cudaMemcpyAsync(&nAP_own, F.APptr + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaMemcpyAsync(&ovf, d_ovf, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    double flag = ovf ? 1.0 : 0.0;   // collective decision
    DDPS_CUDA(cudaMemcpyAsync(&ctx->scal->norm_aux2, &flag, sizeof(double), cudaMemcpyHostToDevice, st));
    if ((rc = comm_allreduce(ctx, &ctx->scal->norm_aux2, 1, COMM_MAX))) return rc;
    DDPS_CUDA(cudaMemcpyAsync(&flag, &ctx->scal->norm_aux2, sizeof(double), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    if (flag > 0.5) { ctx->fail(DDPS_ERR_ARG, "distributed multigrid: a row of A*P is longer than 64 entries"); return DDPS_ERR_ARG; }
    DDPS_CUDA(cudaMalloc(&F.APcol, std::max<size_t>(nAP_own, 1) * sizeof(int)));
    if (n) k_ap_pattern<true><<<nblk(n, T), T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, F.Pptr, F.Pent, nullptr, F.APptr, F.APcol, d_ovf);
    F.ap_ghost_base = nAP_own;
    if ((rc = row_exchange(ctx, tmp, H, n, nl, F.APptr, (void**)&F.APcol, sizeof(int), &apx))) return rc;
  }
  int nAP = 0;
  DDPS_CUDA(cudaMemcpyAsync(&nAP, F.APptr + nl, sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));
  F.nAP = nAP;
  DDPS_CUDA(cudaMalloc(&F.APval, std::max<size_t>(nAP, 1) * sizeof(double)));
  DDPS_CUDA(cudaMemsetAsync(F.APval, 0, std::max<size_t>(nAP, 1) * sizeof(double), st));
  F.ap_send_pos = apx.send_pos;
  F.ap_nsend = apx.nsend_ent;
  F.ap_send_off = apx.send_ent_off;
  F.ap_recv_off = apx.recv_ent_off;
  DDPS_CUDA(cudaMalloc(&F.ap_send_buf, std::max<size_t>(F.ap_nsend, 1) * sizeof(double)));

  // ---- 7. symbolic Galerkin product: keys (own coarse row << 32 | global coarse column) ----
  long long *d_qcnt, *d_qoff;
  DDPS_CUDA(tmp.alloc(&d_qcnt, (size_t)nR + 1));
  DDPS_CUDA(tmp.alloc(&d_qoff, (size_t)nR + 1));
  DDPS_CUDA(cudaMemsetAsync(d_qcnt, 0, ((size_t)nR + 1) * sizeof(long long), st));
  if (nR) k_sym_keys_ap<false><<<nblk(nR, T), T, 0, st>>>(nR, F.Rent, d_rkey, F.APptr, F.APcol, d_qcnt, nullptr, nullptr);
  if ((rc = scan_exclusive(ctx, tmp, d_qcnt, d_qoff, nR + 1))) return rc;
  long long nk = 0;
  DDPS_CUDA(cudaMemcpyAsync(&nk, d_qoff + nR, sizeof(long long), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));
  if (nk >= (1LL << 31) - 1) { ctx->fail(DDPS_ERR_ARG, "distributed multigrid: too many structural products on one rank"); return DDPS_ERR_ARG; }
  unsigned long long *d_k0, *d_k1;
  DDPS_CUDA(tmp.alloc(&d_k0, (size_t)std::max<long long>(nk, 1)));
  DDPS_CUDA(tmp.alloc(&d_k1, (size_t)std::max<long long>(nk, 1)));
  if (nR) k_sym_keys_ap<true><<<nblk(nR, T), T, 0, st>>>(nR, F.Rent, d_rkey, F.APptr, F.APcol, nullptr, d_qoff, d_k0);

  // ---- 8. the ghost aggregates this rank needs: every non-owned coarse id in its coarse rows, in the A*P rows and in the
  //         prolongator rows of its owned fine nodes; sorted by global id = grouped by owner ----
  const int nAP_own = F.ap_ghost_base;
  // (the prolongator rows of the GHOST fine nodes count too: the cycle prolongs onto them redundantly, which saves the halo
  //  exchange of x after the prolongation)
  const long long ncand = nk + nAP_own + nP;
  if (ncand >= (1LL << 31) - 1) { ctx->fail(DDPS_ERR_ARG, "distributed multigrid: too many structural products on one rank"); return DDPS_ERR_ARG; }
  int *d_cand, *d_cand2, *d_ng;
  DDPS_CUDA(tmp.alloc(&d_cand, (size_t)std::max<long long>(ncand, 1)));
  DDPS_CUDA(tmp.alloc(&d_cand2, (size_t)std::max<long long>(ncand, 1)));
  DDPS_CUDA(tmp.alloc(&d_ng, 4));
  if (nk) k_cand_keys<<<nblk(nk, T), T, 0, st>>>(nk, d_k0, off, nc, d_cand);
  if (nAP_own) k_cand_int<<<nblk(nAP_own, T), T, 0, st>>>(nAP_own, F.APcol, off, nc, d_cand + nk);
  if (nP) k_cand_int2<<<nblk(nP, T), T, 0, st>>>(nP, F.Pent, off, nc, d_cand + nk + nAP_own);
  int ngh = 0;
  std::vector<int> gl;
  if (ncand > 0) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, d_cand, d_cand2, (int)ncand, 0, 31, st);
    void* ws;
    DDPS_CUDA(tmp.alloc((char**)&ws, bytes));
    DDPS_CUDA(cub::DeviceRadixSort::SortKeys(ws, bytes, d_cand, d_cand2, (int)ncand, 0, 31, st));
    size_t b2 = 0;
    cub::DeviceSelect::Unique(nullptr, b2, d_cand2, d_cand, d_ng, (int)ncand, st);
    void* ws2;
    DDPS_CUDA(tmp.alloc((char**)&ws2, b2));
    DDPS_CUDA(cub::DeviceSelect::Unique(ws2, b2, d_cand2, d_cand, d_ng, (int)ncand, st));
    DDPS_CUDA(cudaMemcpyAsync(&ngh, d_ng, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    gl.resize((size_t)ngh);
    if (ngh) DDPS_CUDA(cudaMemcpy(gl.data(), d_cand, (size_t)ngh * sizeof(int), cudaMemcpyDeviceToHost));
    if (ngh && gl.back() == 0x7fffffff) { gl.pop_back(); --ngh; }
  }
  const int* d_gl = d_cand;   // sorted ghost ids (first ngh entries)

  // ---- 9. local coarse numbering everywhere, then the coarse pattern ----
  if (nP) k_relabel_int2<<<nblk(nP, T), T, 0, st>>>(nP, F.Pent, off, nc, d_gl, ngh);
  if (nAP) k_relabel_int<<<nblk(nAP, T), T, 0, st>>>(nAP, F.APcol, off, nc, d_gl, ngh);
  if (nk) k_relabel_keys<<<nblk(nk, T), T, 0, st>>>(nk, d_k0, off, nc, d_gl, ngh);
  int cnnz = 0;
  if (nk > 0) {
    const int endbit = 32 + bits_for((uint64_t)nc + 1);
    int* d_num;
    DDPS_CUDA(tmp.alloc(&d_num, 4));
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
    DDPS_CUDA(cudaMemcpyAsync(&cnnz, d_num, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
  }
  AmgLevelDev& C = A->lv[l + 1];
  if ((rc = finish_coarse_level(ctx, tmp, d_k0, cnnz, nc, C))) return rc;
  C.n_loc = nc + ngh;
  C.goff = coff[me];
  F.nc = nc;

  // ---- 10. halo plan of the coarse level: tell every owner which of its aggregates I hold as ghosts ----
  {
    HaloPlan* P = new HaloPlan();
    C.plan = P;
    C.owns_plan = true;
    P->n_own = nc;
    P->n_ghost = ngh;
    std::vector<int> want((size_t)kMaxRanks, 0), wstart((size_t)nr + 1, 0);
    for (int r = 0; r < nr; ++r) {
      wstart[r] = (int)(std::lower_bound(gl.begin(), gl.end(), (int)coff[r]) - gl.begin());
    }
    wstart[nr] = ngh;
    for (int r = 0; r < nr; ++r) want[r] = wstart[r + 1] - wstart[r];
    std::vector<int> wall((size_t)kMaxRanks * nr);
    if ((rc = comm_allgather_host(ctx, want.data(), kMaxRanks * sizeof(int), wall.data()))) return rc;
    // neighbours = ranks I need aggregates from or that need some of mine (the same symmetric set on both sides)
    for (int r = 0; r < nr; ++r) {
      if (r == me) continue;
      const int need = want[r], give = wall[(size_t)r * kMaxRanks + me];
      if (need > 0 || give > 0) P->peer.push_back(r);
    }
    P->nneigh = (int)P->peer.size();
    P->send_off.assign(P->nneigh + 1, 0);
    P->recv_off.assign(P->nneigh + 1, 0);
    for (int j = 0; j < P->nneigh; ++j) {
      const int r = P->peer[j];
      P->send_off[j + 1] = P->send_off[j] + wall[(size_t)r * kMaxRanks + me];
      P->recv_off[j] = wstart[r];
      P->recv_off[j + 1] = wstart[r + 1];
    }
    if (P->nneigh == 0) P->recv_off[0] = 0;
    // (ghosts are sorted by global id = by owner: the block of peer j is [wstart[r], wstart[r+1]); blocks of ranks that own
    //  nothing I need are empty, so the offsets are consistent with the ghost order)
    P->nsend = P->send_off[P->nneigh];
    DDPS_CUDA(cudaMalloc(&P->send_idx, std::max<size_t>(P->nsend, 1) * sizeof(int)));
    DDPS_CUDA(cudaMalloc(&P->send_buf, std::max<size_t>(P->nsend, 1) * sizeof(double)));
    // requests: the owner-local ids of my ghosts, to their owners; what arrives is my send list
    int* d_req;
    DDPS_CUDA(tmp.alloc(&d_req, (size_t)std::max(ngh, 1)));
    std::vector<int> req((size_t)ngh);
    for (int r = 0; r < nr; ++r)
      for (int k = wstart[r]; k < wstart[r + 1]; ++k) req[k] = gl[k] - (int)coff[r];
    if (ngh) DDPS_CUDA(cudaMemcpyAsync(d_req, req.data(), (size_t)ngh * sizeof(int), cudaMemcpyHostToDevice, st));
    std::vector<CommMsg> msgs;
    for (int j = 0; j < P->nneigh; ++j) {
      CommMsg m;
      m.peer = P->peer[j];
      m.send = d_req + P->recv_off[j];
      m.send_bytes = (size_t)(P->recv_off[j + 1] - P->recv_off[j]) * sizeof(int);
      m.recv = P->send_idx + P->send_off[j];
      m.recv_bytes = (size_t)(P->send_off[j + 1] - P->send_off[j]) * sizeof(int);
      msgs.push_back(m);
    }
    if ((rc = comm_exchange(ctx, msgs.data(), (int)msgs.size()))) return rc;
    DDPS_CUDA(cudaStreamSynchronize(st));
    if ((rc = plan_finish(ctx, *P))) return rc;
  }
  // global ids of the coarse level's local columns
  DDPS_CUDA(cudaMalloc(&C.glob, std::max<size_t>((size_t)C.n_loc, 1) * sizeof(int)));
  if (C.n_loc) k_global_ids<<<nblk(C.n_loc, T), T, 0, st>>>(C.n_loc, nc, nullptr, off, C.glob);
  if (ngh) DDPS_CUDA(cudaMemcpyAsync(C.glob + nc, d_gl, (size_t)ngh * sizeof(int), cudaMemcpyDeviceToDevice, st));

  // ---- 11. Galerkin product of the base values (the next level's strength / prolongator smoothing need them) ----
  F.dist = 1;
  if ((rc = amg_launch_rap(ctx, F, fval, C, C.val, C.dinv))) return rc;
  C.cur_val = C.val;
  C.cur_dinv = C.dinv;
  DDPS_CUDA(cudaStreamSynchronize(st));
  DDPS_CUDA(cudaGetLastError());
  *ncg_out = ncg;
  return 0;
}


// Transition: level l is distributed, everything below is REPLICATED.  The level's base operator (a few hundred thousand rows
// at most) is gathered on every rank, the host reference implementation (amg_host.cu) builds the rest of the hierarchy from
// it -- identically on every rank, aggregates may cross rank boundaries freely here -- and every rank installs: the
// prolongator rows of its local (owned + ghost) fine nodes, the restriction rows cut down to its owned fine nodes (partial
// sums, completed by an allreduce in the cycle and in the Galerkin product), and the replicated coarse levels.
int amg_dist_transition(Context* ctx, Amg* A, int l, const AmgParams& prm, int* nlev_out) {
  *nlev_out = 0;
  AmgLevelDev& F = A->lv[l];
  const int n = F.n, nl = F.n_loc;
  const HaloPlan& H = *F.plan;
  const int me = ctx->rank, nr = ctx->nranks;
  cudaStream_t st = ctx->stream;
  const int T = 256;
  Tmp tmp(st);
  int rc;
  // ---- sizes, global numbering ----
  long long sz_me[2] = {n, F.nnz};
  std::vector<long long> sz((size_t)2 * nr);
  if ((rc = comm_allgather_host(ctx, sz_me, sizeof(sz_me), sz.data()))) return rc;
  std::vector<long long> roff((size_t)nr + 1, 0), eoff((size_t)nr + 1, 0);
  for (int r = 0; r < nr; ++r) { roff[r + 1] = roff[r] + sz[2 * r]; eoff[r + 1] = eoff[r] + sz[2 * r + 1]; }
  const long long ng = roff[nr], nnzg = eoff[nr];
  if (ng >= (1LL << 31) - 2 || nnzg >= (1LL << 31) - 2) { ctx->fail(DDPS_ERR_ARG, "transition level too large"); return DDPS_ERR_ARG; }
  F.goff = roff[me];
  if (!F.glob) {
    DDPS_CUDA(cudaMalloc(&F.glob, std::max<size_t>((size_t)nl, 1) * sizeof(int)));
    if (nl) k_global_ids<<<nblk(nl, T), T, 0, st>>>(nl, n, nullptr, (int)roff[me], F.glob);
    if ((rc = halo_exchange_bytes(ctx, H, F.glob, sizeof(int)))) return rc;
  }
  // ---- my rows: CSR with GLOBAL column ids ----
  std::vector<int> rp((size_t)n + 1), sp((size_t)F.nslices + 1), scol((size_t)F.nentries), glob((size_t)nl);
  std::vector<double> sv((size_t)F.nentries);
  std::vector<unsigned char> fix((size_t)n, 0);
  DDPS_CUDA(cudaStreamSynchronize(st));
  DDPS_CUDA(cudaMemcpy(rp.data(), F.rowptr, rp.size() * sizeof(int), cudaMemcpyDeviceToHost));
  DDPS_CUDA(cudaMemcpy(sp.data(), F.slice_ptr, sp.size() * sizeof(int), cudaMemcpyDeviceToHost));
  if (F.nentries) DDPS_CUDA(cudaMemcpy(scol.data(), F.col, scol.size() * sizeof(int), cudaMemcpyDeviceToHost));
  if (F.nentries) DDPS_CUDA(cudaMemcpy(sv.data(), F.cur_val, sv.size() * sizeof(double), cudaMemcpyDeviceToHost));
  if (nl) DDPS_CUDA(cudaMemcpy(glob.data(), F.glob, glob.size() * sizeof(int), cudaMemcpyDeviceToHost));
  if (l == 0 && n) DDPS_CUDA(cudaMemcpy(fix.data(), ctx->isD, (size_t)n, cudaMemcpyDeviceToHost));
  const long long nnz_me = rp[n];
  if (nnz_me != F.nnz) { ctx->fail(DDPS_ERR_STATE, "transition level: inconsistent nonzero count"); return DDPS_ERR_STATE; }
  // packed message: [row lengths int x n][global columns int x nnz][fixed flags, padded to 8][values double x nnz]
  auto msg_bytes = [](long long rows, long long nnz) { return (size_t)(4 * rows + 4 * nnz + 8 * ((rows + 7) / 8) + 8) / 8 * 8 + 8 * (size_t)nnz; };
  size_t maxb = 0;
  for (int r = 0; r < nr; ++r) maxb = std::max(maxb, msg_bytes(sz[2 * r], sz[2 * r + 1]));
  std::vector<char> mine(maxb, 0), all(maxb * (size_t)nr);
  {
    int* plen = (int*)mine.data();
    int* pcol = plen + n;
    unsigned char* pfix = (unsigned char*)(pcol + nnz_me);
    double* pval = (double*)(mine.data() + msg_bytes(n, nnz_me) - 8 * (size_t)nnz_me);
    for (int r = 0; r < n; ++r) {
      plen[r] = rp[r + 1] - rp[r];
      const int base = sp[r >> 5], lane = r & 31;
      for (int k = 0; k < plen[r]; ++k) {
        pcol[rp[r] + k] = glob[scol[(size_t)base + 32 * k + lane]];
        pval[rp[r] + k] = sv[(size_t)base + 32 * k + lane];
      }
      pfix[r] = fix[r];
    }
  }
  if ((rc = comm_allgather_host(ctx, mine.data(), maxb, all.data()))) return rc;
  std::vector<char>().swap(mine);
  std::vector<int> grp((size_t)ng + 1, 0), gcc((size_t)nnzg);
  std::vector<double> gvv((size_t)nnzg);
  std::vector<unsigned char> gfix((size_t)ng, 0);
  for (int r = 0; r < nr; ++r) {
    const long long rows = sz[2 * r], nnz = sz[2 * r + 1];
    const char* base = all.data() + maxb * (size_t)r;
    const int* plen = (const int*)base;
    const int* pcol = plen + rows;
    const unsigned char* pfix = (const unsigned char*)(pcol + nnz);
    const double* pval = (const double*)(base + msg_bytes(rows, nnz) - 8 * (size_t)nnz);
    long long e = eoff[r];
    for (long long i = 0; i < rows; ++i) {
      grp[roff[r] + i + 1] = plen[i];
      gfix[roff[r] + i] = pfix[i];
    }
    // columns sorted by global id inside every row (the local order is owned-first)
    long long k0 = 0;
    std::vector<std::pair<int, double>> row;
    for (long long i = 0; i < rows; ++i) {
      row.clear();
      for (int k = 0; k < plen[i]; ++k) row.emplace_back(pcol[k0 + k], pval[k0 + k]);
      std::sort(row.begin(), row.end());
      for (auto& pr : row) { gcc[e] = pr.first; gvv[e] = pr.second; ++e; }
      k0 += plen[i];
    }
  }
  std::vector<char>().swap(all);
  for (long long i = 0; i < ng; ++i) grp[i + 1] += grp[i];
  // ---- replicated rest of the hierarchy (host reference implementation; the ranks share the host cores) ----
  AmgHost Hh;
  AmgParams sub = prm;
  sub.max_levels = prm.max_levels - l;
  sub.nthreads = std::max(1, (int)std::min(16u, std::thread::hardware_concurrency()) / nr);
  const int nh = amg_host_build(Hh, (int)ng, std::move(grp), std::move(gcc), std::move(gvv), l == 0 ? gfix.data() : nullptr, sub);
  if (nh < 2) return 0;
  const AmgHostLevel& H0 = Hh.lv[0];
  const int ncg = H0.nc;
  // ---- transfers of level l: P rows of my local fine nodes, R rows cut down to my owned fine nodes ----
  {
    std::vector<int> pptr((size_t)nl + 1, 0);
    for (int i = 0; i < nl; ++i) pptr[i + 1] = pptr[i] + (H0.Pptr[glob[i] + 1] - H0.Pptr[glob[i]]);
    std::vector<int2> pent((size_t)pptr[nl]);
    for (int i = 0; i < nl; ++i) {
      int o = pptr[i];
      for (int k = H0.Pptr[glob[i]]; k < H0.Pptr[glob[i] + 1]; ++k, ++o) {
        const float pv = (float)H0.Pval[k];
        int bits;
        memcpy(&bits, &pv, 4);
        pent[o] = make_int2(H0.Pcol[k], bits);
      }
    }
    const int lo = (int)roff[me], hi = lo + n;
    std::vector<int> rptr((size_t)ncg + 1, 0);
    std::vector<int2> rent;
    for (int J = 0; J < ncg; ++J) {
      for (int q = H0.Rptr[J]; q < H0.Rptr[J + 1]; ++q) {
        const int g = H0.Rrow[q];
        if (g < lo || g >= hi) continue;
        const float rv = (float)H0.Rval[q];
        int bits;
        memcpy(&bits, &rv, 4);
        rent.push_back(make_int2(g - lo, bits));
      }
      rptr[J + 1] = (int)rent.size();
    }
    auto up = [&](auto** dptr, const auto& h) -> int {
      using E = typename std::remove_reference<decltype(h[0])>::type;
      DDPS_CUDA(cudaMalloc((void**)dptr, std::max<size_t>(h.size(), 1) * sizeof(E)));
      if (!h.empty()) DDPS_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(E), cudaMemcpyHostToDevice));
      return 0;
    };
    if ((rc = up(&F.Pptr, pptr)) || (rc = up(&F.Pent, pent)) || (rc = up(&F.Rptr, rptr)) || (rc = up(&F.Rent, rent))) return rc;
  }
  F.nc = ncg;
  // ---- pattern of A*P for my owned rows (two-stage Galerkin product) ----
  {
    int *d_cnt, *d_ovf;
    DDPS_CUDA(tmp.alloc(&d_cnt, (size_t)n + 1));
    DDPS_CUDA(tmp.alloc(&d_ovf, 4));
    DDPS_CUDA(cudaMemsetAsync(d_ovf, 0, sizeof(int), st));
    DDPS_CUDA(cudaMemsetAsync(d_cnt, 0, ((size_t)n + 1) * sizeof(int), st));
    if (n) k_ap_pattern<false><<<nblk(n, T), T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, F.Pptr, F.Pent, d_cnt, nullptr, nullptr, d_ovf);
    DDPS_CUDA(cudaMalloc(&F.APptr, ((size_t)n + 1) * sizeof(int)));
    if ((rc = scan_exclusive(ctx, tmp, d_cnt, F.APptr, n + 1))) return rc;
    int nAP = 0, ovf = 0;
    DDPS_CUDA(cudaMemcpyAsync(&nAP, F.APptr + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaMemcpyAsync(&ovf, d_ovf, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    double flag = (ovf || F.wmax > kMaxLocalRow) ? 1.0 : 0.0;
    DDPS_CUDA(cudaMemcpyAsync(&ctx->scal->norm_aux2, &flag, sizeof(double), cudaMemcpyHostToDevice, st));
    if ((rc = comm_allreduce(ctx, &ctx->scal->norm_aux2, 1, COMM_MAX))) return rc;
    DDPS_CUDA(cudaMemcpyAsync(&flag, &ctx->scal->norm_aux2, sizeof(double), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    if (flag > 0.5) { ctx->fail(DDPS_ERR_ARG, "distributed multigrid: a row of A*P is longer than 64 entries"); return DDPS_ERR_ARG; }
    F.nAP = nAP;
    DDPS_CUDA(cudaMalloc(&F.APcol, std::max<size_t>(nAP, 1) * sizeof(int)));
    DDPS_CUDA(cudaMalloc(&F.APval, std::max<size_t>(nAP, 1) * sizeof(double)));
    if (n) k_ap_pattern<true><<<nblk(n, T), T, 0, st>>>(n, F.rowptr, F.slice_ptr, F.col, F.Pptr, F.Pent, nullptr, F.APptr, F.APcol, d_ovf);
  }
  // ---- replicated coarse levels ----
  for (int k = 1; k < nh; ++k) {
    AmgLevelDev& L = A->lv[l + k];
    if ((rc = amg_install_pattern(ctx, L, Hh.lv[k]))) return rc;
    if (k + 1 < nh && (rc = amg_install_transfers(ctx, L, Hh.lv[k]))) return rc;
  }
  F.dist = 2;
  DDPS_CUDA(cudaStreamSynchronize(st));
  *nlev_out = l + nh;
  return 0;
}

}  // namespace ddps
