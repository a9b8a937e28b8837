// Smoothed-aggregation multigrid preconditioner for the hand-written CG ("a better
// preconditioner behind the same PCG interface", the replacement of the reference's `\`, src:394,588,1063).
//
// Split of the work:
//   * once per base operator: aggregates, frozen smoothed prolongators P_l (fp32, packed), R_l = P_l^T stored
//     row-wise, and the coarse sparsity patterns -- on the device for the large levels (amg_setup.cu; only the
//     strong-neighbour lists go to the host for the sequential greedy aggregation), by the host reference
//     implementation (amg_host.cu) for the small rest;
//   * once per operator A = M - J0 (device, k_amg_rap2): the Galerkin products A_{l+1} = R_l A_l P_l on the frozen
//     patterns -- one warp per coarse row, the (fine row, fine entry) pairs dealt to the lanes, lane-private
//     accumulator rows in shared memory, fixed-order reduction: no atomics, bitwise reproducible; then the dense
//     inverse of the coarsest operator (Gauss-Jordan in one CTA, in shared memory).  Coarse operators are kept per
//     KIND of system and reused while the fine operator stays within 1 % of the snapshot they came from (amg_numeric);
//   * per CG iteration (device): one V(1,1) cycle with damped-Jacobi smoothing; all matrices are SELL-32 like the
//     fine operator (the large levels also as an fp32 copy for the cycle's SpMVs; arithmetic is FP64 throughout), the
//     smoother and the residual are fused into the SpMV epilogue, restriction/prolongation are deterministic row
//     gathers.  The cycle is a fixed SPD linear operator, so plain (non-flexible) CG applies.
// The PCG control flow (alpha, beta, stop test, iteration counter) stays on the device like in the Jacobi path;
// every kernel of the cycle early-exits once the `done` flag is up.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "amg.h"
#include "amg_dev.cuh"
#include "ddps_internal.cuh"
#include "device_utils.cuh"

namespace ddps {

namespace {

template <typename T>
int amg_alloc(Context* ctx, T** p, size_t n) {
  DDPS_CUDA(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
  return 0;
}
template <typename T>
int amg_upload(Context* ctx, T** p, const std::vector<T>& h) {
  int rc = amg_alloc(ctx, p, h.size());
  if (rc) return rc;
  if (!h.empty()) DDPS_CUDA(cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
template <typename T>
void amg_release(T*& p) { if (p) { cudaFree(p); p = nullptr; } }

inline int amg_grid(Context* ctx, int64_t work_items, int block) {
  int64_t need = (work_items + block - 1) / block;
  int64_t cap = (int64_t)ctx->num_sms * (2048 / block);
  return (int)std::max<int64_t>(1, std::min(need, cap));
}

double wall_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// ---- iteration timeline (DDPS_PROFILE_ITER) ----
static inline void prof_mark(Context* ctx, const char* name) {
  AmgProf& P = ctx->amg->prof;
  if (!P.armed || P.n >= 96) return;
  P.name[P.n] = name;
  cudaEventRecord(P.ev[P.n++], ctx->stream);
}
static void prof_begin(Context* ctx) {
  AmgProf& P = ctx->amg->prof;
  if (!P.on) return;
  if (!P.have_ev) { for (int i = 0; i < 96; ++i) cudaEventCreate(&P.ev[i]); P.have_ev = true; }
  P.armed = true;
  P.n = 0;
  prof_mark(ctx, "start");
}
static void prof_end(Context* ctx) {
  AmgProf& P = ctx->amg->prof;
  if (!P.armed) return;
  P.armed = false;
  cudaEventSynchronize(P.ev[P.n - 1]);
  for (int i = 1; i < P.n; ++i) {
    float ms = 0;
    cudaEventElapsedTime(&ms, P.ev[i - 1], P.ev[i]);
    P.acc[i] += ms;
  }
  P.samples++;
}
static void prof_report(Context* ctx) {
  AmgProf& P = ctx->amg->prof;
  if (!P.on || !P.samples) return;
  double tot = 0;
  for (int i = 1; i < P.n; ++i) tot += P.acc[i];
  fprintf(stderr, "[ddps prof] rank %d: one CG iteration = %.1f us (mean of %d sampled iterations)\n", ctx->rank, 1e3 * tot / P.samples, P.samples);
  for (int i = 1; i < P.n; ++i) fprintf(stderr, "[ddps prof] rank %d   %-28s %7.1f us\n", ctx->rank, P.name[i], 1e3 * P.acc[i] / P.samples);
  if (P.have_ev) for (int i = 0; i < 96; ++i) cudaEventDestroy(P.ev[i]);
  P.have_ev = false;
}

// ------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------
// SELL-32 SpMV with fused epilogue.  MODE 0: y = A x;  1: y = b - A x;  2: y = x + om * dinv * (b - A x)
// (one damped-Jacobi sweep, out of place).  DOT (MODE 2, finest level): also sums b.y = r.z for the CG.
// VT = float: the cycle's copy of the matrix values in fp32 (the cycle is only a preconditioner; arithmetic stays FP64,
// so the cycle remains an exactly linear, symmetric operator) -- a third less matrix traffic per cycle SpMV.
template <int MODE, bool DOT, typename VT>
__global__ void __launch_bounds__(kVecBlock, 8) k_amg_spmv(int nrows, int nslices, const int* __restrict__ slice_ptr,
                                                        const int* __restrict__ col, const VT* __restrict__ val,
                                                        const double* __restrict__ x, const double* __restrict__ b,
                                                        const double* __restrict__ dinv, double om,
                                                        double* __restrict__ y, double* partials, PcgScalars* scal) {
  pdl_enter();
  __shared__ double sh_red[32];
  if (scal->done) return;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  double dot = 0.0;
  for (int slice = gw; slice < nslices; slice += nw) {
    const int base = slice_ptr[slice];
    const int w = (slice_ptr[slice + 1] - base) >> 5;
    int idx = base + lane;
    double sum = 0.0;
    int k = 0;
    for (; k + 4 <= w; k += 4) {
      const int c0 = ld_stream(col + idx), c1 = ld_stream(col + idx + 32), c2 = ld_stream(col + idx + 64),
                c3 = ld_stream(col + idx + 96);
      const double v0 = (double)ld_stream(val + idx), v1 = (double)ld_stream(val + idx + 32), v2 = (double)ld_stream(val + idx + 64),
                   v3 = (double)ld_stream(val + idx + 96);
      const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
      sum += v0 * x0; sum += v1 * x1; sum += v2 * x2; sum += v3 * x3;
      idx += 128;
    }
    for (; k < w; ++k) {
      const int c0 = ld_stream(col + idx);
      const double v0 = (double)ld_stream(val + idx);
      sum += v0 * __ldg(x + c0);
      idx += 32;
    }
    const int row = slice * 32 + lane;
    if (row < nrows) {
      double out;
      if (MODE == 0) out = sum;
      else if (MODE == 1) out = b[row] - sum;
      else out = __ldg(x + row) + om * (dinv[row] * (b[row] - sum));
      y[row] = out;
      if (DOT) dot += b[row] * out;
    }
  }
  if (DOT) {
    double v[1] = {dot}, tot[1];
    if (grid_reduce<1, false>(v, partials, &scal->cnt[0], sh_red, tot)) {
      if (threadIdx.x == 0) scal->sums[1] = tot[0];
    }
  }
}

// Small levels (a few thousand rows, 15-40 entries per row): one WARP per row instead of one lane per row, so the row's
// entries are loaded by 32 lanes at once (one dependent load chain instead of len/4 of them; these kernels are pure
// latency).  Same SELL-32 addressing, fixed shuffle-tree summation (deterministic).
template <int MODE>
__global__ void __launch_bounds__(256) k_amg_spmv_wpr(int nrows, const int* __restrict__ rowptr, const int* __restrict__ slice_ptr,
                                                      const int* __restrict__ col, const double* __restrict__ val,
                                                      const double* __restrict__ x, const double* __restrict__ b,
                                                      const double* __restrict__ dinv, double om, double* __restrict__ y,
                                                      const PcgScalars* scal) {
  pdl_enter();
  if (scal->done) return;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int row = gw; row < nrows; row += nw) {
    const int len = rowptr[row + 1] - rowptr[row];
    const int base = slice_ptr[row >> 5] + (row & 31);
    double sum = 0.0;
    for (int k = lane; k < len; k += 32) sum += val[base + 32 * k] * __ldg(x + col[base + 32 * k]);
    sum = warp_sum(sum);
    if (lane == 0) {
      double out;
      if (MODE == 0) out = sum;
      else if (MODE == 1) out = b[row] - sum;
      else out = x[row] + om * (dinv[row] * (b[row] - sum));
      y[row] = out;
    }
  }
}

__global__ void __launch_bounds__(256) k_amg_restrict_wpr(int nc, const int* __restrict__ Rptr, const int2* __restrict__ Rent,
                                                          const double* __restrict__ t, double* __restrict__ bc,
                                                          const double* __restrict__ dinv_c, double om, double* __restrict__ xc,
                                                          const PcgScalars* scal) {
  pdl_enter();
  if (scal->done) return;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int I = gw; I < nc; I += nw) {
    const int q0 = Rptr[I], q1 = Rptr[I + 1];
    double s = 0.0;
    for (int q = q0 + lane; q < q1; q += 32) {
      const int2 en = Rent[q];
      s += (double)__int_as_float(en.y) * __ldg(t + en.x);
    }
    s = warp_sum(s);
    if (lane == 0) {
      bc[I] = s;
      if (dinv_c) xc[I] = om * (dinv_c[I] * s);
    }
  }
}

__global__ void k_amg_to_float(size_t n, const double* __restrict__ in, float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __double2float_rn(ld_stream(in + i));
}

// x = om * dinv * b  (first damped-Jacobi sweep from the zero initial guess)
__global__ void k_amg_jac0(int n, double om, const double* __restrict__ dinv, const double* __restrict__ b,
                           double* __restrict__ x, const PcgScalars* scal) {
  pdl_enter();
  if (scal->done) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = om * (dinv[i] * b[i]);
}

// bc = R t   (R = P^T stored row-wise: a deterministic gather); fused: xc = om * dinv_c * bc, the first damped-Jacobi
// sweep of the coarse level from the zero initial guess (dinv_c == nullptr on the coarsest level)
__global__ void __launch_bounds__(256) k_amg_restrict(int nc, const int* __restrict__ Rptr, const int2* __restrict__ Rent,
                               const double* __restrict__ t, double* __restrict__ bc, const double* __restrict__ dinv_c,
                               double om, double* __restrict__ xc, const PcgScalars* scal) {
  pdl_enter();
  if (scal->done) return;
  for (int I = blockIdx.x * blockDim.x + threadIdx.x; I < nc; I += gridDim.x * blockDim.x) {
    double s = 0.0;
    const int e = Rptr[I + 1];
    int q = Rptr[I];
    for (; q + 8 <= e; q += 8) {   // eight entries, then their eight gathers, in flight together; summed in entry order
      int2 en[8];
      double tv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) en[k] = Rent[q + k];
#pragma unroll
      for (int k = 0; k < 8; ++k) tv[k] = __ldg(t + en[k].x);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += (double)__int_as_float(en[k].y) * tv[k];
    }
    for (; q + 4 <= e; q += 4) {
      const int2 e0 = Rent[q], e1 = Rent[q + 1], e2 = Rent[q + 2], e3 = Rent[q + 3];
      const double t0 = __ldg(t + e0.x), t1 = __ldg(t + e1.x), t2 = __ldg(t + e2.x), t3 = __ldg(t + e3.x);
      s += (double)__int_as_float(e0.y) * t0; s += (double)__int_as_float(e1.y) * t1;
      s += (double)__int_as_float(e2.y) * t2; s += (double)__int_as_float(e3.y) * t3;
    }
    for (; q < e; ++q) {
      const int2 en = Rent[q];
      s += (double)__int_as_float(en.y) * __ldg(t + en.x);
    }
    bc[I] = s;
    if (dinv_c) xc[I] = om * (dinv_c[I] * s);
  }
}

// x += P xc
__global__ void k_amg_prolong(int n, const int* __restrict__ Pptr, const int2* __restrict__ Pent,
                              const double* __restrict__ xc, double* __restrict__ x, const PcgScalars* scal) {
  pdl_enter();
  if (scal->done) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    const int m0 = Pptr[i], e = Pptr[i + 1];
    const double xi = x[i];
    if (e - m0 <= 4) {   // the usual row (1-4 aggregates): all entries, then all gathers, in flight together; summed in entry order
      int2 en[4];
      double xv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) en[k] = (m0 + k < e) ? Pent[m0 + k] : make_int2(-1, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) xv[k] = (en[k].x >= 0) ? __ldg(xc + en[k].x) : 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (en[k].x >= 0) s += (double)__int_as_float(en[k].y) * xv[k];
    } else {
#pragma unroll 1
      for (int m = m0; m < e; ++m) {
        const int2 en = Pent[m];
        if (en.x >= 0) s += (double)__int_as_float(en.y) * __ldg(xc + en.x);
      }
    }
    x[i] = xi + s;
  }
}

// Galerkin product on the frozen pattern: C = R A P.  One warp per coarse row I.  The (fine row i of R(I,:), entry k
// of fine row i) pairs are dealt to the 32 lanes entry-major (consecutive lanes = consecutive fine rows at the same
// entry position: aggregates are index-local, so their SELL entries and the prolongator rows of their neighbours
// share sectors); kRapBatch pairs per lane are loaded together (memory-level parallelism).  Every lane multiplies
// its pairs out with the prolongator rows and adds each product into its PRIVATE accumulator row (shared memory,
// slot found by binary search in the staged sorted coarse row); then lane t sums column t of the 32 private rows
// in lane order.  The assignment of products to lanes and both summation orders are fixed functions of the
// pattern: no atomics, bitwise reproducible.  Coarse rows longer than SLOTS entries take the serial walk at the end.
constexpr int kRapBatch = 4;

template <int SLOTS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_amg_rap2(int nc, const int* __restrict__ f_slice_ptr,
                                                        const int* __restrict__ f_col, const double* __restrict__ f_val,
                                                        const int* __restrict__ Pptr, const int2* __restrict__ Pent,
                                                        const int* __restrict__ Rptr, const int2* __restrict__ Rent,
                                                        const int* __restrict__ c_rowptr, const int* __restrict__ c_slice_ptr,
                                                        const int* __restrict__ c_col, double* __restrict__ c_val,
                                                        double* __restrict__ c_dinv) {
  constexpr int STRIDE = SLOTS + 1;   // padded row stride of the private tables: equal slots of different lanes hit different banks
  __shared__ double sh_tab[WARPS][32 * STRIDE];
  __shared__ int sh_col[WARPS][SLOTS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* tab = sh_tab[wid];
  int* ccols = sh_col[wid];
  const int gw = blockIdx.x * WARPS + wid, nw = gridDim.x * WARPS;
  for (int I = gw; I < nc; I += nw) {
    const int len = c_rowptr[I + 1] - c_rowptr[I];
    const int cbase = c_slice_ptr[I >> 5] + (I & 31);
    const int q0 = Rptr[I], nq = Rptr[I + 1] - q0;
    if (len <= SLOTS) {
      for (int t = lane; t < len; t += 32) ccols[t] = c_col[cbase + 32 * t];
      for (int t = 0; t < len; ++t) tab[lane * STRIDE + t] = 0.0;
      __syncwarp();
      // widest fine row among the rows of R(I,:) (lanes beyond it have nothing to do)
      int wloc = 0;
      for (int q = lane; q < nq; q += 32) {
        const int i = Rent[q0 + q].x;
        wloc = max(wloc, (f_slice_ptr[(i >> 5) + 1] - f_slice_ptr[i >> 5]) >> 5);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) wloc = max(wloc, __shfl_xor_sync(0xffffffffu, wloc, o));
      const int npairs = nq * wloc;
      for (int e0 = lane; e0 < npairs; e0 += 32 * kRapBatch) {
        int idx[kRapBatch], jj[kRapBatch], m0[kRapBatch], m1[kRapBatch];
        int2 re[kRapBatch];
        double av[kRapBatch];
#pragma unroll
        for (int u = 0; u < kRapBatch; ++u) {
          const int e = e0 + 32 * u;
          const int k = e / nq;
          idx[u] = k;                                  // entry position for now
          re[u] = e < npairs ? Rent[q0 + (e - k * nq)] : make_int2(-1, 0);
        }
#pragma unroll
        for (int u = 0; u < kRapBatch; ++u) {
          const int i = re[u].x, k = idx[u];
          idx[u] = -1;
          if (i >= 0) {
            const int fb = f_slice_ptr[i >> 5];
            const int w = (f_slice_ptr[(i >> 5) + 1] - fb) >> 5;
            if (k < w) idx[u] = fb + 32 * k + (i & 31);
          }
        }
#pragma unroll
        for (int u = 0; u < kRapBatch; ++u) {
          av[u] = idx[u] >= 0 ? f_val[idx[u]] : 0.0;
          jj[u] = idx[u] >= 0 ? f_col[idx[u]] : 0;
        }
#pragma unroll
        for (int u = 0; u < kRapBatch; ++u) {
          // padding and the exact zeros of right-angle meshes / eliminated entries contribute nothing
          m0[u] = Pptr[jj[u]];
          m1[u] = av[u] != 0.0 ? Pptr[jj[u] + 1] : m0[u];
        }
#pragma unroll
        for (int u = 0; u < kRapBatch; ++u) {
          const double ra = (double)__int_as_float(re[u].y) * av[u];
          for (int m = m0[u]; m < m1[u]; ++m) {
            const int2 pe = Pent[m];
            int lo = 0, hi = len;
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (ccols[mid] < pe.x) lo = mid + 1; else hi = mid;
            }
            tab[lane * STRIDE + lo] += ra * (double)__int_as_float(pe.y);
          }
        }
      }
      __syncwarp();
      for (int t = lane; t < len; t += 32) {
        double s = 0.0;
        for (int l = 0; l < 32; ++l) s += tab[l * STRIDE + t];
        c_val[cbase + 32 * t] = s;
        if (ccols[t] == I) c_dinv[I] = 1.0 / s;
      }
      __syncwarp();
    } else {
      // long coarse row: every lane walks the whole product sequence and keeps its own column(s)
      for (int t0 = 0; t0 < len; t0 += 32) {
        const int t = t0 + lane;
        const int myJ = t < len ? c_col[cbase + 32 * t] : -1;
        double acc = 0.0;
        for (int q = q0; q < q0 + nq; ++q) {
          const int2 en = Rent[q];
          const int i = en.x;
          const double r = (double)__int_as_float(en.y);
          const int fb = f_slice_ptr[i >> 5];
          const int w = (f_slice_ptr[(i >> 5) + 1] - fb) >> 5;
          int idx = fb + (i & 31);
          for (int k = 0; k < w; ++k, idx += 32) {
            const double a = f_val[idx];
            if (a == 0.0) continue;
            const int j = f_col[idx];
            const double ra = r * a;
            const int m1 = Pptr[j + 1];
            for (int m = Pptr[j]; m < m1; ++m) {
              const int2 pe = Pent[m];
              if (pe.x == myJ) acc += ra * (double)__int_as_float(pe.y);
            }
          }
        }
        if (t < len) {
          c_val[cbase + 32 * t] = acc;
          if (myJ == I) c_dinv[I] = 1.0 / acc;
        }
      }
    }
  }
}

// Two-stage Galerkin product (device-built levels).  Stage 1: row i of A*P on its fixed sorted pattern -- one thread per fine
// row (consecutive lanes read neighbouring SELL entries and prolongator rows), products added in entry order into a
// thread-private accumulator (slot by a search in the row's short column list).  Stage 2 (k_amg_rap_ap): C(I,:) = sum_i R(I,i)
// (A*P)(i,:) with the lane-private tables of k_amg_rap2 -- a third of the products of the one-stage kernel and one gather
// level less.  Both orders of summation are fixed functions of the patterns: bitwise reproducible.
constexpr int kApRowMax = 64;   // = kMaxLocalRow of amg_setup.cu (longer rows: no A*P pattern, one-stage kernel)

__global__ void __launch_bounds__(128) k_amg_ap(int n, const int* __restrict__ slice_ptr, const int* __restrict__ col,
                                                const double* __restrict__ val, const int* __restrict__ Pptr,
                                                const int2* __restrict__ Pent, const int* __restrict__ APptr,
                                                const int* __restrict__ APcol, double* __restrict__ APval) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = APptr[i], cnt = APptr[i + 1] - o;
  if (cnt == 0) return;
  int cols[kApRowMax];
  double acc[kApRowMax];
  for (int t = 0; t < cnt; ++t) { cols[t] = APcol[o + t]; acc[t] = 0.0; }
  const int fb = slice_ptr[i >> 5];
  const int w = (slice_ptr[(i >> 5) + 1] - fb) >> 5;
  int idx = fb + (i & 31);
  for (int k = 0; k < w; ++k, idx += 32) {
    const double a = val[idx];
    if (a == 0.0) continue;   // padding and exact zeros
    const int j = col[idx];
    const int m1 = Pptr[j + 1];
    for (int m = Pptr[j]; m < m1; ++m) {
      const int2 pe = Pent[m];
      int t = 0;
      while (t < cnt - 1 && cols[t] != pe.x) ++t;
      acc[t] += a * (double)__int_as_float(pe.y);
    }
  }
  for (int t = 0; t < cnt; ++t) APval[o + t] = acc[t];
}

template <int SLOTS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_amg_rap_ap(int nc, const int* __restrict__ Rptr, const int2* __restrict__ Rent,
                                                          const int* __restrict__ APptr, const int* __restrict__ APcol,
                                                          const double* __restrict__ APval, const int* __restrict__ c_rowptr,
                                                          const int* __restrict__ c_slice_ptr, const int* __restrict__ c_col,
                                                          double* __restrict__ c_val, double* __restrict__ c_dinv) {
  constexpr int STRIDE = SLOTS + 1;
  __shared__ double sh_tab[WARPS][32 * STRIDE];
  __shared__ int sh_col[WARPS][SLOTS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* tab = sh_tab[wid];
  int* ccols = sh_col[wid];
  const int gw = blockIdx.x * WARPS + wid, nw = gridDim.x * WARPS;
  for (int I = gw; I < nc; I += nw) {
    const int len = c_rowptr[I + 1] - c_rowptr[I];
    const int cbase = c_slice_ptr[I >> 5] + (I & 31);
    const int q0 = Rptr[I], q1 = Rptr[I + 1];
    if (len <= SLOTS) {
      for (int t = lane; t < len; t += 32) ccols[t] = c_col[cbase + 32 * t];
      for (int t = 0; t < len; ++t) tab[lane * STRIDE + t] = 0.0;
      __syncwarp();
      for (int q = q0 + lane; q < q1; q += 32) {
        const int2 re = Rent[q];
        const double r = (double)__int_as_float(re.y);
        const int e1 = APptr[re.x + 1];
        for (int e = APptr[re.x]; e < e1; ++e) {
          const int J = APcol[e];
          int lo = 0, hi = len;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (ccols[mid] < J) lo = mid + 1; else hi = mid;
          }
          tab[lane * STRIDE + lo] += r * APval[e];
        }
      }
      __syncwarp();
      for (int t = lane; t < len; t += 32) {
        double s = 0.0;
        for (int l = 0; l < 32; ++l) s += tab[l * STRIDE + t];
        c_val[cbase + 32 * t] = s;
        if (ccols[t] == I) c_dinv[I] = 1.0 / s;
      }
      __syncwarp();
    } else {
      for (int t0 = 0; t0 < len; t0 += 32) {   // long coarse row: every lane walks all products and keeps its own column(s)
        const int t = t0 + lane;
        const int myJ = t < len ? c_col[cbase + 32 * t] : -1;
        double acc = 0.0;
        for (int q = q0; q < q1; ++q) {
          const int2 re = Rent[q];
          const double r = (double)__int_as_float(re.y);
          const int e1 = APptr[re.x + 1];
          for (int e = APptr[re.x]; e < e1; ++e)
            if (APcol[e] == myJ) acc += r * APval[e];
        }
        if (t < len) {
          c_val[cbase + 32 * t] = acc;
          if (myJ == I) c_dinv[I] = 1.0 / acc;
        }
      }
    }
  }
}

// Dense inverse of the coarsest operator: in-place Gauss-Jordan without pivoting (the operator is SPD), one CTA.
// IN_SHARED: the n x n matrix lives in dynamic shared memory (n <= kAmgDenseShared) and is written out at the end;
// otherwise it is eliminated in global memory (fallback for hierarchies whose coarsening stalled early).
constexpr int kAmgDenseShared = 128;

template <bool IN_SHARED>
__global__ void __launch_bounds__(1024) k_amg_dense_inverse(int n, const int* __restrict__ rowptr,
                                                            const int* __restrict__ slice_ptr, const int* __restrict__ col,
                                                            const double* __restrict__ val, double* __restrict__ Ainv) {
  extern __shared__ double sh_dyn[];
  __shared__ double shrow[kAmgDenseMax], shcol[kAmgDenseMax];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nn = n * n;
  double* M = IN_SHARED ? sh_dyn : Ainv;
  for (int e = tid; e < nn; e += nt) M[e] = 0.0;
  __syncthreads();
  for (int i = tid; i < n; i += nt) {
    const int len = rowptr[i + 1] - rowptr[i];
    const int base = slice_ptr[i >> 5] + (i & 31);
    for (int k = 0; k < len; ++k) M[i * n + col[base + 32 * k]] = val[base + 32 * k];
  }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const double p = 1.0 / M[k * n + k];
    for (int j = tid; j < n; j += nt) {
      shrow[j] = (j == k) ? p : M[k * n + j] * p;
      shcol[j] = M[j * n + k];
    }
    __syncthreads();
    for (int e = tid; e < nn; e += nt) {
      const int i = e / n, j = e - i * n;
      double v;
      if (i == k) v = shrow[j];
      else if (j == k) v = -(shcol[i] * p);
      else v = M[e] - shcol[i] * shrow[j];
      M[e] = v;
    }
    __syncthreads();
  }
  if (IN_SHARED)
    for (int e = tid; e < nn; e += nt) Ainv[e] = M[e];
}

// x = Ainv b on the coarsest level: one CTA, one warp per row (lanes stride the row: coalesced), warp-shuffle sum
__global__ void __launch_bounds__(1024) k_amg_dense_apply(int n, const double* __restrict__ Ainv,
                                                          const double* __restrict__ b, double* __restrict__ x,
                                                          const PcgScalars* scal) {
  __shared__ double shb[kAmgDenseMax];
  if (scal->done) return;
  for (int j = threadIdx.x; j < n; j += blockDim.x) shb[j] = b[j];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int i = wid; i < n; i += nwarp) {
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s += Ainv[i * n + j] * shb[j];
    s = warp_sum(s);
    if (lane == 0) x[i] = s;
  }
}

constexpr int kAmgSmallLevel = 32768;   // rows; below this the warp-per-row kernels run (and the fused tail of the cycle starts)

// b = sum over the ranks of the partial restrictions of the transition level, in rank order (bitwise the same on every rank);
// fused: the first damped-Jacobi sweep of the (replicated) level from the zero guess
__global__ void k_amg_tail_sum(int nc, int nr, int me, const double* __restrict__ parts, double* __restrict__ b,
                               const double* __restrict__ dinv, double om, double* __restrict__ x, const PcgScalars* scal) {
  pdl_enter();
  if (scal->done) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < nr; ++q) {
      const double v = parts[(q == me ? 0 : (size_t)(1 + (q < me ? q : q - 1)) * nc) + i];
      s = q == 0 ? v : s + v;
    }
    b[i] = s;
    if (dinv) x[i] = om * (dinv[i] * s);
  }
}

// ------------------------------------------------------------------------------------------------------------
// The bottom of the V-cycle in ONE launch.  Below kAmgSmallLevel rows every cycle kernel is a dependent chain of a few
// memory round trips on a handful of CTAs: 5-11 us per launch, 17 launches for four levels (latency, not work).  This
// kernel walks those levels down and up itself on one CTA per SM (co-resident: cooperative launch), phases separated by a
// grid barrier (one atomic per CTA on a counter the last CTA to leave resets): residual / restriction + first sweep per level
// on the way down, the dense coarsest solve, prolongation / smoothing sweep on the way up.  Vectors written inside the kernel
// are read back through L2 (ld.global.cg): another CTA produced them since this SM's L1 saw the line.  Same arithmetic as the
// per-level kernels; levels between kAmgTailWpr and kAmgSmallLevel rows sum their rows lane-per-row (sequential walk) instead
// of warp-per-row (shuffle tree) -- a fixed order either way, so the cycle stays a fixed linear operator, reproducible run to
// run (DDPS_AMG_NO_TAIL=1 keeps the per-level launches).
// ------------------------------------------------------------------------------------------------------------
struct TailLevel {
  int n, nc, nslices, wpr;   // wpr: warp-per-row kernels (n <= kAmgTailWpr), else lane-per-row
  const int *slice_ptr, *col, *rowptr;
  const double *val, *dinv;
  const int *Pptr, *Rptr;
  const int2 *Pent, *Rent;
  double *b, *x, *x2, *t;
};
constexpr int kAmgTailMax = 8;
constexpr int kAmgTailWpr = 4096;
struct TailArgs {
  int nl;                       // levels in the tail; the last one is the coarsest (dense)
  TailLevel lv[kAmgTailMax];
  const double* dense;
  double om;
  unsigned* bar;                // {phase counter, exit counter}
  const PcgScalars* scal;
};

__device__ __forceinline__ void tail_sync(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (*(volatile unsigned*)bar < target) {}
    __threadfence();
  }
  __syncthreads();
}

// y = b - A x (MODE 1) or y = x + om dinv (b - A x) (MODE 2) on one tail level
template <int MODE>
__device__ __forceinline__ void tail_spmv(const TailLevel& L, double om, const double* x, double* y, int gw, int nw, int lane) {
  if (L.wpr) {
    for (int row = gw; row < L.n; row += nw) {
      const int len = __ldg(L.rowptr + row + 1) - __ldg(L.rowptr + row);
      const int base = __ldg(L.slice_ptr + (row >> 5)) + (row & 31);
      double sum = 0.0;
      for (int k = lane; k < len; k += 32) sum += __ldg(L.val + base + 32 * k) * __ldcg(x + __ldg(L.col + base + 32 * k));
      sum = warp_sum(sum);
      if (lane == 0) {
        double out;
        if (MODE == 1) out = __ldcg(L.b + row) - sum;
        else out = __ldcg(x + row) + om * (__ldg(L.dinv + row) * (__ldcg(L.b + row) - sum));
        y[row] = out;
      }
    }
  } else {
    for (int slice = gw; slice < L.nslices; slice += nw) {
      const int base = __ldg(L.slice_ptr + slice);
      const int w = (__ldg(L.slice_ptr + slice + 1) - base) >> 5;
      int idx = base + lane;
      double sum = 0.0;
      int k = 0;
      for (; k + 4 <= w; k += 4) {
        const int c0 = __ldg(L.col + idx), c1 = __ldg(L.col + idx + 32), c2 = __ldg(L.col + idx + 64), c3 = __ldg(L.col + idx + 96);
        const double v0 = __ldg(L.val + idx), v1 = __ldg(L.val + idx + 32), v2 = __ldg(L.val + idx + 64), v3 = __ldg(L.val + idx + 96);
        const double x0 = __ldcg(x + c0), x1 = __ldcg(x + c1), x2 = __ldcg(x + c2), x3 = __ldcg(x + c3);
        sum += v0 * x0; sum += v1 * x1; sum += v2 * x2; sum += v3 * x3;
        idx += 128;
      }
      for (; k < w; ++k) {
        sum += __ldg(L.val + idx) * __ldcg(x + __ldg(L.col + idx));
        idx += 32;
      }
      const int row = slice * 32 + lane;
      if (row < L.n) {
        double out;
        if (MODE == 1) out = __ldcg(L.b + row) - sum;
        else out = __ldcg(x + row) + om * (__ldg(L.dinv + row) * (__ldcg(L.b + row) - sum));
        y[row] = out;
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_amg_tail(const __grid_constant__ TailArgs A) {
  if (A.scal->done) return;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  unsigned phase = 0;
  const unsigned G = gridDim.x;
  // ---- down: residual, restriction (+ first damped-Jacobi sweep of the coarse level from the zero guess) ----
  for (int l = 0; l + 1 < A.nl; ++l) {
    const TailLevel& L = A.lv[l];
    const TailLevel& C = A.lv[l + 1];
    tail_spmv<1>(L, A.om, L.x, L.t, gw, nw, lane);
    tail_sync(A.bar, ++phase * G);
    const bool sweep = (l + 2 < A.nl);
    {   // warp per coarse row (as k_amg_restrict_wpr)
      for (int I = gw; I < L.nc; I += nw) {
        const int q0 = __ldg(L.Rptr + I), q1 = __ldg(L.Rptr + I + 1);
        double s = 0.0;
        for (int q = q0 + lane; q < q1; q += 32) {
          const int2 en = __ldg(L.Rent + q);
          s += (double)__int_as_float(en.y) * __ldcg(L.t + en.x);
        }
        s = warp_sum(s);
        if (lane == 0) {
          C.b[I] = s;
          if (sweep) C.x[I] = A.om * (__ldg(C.dinv + I) * s);
        }
      }
    }
    tail_sync(A.bar, ++phase * G);
  }
  // ---- coarsest level: x2 = Ainv b, one warp per row ----
  {
    const TailLevel& L = A.lv[A.nl - 1];
    for (int i = gw; i < L.n; i += nw) {
      double s = 0.0;
      for (int j = lane; j < L.n; j += 32) s += __ldg(A.dense + (size_t)i * L.n + j) * __ldcg(L.b + j);
      s = warp_sum(s);
      if (lane == 0) L.x2[i] = s;
    }
    tail_sync(A.bar, ++phase * G);
  }
  // ---- up: prolongation, smoothing sweep ----
  for (int l = A.nl - 2; l >= 0; --l) {
    const TailLevel& L = A.lv[l];
    const TailLevel& C = A.lv[l + 1];
    for (int i = gt; i < L.n; i += nt) {
      double s = 0.0;
      const int e = __ldg(L.Pptr + i + 1);
#pragma unroll 1
      for (int m = __ldg(L.Pptr + i); m < e; ++m) {
        const int2 en = __ldg(L.Pent + m);
        if (en.x >= 0) s += (double)__int_as_float(en.y) * __ldcg(C.x2 + en.x);
      }
      L.x[i] = __ldcg(L.x + i) + s;
    }
    tail_sync(A.bar, ++phase * G);
    tail_spmv<2>(L, A.om, L.x, L.x2, gw, nw, lane);
    if (l > 0) tail_sync(A.bar, ++phase * G);
  }
  // ---- the last CTA to leave resets the counters for the next launch ----
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(A.bar + 1, 1u) == G - 1) { A.bar[0] = 0u; A.bar[1] = 0u; __threadfence(); }
  }
}

// ---- CG with a general preconditioner: control kernels -------------------------------------------------------
// r = rhs - q (q = A x0) or r = rhs; sums {r.r, rhs.rhs, x0.rhs}, max|r|
template <bool HAVE_Q>
__global__ void k_apcg_init(int n, const double* __restrict__ rhs, const double* __restrict__ q, const double* __restrict__ x0,
                            double* __restrict__ r, double* partials, PcgScalars* scal) {
  __shared__ double sh_red[32];
  double s_rr = 0.0, s_bb = 0.0, s_xb = 0.0, s_max = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double bi = rhs[i];
    const double ri = HAVE_Q ? bi - q[i] : bi;
    r[i] = ri;
    s_rr += ri * ri; s_bb += bi * bi;
    if (HAVE_Q) s_xb += x0[i] * bi;
    s_max = fmax(s_max, fabs(ri));
  }
  double v[4] = {s_rr, s_bb, s_xb, s_max}, tot[4];
  if (grid_reduce_mixed<4, 8u>(v, partials, &scal->cnt[1], sh_red, tot)) {
    if (threadIdx.x == 0) { scal->sums[0] = tot[0]; scal->sums[1] = tot[1]; scal->sums[2] = tot[2]; scal->sums[3] = tot[3]; }
    // (this rank's part; multi-rank contexts allreduce sums[0..3] -- three sums and a maximum -- before k_apcg_start)
  }
}

// Stop rule: ||r||_2 <= rtol*||rhs||_2  OR  ||r||_inf <= atol_inf, like the Jacobi path.  A WARM-started solve must
// in addition bring the energy norm of the error, estimated by r.z = r.Br ~ e.Ae, below rtol^2 * ||x||_A^2 (~ x0.rhs):
// the residual of a smooth error (the Gummel update of the previous outer iteration) is far below rtol*||rhs|| long
// before the error itself is, and a solve that returns its initial guess would end the Gummel loop early.
__global__ void k_apcg_start(PcgScalars* scal, double rtol, double atol_inf, int warm, double eps_e) {
  // a warm start that does not even reduce the residual (||b - A x0|| >= ||b||: the earlier solution this guess comes from has
  // little to do with the new system) is dropped: k_apcg_cold restarts the solve from x = 0
  scal->pad0 = (warm && scal->sums[0] >= scal->sums[1]) ? 1 : 0;
  if (scal->pad0) { warm = 0; scal->sums[0] = scal->sums[1]; scal->sums[2] = 0.0; scal->sums[3] = 1e300; }
  scal->rr = scal->sums[0]; scal->bb = scal->sums[1]; scal->xb = scal->sums[2]; scal->rmax = scal->sums[3];
  scal->thresh2 = rtol * rtol * scal->bb;
  scal->thresh_inf = atol_inf;
  scal->thresh_e = 0.0;
  scal->warm = warm;
  scal->eps_e2 = eps_e * eps_e;
  scal->iters = 0;
  scal->breakdown = 0;
  // with the energy rule armed nothing is known about the error yet: only an exactly zero residual ends the solve here
  if (warm || eps_e > 0.0) scal->done = (scal->rr == 0.0) ? 1 : 0;
  else scal->done = (scal->rr <= scal->thresh2 || scal->rmax <= atol_inf) ? 1 : 0;
  scal->conv = scal->done;   // every stop rule met (an exhausted iteration cap leaves it 0 -> DDPS_ERR_NOCONV)
}

// x += alpha p; r -= alpha q; the block that finishes last takes the convergence decision, so that the cycle
// kernels behind it exit immediately once the residual is small enough.
// Fused: xs = om * dinv * r, the first damped-Jacobi sweep of the V-cycle on the finest level (saves a pass over r).
// The convergence decision from the (global) sums {r.r, max|r|}: by the last block of k_apcg_update on one rank, by the
// one-thread kernel k_apcg_decide after the allreduce in multi-rank contexts (identical on every rank).
__global__ void k_apcg_cold(int n, const double* __restrict__ rhs, double* __restrict__ x, double* __restrict__ r,
                            const PcgScalars* scal) {
  if (!scal->pad0) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { x[i] = 0.0; r[i] = rhs[i]; }
}

__device__ __forceinline__ void apcg_decide(PcgScalars* scal, double rr, double rmax, double rz, int it, int max_it) {
  scal->rr = rr;
  scal->rmax = rmax;
  scal->iters = it + 1;
  // rz = r.Br of the residual BEFORE this update bounds the energy of the error after it (CG decreases it)
  const bool small_r = rr <= scal->thresh2 || rmax <= scal->thresh_inf;
  const bool small_e = !(scal->warm || scal->eps_e2 > 0.0) || rz <= scal->thresh_e;
  const bool conv = (small_r && small_e) || rr == 0.0;
  if (conv) scal->conv = 1;
  if (conv || it + 1 >= max_it || !(rr == rr)) scal->done = 1;
}
__global__ void k_apcg_decide(PcgScalars* scal, int par, int it, int max_it) {
  if (scal->done) return;
  apcg_decide(scal, scal->sums[2], scal->sums[3], scal->rz[par], it, max_it);
}

template <bool DECIDE>
__global__ void __launch_bounds__(kVecBlock, 8) k_apcg_update(int n, int par, int it, int max_it,
                                                           const double* __restrict__ p, const double* __restrict__ q,
                                                           double* __restrict__ x, double* __restrict__ r,
                                                           const double* __restrict__ dinv, double om, double* __restrict__ xs,
                                                           double* partials, PcgScalars* scal) {
  __shared__ double sh_red[32];
  pdl_enter();
  if (scal->done) return;
  const double pq = scal->sums[0];
  const double rz = scal->rz[par];
  if (!(pq > 0.0) || !(rz == rz)) {   // operator (or preconditioner) not SPD, or NaN
    if (blockIdx.x == 0 && threadIdx.x == 0) { scal->breakdown = 1; scal->iters = it; }
    // `done` is raised by the last block below so that no block can observe it before taking this branch
    double v[1] = {0.0}, tot[1];
    if (grid_reduce<1, false>(v, partials, &scal->cnt[1], sh_red, tot)) {
      if (threadIdx.x == 0) scal->done = 1;
    }
    return;
  }
  const double alpha = rz / pq;
  double s_rr = 0.0, s_max = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double xi = x[i] + alpha * p[i];
    const double ri = r[i] - alpha * q[i];
    x[i] = xi;
    r[i] = ri;
    xs[i] = om * (dinv[i] * ri);
    s_rr += ri * ri;
    s_max = fmax(s_max, fabs(ri));
  }
  double v[2] = {s_rr, s_max}, tot[2];
  if (grid_reduce_mixed<2, 2u>(v, partials, &scal->cnt[1], sh_red, tot)) {
    if (threadIdx.x == 0) {
      if (DECIDE) apcg_decide(scal, tot[0], tot[1], rz, it, max_it);
      else { scal->sums[2] = tot[0]; scal->sums[3] = tot[1]; }
    }
  }
}

// p = z + beta p  (first: p = z); r.z of this iteration was left in sums[1] by the last kernel of the cycle
__global__ void __launch_bounds__(kVecBlock, 8) k_apcg_direction(int n, int par, int first, double rtol,
                                                              const double* __restrict__ z, double* __restrict__ p,
                                                              PcgScalars* scal) {
  pdl_enter();
  if (scal->done) return;
  const double rz_new = scal->sums[1];
  const double beta = first ? 0.0 : rz_new / scal->rz[par];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal->rz[first ? 0 : (par ^ 1)] = rz_new;
    if (first) {
      // r0.Br0 ~ the energy of what this solve adds to x0, x0.b ~ the energy of x0 itself (warm start): E_ref carries the largest
      // SOLUTION energy across the corrections of a Newton loop (a warm-started correction counts with its full size, not with
      // the little that was left to add)
      const double eref = fmax(scal->e_ref, fmax(rz_new, fabs(scal->xb)));
      scal->e_ref = eref;
      const double e2 = scal->eps_e2 > 0.0 ? scal->eps_e2 : rtol * rtol;
      scal->thresh_e = e2 * eref;
    }
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    p[i] = first ? z[i] : z[i] + beta * p[i];
}

// ------------------------------------------------------------------------------------------------------------
// setup / teardown
// ------------------------------------------------------------------------------------------------------------
bool amg_ready(const Context* ctx) { return ctx->amg && ctx->amg->ready; }

void amg_free(Context* ctx) {
  Amg* A = ctx->amg;
  if (!A) return;
  prof_report(ctx);
  for (int l = 0; l < kAmgMaxLevels; ++l) {
    AmgLevelDev& L = A->lv[l];
    if (L.owns) { amg_release(L.slice_ptr); amg_release(L.col); amg_release(L.rowptr); }
    amg_release(L.val); amg_release(L.dinv); amg_release(L.val32);
    amg_release(L.Pptr); amg_release(L.Pent); amg_release(L.APptr); amg_release(L.APcol); amg_release(L.APval);
    amg_release(L.Rptr); amg_release(L.Rent);
    if (l > 0) { amg_release(L.b); amg_release(L.x2); }
    amg_release(L.x); amg_release(L.t);
    amg_release(L.glob); amg_release(L.ap_send_pos); amg_release(L.ap_send_buf);
    if (L.owns_plan && L.plan) { plan_free(*L.plan); delete L.plan; }
    L.plan = nullptr;
  }
  amg_release(A->dense);
  amg_release(A->z);
  amg_release(A->tail_bar);
  amg_release(A->tail_parts);
  if (A->tail_plan) { plan_free(*A->tail_plan); delete A->tail_plan; A->tail_plan = nullptr; }
  for (int k = 0; k < kAmgKinds; ++k) {
    AmgOpSet& S = A->sets[k];
    auto rel = [&](auto*& p) { if (p) { cudaFreeAsync(p, ctx->stream); p = nullptr; } };
    rel(S.snap); rel(S.dense);
    for (int l = 0; l < kAmgMaxLevels; ++l) { rel(S.val[l]); rel(S.dinv[l]); rel(S.val32[l]); }
  }
  delete A;
  ctx->amg = nullptr;
}

// upload helpers of the host-built part of the hierarchy
int amg_install_pattern(Context* ctx, AmgLevelDev& L, const AmgHostLevel& HL) {
  int rc;
  L = AmgLevelDev();
  L.n = HL.n;
  L.n_loc = HL.n;
  L.nnz = (int64_t)HL.col.size();
  // SELL-32 layout of the coarse pattern (same convention as the fine pattern: padding = value 0, own column)
  L.nslices = (HL.n + 31) / 32;
  std::vector<int> hsp(L.nslices + 1, 0);
  for (int s = 0; s < L.nslices; ++s) {
    int w = 0;
    for (int r = s * 32; r < std::min(HL.n, s * 32 + 32); ++r) w = std::max(w, HL.rowptr[r + 1] - HL.rowptr[r]);
    hsp[s + 1] = hsp[s] + 32 * w;
    L.wmax = std::max(L.wmax, w);
  }
  L.nentries = hsp[L.nslices];
  std::vector<int> hcol((size_t)L.nentries, 0);
  for (int s = 0; s < L.nslices; ++s) {
    const int w = (hsp[s + 1] - hsp[s]) / 32;
    for (int lane = 0; lane < 32; ++lane) {
      const int r = s * 32 + lane;
      const int beg = r < HL.n ? HL.rowptr[r] : 0, len = r < HL.n ? HL.rowptr[r + 1] - beg : 0;
      const int self = r < HL.n ? r : 0;
      for (int k = 0; k < w; ++k) hcol[(size_t)hsp[s] + 32 * k + lane] = k < len ? HL.col[beg + k] : self;
    }
  }
  L.owns = true;
  if ((rc = amg_upload(ctx, &L.slice_ptr, hsp)) || (rc = amg_upload(ctx, &L.col, hcol)) || (rc = amg_upload(ctx, &L.rowptr, HL.rowptr)))
    return rc;
  if ((rc = amg_alloc(ctx, &L.val, (size_t)L.nentries)) || (rc = amg_alloc(ctx, &L.dinv, (size_t)L.n))) return rc;
  DDPS_CUDA(cudaMemsetAsync(L.val, 0, std::max<size_t>((size_t)L.nentries, 1) * sizeof(double), ctx->stream));
  DDPS_CUDA(cudaMemsetAsync(L.dinv, 0, std::max<size_t>((size_t)L.n, 1) * sizeof(double), ctx->stream));
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));   // the host vectors go out of scope
  return 0;
}

int amg_install_transfers(Context* ctx, AmgLevelDev& L, const AmgHostLevel& HL) {
  int rc;
  L.nc = HL.nc;
  std::vector<int2> pe(HL.Pcol.size()), re(HL.Rrow.size());
  for (size_t k = 0; k < pe.size(); ++k) {
    const float pv = (float)HL.Pval[k], rv = (float)HL.Rval[k];   // exact: the host rounded them to fp32 already
    int pb, rb;
    memcpy(&pb, &pv, 4); memcpy(&rb, &rv, 4);
    pe[k] = make_int2(HL.Pcol[k], pb);
    re[k] = make_int2(HL.Rrow[k], rb);
  }
  if ((rc = amg_upload(ctx, &L.Pptr, HL.Pptr)) || (rc = amg_upload(ctx, &L.Pent, pe)) || (rc = amg_upload(ctx, &L.Rptr, HL.Rptr)) ||
      (rc = amg_upload(ctx, &L.Rent, re)))
    return rc;
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Hierarchy of the freshly assembled base operator.  Levels with more than DDPS_AMG_DEV_MIN rows (default 100000) are
// coarsened on the device (amg_setup.cu: only the strong-neighbour lists travel to the host for the greedy
// aggregation); the small rest of the hierarchy is built by the host reference implementation (amg_host.cu) and
// uploaded.  DDPS_AMG_HOST_SETUP=1 forces the host path for every level (A/B and fallback).
// Damping of the cycle's Jacobi sweeps.  omega * rho(D^-1 A) < 2 keeps the sweep convergent and the V-cycle symmetric positive
// definite; the Gershgorin bound rho_G = max_i sum_j |a_ij| / a_ii of the BASE operator (2 for the stiffness matrix of an acute mesh:
// an M-matrix with zero row sums) is computed once per hierarchy, and omega = min(0.8, 1.6 / rho_G): 0.8 is the optimal point-Jacobi
// damping of the 5-point / P1 stencil (smoothing factor 0.6 instead of 0.78 at 2/3), meshes with obtuse triangles (positive
// off-diagonals, rho_G > 2) get less.  Measured at 16.8M triangles: 501 -> 459 CG iterations per solve.  The prolongator smoothing
// keeps prm.omega = 2/3.
__global__ void k_amg_gershgorin(int nrows, const int* __restrict__ slice_ptr, const double* __restrict__ val,
                                 const unsigned short* __restrict__ rowinfo, double* partials, PcgScalars* scal) {
  __shared__ double sh_red[32];
  double m = 0.0;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += gridDim.x * blockDim.x) {
    const int base = slice_ptr[row >> 5] + (row & 31);
    const int len = rowinfo[row] >> 5, sd = rowinfo[row] & 31;
    double sum = 0.0, d = 0.0;
    for (int k = 0; k < len; ++k) {
      const double v = val[base + 32 * k];
      sum += fabs(v);
      if (k == sd) d = v;
    }
    if (d > 0.0) m = fmax(m, sum / d);
  }
  double v[1] = {m}, tot[1];
  if (grid_reduce<1, true>(v, partials, &scal->cnt[3], sh_red, tot)) {
    if (threadIdx.x == 0) scal->sums[0] = tot[0];
  }
}

static int amg_cycle_omega(Context* ctx, Amg* A, double fallback) {
  if (const char* e = getenv("DDPS_AMG_OMEGA_CYCLE")) { A->omega = atof(e); return 0; }   // (experiments)
  A->omega = fallback;
  const Pattern& P = ctx->pat;
  if (!P.rowinfo || !ctx->val_base) return 0;
  int rc;
  if (P.nrows > 0)
    k_amg_gershgorin<<<amg_grid(ctx, P.nrows, 256), 256, 0, ctx->stream>>>(P.nrows, P.slice_ptr, ctx->val_base, P.rowinfo, ctx->partials, ctx->scal);
  else
    DDPS_CUDA(cudaMemsetAsync(&ctx->scal->sums[0], 0, sizeof(double), ctx->stream));
  ctx->stats.kernel_launches++;
  if ((rc = comm_allreduce(ctx, &ctx->scal->sums[0], 1, COMM_MAX))) return rc;   // (collective; no-op on one rank)
  double rho = 0.0;
  DDPS_CUDA(cudaMemcpyAsync(&rho, &ctx->scal->sums[0], sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (rho > 0.0 && rho == rho) A->omega = std::min(0.8, 1.6 / rho);
  if (getenv("DDPS_AMG_TIMING") && ctx->rank == 0) fprintf(stderr, "[ddps amg] Gershgorin bound of D^-1 A (base operator) %.4f -> cycle damping %.4f\n", rho, A->omega);
  return 0;
}

static int amg_finish_setup(Context* ctx, Amg* A, int nlev, double t0);

// Multi-rank contexts: distributed coarsening (amg_setup.cu: amg_dist_coarsen) while the level has more than DDPS_AMG_TAIL
// rows in total (default 200000), then the transition to the replicated rest of the hierarchy (amg_dist_transition).
static int amg_setup_dist(Context* ctx) {
  const double t0 = wall_ms();
  const Pattern& P = ctx->pat;
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  AmgParams prm;
  // (at least 8 x coarse_max rows: the replicated part must still coarsen once, the coarsest level is solved directly)
  const long long tail = std::max<long long>(8LL * prm.coarse_max, getenv("DDPS_AMG_TAIL") ? atoll(getenv("DDPS_AMG_TAIL")) : 200000);
  const bool timing = getenv("DDPS_AMG_TIMING") != nullptr;
  Amg* A = new Amg();
  ctx->amg = A;
  A->prof.on = getenv("DDPS_PROFILE_ITER") != nullptr;
  int rc;
  if ((rc = amg_cycle_omega(ctx, A, prm.omega))) return rc;
  {
    AmgLevelDev& L = A->lv[0];
    L.n = P.nrows; L.n_loc = ctx->n_loc; L.nnz = P.nnz; L.nslices = P.nslices; L.nentries = P.nentries; L.wmax = P.max_row_len;
    L.slice_ptr = P.slice_ptr; L.col = P.col; L.rowptr = P.rowptr; L.owns = false;
    L.cur_val = ctx->val_base;
    L.plan = &ctx->halo; L.owns_plan = false;
  }
  long long ng = ctx->nq_global;
  int l = 0, nlev = 0;
  while (true) {
    if (ng > tail && l + 2 < prm.max_levels) {
      long long ncg = 0;
      const double ta = wall_ms();
      rc = amg_dist_coarsen(ctx, A, l, prm, &ncg);
      if (rc < 0) return rc;
      if (rc > 0 || ncg == 0) {   // does not apply / stalled: Jacobi-PCG stays (same decision on all ranks)
        if (timing) fprintf(stderr, "[ddps amg] rank %d: distributed coarsening of level %d (%lld rows) %s -> no hierarchy\n", ctx->rank, l, ng,
                            rc > 0 ? "does not apply" : "stalled");
        amg_free(ctx);
        return 0;
      }
      if (timing && ctx->rank == 0)
        fprintf(stderr, "[ddps amg] distributed level %d: %lld -> %lld rows (rank 0: %d -> %d owned, %d ghost columns): %.1f ms\n", l, ng, ncg,
                A->lv[l].n, A->lv[l + 1].n, A->lv[l + 1].n_loc - A->lv[l + 1].n, wall_ms() - ta);
      ng = ncg;
      ++l;
      continue;
    }
    const double ta = wall_ms();
    if ((rc = amg_dist_transition(ctx, A, l, prm, &nlev))) return rc;
    if (timing && ctx->rank == 0) fprintf(stderr, "[ddps amg] transition at level %d (%lld rows gathered), %d levels in all: %.1f ms\n", l, ng, nlev, wall_ms() - ta);
    break;
  }
  if (nlev < 2 || A->lv[nlev - 1].n > kAmgDenseMax) {
    if (timing) fprintf(stderr, "[ddps amg] rank %d: %d levels, coarsest %d rows -> no hierarchy\n", ctx->rank, nlev, nlev ? A->lv[nlev - 1].n : 0);
    amg_free(ctx);
    return 0;
  }
  return amg_finish_setup(ctx, A, nlev, t0);
}

int amg_setup(Context* ctx) {
  amg_free(ctx);
  if (!ctx->base_ready) return 0;
  if (ctx->nranks != 1) return amg_setup_dist(ctx);
  const double t0 = wall_ms();
  const Pattern& P = ctx->pat;
  const int n = P.nrows;
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  AmgParams prm;
  if (const char* e = getenv("DDPS_AMG_COARSE_MAX")) prm.coarse_max = std::max(8, std::min(kAmgDenseMax, atoi(e)));   // experiments
  const bool host_only = getenv("DDPS_AMG_HOST_SETUP") != nullptr;
  const int dev_min = getenv("DDPS_AMG_DEV_MIN") ? atoi(getenv("DDPS_AMG_DEV_MIN")) : 100000;
  const bool timing = getenv("DDPS_AMG_TIMING") != nullptr;

  Amg* A = new Amg();
  ctx->amg = A;
  A->prof.on = getenv("DDPS_PROFILE_ITER") != nullptr;
  int rc;
  if ((rc = amg_cycle_omega(ctx, A, prm.omega))) return rc;
  {
    AmgLevelDev& L = A->lv[0];
    L.n = n; L.nnz = P.nnz; L.nslices = P.nslices; L.nentries = P.nentries; L.wmax = P.max_row_len;
    L.slice_ptr = P.slice_ptr; L.col = P.col; L.rowptr = P.rowptr; L.owns = false;
    L.cur_val = ctx->val_base;
  }
  int nlev = 1;
  // ---- device phase ----
  while (!host_only && A->lv[nlev - 1].n > dev_min && A->lv[nlev - 1].n > prm.coarse_max && nlev < prm.max_levels) {
    int nc = 0;
    const double ta = wall_ms();
    rc = amg_dev_coarsen(ctx, A, nlev - 1, nlev == 1 ? ctx->isD : nullptr, prm, &nc);
    if (rc < 0) return rc;
    if (rc > 0 || nc == 0) break;
    if (timing) fprintf(stderr, "[ddps amg] level %d n=%d -> %d on the device: %.1f ms\n", nlev - 1, A->lv[nlev - 1].n, nc, wall_ms() - ta);
    ++nlev;
  }
  const double t1 = wall_ms();
  // ---- host phase: the rest of the hierarchy from level ld ----
  const int ld = nlev - 1;
  if (A->lv[ld].n > prm.coarse_max && nlev < prm.max_levels) {
    const AmgLevelDev& LD = A->lv[ld];
    const int m = LD.n;
    std::vector<int> rp((size_t)m + 1), sp((size_t)LD.nslices + 1), scol((size_t)LD.nentries), cc((size_t)LD.nnz);
    std::vector<double> sv((size_t)LD.nentries), vv((size_t)LD.nnz);
    std::vector<unsigned char> fixed;
    DDPS_CUDA(cudaMemcpy(rp.data(), LD.rowptr, rp.size() * sizeof(int), cudaMemcpyDeviceToHost));
    DDPS_CUDA(cudaMemcpy(sp.data(), LD.slice_ptr, sp.size() * sizeof(int), cudaMemcpyDeviceToHost));
    DDPS_CUDA(cudaMemcpy(scol.data(), LD.col, scol.size() * sizeof(int), cudaMemcpyDeviceToHost));
    DDPS_CUDA(cudaMemcpy(sv.data(), LD.cur_val, sv.size() * sizeof(double), cudaMemcpyDeviceToHost));
    if (ld == 0) {
      fixed.resize(m);
      DDPS_CUDA(cudaMemcpy(fixed.data(), ctx->isD, fixed.size(), cudaMemcpyDeviceToHost));
    }
    {
      const int nt = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
      std::vector<std::thread> th;
      const int chunk = (m + nt - 1) / nt;
      for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t]() {
          for (int r = t * chunk; r < std::min(m, (t + 1) * chunk); ++r) {
            const int base = sp[r >> 5], lane = r & 31;
            for (int k = 0; k < rp[r + 1] - rp[r]; ++k) {
              cc[rp[r] + k] = scol[(size_t)base + 32 * k + lane];
              vv[rp[r] + k] = sv[(size_t)base + 32 * k + lane];
            }
          }
        });
      for (auto& x : th) x.join();
    }
    std::vector<double>().swap(sv);
    std::vector<int>().swap(scol);
    AmgHost H;
    AmgParams sub = prm;
    sub.max_levels = prm.max_levels - ld;
    const int nh = amg_host_build(H, m, std::move(rp), std::move(cc), std::move(vv), ld == 0 ? fixed.data() : nullptr, sub);
    for (int k = 0; k < nh; ++k) {
      AmgLevelDev& L = A->lv[ld + k];
      if (k >= 1 && (rc = amg_install_pattern(ctx, L, H.lv[k]))) return rc;
      if (k + 1 < nh && (rc = amg_install_transfers(ctx, L, H.lv[k]))) return rc;
    }
    nlev = ld + nh;
  }
  const double t2 = wall_ms();
  if (nlev < 2 || A->lv[nlev - 1].n > kAmgDenseMax) { amg_free(ctx); return 0; }   // no hierarchy: Jacobi-PCG stays
  return amg_finish_setup(ctx, A, nlev, t0);
}

// cycle vectors, fp32 copies, coarsest inverse (vectors of distributed levels carry their ghost entries)
static int amg_finish_setup(Context* ctx, Amg* A, int nlev, double t0) {
  int rc;
  A->nlev = nlev;
  const bool cycle_fp32 = getenv("DDPS_AMG_FP64_CYCLE") == nullptr;   // A/B switch: keep the cycle's matrices in FP64
  const int fp32_min = getenv("DDPS_AMG_FP32_MIN") ? atoi(getenv("DDPS_AMG_FP32_MIN")) : 50000;   // (tests lower it)
  for (int l = 0; l < nlev; ++l) {
    AmgLevelDev& L = A->lv[l];
    if (L.n_loc < L.n) L.n_loc = L.n;
    if (cycle_fp32 && L.n > fp32_min && (rc = amg_alloc(ctx, &L.val32, (size_t)L.nentries))) return rc;
    if ((rc = amg_alloc(ctx, &L.x, (size_t)L.n_loc)) || (rc = amg_alloc(ctx, &L.t, (size_t)L.n_loc))) return rc;
    DDPS_CUDA(cudaMemsetAsync(L.x, 0, std::max<size_t>((size_t)L.n_loc, 1) * sizeof(double), ctx->stream));
    DDPS_CUDA(cudaMemsetAsync(L.t, 0, std::max<size_t>((size_t)L.n_loc, 1) * sizeof(double), ctx->stream));
    if (l >= 1) {
      if ((rc = amg_alloc(ctx, &L.b, (size_t)L.n)) || (rc = amg_alloc(ctx, &L.x2, (size_t)L.n_loc))) return rc;
      DDPS_CUDA(cudaMemsetAsync(L.x2, 0, std::max<size_t>((size_t)L.n_loc, 1) * sizeof(double), ctx->stream));
    }
  }
  const int nL = A->lv[nlev - 1].n;
  if ((rc = amg_alloc(ctx, &A->dense, (size_t)nL * nL))) return rc;
  DDPS_CUDA(cudaFuncSetAttribute(k_amg_dense_inverse<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kAmgDenseShared * kAmgDenseShared * (int)sizeof(double)));
  if ((rc = amg_alloc(ctx, &A->z, (size_t)A->lv[0].n))) return rc;
  // transition level of a distributed hierarchy: staged all-to-all of the partial right-hand sides (collective: every rank
  // takes the same decision -- sizes, capacity and the switch are identical everywhere).  OPT-IN (DDPS_AMG_TAIL_P2P=1): measured
  // at 16.8M triangles it is no faster than the library allreduce it replaces (8 GPUs: 47.9 vs 42.9 us, 2 GPUs: 25.8 vs 20 us).
  if (ctx->nranks > 1 && getenv("DDPS_AMG_TAIL_P2P")) {
    for (int l = 0; l + 1 < nlev; ++l) {
      if (A->lv[l].dist != 2) continue;
      const int nc = A->lv[l + 1].n, nr = ctx->nranks, me = ctx->rank;
      if (!ctx->p2p.xenabled || (size_t)(nr - 1) * nc > ctx->p2p.stage_cap) break;
      HaloPlan* P = new HaloPlan();
      P->n_own = nc; P->n_ghost = (nr - 1) * nc; P->nneigh = nr - 1;
      P->send_off.assign(nr, 0); P->recv_off.assign(nr, 0);
      for (int q = 0, j = 0; q < nr; ++q) {
        if (q == me) continue;
        P->peer.push_back(q);
        ++j;
        P->send_off[j] = j * nc; P->recv_off[j] = j * nc;
      }
      P->nsend = (nr - 1) * nc;
      std::vector<int> idx((size_t)P->nsend);
      for (int k = 0; k < P->nsend; ++k) idx[k] = k % nc;
      DDPS_CUDA(cudaMalloc(&P->send_idx, std::max<size_t>(idx.size(), 1) * sizeof(int)));
      DDPS_CUDA(cudaMalloc(&P->send_buf, std::max<size_t>(idx.size(), 1) * sizeof(double)));
      DDPS_CUDA(cudaMemcpyAsync(P->send_idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
      DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
      A->tail_plan = P;
      if ((rc = plan_finish(ctx, *P))) return rc;
      if ((rc = amg_alloc(ctx, &A->tail_parts, (size_t)nr * nc))) return rc;
      DDPS_CUDA(cudaMemsetAsync(A->tail_parts, 0, (size_t)nr * nc * sizeof(double), ctx->stream));
      break;
    }
  }
  // fused tail of the cycle: from the first small, replicated, FP64-only level down (k_amg_tail)
  A->tail_from = 0;
  if (!getenv("DDPS_AMG_NO_TAIL")) {
    int lt = nlev;
    while (lt > 1 && A->lv[lt - 1].n <= kAmgSmallLevel && A->lv[lt - 1].dist == 0 && !A->lv[lt - 1].val32 && nlev - (lt - 1) <= kAmgTailMax) --lt;
    int coop = 0, per_sm = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_amg_tail, 256, 0);
    if (coop && per_sm >= 1 && nlev - lt >= 2) {   // (a tail of the coarsest level alone is the dense solve: nothing to fuse)
      A->tail_from = lt;
      // two CTAs per SM (measured at S16: 0.819 ms per CG iteration; one per SM 0.825, four 0.830, 64 CTAs in all 0.844): the phases are
      // latency-bound, more warps hide more of it until the barrier's cost takes over.  In-process ranks share one device: one per SM.
      A->tail_grid = (per_sm >= 2 && !ctx->lgroup) ? 2 * ctx->num_sms : ctx->num_sms;
      if (const char* e = getenv("DDPS_AMG_TAIL_GRID")) A->tail_grid = std::max(1, std::min(ctx->num_sms * per_sm, atoi(e)));   // (experiments)
      if (!A->tail_bar && (rc = amg_alloc(ctx, &A->tail_bar, 2))) return rc;
      DDPS_CUDA(cudaMemsetAsync(A->tail_bar, 0, 2 * sizeof(unsigned), ctx->stream));
    }
  }
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  A->ready = true;
  A->setup_ms = wall_ms() - t0;
  if (getenv("DDPS_AMG_TIMING") && ctx->rank == 0) {
    fprintf(stderr, "[ddps amg] setup %.1f ms, levels (rows owned by rank 0):", A->setup_ms);
    for (int l = 0; l < nlev; ++l) fprintf(stderr, " %d%s", A->lv[l].n, A->lv[l].dist == 1 ? "d" : A->lv[l].dist == 2 ? "t" : "");
    fprintf(stderr, "\n");
  }
  return 0;
}

int amg_debug_levels(Context* ctx, int64_t* nlevels, int64_t* sizes, int cap) {
  if (nlevels) *nlevels = amg_ready(ctx) ? ctx->amg->nlev : 0;
  if (sizes && amg_ready(ctx))
    for (int l = 0; l < ctx->amg->nlev && l < cap; ++l) sizes[l] = ctx->amg->lv[l].n;
  return 0;
}

// Pattern (CSR, sorted columns) and prolongator (CSR) of a level as they live on the device.
// sizes = {n, nnz, n_coarse (0 on the coarsest level), nnz(P)}; array pointers may be NULL.
int amg_debug_get_level(Context* ctx, int level, int64_t* sizes, int64_t* rowptr, int64_t* col, int64_t* Pptr, int64_t* Pcol,
                        double* Pval) {
  if (!amg_ready(ctx) || level < 0 || level >= ctx->amg->nlev) { ctx->fail(DDPS_ERR_ARG, "no such multigrid level"); return DDPS_ERR_ARG; }
  const AmgLevelDev& L = ctx->amg->lv[level];
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<int> pp;
  if (L.Pptr) {
    pp.resize((size_t)L.n + 1);
    DDPS_CUDA(cudaMemcpy(pp.data(), L.Pptr, pp.size() * sizeof(int), cudaMemcpyDeviceToHost));
  }
  if (sizes) { sizes[0] = L.n; sizes[1] = L.nnz; sizes[2] = L.Pptr ? L.nc : 0; sizes[3] = L.Pptr ? pp[L.n] : 0; }
  if (rowptr || col) {
    std::vector<int> rp((size_t)L.n + 1), sp((size_t)L.nslices + 1), sc((size_t)L.nentries);
    DDPS_CUDA(cudaMemcpy(rp.data(), L.rowptr, rp.size() * sizeof(int), cudaMemcpyDeviceToHost));
    DDPS_CUDA(cudaMemcpy(sp.data(), L.slice_ptr, sp.size() * sizeof(int), cudaMemcpyDeviceToHost));
    DDPS_CUDA(cudaMemcpy(sc.data(), L.col, sc.size() * sizeof(int), cudaMemcpyDeviceToHost));
    if (rowptr) for (size_t i = 0; i < rp.size(); ++i) rowptr[i] = rp[i];
    if (col)
      for (int r = 0; r < L.n; ++r)
        for (int k = 0; k < rp[r + 1] - rp[r]; ++k) col[rp[r] + k] = sc[(size_t)sp[r >> 5] + 32 * k + (r & 31)];
  }
  if (L.Pptr && (Pptr || Pcol || Pval)) {
    std::vector<int2> pe((size_t)pp[L.n]);
    DDPS_CUDA(cudaMemcpy(pe.data(), L.Pent, pe.size() * sizeof(int2), cudaMemcpyDeviceToHost));
    if (Pptr) for (size_t i = 0; i < pp.size(); ++i) Pptr[i] = pp[i];
    for (size_t k = 0; k < pe.size(); ++k) {
      float f;
      memcpy(&f, &pe[k].y, 4);
      if (Pcol) Pcol[k] = pe[k].x;
      if (Pval) Pval[k] = (double)f;
    }
  }
  return 0;
}

// values (CSR order of the level's pattern) and inverse diagonal of level >= 1 after the last amg_numeric
int amg_debug_get_values(Context* ctx, int level, double* val_csr, double* dinv) {
  if (!amg_ready(ctx) || level < 1 || level >= ctx->amg->nlev) { ctx->fail(DDPS_ERR_ARG, "no such multigrid level"); return DDPS_ERR_ARG; }
  Amg& A = *ctx->amg;
  const AmgLevelDev& L = A.lv[level];
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<double> sv((size_t)L.nentries);
  std::vector<int> rp((size_t)L.n + 1), sp((size_t)L.nslices + 1);
  DDPS_CUDA(cudaMemcpy(sv.data(), L.cur_val, sv.size() * sizeof(double), cudaMemcpyDeviceToHost));
  DDPS_CUDA(cudaMemcpy(rp.data(), L.rowptr, rp.size() * sizeof(int), cudaMemcpyDeviceToHost));
  DDPS_CUDA(cudaMemcpy(sp.data(), L.slice_ptr, sp.size() * sizeof(int), cudaMemcpyDeviceToHost));
  if (val_csr)
    for (int r = 0; r < L.n; ++r)
      for (int k = 0; k < rp[r + 1] - rp[r]; ++k) val_csr[rp[r] + k] = sv[(size_t)sp[r >> 5] + 32 * k + (r & 31)];
  if (dinv) DDPS_CUDA(cudaMemcpy(dinv, L.cur_dinv, (size_t)L.n * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// per-operator numeric phase and the V-cycle
// ------------------------------------------------------------------------------------------------------------
// How far is the operator from the snapshot its coarse operators were computed from?  out (scal->sums[0..3]) =
// {max|a - s|, max|s|, max|rowsum(a) - rowsum(s)|, max|rowsum(s)|} (entries: the stiffness part; row sums: the mass part)
__global__ void __launch_bounds__(kVecBlock, 8) k_amg_change(int nrows, int nslices, const int* __restrict__ slice_ptr,
                                                          const double* __restrict__ a, const double* __restrict__ s,
                                                          double* partials, PcgScalars* scal) {
  __shared__ double sh_red[32];
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  double m1 = 0.0, n1 = 0.0, m2 = 0.0, n2 = 0.0;
  for (int slice = gw; slice < nslices; slice += nw) {
    const int base = slice_ptr[slice];
    const int w = (slice_ptr[slice + 1] - base) >> 5;
    double ra = 0.0, rs = 0.0;
    for (int k = 0; k < w; ++k) {
      const double va = ld_stream(a + base + 32 * k + lane), vs = ld_stream(s + base + 32 * k + lane);
      m1 = fmax(m1, fabs(va - vs));
      n1 = fmax(n1, fabs(vs));
      ra += va; rs += vs;
    }
    if (slice * 32 + lane < nrows) { m2 = fmax(m2, fabs(ra - rs)); n2 = fmax(n2, fabs(rs)); }
  }
  double v[4] = {m1, n1, m2, n2}, tot[4];
  if (grid_reduce_mixed<4, 15u>(v, partials, &scal->cnt[3], sh_red, tot)) {
    if (threadIdx.x == 0) { scal->sums[0] = tot[0]; scal->sums[1] = tot[1]; scal->sums[2] = tot[2]; scal->sums[3] = tot[3]; }
  }
}

// inverse diagonal from the (completed) values of a level
__global__ void k_amg_dinv(int n, const int* __restrict__ rowptr, const int* __restrict__ slice_ptr, const int* __restrict__ col,
                           const double* __restrict__ val, double* __restrict__ dinv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int len = rowptr[i + 1] - rowptr[i];
  const int base = slice_ptr[i >> 5] + (i & 31);
  for (int k = 0; k < len; ++k)
    if (col[base + 32 * k] == i) dinv[i] = 1.0 / val[base + 32 * k];
}

static int amg_launch_rap_impl(Context* ctx, const AmgLevelDev& F, const double* f_val, const AmgLevelDev& C, double* out_val, double* out_dinv);

int amg_launch_rap(Context* ctx, const AmgLevelDev& F, const double* f_val, const AmgLevelDev& C, double* out_val, double* out_dinv) {
  int rc = amg_launch_rap_impl(ctx, F, f_val, C, out_val, out_dinv);
  if (rc) return rc;
  if (F.dist == 2) {
    // transition level: every rank summed over ITS fine rows only; the replicated coarse operator is the sum over the ranks
    if ((rc = comm_allreduce(ctx, out_val, (size_t)C.nentries, COMM_SUM))) return rc;
    k_amg_dinv<<<(C.n + 255) / 256, 256, 0, ctx->stream>>>(C.n, C.rowptr, C.slice_ptr, C.col, out_val, out_dinv);
    ctx->stats.kernel_launches++;
  }
  return 0;
}

static int amg_launch_rap_impl(Context* ctx, const AmgLevelDev& F, const double* f_val, const AmgLevelDev& C, double* out_val, double* out_dinv) {
  if (F.APptr) {   // two-stage product
    if (F.n) k_amg_ap<<<(F.n + 127) / 128, 128, 0, ctx->stream>>>(F.n, F.slice_ptr, F.col, f_val, F.Pptr, F.Pent, F.APptr, F.APcol, F.APval);
    // distributed level: the A*P rows of the neighbours' boundary rows (my ghost rows) arrive by one exchange of values
    if (F.dist == 1) {
      int rc = amg_dist_ap_halo(ctx, const_cast<AmgLevelDev&>(F));
      if (rc) return rc;
    }
    if (C.n == 0) { ctx->stats.kernel_launches++; return 0; }
    if (C.wmax <= 16) {
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)C.n + 7) / 8, (int64_t)ctx->num_sms * 6));
      k_amg_rap_ap<16, 8><<<grid, 256, 0, ctx->stream>>>(C.n, F.Rptr, F.Rent, F.APptr, F.APcol, F.APval, C.rowptr, C.slice_ptr, C.col, out_val,
                                                         out_dinv);
    } else if (C.wmax <= 32) {
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)C.n + 3) / 4, (int64_t)ctx->num_sms * 6));
      k_amg_rap_ap<32, 4><<<grid, 128, 0, ctx->stream>>>(C.n, F.Rptr, F.Rent, F.APptr, F.APcol, F.APval, C.rowptr, C.slice_ptr, C.col, out_val,
                                                         out_dinv);
    } else {
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)C.n + 1) / 2, (int64_t)ctx->num_sms * 6));
      k_amg_rap_ap<64, 2><<<grid, 64, 0, ctx->stream>>>(C.n, F.Rptr, F.Rent, F.APptr, F.APcol, F.APval, C.rowptr, C.slice_ptr, C.col, out_val,
                                                        out_dinv);
    }
    ctx->stats.kernel_launches += 2;
    return 0;
  }
  if (F.dist) { ctx->fail(DDPS_ERR_STATE, "distributed multigrid level without an A*P pattern"); return DDPS_ERR_STATE; }
  if (C.wmax <= 16) {   // short coarse rows: small private tables, more resident warps per SM
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)C.n + 7) / 8, (int64_t)ctx->num_sms * 5));
    k_amg_rap2<16, 8><<<grid, 256, 0, ctx->stream>>>(C.n, F.slice_ptr, F.col, f_val, F.Pptr, F.Pent, F.Rptr, F.Rent, C.rowptr,
                                                     C.slice_ptr, C.col, out_val, out_dinv);
  } else if (C.wmax <= 32) {
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)C.n + 3) / 4, (int64_t)ctx->num_sms * 6));
    k_amg_rap2<32, 4><<<grid, 128, 0, ctx->stream>>>(C.n, F.slice_ptr, F.col, f_val, F.Pptr, F.Pent, F.Rptr, F.Rent, C.rowptr,
                                                     C.slice_ptr, C.col, out_val, out_dinv);
  } else {              // the last, gently coarsened levels have long rows; longer ones still take the serial walk inside
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)C.n + 1) / 2, (int64_t)ctx->num_sms * 6));
    k_amg_rap2<64, 2><<<grid, 64, 0, ctx->stream>>>(C.n, F.slice_ptr, F.col, f_val, F.Pptr, F.Pent, F.Rptr, F.Rent, C.rowptr,
                                                    C.slice_ptr, C.col, out_val, out_dinv);
  }
  ctx->stats.kernel_launches++;
  return 0;
}

// The per-kind copies come from the stream-ordered pool (only the library stream touches them): when lagging is armed, a
// few outer iterations into the solve, the pool still holds the setup's freed temporaries, so nothing new is mapped (a
// legacy cudaMalloc of these ~3 GB was measured at 0.3 s inside one Gummel iteration).
template <typename T>
static int amg_alloc_pooled(Context* ctx, T** p, size_t n) {
  DDPS_CUDA(cudaMallocAsync((void**)p, std::max<size_t>(n, 1) * sizeof(T), ctx->stream));
  return 0;
}

static int amg_alloc_set(Context* ctx, Amg& A, AmgOpSet& S) {
  int rc;
  if ((rc = amg_alloc_pooled(ctx, &S.snap, (size_t)A.lv[0].nentries))) return rc;
  for (int l = 1; l < A.nlev; ++l) {
    const AmgLevelDev& L = A.lv[l];
    if ((rc = amg_alloc_pooled(ctx, &S.val[l], (size_t)L.nentries)) || (rc = amg_alloc_pooled(ctx, &S.dinv[l], (size_t)L.n))) return rc;
    DDPS_CUDA(cudaMemsetAsync(S.val[l], 0, std::max<size_t>((size_t)L.nentries, 1) * sizeof(double), ctx->stream));   // SELL padding stays 0
    DDPS_CUDA(cudaMemsetAsync(S.dinv[l], 0, std::max<size_t>((size_t)L.n, 1) * sizeof(double), ctx->stream));
    if (L.val32 && (rc = amg_alloc_pooled(ctx, &S.val32[l], (size_t)L.nentries))) return rc;
  }
  const int nL = A.lv[A.nlev - 1].n;
  if ((rc = amg_alloc_pooled(ctx, &S.dense, (size_t)nL * nL))) return rc;
  S.allocated = true;
  return 0;
}

// kind: which operator this is (0: first Newton system of a Vsolve / semlin, 3: later Newton systems, 1: u-, 2: v-continuity
// system; < 0: unknown -> always recompute).  Returns through the level structs the matrices the cycle will use.
int amg_numeric(Context* ctx, const double* val, const double* dinv, int kind) {
  Amg& A = *ctx->amg;
  int rc;
  A.lv[0].cur_val = val;
  A.lv[0].cur_dinv = dinv;
  const bool lag = kind >= 0 && kind < kAmgKinds && !ctx->opts.amg_no_lag;
  AmgOpSet* S = lag ? &A.sets[kind] : nullptr;
  if (S && !S->allocated && (rc = amg_alloc_set(ctx, A, *S))) return rc;
  bool reuse = false;
  if (S && S->valid) {
    const AmgLevelDev& L0 = A.lv[0];
    k_amg_change<<<amg_grid(ctx, (int64_t)L0.nslices * 32, kVecBlock), kVecBlock, 0, ctx->stream>>>(L0.n, L0.nslices, L0.slice_ptr, val, S->snap,
                                                                                               ctx->partials, ctx->scal);
    ctx->stats.kernel_launches++;
    if ((rc = scalar_allreduce(ctx, ctx->scal->sums, 0, 4, false))) return rc;   // every rank takes the same decision
    DDPS_CUDA(cudaMemcpyAsync(ctx->scal_host, ctx->scal, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream));
    DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
    const double* m = ctx->scal_host->sums;
    const double tau = 0.01;
    reuse = (m[0] <= tau * m[1]) && (m[2] <= tau * m[3] || m[2] <= 1e-12 * m[1]);
  }
  for (int l = 1; l < A.nlev; ++l) {
    AmgLevelDev& L = A.lv[l];
    L.cur_val = S ? S->val[l] : L.val;
    L.cur_dinv = S ? S->dinv[l] : L.dinv;
    L.cur32 = S ? S->val32[l] : L.val32;
  }
  A.cur_dense = S ? S->dense : A.dense;
  if (!reuse) {
    for (int l = 0; l + 1 < A.nlev; ++l) {
      const AmgLevelDev& F = A.lv[l];
      const AmgLevelDev& C = A.lv[l + 1];
      if ((rc = amg_launch_rap(ctx, F, F.cur_val, C, const_cast<double*>(C.cur_val), const_cast<double*>(C.cur_dinv)))) return rc;
    }
    for (int l = 1; l < A.nlev; ++l) {   // fp32 copies of the large coarse levels' values for the cycle's SpMVs
      AmgLevelDev& L = A.lv[l];
      if (!L.cur32) continue;
      k_amg_to_float<<<amg_grid(ctx, L.nentries, 256), 256, 0, ctx->stream>>>((size_t)L.nentries, L.cur_val, const_cast<float*>(L.cur32));
      ctx->stats.kernel_launches++;
    }
    const AmgLevelDev& L = A.lv[A.nlev - 1];
    double* dense = const_cast<double*>(A.cur_dense);
    if (L.n <= kAmgDenseShared)
      k_amg_dense_inverse<true><<<1, 1024, (size_t)L.n * L.n * sizeof(double), ctx->stream>>>(L.n, L.rowptr, L.slice_ptr, L.col, L.cur_val, dense);
    else
      k_amg_dense_inverse<false><<<1, 1024, 0, ctx->stream>>>(L.n, L.rowptr, L.slice_ptr, L.col, L.cur_val, dense);
    ctx->stats.kernel_launches++;
    if (S) {
      DDPS_CUDA(cudaMemcpyAsync(S->snap, val, (size_t)A.lv[0].nentries * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      S->valid = true;
    }
    A.numerics++;
  } else {
    A.reuses++;
  }
  ctx->stats.amg_reuses = (int32_t)A.reuses;
  // the finest level's fp32 copy always follows the current operator
  AmgLevelDev& L0 = A.lv[0];
  L0.cur32 = L0.val32;
  if (L0.val32) {
    k_amg_to_float<<<amg_grid(ctx, L0.nentries, 256), 256, 0, ctx->stream>>>((size_t)L0.nentries, val, L0.val32);
    ctx->stats.kernel_launches++;
  }
  DDPS_CUDA(cudaGetLastError());
  return 0;
}


template <int MODE, bool DOT>
static void launch_amg_spmv(Context* ctx, const AmgLevelDev& L, const double* x, const double* b, double om, double* y) {
  if (!DOT && L.n <= kAmgSmallLevel && !L.cur32) {
    (void)(launch_pdl(k_amg_spmv_wpr<MODE>, amg_grid(ctx, (int64_t)L.n * 32, 256), 256, 0, ctx->stream, L.n, L.rowptr, L.slice_ptr, L.col, L.cur_val, x, b,
                                                                                        L.cur_dinv, om, y, ctx->scal));
    ctx->stats.kernel_launches++;
    ctx->stats.spmv_total++;
    return;
  }
  const int grid = amg_grid(ctx, (int64_t)L.nslices * 32, kVecBlock);
  if (L.cur32)
    (void)(launch_pdl(k_amg_spmv<MODE, DOT, float>, grid, kVecBlock, 0, ctx->stream, L.n, L.nslices, L.slice_ptr, L.col, L.cur32, x, b, L.cur_dinv,
                                                                     om, y, ctx->partials, ctx->scal));
  else
    (void)(launch_pdl(k_amg_spmv<MODE, DOT, double>, grid, kVecBlock, 0, ctx->stream, L.n, L.nslices, L.slice_ptr, L.col, L.cur_val, x, b, L.cur_dinv,
                                                            om, y, ctx->partials, ctx->scal));
  ctx->stats.kernel_launches++;
  ctx->stats.spmv_total++;
}

// z = B r: one V(1,1) cycle from the zero initial guess.  dot_rz: also leave r.z in scal->sums[1].
// Distributed levels (multi-rank contexts): the level's x, t, x2 carry ghost entries that are refreshed right before the
// kernel that gathers them -- x before the residual and before the smoothing sweep, t before the restriction, the coarse
// correction before the prolongation (amg_dist_prototype.py messages 5-8); on the transition level restriction sums over
// the owned fine rows only and one allreduce completes the replicated right-hand side.
static int amg_vcycle_impl(Context* ctx, const double* r, double* z, bool dot_rz, bool have_first_sweep) {
  Amg& A = *ctx->amg;
  const double om = A.omega;
  cudaStream_t st = ctx->stream;
  int rc;
  A.lv[0].b = const_cast<double*>(r);
  A.lv[0].x2 = z;
  const int lt = (A.tail_from > 0 && A.tail_from < A.nlev) ? A.tail_from : A.nlev;   // first level of the fused tail
  const int ltop = std::min(A.nlev - 1, lt);                                          // per-level launches for l < ltop
  for (int l = 0; l < ltop; ++l) {
    AmgLevelDev& L = A.lv[l];
    if (l == 0 && !have_first_sweep) {   // (inside the CG the update kernel already produced the first sweep of the finest
      DDPS_CUDA(launch_pdl(k_amg_jac0, amg_grid(ctx, L.n, 256), 256, 0, st, L.n, om, L.cur_dinv, L.b, L.x, ctx->scal));   // level; coarser
      ctx->stats.kernel_launches++;                                                                    // levels: fused in restrict)
    }
    if (L.dist && (rc = halo_xchg(ctx, *L.plan, L.x, true))) return rc;
    if (L.dist) prof_mark(ctx, l == 0 ? "L0 halo(x)" : "L>0 halo(x)");
    launch_amg_spmv<1, false>(ctx, L, L.x, L.b, om, L.t);
    prof_mark(ctx, l == 0 ? "L0 residual spmv" : l == 1 ? "L1 residual spmv" : "L>1 residual spmv");
    if (L.dist == 1 && (rc = halo_xchg(ctx, *L.plan, L.t, true))) return rc;
    if (L.dist == 1) prof_mark(ctx, l == 0 ? "L0 halo(t)" : "L>0 halo(t)");
    AmgLevelDev& C = A.lv[l + 1];
    const bool coarsest = (l + 2 == A.nlev);
    const bool fuse_sweep = !coarsest && L.dist != 2;   // transition: the sweep needs the COMPLETED right-hand side
    const bool staged_sum = L.dist == 2 && A.tail_plan;   // partial sums into my slot of the exchange buffer
    double* bc = staged_sum ? A.tail_parts : C.b;
    if (L.nc <= kAmgSmallLevel)
      DDPS_CUDA(launch_pdl(k_amg_restrict_wpr, amg_grid(ctx, (int64_t)L.nc * 32, 256), 256, 0, st, L.nc, L.Rptr, L.Rent, L.t, bc, fuse_sweep ? C.cur_dinv : nullptr,
                                                                                om, C.x, ctx->scal));
    else
      DDPS_CUDA(launch_pdl(k_amg_restrict, amg_grid(ctx, L.nc, 256), 256, 0, st, L.nc, L.Rptr, L.Rent, L.t, bc, fuse_sweep ? C.cur_dinv : nullptr, om, C.x,
                                                              ctx->scal));
    ctx->stats.kernel_launches++;
    prof_mark(ctx, l == 0 ? "L0 restrict" : l == 1 ? "L1 restrict" : "L>1 restrict");
    if (staged_sum) {
      if ((rc = halo_xchg(ctx, *A.tail_plan, A.tail_parts, true))) return rc;
      DDPS_CUDA(launch_pdl(k_amg_tail_sum, amg_grid(ctx, C.n, 256), 256, 0, st, C.n, ctx->nranks, ctx->rank, A.tail_parts, C.b,
                           coarsest ? nullptr : C.cur_dinv, om, C.x, ctx->scal));
      ctx->stats.kernel_launches++;
      prof_mark(ctx, "tail exchange + sum + sweep");
    } else if (L.dist == 2) {
      if ((rc = comm_allreduce(ctx, C.b, (size_t)C.n, COMM_SUM))) return rc;
      if (!coarsest) {
        DDPS_CUDA(launch_pdl(k_amg_jac0, amg_grid(ctx, C.n, 256), 256, 0, st, C.n, om, C.cur_dinv, C.b, C.x, ctx->scal));
        ctx->stats.kernel_launches++;
      }
      prof_mark(ctx, "tail allreduce + sweep");
    }
  }
  if (lt < A.nlev) {
    TailArgs T;
    T.nl = A.nlev - lt;
    for (int l = lt; l < A.nlev; ++l) {
      const AmgLevelDev& L = A.lv[l];
      TailLevel& t = T.lv[l - lt];
      t.n = L.n; t.nc = L.nc; t.nslices = L.nslices; t.wpr = L.n <= kAmgTailWpr;
      t.slice_ptr = L.slice_ptr; t.col = L.col; t.rowptr = L.rowptr; t.val = L.cur_val; t.dinv = L.cur_dinv;
      t.Pptr = L.Pptr; t.Rptr = L.Rptr; t.Pent = L.Pent; t.Rent = L.Rent;
      t.b = L.b; t.x = L.x; t.x2 = L.x2; t.t = L.t;
    }
    T.dense = A.cur_dense; T.om = om; T.bar = A.tail_bar; T.scal = ctx->scal;
    void* args[] = {&T};
    DDPS_CUDA(cudaLaunchCooperativeKernel((const void*)k_amg_tail, dim3(A.tail_grid), dim3(256), args, 0, st));
    ctx->stats.kernel_launches++;
    prof_mark(ctx, "fused tail of the cycle");
  } else {
    AmgLevelDev& L = A.lv[A.nlev - 1];
    k_amg_dense_apply<<<1, 1024, 0, st>>>(L.n, A.cur_dense, L.b, L.x2, ctx->scal);
    ctx->stats.kernel_launches++;
    prof_mark(ctx, "coarsest solve");
  }
  for (int l = ltop - 1; l >= 0; --l) {
    AmgLevelDev& L = A.lv[l];
    AmgLevelDev& C = A.lv[l + 1];
    if (L.dist == 1 && (rc = halo_xchg(ctx, *C.plan, C.x2, true))) return rc;
    if (L.dist == 1) prof_mark(ctx, l == 0 ? "L0 halo(xc)" : "L>0 halo(xc)");
    // distributed levels: the ghost rows are prolonged too (their prolongator rows came from their owners at setup, their
    // pre-smoothed x from the exchange on the way down): same numbers in the same order as on the owner, so x needs no
    // exchange before the smoothing sweep
    const int np = L.dist ? L.n_loc : L.n;
    DDPS_CUDA(launch_pdl(k_amg_prolong, amg_grid(ctx, np, 256), 256, 0, st, np, L.Pptr, L.Pent, C.x2, L.x, ctx->scal));
    ctx->stats.kernel_launches++;
    prof_mark(ctx, l == 0 ? "L0 prolong" : l == 1 ? "L1 prolong" : "L>1 prolong");
    if (l == 0 && dot_rz) launch_amg_spmv<2, true>(ctx, L, L.x, L.b, om, L.x2);
    else launch_amg_spmv<2, false>(ctx, L, L.x, L.b, om, L.x2);
    prof_mark(ctx, l == 0 ? "L0 smoothing spmv" : l == 1 ? "L1 smoothing spmv" : "L>1 smoothing spmv");
  }
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

int amg_vcycle(Context* ctx, const double* r, double* z, bool dot_rz) { return amg_vcycle_impl(ctx, r, z, dot_rz, false); }

// Single kernels of the multigrid-preconditioned CG iteration, launched alone for ddps_time_kernel (ids 9..13), with the
// ALGORITHMIC bytes of one launch (every input array read once, every output written once; the cycle's matrix values in
// the width it stores them: fp32 on the large levels):
//   9  finest-level residual t = b - A x          (k_amg_spmv<1,false,VT>)
//   10 finest-level smoothing sweep + r.z         (k_amg_spmv<2,true,VT>)
//   11 CG update x,r + first sweep                (k_apcg_update)
//   12 restriction 0 -> 1 + coarse first sweep    (k_amg_restrict)
//   13 prolongation 1 -> 0                        (k_amg_prolong)
int amg_time_launch(Context* ctx, int which, int it, double* bytes) {
  Amg& A = *ctx->amg;
  AmgLevelDev& L = A.lv[0];
  AmgLevelDev& C = A.lv[1];
  const double n = L.n, nnz = (double)L.nnz, nc = L.nc, vw = L.cur32 ? 4.0 : 8.0;
  L.b = ctx->r; L.x2 = A.z;
  double nnzP = 0.0;
  if (bytes) {
    int last = 0;
    cudaMemcpy(&last, L.Pptr + L.n, sizeof(int), cudaMemcpyDeviceToHost);
    nnzP = last;
  }
  switch (which) {
    case 9:
      launch_amg_spmv<1, false>(ctx, L, L.x, L.b, A.omega, L.t);
      if (bytes) *bytes = (vw + 4.0) * nnz + 4.0 * (n + 1) + 3 * 8.0 * n;
      return 0;
    case 10:
      launch_amg_spmv<2, true>(ctx, L, L.x, L.b, A.omega, L.x2);
      if (bytes) *bytes = (vw + 4.0) * nnz + 4.0 * (n + 1) + 4 * 8.0 * n;
      return 0;
    case 11:
      DDPS_CUDA(launch_pdl(k_apcg_update<true>, amg_grid(ctx, L.n, kVecBlock), kVecBlock, 0, ctx->stream, L.n, it & 1, it, 1 << 30, ctx->p, ctx->q, ctx->x, ctx->r, L.cur_dinv,
                                                                                A.omega, L.x, ctx->partials, ctx->scal));
      if (bytes) *bytes = 8 * 8.0 * n;
      return 0;
    case 12:
      DDPS_CUDA(launch_pdl(k_amg_restrict, amg_grid(ctx, L.nc, 256), 256, 0, ctx->stream, L.nc, L.Rptr, L.Rent, L.t, C.b, C.cur_dinv, A.omega, C.x, ctx->scal));
      if (bytes) *bytes = 8.0 * nnzP + 4.0 * (nc + 1) + 8.0 * n + 3 * 8.0 * nc;
      return 0;
    case 13:
      DDPS_CUDA(launch_pdl(k_amg_prolong, amg_grid(ctx, L.n, 256), 256, 0, ctx->stream, L.n, L.Pptr, L.Pent, C.x2, L.x, ctx->scal));
      if (bytes) *bytes = 8.0 * nnzP + 4.0 * (n + 1) + 8.0 * nc + 2 * 8.0 * n;
      return 0;
  }
  return DDPS_ERR_ARG;
}

// algorithmic bytes of one whole multigrid-preconditioned CG iteration: the CG's own SpMV + vector kernels and, per level,
// the two cycle SpMVs and the two transfers (same counting rules as above)
double amg_iteration_bytes(Context* ctx) {
  Amg& A = *ctx->amg;
  const double n0 = A.lv[0].n;
  double by = 12.0 * (double)A.lv[0].nnz + 4.0 * (n0 + 1) + 16.0 * n0;   // k_spmv<true>
  by += 8 * 8.0 * n0 + 3 * 8.0 * n0;                                     // k_apcg_update, k_apcg_direction
  for (int l = 0; l + 1 < A.nlev; ++l) {
    const AmgLevelDev& L = A.lv[l];
    const double n = L.n, nnz = (double)L.nnz, nc = L.nc, vw = L.cur32 ? 4.0 : 8.0;
    int last = 0;
    cudaMemcpy(&last, L.Pptr + L.n, sizeof(int), cudaMemcpyDeviceToHost);
    by += 2 * ((vw + 4.0) * nnz + 4.0 * (n + 1)) + 7 * 8.0 * n;           // residual + smoothing sweep
    by += 2 * 8.0 * (double)last + 4.0 * (nc + 1) + 4.0 * (n + 1) + 8.0 * n + 4 * 8.0 * nc + 2 * 8.0 * n;   // restrict + prolong
  }
  const double nL = A.lv[A.nlev - 1].n;
  by += 8.0 * nL * nL;
  return by;
}

// ------------------------------------------------------------------------------------------------------------
// CG preconditioned with the V-cycle
// ------------------------------------------------------------------------------------------------------------
int amg_pcg_iteration_launch(Context* ctx, double* x, int it, int max_it) {
  Amg& A = *ctx->amg;
  const int n = ctx->n_own;
  const int par = it & 1;
  const bool mr = ctx->nranks > 1;
  int rc;
  const int g2 = amg_grid(ctx, n, kVecBlock);
  if (it == 3) prof_begin(ctx);
  if (mr && (rc = halo_xchg(ctx, ctx->halo, ctx->p, true))) return rc;              // ghost entries of the search direction
  if (mr) prof_mark(ctx, "halo(p)");
  if ((rc = launch_spmv_dot(ctx, A.lv[0].cur_val, ctx->p, ctx->q))) return rc;    // q = A p, p.q -> sums[0]
  prof_mark(ctx, "CG spmv + p.Ap");
  if (mr) {
    if ((rc = scalar_allreduce(ctx, &ctx->scal->sums[0], 1, 0, true))) return rc;
    prof_mark(ctx, "allreduce p.Ap");
    DDPS_CUDA(launch_pdl(k_apcg_update<false>, g2, kVecBlock, 0, ctx->stream, n, par, it, max_it, ctx->p, ctx->q, x, ctx->r, A.lv[0].cur_dinv, A.omega,
                                                            A.lv[0].x, ctx->partials, ctx->scal));
    if ((rc = scalar_allreduce(ctx, &ctx->scal->sums[2], 1, 1, true))) return rc;   // {r.r, max|r|}
    k_apcg_decide<<<1, 1, 0, ctx->stream>>>(ctx->scal, par, it, max_it);
    ctx->stats.kernel_launches++;
    prof_mark(ctx, "update + allreduce + decide");
  } else {
    DDPS_CUDA(launch_pdl(k_apcg_update<true>, g2, kVecBlock, 0, ctx->stream, n, par, it, max_it, ctx->p, ctx->q, x, ctx->r, A.lv[0].cur_dinv, A.omega,
                                                           A.lv[0].x, ctx->partials, ctx->scal));
    prof_mark(ctx, "update");
  }
  if ((rc = amg_vcycle_impl(ctx, ctx->r, A.z, true, true))) return rc;             // z = B r, r.z -> sums[1]
  if (mr && (rc = scalar_allreduce(ctx, &ctx->scal->sums[1], 1, 0, true))) return rc;
  DDPS_CUDA(launch_pdl(k_apcg_direction, g2, kVecBlock, 0, ctx->stream, n, par, 0, 0.0, A.z, ctx->p, ctx->scal));
  prof_mark(ctx, mr ? "allreduce r.z + direction" : "direction");
  ctx->stats.kernel_launches += 2;
  if (it == 3) prof_end(ctx);
  return 0;
}

int amg_pcg_solve(Context* ctx, const double* val, const double* dinv, const double* rhs, double* x, bool x0_nonzero,
                  double rtol, double atol_inf, int64_t max_it, int64_t* iters_out, double* relres_out) {
  Amg& A = *ctx->amg;
  const int n = ctx->n_own;
  const bool mr = ctx->nranks > 1;
  int rc;
  if (max_it <= 0) max_it = 200000;
  if (max_it > 2000000000LL) max_it = 2000000000LL;
  cudaStream_t st = ctx->stream;
  if (ctx->p2p.win_local) DDPS_CUDA(cudaMemsetAsync(&ctx->p2p.win_local->error, 0, sizeof(int), st));
  if ((rc = amg_numeric(ctx, val, dinv, ctx->amg_kind))) return rc;
  const int g = amg_grid(ctx, std::max(n, 1), kVecBlock);
  if (x0_nonzero) {
    DDPS_CUDA(cudaMemcpyAsync(ctx->p, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (mr && (rc = halo_xchg(ctx, ctx->halo, ctx->p, false))) return rc;
    if ((rc = launch_spmv(ctx, val, ctx->p, ctx->q))) return rc;
    k_apcg_init<true><<<g, kVecBlock, 0, st>>>(n, rhs, ctx->q, ctx->p, ctx->r, ctx->partials, ctx->scal);
  } else {
    DDPS_CUDA(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), st));
    k_apcg_init<false><<<g, kVecBlock, 0, st>>>(n, rhs, nullptr, nullptr, ctx->r, ctx->partials, ctx->scal);
  }
  if (mr && (rc = scalar_allreduce(ctx, &ctx->scal->sums[0], 3, 1, false))) return rc;   // {r.r, b.b, x0.b, max|r|}
  k_apcg_start<<<1, 1, 0, st>>>(ctx->scal, rtol, atol_inf, x0_nonzero ? 1 : 0, ctx->eps_e);
  if (x0_nonzero) { k_apcg_cold<<<g, kVecBlock, 0, st>>>(n, rhs, x, ctx->r, ctx->scal); ctx->stats.kernel_launches++; }
  if ((rc = amg_vcycle(ctx, ctx->r, A.z, true))) return rc;
  if (mr && (rc = scalar_allreduce(ctx, &ctx->scal->sums[1], 1, 0, true))) return rc;
  DDPS_CUDA(launch_pdl(k_apcg_direction, g, kVecBlock, 0, st, n, 0, 1, rtol, A.z, ctx->p, ctx->scal));
  ctx->stats.kernel_launches += 3;
  DDPS_CUDA(cudaGetLastError());

  int64_t it = 0;
  const int chunk = 4;   // iterations enqueued between two polls of the device-side done flag
  while (true) {
    DDPS_CUDA(cudaMemcpyAsync(ctx->scal_host, ctx->scal, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    if (ctx->scal_host->done || it >= max_it) break;
    const int nrun = (int)std::min<int64_t>(chunk, max_it - it);
    for (int j = 0; j < nrun; ++j, ++it)
      if ((rc = amg_pcg_iteration_launch(ctx, x, (int)it, (int)max_it))) return rc;
    DDPS_CUDA(cudaGetLastError());
  }
  const PcgScalars& S = *ctx->scal_host;
  if (iters_out) *iters_out = S.iters;
  if (relres_out) *relres_out = (S.bb > 0) ? sqrt(S.rr / S.bb) : 0.0;
  ctx->stats.pcg_total += S.iters;
  ctx->hist_pcg.push_back((double)S.iters);
  if (S.breakdown == 2) {
    ctx->fail(DDPS_ERR_COMM, "peer-memory exchange timed out (a rank stopped participating)");
    return DDPS_ERR_COMM;
  }
  if (S.breakdown) {
    ctx->fail(DDPS_ERR_BREAKDOWN, "PCG breakdown: p.Ap <= 0 (operator not SPD; u or v negative? SURVEY H3)");
    return DDPS_ERR_BREAKDOWN;
  }
  if (!S.conv && !ctx->opts.lin_cap_ok) {   // residual rule AND (warm starts) the energy rule: a cap hit on either is reported
    ctx->fail(DDPS_ERR_NOCONV, "AMG-PCG hit the iteration cap (" + std::to_string(S.iters) + " its, relres " +
                                   std::to_string(S.bb > 0 ? sqrt(S.rr / S.bb) : 0.0) + ")");
    return DDPS_ERR_NOCONV;
  }
  return 0;
}

}  // namespace ddps
