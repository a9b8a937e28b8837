// Hand-written sm_100a FP64 kernels of the hot path (compiled with -fmad=false so the element
// formulas round exactly like the reference's IEEE expressions).
//
//   (the fused assembly kernels and the nodal prep kernels live in assemble.cu)
//   k_spmv_dot                        SELL-32 SpMV fused with the p.Ap partial dot (K8/K9)
//   k_pcg_update / k_pcg_direction    fused vector updates + dots of Jacobi-PCG (K9)
#include <cuda_pipeline.h>

#include "ddps_internal.cuh"
#include "device_utils.cuh"

namespace ddps {

// ---- peer-memory collectives (multi-GPU): flags + payload live in every rank's P2PWin -----------------
// Writer: payload stores, __threadfence_system, flag store.  Reader: spin on the local flag (peers write
// it over NVLink), fence, read payload.  Tags increase monotonically, one per use, so a slot is never
// ambiguous; the data dependencies of CG (every rank needs every other rank's previous partial before it
// can produce the next one) make single buffering safe.  A spin gives up after ~10 s and raises an error
// flag instead of hanging the GPU.
// `err` is the window's sticky error flag: once any wait of this solve has timed out, later waits give up at once
// (a dead rank then costs one timeout, not one per kernel) and the solve ends with DDPS_ERR_COMM.
__device__ __forceinline__ bool spin_until(const volatile unsigned long long* flag, unsigned long long tag, volatile int* err) {
  const long long t0 = clock64();
  while (*flag < tag) {
    if (*err) return false;
    if (clock64() - t0 > 20000000000LL) { *err = 1; return false; }
  }
  return true;
}

// thread 0 of the calling (last) block publishes NV doubles to every rank's window
template <int NV>
__device__ __forceinline__ void p2p_publish(const P2PDev& pd, int phase, unsigned long long tag, const double (&v)[NV]) {
  for (int r = 0; r < pd.nranks; ++r) {
    P2PWin* W = pd.win[r];
    volatile double* dst = phase == 0 ? W->redA[pd.rank] : W->redB[pd.rank];
#pragma unroll
    for (int i = 0; i < NV; ++i) dst[i] = v[i];
  }
  __threadfence_system();
  for (int r = 0; r < pd.nranks; ++r) {
    P2PWin* W = pd.win[r];
    volatile unsigned long long* f = phase == 0 ? &W->redA_flag[pd.rank] : &W->redB_flag[pd.rank];
    *f = tag;
  }
}

// every block: thread 0 waits for all ranks' partials and combines them in RANK ORDER (deterministic and
// identical on every rank); result broadcast through shared memory.  Returns false on timeout.
template <int NV, unsigned MAXMASK>
__device__ __forceinline__ bool p2p_gather(const P2PDev& pd, int phase, unsigned long long tag, double (&out)[NV],
                                           double* sh) {
  __shared__ int s_ok;
  if (threadIdx.x < 32) {   // lane r waits for rank r (all flags polled concurrently), lane 0 combines in rank order
    P2PWin* W = pd.win[pd.rank];
    const int r = threadIdx.x;
    bool ok = true;
    double mine[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) mine[i] = 0.0;
    if (r < pd.nranks) {
      ok = spin_until(phase == 0 ? &W->redA_flag[r] : &W->redB_flag[r], tag, &W->error);
      __threadfence();
      const volatile double* src = phase == 0 ? W->redA[r] : W->redB[r];
#pragma unroll
      for (int i = 0; i < NV; ++i) mine[i] = src[i];
    }
    ok = __all_sync(0xffffffffu, ok);
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    for (int q = 0; q < pd.nranks; ++q) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const double t = __shfl_sync(0xffffffffu, mine[i], q);
        acc[i] = ((MAXMASK >> i) & 1) ? fmax(acc[i], t) : acc[i] + t;
      }
    }
    if (r == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) sh[i] = acc[i];
      s_ok = ok ? 1 : 0;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) out[i] = sh[i];
  const bool ok = s_ok != 0;
  __syncthreads();
  return ok;
}

// push the owned boundary values of the search direction straight into the neighbours' ghost blocks
__global__ void __launch_bounds__(1024) k_halo_push(const PushArgs A, const int* __restrict__ send_idx,
                                                     const double* __restrict__ p, const PcgScalars* scal) {
  if (scal->done) return;
  const int j = blockIdx.x;
  const int n = A.off[j + 1] - A.off[j];
  double* dst = A.dst[j];
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = p[send_idx[A.off[j] + i]];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *(volatile unsigned long long*)A.flag[j] = A.tag;
}


// ------------------------------------------------------------------------------------------------
// SpMV (SELL-32) with fused dot, persistent grid-stride over slices
// ------------------------------------------------------------------------------------------------
// gather of x: the read-only path for columns this rank owns.  MR (fused peer-memory halo): the ghost block of x is
// written by the peers over NVLink while this kernel may already be resident (it waits on the halo flags itself), so
// ghost columns (c >= nrows) must not go through the non-coherent path: ld.global.cg reads them from L2.
template <bool MR>
__device__ __forceinline__ double ld_x(const double* __restrict__ x, int c, int nrows) {
  if (MR && c >= nrows) return __ldcg(x + c);
  return __ldg(x + c);
}

template <bool DOT, bool MR>
__global__ void __launch_bounds__(kVecBlock, 8) k_spmv(int nrows, int nslices, const int* __restrict__ slice_ptr,
                                                    const int* __restrict__ col, const double* __restrict__ val,
                                                    const double* __restrict__ x, double* __restrict__ y,
                                                    double* partials, PcgScalars* scal, const P2PDev pd, const int rev) {
  __shared__ double sh_red[32];
  pdl_enter();
  if (DOT && scal->done) return;
  if (MR) {   // the neighbours' halo pushes for this iteration must have landed
    if (threadIdx.x < pd.nneigh) {
      P2PWin* W = pd.win[pd.rank];
      if (!spin_until(&W->halo_flag[pd.neigh[threadIdx.x]], pd.tag_halo, &W->error)) scal->breakdown = 2;   // -> DDPS_ERR_COMM
      __threadfence();
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  double dot = 0.0;
  // `rev` alternates the traversal direction between consecutive PCG kernels: the vector tail the previous
  // kernel wrote last is still in the 126 MB L2 and is consumed first.
  for (int s0 = gw; s0 < nslices; s0 += nw) {
    const int slice = rev ? nslices - 1 - s0 : s0;
    const int base = slice_ptr[slice];
    const int w = (slice_ptr[slice + 1] - base) >> 5;
    int idx = base + lane;
    double sum = 0.0;
    int k = 0;
    // groups of 4 entries: 8 independent coalesced loads in flight per warp before the 4 gathers; measured
    // faster than whole-row unrolling (which halves the occupancy) and than 16-bit column deltas (no gain:
    // the kernel already keeps HBM ~72% busy = the practical read limit)
    for (; k + 4 <= w; k += 4) {
      const int c0 = ld_stream(col + idx), c1 = ld_stream(col + idx + 32), c2 = ld_stream(col + idx + 64),
                c3 = ld_stream(col + idx + 96);
      const double v0 = ld_stream(val + idx), v1 = ld_stream(val + idx + 32), v2 = ld_stream(val + idx + 64),
                   v3 = ld_stream(val + idx + 96);
      const double x0 = ld_x<MR>(x, c0, nrows), x1 = ld_x<MR>(x, c1, nrows), x2 = ld_x<MR>(x, c2, nrows), x3 = ld_x<MR>(x, c3, nrows);
      sum += v0 * x0; sum += v1 * x1; sum += v2 * x2; sum += v3 * x3;
      idx += 128;
    }
    for (; k < w; ++k) {
      const int c0 = ld_stream(col + idx);
      const double v0 = ld_stream(val + idx);
      sum += v0 * ld_x<MR>(x, c0, nrows);
      idx += 32;
    }
    const int row = slice * 32 + lane;
    if (row < nrows) {
      y[row] = sum;
      if (DOT) dot += __ldg(x + row) * sum;
    }
  }
  if (DOT) {
    double v[1] = {dot}, tot[1];
    if (grid_reduce<1, false>(v, partials, &scal->cnt[0], sh_red, tot)) {
      if (threadIdx.x == 0) {
        if (MR) p2p_publish<1>(pd, 0, pd.tagA, tot);
        else scal->sums[0] = tot[0];
      }
    }
  }
}

static inline int vec_grid(Context* ctx, int64_t work_items, int block) {
  int64_t need = (work_items + block - 1) / block;
  int64_t cap = (int64_t)ctx->num_sms * (2048 / block);
  return (int)std::max<int64_t>(1, std::min(need, cap));
}

int launch_spmv(Context* ctx, const double* val, const double* x, double* y) {
  const Pattern& P = ctx->pat;
  if (!P.nslices) return 0;
  int grid = vec_grid(ctx, (int64_t)P.nslices * 32, kVecBlock);
  P2PDev none; memset(&none, 0, sizeof(none)); none.nranks = 1;
  DDPS_CUDA(launch_pdl(k_spmv<false, false>, grid, kVecBlock, 0, ctx->stream, P.nrows, P.nslices, P.slice_ptr, P.col, val, x, y, nullptr, nullptr,
                       none, 0));
  ctx->stats.kernel_launches++;
  ctx->stats.spmv_total++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

int launch_spmv_dot(Context* ctx, const double* val, const double* x, double* y) {
  const Pattern& P = ctx->pat;
  if (!P.nslices) return 0;
  int grid = vec_grid(ctx, (int64_t)P.nslices * 32, kVecBlock);
  P2PDev none; memset(&none, 0, sizeof(none)); none.nranks = 1;
  DDPS_CUDA(launch_pdl(k_spmv<true, false>, grid, kVecBlock, 0, ctx->stream, P.nrows, P.nslices, P.slice_ptr, P.col, val, x, y, ctx->partials,
                       ctx->scal, none, 0));
  ctx->stats.kernel_launches++;
  ctx->stats.spmv_total++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Vector kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_axpy(int n, double a, const double* __restrict__ x, double* __restrict__ y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] += a * x[i];
}
__global__ void k_scale_copy(int n, double s, const double* __restrict__ x, double* __restrict__ y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = s * x[i];
}

int launch_axpy(Context* ctx, double a, const double* x, double* y, int n) {
  if (!n) return 0;
  k_axpy<<<vec_grid(ctx, n, 256), 256, 0, ctx->stream>>>(n, a, x, y);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}
int launch_negate_copy(Context* ctx, const double* x, double* y, double s, int n) {
  if (!n) return 0;
  k_scale_copy<<<vec_grid(ctx, n, 256), 256, 0, ctx->stream>>>(n, s, x, y);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

__global__ void k_max_abs_diff3(int n, const double* __restrict__ a0, const double* __restrict__ a1,
                                const double* __restrict__ b0, const double* __restrict__ b1,
                                const double* __restrict__ c0, const double* __restrict__ c1, double* partials,
                                PcgScalars* scal) {
  __shared__ double sh_red[32];
  double m = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    m = fmax(m, fabs(a0[i] - a1[i]));
    if (b0) m = fmax(m, fabs(b0[i] - b1[i]));
    if (c0) m = fmax(m, fabs(c0[i] - c1[i]));
  }
  double v[1] = {m}, tot[1];
  if (grid_reduce<1, true>(v, partials, &scal->cnt[3], sh_red, tot)) {
    if (threadIdx.x == 0) scal->norm_inf = tot[0];
  }
}

__global__ void k_max_abs(int n, const double* __restrict__ a, double* partials, PcgScalars* scal) {
  __shared__ double sh_red[32];
  double m = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmax(m, fabs(a[i]));
  double v[1] = {m}, tot[1];
  if (grid_reduce<1, true>(v, partials, &scal->cnt[3], sh_red, tot)) {
    if (threadIdx.x == 0) scal->norm_inf = tot[0];
  }
}

static int fetch_norm_inf(Context* ctx, double* host_out) {
  int rc = allreduce_max(ctx, &ctx->scal->norm_inf, 1);
  if (rc) return rc;
  DDPS_CUDA(cudaMemcpyAsync(&ctx->scal_host->norm_inf, &ctx->scal->norm_inf, sizeof(double), cudaMemcpyDeviceToHost,
                            ctx->stream));
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  *host_out = ctx->scal_host->norm_inf;
  return 0;
}

int launch_max_abs_diff3(Context* ctx, const double* a0, const double* a1, const double* b0, const double* b1,
                         const double* c0, const double* c1, int n, double* host_out) {
  k_max_abs_diff3<<<vec_grid(ctx, std::max(n, 1), 256), 256, 0, ctx->stream>>>(n, a0, a1, b0, b1, c0, c1, ctx->partials,
                                                                              ctx->scal);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return fetch_norm_inf(ctx, host_out);
}

int launch_max_abs(Context* ctx, const double* a, int n, double* host_out) {
  k_max_abs<<<vec_grid(ctx, std::max(n, 1), 256), 256, 0, ctx->stream>>>(n, a, ctx->partials, ctx->scal);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return fetch_norm_inf(ctx, host_out);
}

int fetch_resid_norm(Context* ctx, double* host_out) { return launch_max_abs(ctx, ctx->resid, ctx->n_own, host_out); }

// ------------------------------------------------------------------------------------------------
// Jacobi-preconditioned CG (K9), all control flow on the device
// ------------------------------------------------------------------------------------------------
// r = rhs - q (q = A*x0) or r = rhs;  p = dinv*r;  local sums {r.z, r.r, rhs.rhs}
// (x0 = the initial guess when HAVE_Q: x0.rhs ~ ||x||_A^2 is the reference energy of the error-based stop rule)
template <bool HAVE_Q>
__global__ void k_pcg_init(int n, const double* __restrict__ rhs, const double* __restrict__ q, const double* __restrict__ x0,
                           const double* __restrict__ dinv, double* __restrict__ r, double* __restrict__ p,
                           double* partials, PcgScalars* scal) {
  __shared__ double sh_red[32];
  double s_rz = 0.0, s_rr = 0.0, s_bb = 0.0, s_xb = 0.0, s_max = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double bi = rhs[i];
    const double ri = HAVE_Q ? bi - q[i] : bi;
    const double zi = dinv ? dinv[i] * ri : ri;   // dinv == nullptr: symmetrically pre-scaled system, M = I
    r[i] = ri;
    p[i] = zi;
    s_rz += ri * zi; s_rr += ri * ri; s_bb += bi * bi;
    if (HAVE_Q) s_xb += x0[i] * bi;
    s_max = fmax(s_max, fabs(ri));
  }
  double v[4] = {s_rz, s_rr, s_bb, s_xb}, tot[4];
  const bool last = grid_reduce<4, false>(v, partials, &scal->cnt[1], sh_red, tot);
  if (last && threadIdx.x == 0) { scal->sums[0] = tot[0]; scal->sums[1] = tot[1]; scal->sums[2] = tot[2]; scal->sums[3] = tot[3]; }
  double vm[1] = {s_max}, totm[1];
  if (grid_reduce<1, true>(vm, partials + 4 * gridDim.x, &scal->cnt[2], sh_red, totm)) {
    if (threadIdx.x == 0) scal->rmax = totm[0];
  }
}

// stop rule: ||r||_2 <= rtol*||rhs||_2  OR  ||r||_inf <= atol_inf
__global__ void k_pcg_start(PcgScalars* scal, double rtol, double atol_inf, double eps_e) {
  scal->rz[0] = scal->sums[0];
  scal->rr = scal->sums[1];
  scal->bb = scal->sums[2];
  scal->xb = scal->sums[3];
  scal->thresh2 = rtol * rtol * scal->sums[2];
  scal->thresh_inf = atol_inf;
  scal->eps_e2 = eps_e * eps_e;
  scal->hs_total = 0.0;
  scal->hs_d = kHsDelay;
  scal->iters = 0;
  scal->breakdown = 0;
  if (eps_e > 0.0) scal->done = (scal->sums[1] == 0.0) ? 1 : 0;   // error rule armed: only a zero residual ends the solve here
  else scal->done = (scal->sums[1] <= scal->thresh2 || scal->rmax <= atol_inf) ? 1 : 0;
  scal->conv = scal->done;
}

// x += alpha p; r -= alpha q; partial sums {r.D^-1 r, r.r, max|r|}
__global__ void __launch_bounds__(kVecBlock, 8) k_pcg_update(int n, int par, int rev, const double* __restrict__ p,
                                                          const double* __restrict__ q,
                                                          const double* __restrict__ dinv, double* __restrict__ x,
                                                          double* __restrict__ r, double* partials, PcgScalars* scal,
                                                          const P2PDev pd) {
  __shared__ double sh_red[32];
  if (scal->done) return;
  double pq;
  if (pd.nranks > 1) {
    double t[1];
    const bool ok = p2p_gather<1, 0u>(pd, 0, pd.tagA, t, sh_red);
    pq = t[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) scal->sums[0] = pq;
    if (!ok) {   // a peer's partial never arrived: alpha would be garbage -> stop, DDPS_ERR_COMM (raised by k_pcg_direction)
      if (blockIdx.x == 0 && threadIdx.x == 0) scal->breakdown = 2;
      return;
    }
  } else {
    pq = scal->sums[0];
  }
  const double rz = scal->rz[par];
  if (!(pq > 0.0) || !(rz == rz)) {   // operator not SPD or NaN: stop
    if (blockIdx.x == 0 && threadIdx.x == 0) { scal->breakdown = 1; }
    // every block (and every rank: pq is identical everywhere) takes this branch; `done` is raised by k_pcg_direction
    return;
  }
  const double alpha = rz / pq;
  double s_rz = 0.0, s_rr = 0.0, s_max = 0.0;
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += gridDim.x * blockDim.x) {
    const int i = rev ? n - 1 - i0 : i0;
    const double pi = p[i], qi = q[i];
    const double xi = x[i] + alpha * pi;
    const double ri = r[i] - alpha * qi;
    x[i] = xi;
    r[i] = ri;
    s_rz += dinv ? ri * (dinv[i] * ri) : ri * ri;
    s_rr += ri * ri;
    s_max = fmax(s_max, fabs(ri));
  }
  double v[3] = {s_rz, s_rr, s_max}, tot[3];
  if (grid_reduce_mixed<3, 4u>(v, partials, &scal->cnt[1], sh_red, tot)) {
    if (threadIdx.x == 0) {
      if (pd.nranks > 1) p2p_publish<3>(pd, 1, pd.tagB, tot);
      else { scal->sums[1] = tot[0]; scal->sums[2] = tot[1]; scal->sums[3] = tot[2]; }
    }
  }
}

// p = D^-1 r + beta p; convergence decision (identical in every block, written once by block 0)
__global__ void __launch_bounds__(kVecBlock, 8) k_pcg_direction(int n, int par, int rev, int it, int max_it,
                                                             const double* __restrict__ r,
                                                             const double* __restrict__ dinv, double* __restrict__ p,
                                                             PcgScalars* scal, const P2PDev pd, const PushArgs push,
                                                             const int* __restrict__ send_idx) {
  __shared__ double sh_red[32];
  __shared__ bool s_last_dir;
  if (scal->done) return;
  if (scal->breakdown) {   // (1: operator not SPD; 2: a peer-memory wait timed out -- every rank sees its own timeout)
    if (blockIdx.x == 0 && threadIdx.x == 0) { scal->done = 1; scal->iters = it; }
    return;
  }
  double rz_new, rr, rmax;
  bool comm_ok = true;
  if (pd.nranks > 1) {
    double t[3];
    comm_ok = p2p_gather<3, 4u>(pd, 1, pd.tagB, t, sh_red);
    rz_new = t[0]; rr = t[1]; rmax = t[2];
  } else {
    rz_new = scal->sums[1]; rr = scal->sums[2]; rmax = scal->sums[3];
  }
  const double rz_old = scal->rz[par];
  const bool small_r = (rr <= scal->thresh2) || (rmax <= scal->thresh_inf);
  // error-based rule (Hestenes-Stiefel): the energy the last hs_d updates added, sum_j alpha_j r_j.z_j = ||x_k - x_{k-d}||_A^2,
  // a lower bound of ||x - x_{k-d}||_A^2, must be below eps_e^2 * E_ref.  Every block evaluates it from the ring the earlier
  // launches wrote + this iteration's own term (identical everywhere); block 0 alone stores.
  bool small_e = true;
  double hs_term = 0.0, hs_tot = 0.0;
  if (scal->eps_e2 > 0.0) {
    const int d = scal->hs_d;
    hs_term = rz_old * (rz_old / scal->sums[0]);   // alpha * r.z, alpha = r.z / p.Ap
    hs_tot = scal->hs_total + hs_term;
    double est = hs_term;
    for (int j = 1; j < d; ++j) est += (it - j >= 0) ? scal->hs_ring[(it - j) % d] : 0.0;
    small_e = (it + 1 >= d) && est <= scal->eps_e2 * fmax(fmax(scal->e_ref, fabs(scal->xb)), hs_tot);
  }
  const bool conv = (small_r && small_e) || rr == 0.0;
  const bool stop = conv || (it + 1 >= max_it) || !(rr == rr) || !comm_ok;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal->rz[par ^ 1] = rz_new;
    scal->rmax = rmax;
    if (!comm_ok) scal->breakdown = 2;
    scal->rr = rr;
    scal->iters = it + 1;
    if (scal->eps_e2 > 0.0) {
      scal->hs_ring[it % scal->hs_d] = hs_term;
      scal->hs_total = hs_tot;
      if (stop) scal->e_ref = fmax(scal->e_ref, fmax(hs_tot, fabs(scal->xb)));   // later corrections of the same Newton loop refer to the largest one (a warm-started one counts with its full size)
    }
    if (conv) scal->conv = 1;
    if (stop) scal->done = 1;
  }
  if (stop) return;
  const double beta = rz_new / rz_old;
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += gridDim.x * blockDim.x) {
    const int i = rev ? n - 1 - i0 : i0;
    p[i] = (dinv ? dinv[i] * r[i] : r[i]) + beta * p[i];
  }
  if (pd.nranks > 1 && push.nneigh > 0) {
    // Fused halo push: the block that finishes last (all of p is final) writes this rank's boundary values
    // straight into the neighbours' ghost blocks over NVLink and raises their flags for the NEXT iteration.
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned t = atomicInc(&scal->cnt[2], gridDim.x - 1);
      s_last_dir = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last_dir) {
      __threadfence();
      for (int j = 0; j < push.nneigh; ++j) {
        const int cnt = push.off[j + 1] - push.off[j];
        double* dst = push.dst[j];
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = __ldcg(p + send_idx[push.off[j] + i]);
      }
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x < push.nneigh) *(volatile unsigned long long*)push.flag[threadIdx.x] = push.tag;
    }
  }
}


// ---- symmetric Jacobi scaling: A' = S A S, S = D^-1/2.  CG on A' is Jacobi-PCG on A (same Krylov iterates) without
// the two per-iteration reads of D^-1: 11 instead of 13 vector passes per iteration.
__global__ void k_make_scale(int n, const double* __restrict__ dinv, double* __restrict__ sc, double* partials,
                             PcgScalars* scal) {
  __shared__ double sh_red[32];
  double m = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double d = dinv[i];
    sc[i] = sqrt(d);
    m = fmax(m, 1.0 / d);   // largest diagonal entry -> smallest scale factor
  }
  double v[1] = {m}, tot[1];
  if (grid_reduce<1, true>(v, partials, &scal->cnt[3], sh_red, tot)) {
    if (threadIdx.x == 0) scal->norm_inf = tot[0];
  }
}

__global__ void __launch_bounds__(kVecBlock) k_scale_matrix(int nslices, const int* __restrict__ slice_ptr,
                                                            const int* __restrict__ col, const double* __restrict__ val,
                                                            const double* __restrict__ sc, int nrows, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int slice = gw; slice < nslices; slice += nw) {
    const int base = slice_ptr[slice];
    const int w = (slice_ptr[slice + 1] - base) >> 5;
    const int row = slice * 32 + lane;
    const double sr = row < nrows ? sc[row] : 0.0;
    for (int k = 0; k < w; ++k) {
      const int idx = base + k * 32 + lane;
      out[idx] = sr * ld_stream(val + idx) * __ldg(sc + ld_stream(col + idx));
    }
  }
}

__global__ void k_vec_mul(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = a[i] * b[i];
}
__global__ void k_vec_div(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = a[i] / b[i];
}

int pcg_iteration_launch(Context* ctx, const double* val, const double* dinv, double* x, int it, int max_it, bool first);

// Solve val*x = rhs (val, dinv = the last assembled operator and its inverse diagonal).
int solve_system(Context* ctx, const double* val, const double* dinv, const double* rhs, double* x, bool x0_nonzero,
                 double rtol, double atol_inf, int64_t max_it, int64_t* iters_out, double* relres_out) {
  const int n = ctx->n_own;
  ctx->last_val = val; ctx->last_dinv = dinv;
  if (amg_ready(ctx) && n > 0)   // opt-in multigrid preconditioner, unscaled system
    return amg_pcg_solve(ctx, val, dinv, rhs, x, x0_nonzero, rtol, atol_inf, max_it, iters_out, relres_out);
  if (ctx->opts.no_sym_scaling || n == 0)
    return pcg_solve(ctx, val, dinv, rhs, x, x0_nonzero, rtol, atol_inf, max_it, iters_out, relres_out);
  const Pattern& P = ctx->pat;
  int rc;
  const int g = vec_grid(ctx, n, kVecBlock);
  if (!ctx->val_s) {   // the scaled copy of the operator exists only for contexts that take this path (0.47 GB at 16.8M triangles)
    const size_t ne = std::max<size_t>((size_t)P.nentries, 1);
    DDPS_CUDA(cudaMalloc(&ctx->val_s, ne * sizeof(double)));
    DDPS_CUDA(cudaMemsetAsync(ctx->val_s, 0, ne * sizeof(double), ctx->stream));
  }
  k_make_scale<<<g, kVecBlock, 0, ctx->stream>>>(n, dinv, ctx->sc, ctx->partials, ctx->scal);
  if ((rc = halo_exchange(ctx, ctx->sc))) return rc;
  k_scale_matrix<<<vec_grid(ctx, (int64_t)P.nslices * 32, kVecBlock), kVecBlock, 0, ctx->stream>>>(
      P.nslices, P.slice_ptr, P.col, val, ctx->sc, P.nrows, ctx->val_s);
  k_vec_mul<<<g, kVecBlock, 0, ctx->stream>>>(n, rhs, ctx->sc, ctx->bs);
  if (x0_nonzero) k_vec_div<<<g, kVecBlock, 0, ctx->stream>>>(n, x, ctx->sc, x);
  ctx->stats.kernel_launches += 3 + (x0_nonzero ? 1 : 0);
  double maxdiag = 0.0;
  if ((rc = allreduce_max(ctx, &ctx->scal->norm_inf, 1))) return rc;
  DDPS_CUDA(cudaMemcpyAsync(&ctx->scal_host->norm_inf, &ctx->scal->norm_inf, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  maxdiag = ctx->scal_host->norm_inf;
  // |r_i| = |r'_i| / s_i <= max|r'| * sqrt(max diag): scale the absolute threshold so that it still bounds the
  // UNscaled residual (conservative)
  const double atol_s = (atol_inf > 0 && maxdiag > 0) ? atol_inf / sqrt(maxdiag) : atol_inf;
  ctx->last_val = ctx->val_s; ctx->last_dinv = nullptr;
  rc = pcg_solve(ctx, ctx->val_s, nullptr, ctx->bs, x, x0_nonzero, rtol, atol_s, max_it, iters_out, relres_out);
  k_vec_mul<<<g, kVecBlock, 0, ctx->stream>>>(n, x, ctx->sc, x);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return rc;
}

int pcg_iteration_launch(Context* ctx, const double* val, const double* dinv, double* x, int it, int max_it, bool first) {
  const Pattern& P = ctx->pat;
  const int n = ctx->n_own;
  const int par = it & 1;
  int rc;
  const int g1 = vec_grid(ctx, (int64_t)P.nslices * 32, kVecBlock);
  const int g2 = vec_grid(ctx, n, kVecBlock);
  // traversal directions: the three kernels of an iteration alternate (opts.no_alt_traversal disables the trick)
  const int r0 = ctx->opts.no_alt_traversal ? 0 : (it & 1);
  const int r1 = ctx->opts.no_alt_traversal ? 0 : (r0 ^ 1);
  P2PDev pd;
  if (ctx->p2p.enabled) {
    // Fused compute + collective over peer memory (NVLink): no NCCL call inside the PCG iteration.
    P2PHost& H = ctx->p2p;
    pd = H.dev;
    pd.tag_halo = ++H.tag; pd.tagA = ++H.tag; pd.tagB = ++H.tag;
    PushArgs pa = H.push;
    if (first && pa.nneigh > 0) {   // later iterations: the previous k_pcg_direction already pushed (fused)
      pa.tag = pd.tag_halo;
      k_halo_push<<<pa.nneigh, 1024, 0, ctx->stream>>>(pa, ctx->halo.send_idx, ctx->p, ctx->scal);
      ctx->stats.kernel_launches++;
    }
    pa.tag = pd.tagB + 1;           // = tag_halo of the next iteration
    k_spmv<true, true><<<g1, kVecBlock, 0, ctx->stream>>>(P.nrows, P.nslices, P.slice_ptr, P.col, val, ctx->p, ctx->q,
                                                          ctx->partials, ctx->scal, pd, r0);
    k_pcg_update<<<g2, kVecBlock, 0, ctx->stream>>>(n, par, r1, ctx->p, ctx->q, dinv, x, ctx->r, ctx->partials, ctx->scal, pd);
    k_pcg_direction<<<g2, kVecBlock, 0, ctx->stream>>>(n, par, r0, it, max_it, ctx->r, dinv, ctx->p, ctx->scal, pd, pa,
                                                       ctx->halo.send_idx);
  } else {
    memset(&pd, 0, sizeof(pd));
    pd.nranks = 1;
    if ((rc = halo_xchg(ctx, ctx->halo, ctx->p, true))) return rc;
    k_spmv<true, false><<<g1, kVecBlock, 0, ctx->stream>>>(P.nrows, P.nslices, P.slice_ptr, P.col, val, ctx->p, ctx->q,
                                                           ctx->partials, ctx->scal, pd, r0);
    if ((rc = scalar_allreduce(ctx, &ctx->scal->sums[0], 1, 0, true))) return rc;
    k_pcg_update<<<g2, kVecBlock, 0, ctx->stream>>>(n, par, r1, ctx->p, ctx->q, dinv, x, ctx->r, ctx->partials, ctx->scal, pd);
    if ((rc = scalar_allreduce(ctx, &ctx->scal->sums[1], 2, 1, true))) return rc;
    PushArgs none;
    memset(&none, 0, sizeof(none));
    k_pcg_direction<<<g2, kVecBlock, 0, ctx->stream>>>(n, par, r0, it, max_it, ctx->r, dinv, ctx->p, ctx->scal, pd, none, nullptr);
  }
  ctx->stats.kernel_launches += 3;
  ctx->stats.spmv_total++;
  return 0;
}

int pcg_solve(Context* ctx, const double* val, const double* dinv, const double* rhs, double* x, bool x0_nonzero,
              double rtol, double atol_inf, int64_t max_it, int64_t* iters_out, double* relres_out) {
  const int n = ctx->n_own;
  int rc;
  if (max_it <= 0) max_it = 200000;
  if (max_it > 2000000000LL) max_it = 2000000000LL;
  cudaStream_t st = ctx->stream;
  int g = vec_grid(ctx, std::max(n, 1), kVecBlock);
  if (ctx->p2p.enabled)   // the sticky timeout flag of the peer-memory waits is per solve
    DDPS_CUDA(cudaMemsetAsync(&ctx->p2p.win_local->error, 0, sizeof(int), st));
  if (x0_nonzero) {
    DDPS_CUDA(cudaMemcpyAsync(ctx->p, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if ((rc = halo_exchange(ctx, ctx->p))) return rc;
    if ((rc = launch_spmv(ctx, val, ctx->p, ctx->q))) return rc;
    k_pcg_init<true><<<g, kVecBlock, 0, st>>>(n, rhs, ctx->q, x, dinv, ctx->r, ctx->p, ctx->partials, ctx->scal);
  } else {
    DDPS_CUDA(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), st));
    k_pcg_init<false><<<g, kVecBlock, 0, st>>>(n, rhs, nullptr, nullptr, dinv, ctx->r, ctx->p, ctx->partials, ctx->scal);
  }
  if ((rc = allreduce_sum(ctx, &ctx->scal->sums[0], 4))) return rc;
  if ((rc = allreduce_max(ctx, &ctx->scal->rmax, 1))) return rc;
  k_pcg_start<<<1, 1, 0, st>>>(ctx->scal, rtol, atol_inf, ctx->eps_e);
  ctx->stats.kernel_launches += 2;
  DDPS_CUDA(cudaGetLastError());

  int64_t it = 0;
  int chunk = 8;
  const int chunk_max = std::max(1, ctx->opts.check_every);
  while (true) {
    DDPS_CUDA(cudaMemcpyAsync(ctx->scal_host, ctx->scal, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    if (ctx->scal_host->done || it >= max_it) break;
    int nrun = (int)std::min<int64_t>(chunk, max_it - it);
    for (int j = 0; j < nrun; ++j, ++it)
      if ((rc = pcg_iteration_launch(ctx, val, dinv, x, (int)it, (int)max_it, it == 0))) return rc;
    DDPS_CUDA(cudaGetLastError());
    chunk = std::min(chunk * 2, chunk_max);
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
  if (!S.conv && !ctx->opts.lin_cap_ok) {
    ctx->fail(DDPS_ERR_NOCONV, "PCG hit the iteration cap (" + std::to_string(S.iters) + " its, relres " +
                                   std::to_string(S.bb > 0 ? sqrt(S.rr / S.bb) : 0.0) + ")");
    return DDPS_ERR_NOCONV;
  }
  return 0;
}

}  // namespace ddps
