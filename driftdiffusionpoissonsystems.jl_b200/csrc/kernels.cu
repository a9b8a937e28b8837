// Hand-written sm_100a FP64 kernels of the hot path (compiled with -fmad=false so the element
// formulas round exactly like the reference's IEEE expressions).
//
//   k_assemble<STIFF,MASS,RHS,RESID>  row-major fused assembly (K1-K7 of SURVEY section 2.1):
//       one thread per matrix row walks the row's incident triangles in increasing triangle order
//       (= the duplicate-summation order of the reference's sparse()), recomputes the triangle
//       geometry from coalesced 16-byte loads, and accumulates its row of every local 3x3 matrix
//       into private shared-memory slots that were resolved at setup time: no atomics, no staging
//       buffer, bitwise reproducible run to run.  Dirichlet elimination (src:271-284), boundary
//       lifting (src:362), the lumped edge-midpoint load vector (src:345-350) and -- for the
//       Newton systems -- the residual M*V-b (src:368) are fused into the same pass.
//   k_spmv_dot                        SELL-32 SpMV fused with the p.Ap partial dot (K8/K9)
//   k_pcg_update / k_pcg_direction    fused vector updates + dots of Jacobi-PCG (K9)
//   nodal prep kernels                exp(+-V/2), exp(+-V/3) and the nodal mass weights g_k, so
//                                     the element loops contain no transcendental at all.
#include "ddps_internal.cuh"

namespace ddps {

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Deterministic block reduction (fixed shuffle tree + fixed-order sum over warps).  Result valid in
// thread 0.  `sh` must hold >= 32 doubles; contains a trailing __syncthreads so it can be reused.
template <bool IS_MAX>
__device__ __forceinline__ double block_reduce(double v, double* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = IS_MAX ? warp_max(v) : warp_sum(v);
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  double out = 0.0;
  if (wid == 0) {
    double t = (lane < nw) ? sh[lane] : (IS_MAX ? 0.0 : 0.0);
    t = IS_MAX ? warp_max(t) : warp_sum(t);
    out = t;
  }
  __syncthreads();
  return out;
}

// "last block finishes" grid reduction: every block stores its partial, the block that takes the
// last ticket sums all partials in a FIXED order (independent of which block that is).
// Returns true (in all threads of the last block) with the total in `total` of thread 0.
template <int NV, bool IS_MAX>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double* partials, unsigned* ticket, double* sh,
                                            double (&total)[NV]) {
  __shared__ bool s_last;
  double bv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) bv[i] = block_reduce<IS_MAX>(v[i], sh);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) partials[(size_t)i * gridDim.x + blockIdx.x] = bv[i];
    __threadfence();
    unsigned t = atomicInc(ticket, gridDim.x - 1);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
    const volatile double* pp = partials + (size_t)i * gridDim.x;
    for (unsigned j = threadIdx.x; j < gridDim.x; j += blockDim.x) s = IS_MAX ? fmax(s, pp[j]) : s + pp[j];
    total[i] = block_reduce<IS_MAX>(s, sh);
  }
  return true;
}

__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }

// ------------------------------------------------------------------------------------------------
// Row-major fused assembly.
//
// Thread mapping: one CTA per 32-row slice; warp k of the CTA handles incidence k (the k-th incident
// triangle, in increasing triangle order) of ALL 32 rows, lane = row.  Every stream that is stored
// slice-interleaved (incidence records, matrix values, columns) is therefore read/written as one
// fully coalesced 256-byte warp access, neighbouring lanes gather neighbouring triangles/nodes, and
// a row's ~6 dependent load chains run in 6 different warps instead of back to back in one thread.
//
// Deterministic, atomic-free scatter: an off-diagonal entry (i,j) receives at most two contributions
// (the two triangles sharing edge ij; checked at setup).  The setup marks every contribution as the
// first or the second one in increasing triangle order, the kernel stores them into two separate
// shared-memory planes (accA/accB) and adds A+B after ONE barrier.  The diagonal, the load and the
// lifting receive one contribution per incident triangle; those are staged per incidence index and
// summed in incidence order.  The result is bitwise reproducible and uses the same summation order
// as the reference's sparse() (increasing triangle id per entry).
// ------------------------------------------------------------------------------------------------
template <int STIFF, bool MASS, int RHS, bool RESID>
__global__ void __launch_bounds__(256) k_assemble(const AsmArgs P) {
  extern __shared__ double sm[];
  __shared__ int s_diag[32];
  __shared__ double s_dsum[32], s_bsum[32], s_lsum[32], s_dval[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int slice = blockIdx.x;
  const int row = slice * 32 + lane;
  const bool live = row < P.nrows;
  const int base = P.slice_ptr[slice];
  const int w = (P.slice_ptr[slice + 1] - base) >> 5;
  const int ibase = P.inc_slice_ptr[slice];
  const int iw = (P.inc_slice_ptr[slice + 1] - ibase) >> 5;
  double* accA = sm;                    // [w][32] first contribution of every off-diagonal slot
  double* accB = accA + w * 32;         // [w][32] second contribution
  double* stS = accB + w * 32;          // [nw][32] diagonal contribution of incidence k
  double* stB = stS + nw * 32;          // [nw][32] load contribution
  double* stL = stB + nw * 32;          // [nw][32] lifting contribution
  double* prod = stL + nw * 32;         // [w][32] base*x products (RESID)

  // ---- load stage 1: everything that does not depend on another load is issued up front -----------
  // (the kernel is latency-bound: a CTA's life is [#dependent load stages] x [DRAM latency], so the
  //  chain is kept at two stages: {records, own-node data, matrix slot} -> {neighbour data, x gather})
  const int k_first = warp;                                  // incidence handled in pass 0
  int4 rec = make_int4(-1, 0, 0, 0);
  if (k_first < iw && live) rec = P.inc[ibase + k_first * 32 + lane];
  const int rr = live ? row : 0;
  const double2 q_own = P.xy[rr];
  const bool rowD = live && P.isD[rr] != 0;
  double own_g = 0.0, own_e3 = 0.0, own_u = 0.0, own_v = 0.0, own_eh = 0.0, own_ehi = 0.0;
  if (MASS) own_g = P.gw[rr];
  if (STIFF == STIFF_EXPV) own_e3 = P.E3[rr];
  if (RHS == RHS_DDP || RHS == RHS_SRH) { own_u = P.fu[rr]; own_v = P.fv[rr]; own_eh = P.EH[rr]; own_ehi = P.EHi[rr]; }
  if (RHS == RHS_SINH || RHS == RHS_POLY) own_eh = P.EH[rr];
  double m_slot = 0.0;   // base matrix entry of the slot this warp writes first (slot = warp)
  int c_slot = 0;
  if (STIFF == STIFF_BASE && warp < w) {
    m_slot = ld_stream(P.base_val + base + warp * 32 + lane);
    if (RESID) c_slot = ld_stream(P.col + base + warp * 32 + lane);
  }
  double fin_lift = 0.0, fin_gd = 0.0, fin_nc = 0.0;
  if (warp == 0 && RHS != RHS_NONE) {
    if (STIFF == STIFF_BASE) fin_lift = P.lift_in[rr];
    fin_gd = P.gD[rr];
    if (P.nodal_const) fin_nc = P.nodal_const[rr];
  }
  for (int k = warp; k < 2 * w; k += nw) accA[k * 32 + lane] = 0.0;
  if (warp == 0) { s_diag[lane] = 0; s_dsum[lane] = 0.0; s_bsum[lane] = 0.0; s_lsum[lane] = 0.0; }
  // ---- load stage 2 (needs the column index) ----
  double x_slot = 0.0;
  if (STIFF == STIFF_BASE && RESID && warp < w) x_slot = P.xvec[c_slot];
  __syncthreads();

  for (int k0 = 0; k0 < iw; k0 += nw) {
    const int k = k0 + warp;
    if (k0 > 0) { rec = make_int4(-1, 0, 0, 0); if (k < iw && live) rec = P.inc[ibase + k * 32 + lane]; }
    double c_self = 0.0, bv = 0.0, lv = 0.0;
    if (rec.x >= 0) {
      const int e = rec.x >> 2, a = rec.x & 3;
      const int nb = rec.z, nc = rec.w;                      // next / previous corner (cyclic)
      const double2 q_b = P.xy[nb], q_c = P.xy[nc];
      // canonical corner order (1,2,3) of the triangle: corner a is this row
      const double2 q1 = (a == 0) ? q_own : (a == 1) ? q_c : q_b;
      const double2 q2 = (a == 0) ? q_b : (a == 1) ? q_own : q_c;
      const double2 q3 = (a == 0) ? q_c : (a == 1) ? q_b : q_own;
      // edge vectors and signed area, src:243-246
      const double ux = q2.x - q3.x, uy = q2.y - q3.y;
      const double vx = q3.x - q1.x, vy = q3.y - q1.y;
      const double wx = q1.x - q2.x, wy = q1.y - q2.y;
      const double ar = (ux * vy - uy * vx) * 0.5;
      const int s_self = rec.y & 255, s_next = (rec.y >> 8) & 255, s_prev = (rec.y >> 16) & 255;
      double c_next = 0.0, c_prev = 0.0;
      if (STIFF != STIFF_BASE) {
        double phx, phy;
        if (STIFF == STIFF_COEF) { phx = P.phx[e]; phy = P.phy[e]; }
        else {
          // exp((V1+V2+V3)/3) as the product of the nodal factors in canonical corner order, src:469
          const double eb = P.E3[nb], ec = P.E3[nc];
          const double e1 = (a == 0) ? own_e3 : (a == 1) ? ec : eb;
          const double e2 = (a == 0) ? eb : (a == 1) ? own_e3 : ec;
          const double e3 = (a == 0) ? ec : (a == 1) ? eb : own_e3;
          phx = P.phx[e] * (e1 * e2 * e3); phy = phx;
        }
        const double inv4 = 1.0 / fabs(ar * 4.0);  // src:257 (one division per triangle; <= 1 ulp vs x/ar4)
        // canonical lower-triangle entries K(i,j) = (e_i.x*phx*e_j.x + e_i.y*phy*e_j.y)/|4ar|, i>=j
        // (src:260-266); select the row of corner a: self, next, prev.
        const double sx = (a == 0) ? ux : (a == 1) ? vx : wx, sy = (a == 0) ? uy : (a == 1) ? vy : wy;
        const double nhx = (a == 0) ? vx : wx, nhy = (a == 0) ? vy : wy;   // (a,next): (2,1),(3,2),(3,1)
        const double nlx = (a == 1) ? vx : ux, nly = (a == 1) ? vy : uy;
        const double khx = (a == 1) ? vx : wx, khy = (a == 1) ? vy : wy;   // (a,prev): (3,1),(2,1),(3,2)
        const double klx = (a == 2) ? vx : ux, kly = (a == 2) ? vy : uy;
        const double k_self = (sx * phx * sx + sy * phy * sy) * inv4;
        const double k_next = (nhx * phx * nlx + nhy * phy * nly) * inv4;
        const double k_prev = (khx * phx * klx + khy * phy * kly) * inv4;
        const bool nbD = P.isD[nb] != 0, ncD = P.isD[nc] != 0;
        // Dirichlet elimination src:271-278: rows and columns of D are zeroed (kept in the pattern)
        c_self = rowD ? 0.0 : k_self;
        c_next = (rowD || nbD) ? 0.0 : k_next;
        c_prev = (rowD || ncD) ? 0.0 : k_prev;
        // boundary lifting b -= K[:,D]*gD on rows outside D, src:362
        if (!rowD) {
          if (nbD) lv += k_next * P.gD[nb];
          if (ncD) lv += k_prev * P.gD[nc];
        }
      }
      if (MASS) {
        // weighted mass with nodal weights W_k = g_k*ar/30, SIGNED ar (src:375-386, Q2); J0 is never
        // eliminated (Q3).  In the row's own frame: self = 3Wa+Wb+Wc, (a,b) = Wa+Wb+Wc/2.
        const double ar30 = ar * (1.0 / 30.0);
        const double Wa = own_g * ar30, Wb = P.gw[nb] * ar30, Wc = P.gw[nc] * ar30;
        // canonical addition order of src:380-386 (W1,W2,W3 left to right)
        const double W1 = (a == 0) ? Wa : (a == 1) ? Wc : Wb;
        const double W2 = (a == 0) ? Wb : (a == 1) ? Wa : Wc;
        const double W3 = (a == 0) ? Wc : (a == 1) ? Wb : Wa;
        double m_self, m_next, m_prev;
        if (a == 0) { m_self = 3.0 * W1 + W2 + W3; m_next = W1 + W2 + W3 * 0.5; m_prev = W1 + W2 * 0.5 + W3; }
        else if (a == 1) { m_self = W1 + 3.0 * W2 + W3; m_next = W1 * 0.5 + W2 + W3; m_prev = W1 + W2 + W3 * 0.5; }
        else { m_self = W1 + W2 + 3.0 * W3; m_next = W1 + W2 * 0.5 + W3; m_prev = W1 * 0.5 + W2 + W3; }
        c_self -= m_self; c_next -= m_next; c_prev -= m_prev;
      }
      // lumped edge-midpoint load: corner a takes the midpoint between a and next (src:345-350, Q1)
      if (RHS != RHS_NONE && RHS != RHS_ZERO) {
        double f;
        if (RHS == RHS_DDP) {   // src:341-343
          const double ua = (own_u + P.fu[nb]) * 0.5, va = (own_v + P.fv[nb]) * 0.5;
          const double eS = own_eh * P.EH[nb], eSi = own_ehi * P.EHi[nb];
          f = P.par0 * (eSi * va - eS * ua) + P.mid[3 * e + a];
        } else if (RHS == RHS_SRH) {  // src:517-519
          const double ua = (own_u + P.fu[nb]) * 0.5, va = (own_v + P.fv[nb]) * 0.5;
          const double eS = own_eh * P.EH[nb], eSi = own_ehi * P.EHi[nb];
          f = 1.0 / (P.par0 * (eS * ua + 1.0) + P.par1 * (eSi * va + 1.0));
        } else if (RHS == RHS_SINH) {  // registered functor, test:95
          const double S = (own_eh + P.EH[nb]) * 0.5;
          f = P.mid[3 * e + a] + P.par0 * sinh(P.par1 * S);
        } else {  // RHS_POLY, README:107-108
          const double S = (own_eh + P.EH[nb]) * 0.5;
          const int deg = (int)P.par0;
          f = P.mid_poly[deg][3 * e + a];
          for (int d = deg - 1; d >= 0; --d) f = f * S + P.mid_poly[d][3 * e + a];
        }
        bv = f * (fabs(ar) * (1.0 / 3.0));  // src:349
      }
      if (STIFF != STIFF_BASE || MASS) {
        // first / second contributor planes (ordinal bits resolved at setup)
        double* pn = (rec.y & (1 << 24)) ? accB : accA;
        double* pp = (rec.y & (1 << 25)) ? accB : accA;
        pn[s_next * 32 + lane] = c_next;
        pp[s_prev * 32 + lane] = c_prev;
      }
      s_diag[lane] = s_self;
    }
    stS[warp * 32 + lane] = c_self;
    stB[warp * 32 + lane] = bv;
    stL[warp * 32 + lane] = lv;
    __syncthreads();
    if (warp == 0) {   // ordered sums of this pass (incidence order)
      double ds = s_dsum[lane], bs = s_bsum[lane], ls = s_lsum[lane];
      const int cnt = min(nw, iw - k0);
      for (int kk = 0; kk < cnt; ++kk) { ds += stS[kk * 32 + lane]; bs += stB[kk * 32 + lane]; ls += stL[kk * 32 + lane]; }
      s_dsum[lane] = ds; s_bsum[lane] = bs; s_lsum[lane] = ls;
    }
    __syncthreads();
  }

  // ---- write the slice: warp k streams slot k of all 32 rows ----
  for (int k = warp; k < w; k += nw) {
    const int idx = base + k * 32 + lane;
    const bool isdiag = live && (k == s_diag[lane]);
    double a_k = isdiag ? s_dsum[lane] : accA[k * 32 + lane] + accB[k * 32 + lane];
    if (STIFF == STIFF_BASE) {
      double m_k = m_slot, x_k = x_slot;
      if (k != warp) {   // slots beyond the first nw (rows longer than the CTA's warp count)
        m_k = ld_stream(P.base_val + idx);
        if (RESID) x_k = P.xvec[ld_stream(P.col + idx)];
      }
      if (RESID) prod[k * 32 + lane] = live ? m_k * x_k : 0.0;
      a_k += m_k;
    } else if (isdiag && rowD) {
      a_k += 1.0;   // src:279-281
    }
    if (!live) a_k = 0.0;
    P.out_val[idx] = a_k;
    if (isdiag) s_dval[lane] = a_k;
  }
  __syncthreads();
  if (warp == 0 && live) {
    if (P.dinv) P.dinv[row] = 1.0 / s_dval[lane];
    double lift = s_lsum[lane];
    if (STIFF != STIFF_BASE && P.lift_out) P.lift_out[row] = lift;
    if (RHS != RHS_NONE) {
      if (STIFF == STIFF_BASE) lift = fin_lift;
      double bi = s_bsum[lane] + fin_nc;             // Neumann loads, src:994-996
      bi -= lift;                                    // src:362
      if (rowD) bi = fin_gd;                         // src:363
      P.b[row] = bi;
      if (RESID) {
        double mv = 0.0;
        for (int k = 0; k < w; ++k) mv += prod[k * 32 + lane];
        P.resid[row] = mv - bi;                      // src:368
      }
    }
  }
}

template <int STIFF, bool MASS, int RHS, bool RESID>
static int launch_asm(Context* ctx, const AsmArgs& a) {
  const Pattern& P = ctx->pat;
  if (P.nslices == 0) return 0;
  const int nw = std::max(1, std::min(8, P.max_inc));
  size_t smem = (size_t)((RESID ? 3 : 2) * P.max_row_len + 3 * nw) * 32 * sizeof(double);
  k_assemble<STIFF, MASS, RHS, RESID><<<P.nslices, 32 * nw, smem, ctx->stream>>>(a);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

int launch_assemble_base(Context* ctx, const AsmArgs& a) { return launch_asm<STIFF_COEF, false, RHS_NONE, false>(ctx, a); }

int launch_assemble_uv(Context* ctx, const AsmArgs& a, int rhs_kind, bool mass) {
  if (rhs_kind == RHS_SRH && mass) return launch_asm<STIFF_EXPV, true, RHS_SRH, false>(ctx, a);
  if (rhs_kind == RHS_ZERO && !mass) return launch_asm<STIFF_EXPV, false, RHS_ZERO, false>(ctx, a);
  ctx->fail(DDPS_ERR_ARG, "unsupported uv assembly mode");
  return DDPS_ERR_ARG;
}

int launch_assemble_newton(Context* ctx, const AsmArgs& a, int rhs_kind) {
  switch (rhs_kind) {
    case RHS_DDP: return launch_asm<STIFF_BASE, true, RHS_DDP, true>(ctx, a);
    case RHS_SINH: return launch_asm<STIFF_BASE, true, RHS_SINH, true>(ctx, a);
    case RHS_POLY: return launch_asm<STIFF_BASE, true, RHS_POLY, true>(ctx, a);
  }
  ctx->fail(DDPS_ERR_ARG, "unsupported Newton rhs functor");
  return DDPS_ERR_ARG;
}

// ------------------------------------------------------------------------------------------------
// Nodal prep kernels (all transcendental work of an assembly, one pass over the nodes)
// ------------------------------------------------------------------------------------------------
__global__ void k_prep_poisson(int n, const double* __restrict__ V, const double* __restrict__ u,
                               const double* __restrict__ v, double delta2, double* __restrict__ EH,
                               double* __restrict__ EHi, double* __restrict__ gw) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double Vi = V[i];
  EH[i] = exp(Vi / 2.0);
  EHi[i] = exp(-Vi / 2.0);
  gw[i] = -delta2 * (u[i] * exp(Vi) + v[i] * exp(-Vi));  // src:375-377
}

// sign = +1: u-solve with (V, tau_p, tau_n, u0, v0); sign = -1: v-solve, the reference passes
// (-V, tau_n, tau_p, v0, u0) (src:598-599).  V here is always the true potential.
__global__ void k_prep_uv(int n, double sign, const double* __restrict__ V, const double* __restrict__ u0,
                          const double* __restrict__ v0, double tau_p, double tau_n, int shockley,
                          double* __restrict__ EH, double* __restrict__ EHi, double* __restrict__ E3,
                          double* __restrict__ gw) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double Vs = sign * V[i];   // the "V" the reference's uvsolve sees
  EH[i] = exp(Vs / 2.0);
  EHi[i] = exp(-Vs / 2.0);
  E3[i] = exp(Vs / 3.0);           // exp((V1+V2+V3)/3) = product of the three nodal factors, src:469
  if (shockley) {
    // src:568: g_k = -v0_k / (tau_p(e^{V}u0+1) + tau_n(e^{-V}v0+1)) in the callee's own naming
    const double a = (sign > 0) ? u0[i] : v0[i];   // callee's u0
    const double b = (sign > 0) ? v0[i] : u0[i];   // callee's v0
    const double tp = (sign > 0) ? tau_p : tau_n, tn = (sign > 0) ? tau_n : tau_p;
    gw[i] = -b / (tp * (exp(Vs) * a + 1.0) + tn * (exp(-Vs) * b + 1.0));
  } else {
    gw[i] = 0.0;
  }
}

struct PolyNodal { const double* c[kMaxPoly]; };

__global__ void k_prep_semlin(int n, const double* __restrict__ V, int kind, double p0, double p1, PolyNodal pn,
                              double* __restrict__ gw) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double u = V[i];
  double g;
  if (kind == DDPS_FUNCTOR_SINH) {
    g = p0 * p1 * cosh(p1 * u);   // g = df/du, test:96
  } else {
    const int deg = (int)p0;
    g = 0.0;
    for (int d = deg; d >= 1; --d) g = g * u + d * pn.c[d][i];
  }
  gw[i] = g;   // src:1046-1048
}

int launch_prep_poisson(Context* ctx, const double* V, const double* u, const double* v, double delta2) {
  int n = ctx->n_loc;
  if (!n) return 0;
  k_prep_poisson<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, V, u, v, delta2, ctx->EH, ctx->EHi, ctx->gw);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

int launch_prep_uv(Context* ctx, double sign, const double* V, const double* /*unused*/, const double* u0,
                   const double* v0, double tau_p, double tau_n, int rec_mode) {
  int n = ctx->n_loc;
  if (!n) return 0;
  k_prep_uv<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, sign, V, u0, v0, tau_p, tau_n, rec_mode == DDPS_REC_SHOCKLEY,
                                                    ctx->EH, ctx->EHi, ctx->E3, ctx->gw);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

int launch_prep_semlin(Context* ctx, const double* V) {
  int n = ctx->n_loc;
  if (!n) return 0;
  PolyNodal pn;
  for (int k = 0; k < kMaxPoly; ++k) pn.c[k] = ctx->coeff[DDPS_COEFF_POLY_NODE0 + k];
  k_prep_semlin<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, V, ctx->functor_kind, ctx->functor_par[0],
                                                        ctx->functor_par[1], pn, ctx->gw);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// SpMV (SELL-32) with fused dot, persistent grid-stride over slices
// ------------------------------------------------------------------------------------------------
template <bool DOT>
__global__ void __launch_bounds__(kVecBlock) k_spmv(int nrows, int nslices, const int* __restrict__ slice_ptr,
                                                    const int* __restrict__ col, const double* __restrict__ val,
                                                    const double* __restrict__ x, double* __restrict__ y,
                                                    double* partials, PcgScalars* scal) {
  __shared__ double sh_red[32];
  if (DOT && scal->done) return;
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
      const double v0 = ld_stream(val + idx), v1 = ld_stream(val + idx + 32), v2 = ld_stream(val + idx + 64),
                   v3 = ld_stream(val + idx + 96);
      const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
      sum += v0 * x0; sum += v1 * x1; sum += v2 * x2; sum += v3 * x3;
      idx += 128;
    }
    for (; k < w; ++k) {
      const int c0 = ld_stream(col + idx);
      const double v0 = ld_stream(val + idx);
      sum += v0 * __ldg(x + c0);
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
      if (threadIdx.x == 0) scal->sums[0] = tot[0];
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
  k_spmv<false><<<grid, kVecBlock, 0, ctx->stream>>>(P.nrows, P.nslices, P.slice_ptr, P.col, val, x, y, nullptr, nullptr);
  ctx->stats.kernel_launches++;
  ctx->stats.spmv_total++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

int launch_spmv_dot(Context* ctx, const double* val, const double* x, double* y) {
  const Pattern& P = ctx->pat;
  if (!P.nslices) return 0;
  int grid = vec_grid(ctx, (int64_t)P.nslices * 32, kVecBlock);
  k_spmv<true><<<grid, kVecBlock, 0, ctx->stream>>>(P.nrows, P.nslices, P.slice_ptr, P.col, val, x, y, ctx->partials, ctx->scal);
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
template <bool HAVE_Q>
__global__ void k_pcg_init(int n, const double* __restrict__ rhs, const double* __restrict__ q,
                           const double* __restrict__ dinv, double* __restrict__ r, double* __restrict__ p,
                           double* partials, PcgScalars* scal) {
  __shared__ double sh_red[32];
  double s_rz = 0.0, s_rr = 0.0, s_bb = 0.0, s_max = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double bi = rhs[i];
    const double ri = HAVE_Q ? bi - q[i] : bi;
    const double zi = dinv[i] * ri;
    r[i] = ri;
    p[i] = zi;
    s_rz += ri * zi; s_rr += ri * ri; s_bb += bi * bi;
    s_max = fmax(s_max, fabs(ri));
  }
  double v[3] = {s_rz, s_rr, s_bb}, tot[3];
  const bool last = grid_reduce<3, false>(v, partials, &scal->cnt[1], sh_red, tot);
  if (last && threadIdx.x == 0) { scal->sums[0] = tot[0]; scal->sums[1] = tot[1]; scal->sums[2] = tot[2]; }
  double vm[1] = {s_max}, totm[1];
  if (grid_reduce<1, true>(vm, partials + 3 * gridDim.x, &scal->cnt[2], sh_red, totm)) {
    if (threadIdx.x == 0) scal->rmax = totm[0];
  }
}

// stop rule: ||r||_2 <= rtol*||rhs||_2  OR  ||r||_inf <= atol_inf
__global__ void k_pcg_start(PcgScalars* scal, double rtol, double atol_inf) {
  scal->rz[0] = scal->sums[0];
  scal->rr = scal->sums[1];
  scal->bb = scal->sums[2];
  scal->thresh2 = rtol * rtol * scal->sums[2];
  scal->thresh_inf = atol_inf;
  scal->iters = 0;
  scal->breakdown = 0;
  scal->done = (scal->sums[1] <= scal->thresh2 || scal->rmax <= atol_inf) ? 1 : 0;
}

// x += alpha p; r -= alpha q; local sums {r.D^-1 r, r.r}
__global__ void __launch_bounds__(kVecBlock) k_pcg_update(int n, int par, const double* __restrict__ p,
                                                          const double* __restrict__ q,
                                                          const double* __restrict__ dinv, double* __restrict__ x,
                                                          double* __restrict__ r, double* partials, PcgScalars* scal) {
  __shared__ double sh_red[32];
  if (scal->done) return;
  const double pq = scal->sums[0];
  const double rz = scal->rz[par];
  if (!(pq > 0.0) || !(rz == rz)) {   // operator not SPD or NaN: stop
    if (blockIdx.x == 0 && threadIdx.x == 0) { scal->breakdown = 1; }
    // every block takes this branch consistently; `done` is raised by k_pcg_direction
    return;
  }
  const double alpha = rz / pq;
  double s_rz = 0.0, s_rr = 0.0, s_max = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double pi = p[i], qi = q[i];
    const double xi = x[i] + alpha * pi;
    const double ri = r[i] - alpha * qi;
    x[i] = xi;
    r[i] = ri;
    s_rz += ri * (dinv[i] * ri);
    s_rr += ri * ri;
    s_max = fmax(s_max, fabs(ri));
  }
  double v[2] = {s_rz, s_rr}, tot[2];
  const bool last = grid_reduce<2, false>(v, partials, &scal->cnt[1], sh_red, tot);
  if (last && threadIdx.x == 0) { scal->sums[1] = tot[0]; scal->sums[2] = tot[1]; }
  double vm[1] = {s_max}, totm[1];
  if (grid_reduce<1, true>(vm, partials + 2 * gridDim.x, &scal->cnt[2], sh_red, totm)) {
    if (threadIdx.x == 0) scal->rmax = totm[0];
  }
}

// p = D^-1 r + beta p; convergence decision (identical in every block, written once by block 0)
__global__ void __launch_bounds__(kVecBlock) k_pcg_direction(int n, int par, int it, int max_it,
                                                             const double* __restrict__ r,
                                                             const double* __restrict__ dinv, double* __restrict__ p,
                                                             PcgScalars* scal) {
  if (scal->done) return;
  if (scal->breakdown) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { scal->done = 1; scal->iters = it; }
    return;
  }
  const double rz_new = scal->sums[1], rr = scal->sums[2];
  const double rz_old = scal->rz[par];
  const bool stop = (rr <= scal->thresh2) || (scal->rmax <= scal->thresh_inf) || (it + 1 >= max_it) || !(rr == rr);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal->rz[par ^ 1] = rz_new;
    scal->rr = rr;
    scal->iters = it + 1;
    if (stop) scal->done = 1;
  }
  if (stop) return;
  const double beta = rz_new / rz_old;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    p[i] = dinv[i] * r[i] + beta * p[i];
}

int pcg_iteration_launch(Context* ctx, const double* val, const double* dinv, double* x, int it, int max_it) {
  const Pattern& P = ctx->pat;
  const int n = ctx->n_own;
  const int par = it & 1;
  int rc;
  if ((rc = halo_exchange(ctx, ctx->p))) return rc;
  int g1 = vec_grid(ctx, (int64_t)P.nslices * 32, kVecBlock);
  k_spmv<true><<<g1, kVecBlock, 0, ctx->stream>>>(P.nrows, P.nslices, P.slice_ptr, P.col, val, ctx->p, ctx->q,
                                                  ctx->partials, ctx->scal);
  if ((rc = allreduce_sum(ctx, &ctx->scal->sums[0], 1))) return rc;
  int g2 = vec_grid(ctx, n, kVecBlock);
  k_pcg_update<<<g2, kVecBlock, 0, ctx->stream>>>(n, par, ctx->p, ctx->q, dinv, x, ctx->r, ctx->partials, ctx->scal);
  if ((rc = allreduce_sum(ctx, &ctx->scal->sums[1], 2))) return rc;
  if ((rc = allreduce_max(ctx, &ctx->scal->rmax, 1))) return rc;
  k_pcg_direction<<<g2, kVecBlock, 0, ctx->stream>>>(n, par, it, max_it, ctx->r, dinv, ctx->p, ctx->scal);
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
  if (x0_nonzero) {
    DDPS_CUDA(cudaMemcpyAsync(ctx->p, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if ((rc = halo_exchange(ctx, ctx->p))) return rc;
    if ((rc = launch_spmv(ctx, val, ctx->p, ctx->q))) return rc;
    k_pcg_init<true><<<g, kVecBlock, 0, st>>>(n, rhs, ctx->q, dinv, ctx->r, ctx->p, ctx->partials, ctx->scal);
  } else {
    DDPS_CUDA(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), st));
    k_pcg_init<false><<<g, kVecBlock, 0, st>>>(n, rhs, nullptr, dinv, ctx->r, ctx->p, ctx->partials, ctx->scal);
  }
  if ((rc = allreduce_sum(ctx, &ctx->scal->sums[0], 3))) return rc;
  if ((rc = allreduce_max(ctx, &ctx->scal->rmax, 1))) return rc;
  k_pcg_start<<<1, 1, 0, st>>>(ctx->scal, rtol, atol_inf);
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
      if ((rc = pcg_iteration_launch(ctx, val, dinv, x, (int)it, (int)max_it))) return rc;
    DDPS_CUDA(cudaGetLastError());
    chunk = std::min(chunk * 2, chunk_max);
  }
  const PcgScalars& S = *ctx->scal_host;
  if (iters_out) *iters_out = S.iters;
  if (relres_out) *relres_out = (S.bb > 0) ? sqrt(S.rr / S.bb) : 0.0;
  ctx->stats.pcg_total += S.iters;
  ctx->hist_pcg.push_back((double)S.iters);
  if (S.breakdown) {
    ctx->fail(DDPS_ERR_BREAKDOWN, "PCG breakdown: p.Ap <= 0 (operator not SPD; u or v negative? SURVEY H3)");
    return DDPS_ERR_BREAKDOWN;
  }
  if (!(S.rr <= S.thresh2 || S.rmax <= S.thresh_inf) && !ctx->opts.lin_cap_ok) {
    ctx->fail(DDPS_ERR_NOCONV, "PCG hit the iteration cap (" + std::to_string(S.iters) + " its, relres " +
                                   std::to_string(S.bb > 0 ? sqrt(S.rr / S.bb) : 0.0) + ")");
    return DDPS_ERR_NOCONV;
  }
  return 0;
}

}  // namespace ddps
