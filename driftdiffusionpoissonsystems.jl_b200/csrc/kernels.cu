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
// Row-major fused assembly with block-cooperative staging.
//
// One CTA per block of 128 consecutive rows, one thread per row.
//   level 1 (coalesced):  the block's sorted node list (built at setup) and each row's packed 8-byte
//                         incidence records;
//   level 2 (gathers):    the nodal fields of every DISTINCT node the block touches are gathered ONCE
//                         into shared memory (a 128-row block of a P1 mesh touches ~3x128 nodes but its
//                         rows reference them ~13x128 times), plus the per-triangle coefficients;
//   compute:              each thread walks its row's incident triangles in increasing triangle order,
//                         reads the two other corners from shared memory through the block-local
//                         indices packed in the record, and accumulates its row of every local 3x3
//                         matrix into its PRIVATE shared-memory column (slots resolved at setup).
// No atomics, no staging buffer in HBM, bitwise reproducible, same summation order as the reference's
// sparse() (increasing triangle id per entry).  Dirichlet elimination (src:271-284), lifting
// (src:362), the lumped edge-midpoint load (src:345-350) and, for the Newton systems, the residual
// M*V-b (src:368) are fused into the same pass; all transcendentals live in the nodal prep kernels.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxIncReg = 8;   // incidence records held in registers; longer rows take the tail loop
#ifndef DDPS_ASM_MINBLOCKS
#define DDPS_ASM_MINBLOCKS 1
#endif

template <int STIFF, bool MASS, int RHS, bool RESID>
__global__ void __launch_bounds__(kAsmBlock, DDPS_ASM_MINBLOCKS) k_assemble(const AsmArgs P) {
  extern __shared__ double sm[];
  constexpr bool NEED_UV = (RHS == RHS_DDP || RHS == RHS_SRH);
  constexpr bool NEED_EH = NEED_UV || RHS == RHS_SINH || RHS == RHS_POLY;
  const int tid = threadIdx.x, lane = tid & 31;
  const int lmax = P.lmax;
  // shared memory: acc[wmax][128] | X | Y | E3 | GW | FU | FV | EH | EHI  (each [lmax]) | node ids | isD
  double* acc = sm;
  double* sX = acc + (size_t)P.wmax * kAsmBlock;
  double* sY = sX + lmax;
  double* sE3 = sY + lmax;
  double* sGW = sE3 + (STIFF == STIFF_EXPV ? lmax : 0);
  double* sFU = sGW + (MASS ? lmax : 0);
  double* sFV = sFU + (NEED_UV ? lmax : 0);
  double* sEH = sFV + (NEED_UV ? lmax : 0);
  double* sEHI = sEH + (NEED_EH ? lmax : 0);
  int* sNode = reinterpret_cast<int*>(sEHI + (NEED_UV ? lmax : 0));
  unsigned char* sD = reinterpret_cast<unsigned char*>(sNode + lmax);

  const int blk = blockIdx.x;
  const int row = blk * kAsmRows + tid;
  const bool live = row < P.nrows;
  const int rr = live ? row : 0;
  const int slice = row >> 5;
  const bool in_pat = slice < P.nslices;
  int base = 0, w = 0, ibase = 0, iw = 0;
  if (in_pat) {
    base = P.slice_ptr[slice]; w = (P.slice_ptr[slice + 1] - base) >> 5;
    ibase = P.inc_slice_ptr[slice]; iw = (P.inc_slice_ptr[slice + 1] - ibase) >> 5;
  }
  // ---- level 1: incidence records (coalesced) ----
  int2 rec[kMaxIncReg];
#pragma unroll
  for (int k = 0; k < kMaxIncReg; ++k) {
    rec[k] = make_int2(-1, 0);
    if (k < iw && live) rec[k] = P.inc[ibase + k * 32 + lane];
  }
  // ---- stage the block's distinct nodes (level 1: ids, level 2: fields) ----
  const int nb0 = P.bn_ptr[blk], nn = P.bn_ptr[blk + 1] - nb0;
  for (int i0 = tid; i0 < nn; i0 += 4 * kAsmBlock) {   // 4 list entries per thread per pass: ids first, then all gathers
    int g[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int i = i0 + j * kAsmBlock; g[j] = (i < nn) ? P.bn_nodes[nb0 + i] : -1; }
    double2 q[4];
    double f_e3[4], f_gw[4], f_u[4], f_v[4], f_eh[4], f_ehi[4];
    unsigned char f_d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (g[j] < 0) continue;
      q[j] = P.xy[g[j]];
      if (STIFF == STIFF_EXPV) f_e3[j] = P.E3[g[j]];
      if (STIFF != STIFF_BASE) f_d[j] = P.isD[g[j]];
      if (MASS) f_gw[j] = P.gw[g[j]];
      if (NEED_UV) { f_u[j] = P.fu[g[j]]; f_v[j] = P.fv[g[j]]; f_ehi[j] = P.EHi[g[j]]; }
      if (NEED_EH) f_eh[j] = P.EH[g[j]];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (g[j] < 0) continue;
      const int i = i0 + j * kAsmBlock;
      sX[i] = q[j].x; sY[i] = q[j].y;
      sNode[i] = g[j];
      if (STIFF == STIFF_EXPV) sE3[i] = f_e3[j];
      if (STIFF != STIFF_BASE) sD[i] = f_d[j];
      if (MASS) sGW[i] = f_gw[j];
      if (NEED_UV) { sFU[i] = f_u[j]; sFV[i] = f_v[j]; sEHI[i] = f_ehi[j]; }
      if (NEED_EH) sEH[i] = f_eh[j];
    }
  }
  // ---- own-node data (coalesced) and per-triangle coefficients (level 2) ----
  const double2 q_own = P.xy[rr];
  const bool rowD = live && P.isD[rr] != 0;
  double own_g = 0.0, own_e3 = 0.0, own_u = 0.0, own_v = 0.0, own_eh = 0.0, own_ehi = 0.0;
  if (MASS) own_g = P.gw[rr];
  if (STIFF == STIFF_EXPV) own_e3 = P.E3[rr];
  if (NEED_UV) { own_u = P.fu[rr]; own_v = P.fv[rr]; own_ehi = P.EHi[rr]; }
  if (NEED_EH) own_eh = P.EH[rr];
  double cphx[kMaxIncReg], cphy[kMaxIncReg], cmid[kMaxIncReg];
#pragma unroll
  for (int k = 0; k < kMaxIncReg; ++k) {
    cphx[k] = 0.0; cphy[k] = 0.0; cmid[k] = 0.0;
    if (rec[k].x >= 0) {
      const int e = rec[k].x >> 2, a = rec[k].x & 3;
      if (STIFF != STIFF_BASE) cphx[k] = P.phx[e];
      if (STIFF == STIFF_COEF) cphy[k] = P.phy[e];
      if (RHS == RHS_DDP || RHS == RHS_SINH) cmid[k] = P.mid[3 * e + a];
    }
  }
  for (int k = 0; k < w; ++k) acc[k * kAsmBlock + tid] = 0.0;
  const int sdiag = live ? P.diag_slot[row] : 0;
  // Newton systems: base*x for the residual -- all values + columns of the row first, then all x gathers
  double mv = 0.0;
  if (STIFF == STIFF_BASE && RESID && in_pat) {
    for (int k0 = 0; k0 < w; k0 += 8) {
      double m[8];
      int c[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        m[j] = 0.0; c[j] = rr;
        if (k0 + j < w) { m[j] = ld_stream(P.base_val + base + (k0 + j) * 32 + lane); c[j] = ld_stream(P.col + base + (k0 + j) * 32 + lane); }
      }
      double xv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) xv[j] = (k0 + j < w) ? P.xvec[c[j]] : 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) if (k0 + j < w) mv += m[j] * xv[j];
    }
  }
  __syncthreads();

  double bsum = 0.0, lift = 0.0;
  auto one = [&](const int2 r, const double phx_in, const double phy_in, const double mid_in) {
    const int e = r.x >> 2, a = r.x & 3;
    const unsigned y = (unsigned)r.y;
    const int s_next = y & 31, s_prev = (y >> 5) & 31, lb = (y >> 10) & 2047, lc = y >> 21;
    // Own frame of the row: A = this row's corner, B = next, C = previous corner (cyclic).  Edge vectors opposite
    // each corner (src:243-245 up to the cyclic relabelling) and the signed area (src:246); every product is
    // written so that entry (i,j) computed for row i and entry (j,i) computed for row j use the same operands
    // in a commutative order (no 3-way selects on the corner index).
    const double bx = sX[lb], by = sY[lb], cx = sX[lc], cy = sY[lc];
    const double eax = bx - cx, eay = by - cy;            // e_A = q_B - q_C
    const double ebx = cx - q_own.x, eby = cy - q_own.y;  // e_B = q_C - q_A
    const double ecx = q_own.x - bx, ecy = q_own.y - by;  // e_C = q_A - q_B
    const double ar = (eax * eby - eay * ebx) * 0.5;
    double c_self = 0.0, c_next = 0.0, c_prev = 0.0;
    if (STIFF != STIFF_BASE) {
      double phx = phx_in, phy = phy_in;
      if (STIFF == STIFF_EXPV) { phx = phx_in * (own_e3 * sE3[lb] * sE3[lc]); phy = phx; }   // src:469
      const double inv4 = 1.0 / fabs(ar * 4.0);  // src:257 (one division per triangle; <= 1 ulp vs x/ar4)
      // K(i,j) = (e_i.x*phx*e_j.x + e_i.y*phy*e_j.y)/|4ar| (src:260-266)
      const double k_self = (phx * (eax * eax) + phy * (eay * eay)) * inv4;
      const double k_next = (phx * (eax * ebx) + phy * (eay * eby)) * inv4;
      const double k_prev = (phx * (eax * ecx) + phy * (eay * ecy)) * inv4;
      const bool bD = sD[lb] != 0, cD = sD[lc] != 0;
      // Dirichlet elimination src:271-278: rows and columns of D are zeroed (kept in the pattern)
      c_self = rowD ? 0.0 : k_self;
      c_next = (rowD || bD) ? 0.0 : k_next;
      c_prev = (rowD || cD) ? 0.0 : k_prev;
      // boundary lifting b -= K[:,D]*gD on rows outside D, src:362 (rare: rows next to a contact)
      if (!rowD) {
        if (bD) lift += k_next * P.gD[sNode[lb]];
        if (cD) lift += k_prev * P.gD[sNode[lc]];
      }
    }
    if (MASS) {
      // weighted mass with nodal weights W_k = g_k*ar/30, SIGNED ar (src:375-386, Q2); J0 is never eliminated (Q3).
      // diag = 3Wa+Wb+Wc, (a,b) = Wa+Wb+Wc/2: the pair sum first, so that (a,b) and (b,a) round identically
      const double ar30 = ar * (1.0 / 30.0);
      const double Wa = own_g * ar30, Wb = sGW[lb] * ar30, Wc = sGW[lc] * ar30;
      c_self -= 3.0 * Wa + (Wb + Wc);
      c_next -= (Wa + Wb) + Wc * 0.5;
      c_prev -= (Wa + Wc) + Wb * 0.5;
    }
    if (STIFF != STIFF_BASE || MASS) {
      acc[sdiag * kAsmBlock + tid] += c_self;
      acc[s_next * kAsmBlock + tid] += c_next;
      acc[s_prev * kAsmBlock + tid] += c_prev;
    }
    // lumped edge-midpoint load: corner a takes the midpoint between a and next (src:345-350, Q1)
    if (RHS != RHS_NONE && RHS != RHS_ZERO) {
      double f;
      if (RHS == RHS_DDP) {   // src:341-343
        const double ua = (own_u + sFU[lb]) * 0.5, va = (own_v + sFV[lb]) * 0.5;
        const double eS = own_eh * sEH[lb], eSi = own_ehi * sEHI[lb];
        f = P.par0 * (eSi * va - eS * ua) + mid_in;
      } else if (RHS == RHS_SRH) {  // src:517-519
        const double ua = (own_u + sFU[lb]) * 0.5, va = (own_v + sFV[lb]) * 0.5;
        const double eS = own_eh * sEH[lb], eSi = own_ehi * sEHI[lb];
        f = 1.0 / (P.par0 * (eS * ua + 1.0) + P.par1 * (eSi * va + 1.0));
      } else if (RHS == RHS_SINH) {  // registered functor, test:95
        const double S = (own_eh + sEH[lb]) * 0.5;
        f = mid_in + P.par0 * sinh(P.par1 * S);
      } else {  // RHS_POLY, README:107-108
        const double S = (own_eh + sEH[lb]) * 0.5;
        const int deg = (int)P.par0;
        f = P.mid_poly[deg][3 * e + a];
        for (int dd = deg - 1; dd >= 0; --dd) f = f * S + P.mid_poly[dd][3 * e + a];
      }
      bsum += f * (fabs(ar) * (1.0 / 3.0));  // src:349
    }
  };
#pragma unroll
  for (int k = 0; k < kMaxIncReg; ++k)
    if (rec[k].x >= 0) one(rec[k], cphx[k], cphy[k], cmid[k]);
  for (int k = kMaxIncReg; k < iw; ++k) {   // rows with more than kMaxIncReg triangles (rare)
    if (!live) break;
    const int2 r = P.inc[ibase + k * 32 + lane];
    if (r.x < 0) continue;
    const int e = r.x >> 2, a = r.x & 3;
    one(r, STIFF != STIFF_BASE ? P.phx[e] : 0.0, STIFF == STIFF_COEF ? P.phy[e] : 0.0,
        (RHS == RHS_DDP || RHS == RHS_SINH) ? P.mid[3 * e + a] : 0.0);
  }

  // ---- write the rows (coalesced per slice) ----
  if (!in_pat) return;
  if (STIFF != STIFF_BASE && rowD) acc[sdiag * kAsmBlock + tid] += 1.0;  // src:279-281
  double diag = 0.0;
  for (int k0 = 0; k0 < w; k0 += 8) {
    double m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {   // base values again (L2-resident from the pass above), all loads in flight together
      m[j] = 0.0;
      if (STIFF == STIFF_BASE && k0 + j < w) m[j] = P.base_val[base + (k0 + j) * 32 + lane];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      if (k < w) {
        const double a_k = acc[k * kAsmBlock + tid] + m[j];
        if (k == sdiag) diag = a_k;
        P.out_val[base + k * 32 + lane] = live ? a_k : 0.0;
      }
    }
  }
  if (live) {
    if (P.dinv) P.dinv[row] = 1.0 / diag;
    if (STIFF != STIFF_BASE && P.lift_out) P.lift_out[row] = lift;
    if (RHS != RHS_NONE) {
      if (STIFF == STIFF_BASE) lift = P.lift_in[row];
      double bi = bsum;
      if (P.nodal_const) bi += P.nodal_const[row];   // Neumann loads, src:994-996
      bi -= lift;                                    // src:362
      if (rowD) bi = P.gD[row];                      // src:363
      P.b[row] = bi;
      if (RESID) P.resid[row] = mv - bi;             // src:368
    }
  }
}

template <int STIFF, bool MASS, int RHS, bool RESID>
static int launch_asm(Context* ctx, const AsmArgs& a) {
  const Pattern& P = ctx->pat;
  if (P.nblocks == 0) return 0;
  constexpr bool NEED_UV = (RHS == RHS_DDP || RHS == RHS_SRH);
  constexpr bool NEED_EH = NEED_UV || RHS == RHS_SINH || RHS == RHS_POLY;
  const int nf = 2 + (STIFF == STIFF_EXPV ? 1 : 0) + (MASS ? 1 : 0) + (NEED_UV ? 3 : 0) + (NEED_EH ? 1 : 0);
  size_t smem = (size_t)P.max_row_len * kAsmBlock * sizeof(double) + (size_t)P.max_bn * (nf * sizeof(double) + sizeof(int) + 1);
  smem = (smem + 15) & ~(size_t)15;
  auto kern = k_assemble<STIFF, MASS, RHS, RESID>;
  // the opt-in limit is a property of the FUNCTION in the context, shared by every host thread: always raise it to the
  // device maximum (idempotent) -- setting the exact size per launch lets another context's thread (in-process ranks with
  // different meshes) lower it between this call and the launch
  if (smem > 48 * 1024) DDPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin));
  if (smem > (size_t)ctx->max_smem_optin) { ctx->fail(DDPS_ERR_ARG, "assembly block needs more shared memory than the device offers (node of very high degree?)"); return DDPS_ERR_ARG; }
  kern<<<P.nblocks, kAsmBlock, smem, ctx->stream>>>(a);
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
  k_spmv<false, false><<<grid, kVecBlock, 0, ctx->stream>>>(P.nrows, P.nslices, P.slice_ptr, P.col, val, x, y, nullptr, nullptr, none, 0);
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
  k_spmv<true, false><<<grid, kVecBlock, 0, ctx->stream>>>(P.nrows, P.nslices, P.slice_ptr, P.col, val, x, y, ctx->partials, ctx->scal, none, 0);
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
      if (stop) scal->e_ref = fmax(scal->e_ref, hs_tot);   // later corrections of the same Newton loop refer to the largest one
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
