// Fused FEM assembly of the hot path (K1-K7 of SURVEY section 2.1), hand-written for sm_100a, FP64, -fmad=false (the
// element formulas round like the reference's IEEE expressions).
//
// Row-major, one WARP per slice of 32 consecutive matrix rows, one thread per row; the four warps of a CTA are fully
// independent (no block barrier anywhere), so a warp waiting for memory never holds another one back.
//
//   level 0   one 48-byte descriptor per slice: where the slice's matrix block and incidence streams start, and the first
//             three RUNS of consecutive node ids of the slice's node list (a regular numbering has exactly three: the mesh
//             row above, the own one, the one below; further runs come from a list).
//   level 1   (a) one bulk asynchronous copy (`cp.async.bulk`, the TMA engine) per run brings the 48-byte nodal records
//             {gw, s1 | u, v | eh, ehi} (+ the 16-byte coordinates for the stiffness kernels) of the run into shared memory,
//             completion on the warp's own mbarrier: no registers, no LSU wavefronts, no per-node address arithmetic;
//             (b) meanwhile the per-(row, triangle) streams are read COALESCED: packed scatter slots + local node positions
//             (4 B) and, for the Newton systems, the signed area and the midpoint sample of the doping / source term in
//             incidence order (8 B each; permuted once per solve) -- no gather at all in the Newton kernel's triangle loop;
//             the stiffness kernels gather their per-triangle coefficient (mu, lambda^2).  The loads of the next three
//             triangles are always in flight.
//   compute   each thread walks its row's incident triangles in increasing triangle order (= the duplicate-summation order
//             of the reference's sparse()), reads the two other corners' records from shared memory with
//             conflict-free 16-byte loads and accumulates its row of every local 3x3 matrix: the diagonal in a register,
//             the off-diagonals in its PRIVATE shared-memory column (slots resolved at setup; an edge has at most two
//             triangles, so the first contribution is a plain store and only the second reads the column back).  No atomics,
//             no staging buffer in HBM, bitwise reproducible.
//   epilogue  coalesced row writes.  Newton systems: out = base - J0 (src:391) and, from the same base values, the residual
//             F = base*V - b (src:368) -- the neighbour values of V were dropped into a second private column during the
//             triangle walk, so neither the column indices nor a gather of V are read.
//
// Dirichlet elimination (src:271-284), boundary lifting (src:362), the lumped edge-midpoint load (src:345-350) are fused
// into the same pass; all transcendentals live in the nodal prep kernels (one pass over the nodes, which also writes the
// nodal records).
#include "ddps_internal.cuh"
#include "device_utils.cuh"

namespace ddps {

namespace {

constexpr int kWarps = kAsmBlock / 32;
#define DDPS_PRAGMA_(x) _Pragma(#x)
#define DDPS_PRAGMA(x) DDPS_PRAGMA_(x)
#ifndef DDPS_STIFF_UNROLL
#define DDPS_STIFF_UNROLL 1
#endif
#ifndef DDPS_STIFF_MINB
#define DDPS_STIFF_MINB 6   // resident CTAs per SM the stiffness kernels are compiled for (register cap 80)
#endif
constexpr int kRecB = kRecD * 8;   // bytes per nodal record (48): 16-byte multiples, so any run of records is a legal bulk copy

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// bulk asynchronous copy global -> shared (TMA engine), completion counted in bytes on the warp's mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p, unsigned bytes) {   // bytes: multiple of 16, p 16-byte aligned
  if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(p), "r"(bytes) : "memory");
}

struct SliceDesc { int base, ibase, nbase, w, iw, nn, lfirst; };
__device__ __forceinline__ SliceDesc decode_desc(const int4 d) {
  const unsigned dw = (unsigned)d.w;
  return {d.x, d.y, d.z, (int)(dw & 63), (int)((dw >> 6) & 63), (int)((dw >> 12) & 1023), (int)(dw >> 22)};
}

// The slice's nodal records (and coordinates) into shared memory, one bulk copy per run of consecutive nodes, one run per lane
template <bool XY>
__device__ __forceinline__ void stage_runs(const AsmArgs& P, const int4 d1, const int4 d2, int nn, int lane, unsigned rec_s,
                                           unsigned xy_s, unsigned bar, unsigned extra_bytes = 0) {
  if (lane == 0) mbar_expect_tx(bar, (unsigned)nn * (kRecB + (XY ? 16 : 0)) + extra_bytes);
  const int rb = d1.x, nr = d1.y;
  for (int r = lane; r < nr; r += 32) {
    int2 run;
    if (r == 0) run = make_int2(d1.z, d1.w);
    else if (r == 1) run = make_int2(d2.x, d2.y);
    else if (r == 2) run = make_int2(d2.z, d2.w);
    else run = __ldg(P.runs + rb + r);
    const unsigned pos = (unsigned)run.y & 0xffffu, len = (unsigned)run.y >> 16;
    bulk_g2s(rec_s + pos * kRecB, P.rec + (size_t)run.x * kRecD, len * kRecB, bar);
    if (XY) bulk_g2s(xy_s + pos * 16, P.xy + run.x, len * 16, bar);
  }
}

// The slice `pf_dist` slices ahead (about one wave of resident warps) is pulled towards L2 by bulk prefetches, one contiguous
// range per lane: its warp then finds the head of its dependent chain (records, first stream entries, base values) in L2.
template <bool NEWTON>
__device__ __forceinline__ void prefetch_slice(const AsmArgs& P, int slice2, const int4 d, int lane) {
  const SliceDesc F = decode_desc(d);
  const size_t row2 = (size_t)slice2 * 32;
  const bool full = slice2 + 1 < P.nslices;   // row-indexed ranges: the last slice is ragged
  if (lane == 0) prefetch_l2(P.inc_pack + F.ibase, F.iw * 128);
  if (lane == 1 && full) prefetch_l2(P.rec + row2 * kRecD, 32 * kRecB);
  if (NEWTON) {
    if (lane == 2) prefetch_l2(P.arinc + F.ibase, F.iw * 256);
    if (lane == 3 && P.cinc) prefetch_l2(P.cinc + F.ibase, F.iw * 256);
    if (lane == 4) prefetch_l2(P.base_val + F.base, F.w * 256);
    if (lane == 5 && full) prefetch_l2(P.lift_in + row2, 256);
  } else {
    if (lane == 2) prefetch_l2(P.inc_tri + F.ibase, F.iw * 128);
    if (lane == 3 && full) prefetch_l2(P.xy + row2, 512);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Newton systems (Vsolve src:368-391, solve_semlin_poisson src:1040-1062): A = base - J0(g), b(V), F = base*V - b
// Shared memory per warp: mbarrier (16 B) | records [snmax][48 B] | acc [wmax][32] | base values [wmax][32]
// ------------------------------------------------------------------------------------------------
template <int RHS>
__global__ void __launch_bounds__(kAsmBlock) k_asm_newton(const AsmArgs P) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int slice = blockIdx.x * kWarps + wib;
  if (slice >= P.nslices) return;
  double* wsm = sm + (size_t)wib * (2 + P.snmax * kRecD + 2 * P.wmax * 32);
  double* nodes = wsm + 2;
  double* acc = nodes + P.snmax * kRecD + lane;   // private column: acc[k*32]
  const double* bv = acc + P.wmax * 32;           // the row's base values (one bulk copy of the slice's block), bv[k*32]
  const unsigned bar = smem_u32(wsm);
  if (lane == 0) mbar_init(bar, 1);
  const int4* dp = P.sdesc + 3 * (size_t)slice;
  const int4 d0 = __ldg(dp), d1 = __ldg(dp + 1), d2 = __ldg(dp + 2);
  const SliceDesc D = decode_desc(d0);
  __syncwarp();
  stage_runs<false>(P, d1, d2, D.nn, lane, smem_u32(nodes), 0u, bar, (unsigned)D.w * 256);
  if (lane == 31) bulk_g2s(smem_u32(bv - lane), P.base_val + D.base, (unsigned)D.w * 256, bar);   // the slice's block of the base matrix
  // the slice's own streams towards L2 in one request each: the loop below then reads them with L2 latency
  if (lane == 30) prefetch_l2(P.arinc + D.ibase, (unsigned)D.iw * 256);
  if (lane == 29 && (RHS == RHS_DDP || RHS == RHS_SINH)) prefetch_l2(P.cinc + D.ibase, (unsigned)D.iw * 256);
  if (lane == 28) prefetch_l2(P.inc_pack + D.ibase, (unsigned)D.iw * 128);
  if (lane == 27 && slice + P.pf_desc < P.nslices) prefetch_l2(P.sdesc + 3 * (size_t)(slice + P.pf_desc), 48);

  const int row = slice * 32 + lane;
  const bool live = row < P.nrows;
  // ---- the row's streams (coalesced): a rolled loop with the loads of the next three triangles always in flight ----
  struct Inc { unsigned y; double cm, ar; };
  auto ld_inc = [&](const int k) {
    Inc r{0xffffffffu, 0.0, 0.0};
    if (k < D.iw) {
      const int idx = D.ibase + k * 32 + lane;
      r.y = __ldcs(P.inc_pack + idx);
      r.ar = __ldcs(P.arinc + idx);
      if (RHS == RHS_DDP || RHS == RHS_SINH) r.cm = __ldcs(P.cinc + idx);
    }
    return r;
  };
  Inc q0 = ld_inc(0), q1 = ld_inc(1), q2 = ld_inc(2);
  // per-row inputs of the epilogue, all independent of each other, requested now
  int sdiag = 0, len = 0;
  bool rowD = false;
  double b_fix = 0.0, lift_in = 0.0;
  if (live) {
    const unsigned ri = P.rowinfo[row];
    sdiag = ri & 31; len = ri >> 5;
    rowD = P.isD[row] != 0;
    lift_in = P.lift_in[row];
    if (P.nodal_const) b_fix = P.nodal_const[row];   // Neumann loads, src:994-996
  }
  const int slice2 = slice + P.pf_dist;
  int4 e0 = make_int4(0, 0, 0, 0);
  if (slice2 < P.nslices) e0 = __ldg(P.sdesc + 3 * (size_t)slice2);
  mbar_wait(bar, 0);

  const double* own = nodes + (D.lfirst + (live ? lane : 0)) * kRecD;
  const double2 o0 = *reinterpret_cast<const double2*>(own);
  const double2 o1 = *reinterpret_cast<const double2*>(own + 2);
  const double2 o2 = *reinterpret_cast<const double2*>(own + 4);
  const double own_g = o0.x, own_V = o0.y, own_u = o1.x, own_v = o1.y, own_eh = o2.x, own_ehi = o2.y;
  // the residual's base*V (src:368) is summed along the triangle walk: the diagonal term here, every off-diagonal term at
  // the FIRST visit of its slot, with the neighbour's V that is in flight anyway -- no column indices, no gather of V
  double mv = bv[sdiag * 32] * own_V;
  double bsum = 0.0, dacc = 0.0;   // the diagonal accumulates in a register (its slot is fixed per row)

  auto one = [&](const unsigned y, const double mid_in, const double a_r, const int idx) {
    const int s_next = y & 31, s_prev = (y >> 5) & 31, lb = (y >> 10) & 1023, lc = (y >> 20) & 1023;
    const double* B = nodes + lb * kRecD;
    const double* C = nodes + lc * kRecD;
    const double2 b0 = *reinterpret_cast<const double2*>(B);   // {g, V} of the next corner
    const double2 c0 = *reinterpret_cast<const double2*>(C);   // ... of the previous corner
    // weighted mass with nodal weights W_k = g_k*ar/30, SIGNED ar (src:375-386, Q2); J0 is never eliminated (Q3).
    // diag = 3Wa+Wb+Wc, (a,b) = Wa+Wb+Wc/2: the pair sum first, so that (a,b) and (b,a) round identically
    const double ar30 = a_r * (1.0 / 30.0);
    const double Wa = own_g * ar30, Wb = b0.x * ar30, Wc = c0.x * ar30;
    dacc -= 3.0 * Wa + (Wb + Wc);
    // an off-diagonal slot receives at most two contributions (the two triangles of its edge): the first one is a plain
    // store (the column is never zeroed), only the second one reads it back
    const double m_next = (Wa + Wb) + Wc * 0.5, m_prev = (Wa + Wc) + Wb * 0.5;
    double a_next = 0.0, a_prev = 0.0;
    if (y & 0x40000000u) a_next = acc[s_next * 32]; else mv += bv[s_next * 32] * b0.y;
    if (y & 0x80000000u) a_prev = acc[s_prev * 32]; else mv += bv[s_prev * 32] * c0.y;
    acc[s_next * 32] = a_next - m_next;
    acc[s_prev * 32] = a_prev - m_prev;
    // lumped edge-midpoint load: corner a takes the midpoint between a and next (src:345-350, Q1)
    double f;
    if (RHS == RHS_DDP) {   // src:341-343
      const double2 b1 = *reinterpret_cast<const double2*>(B + 2);   // {u, v}
      const double2 b2 = *reinterpret_cast<const double2*>(B + 4);   // {exp(V/2), exp(-V/2)}
      const double ua = (own_u + b1.x) * 0.5, va = (own_v + b1.y) * 0.5;
      const double eS = own_eh * b2.x, eSi = own_ehi * b2.y;
      f = P.par0 * (eSi * va - eS * ua) + mid_in;
    } else if (RHS == RHS_SINH) {  // registered functor, test:95
      const double S = (own_V + b0.y) * 0.5;
      f = mid_in + P.par0 * sinh(P.par1 * S);
    } else {  // RHS_POLY, README:107-108
      const double S = (own_V + b0.y) * 0.5;
      const int deg = (int)P.par0;
      f = P.cinc_poly[deg][idx];
      for (int dd = deg - 1; dd >= 0; --dd) f = f * S + P.cinc_poly[dd][idx];
    }
    bsum += f * (fabs(a_r) * (1.0 / 3.0));  // src:349
  };
#pragma unroll 3
  for (int k = 0; k < D.iw; ++k) {
    const Inc c = q0;
    q0 = q1; q1 = q2; q2 = ld_inc(k + 3);
    if (c.y != 0xffffffffu) one(c.y, c.cm, c.ar, D.ibase + k * 32 + lane);
  }

  // ---- the rows (coalesced per slice): out = base - J0 ----
  double diag = 0.0;
  for (int k = 0; k < D.w; ++k) {
    double a_k = 0.0;   // padding entries of the slice (k >= len) stay 0
    if (k < len) a_k = ((k == sdiag) ? dacc : acc[k * 32]) + bv[k * 32];
    if (k == sdiag) diag = a_k;
    P.out_val[D.base + k * 32 + lane] = a_k;
  }
  if (live) {
    if (P.dinv) P.dinv[row] = 1.0 / diag;
    double bi = (bsum + b_fix) - lift_in;            // src:362 (b_fix = 0 without Neumann loads)
    if (rowD) bi = P.gD[row];                        // src:363
    P.b[row] = bi;
    P.resid[row] = mv - bi;                          // src:368
  }
  if (slice2 < P.nslices) prefetch_slice<true>(P, slice2, e0, lane);
}

// ------------------------------------------------------------------------------------------------
// Stiffness systems: MV_assemble (src:229-287, STIFF_COEF), uvsolve (src:433-588, STIFF_EXPV [+ SRH mass and load])
// Shared memory per warp: mbarrier (16 B) | records [snmax][48 B] | coordinates [snmax][16 B] | acc [wmax][32]
// ------------------------------------------------------------------------------------------------
template <int STIFF, bool MASS, int RHS>
__global__ void __launch_bounds__(kAsmBlock, DDPS_STIFF_MINB) k_asm_stiff(const AsmArgs P) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int slice = blockIdx.x * kWarps + wib;
  if (slice >= P.nslices) return;
  double* wsm = sm + (size_t)wib * (2 + P.snmax * (kRecD + 2) + P.wmax * 32);
  double* nodes = wsm + 2;
  const double2* sxy = reinterpret_cast<const double2*>(nodes + P.snmax * kRecD);
  double* acc = nodes + P.snmax * (kRecD + 2) + lane;
  const unsigned bar = smem_u32(wsm);
  if (lane == 0) mbar_init(bar, 1);
  const int4* dp = P.sdesc + 3 * (size_t)slice;
  const int4 d0 = __ldg(dp), d1 = __ldg(dp + 1), d2 = __ldg(dp + 2);
  const SliceDesc D = decode_desc(d0);
  __syncwarp();
  stage_runs<true>(P, d1, d2, D.nn, lane, smem_u32(nodes), smem_u32(sxy), bar);

  const int row = slice * 32 + lane;
  const bool live = row < P.nrows;
  // ---- the row's streams (coalesced) and per-triangle coefficients (gathers): a rolled loop with the records of the next
  //      three triangles and the coefficients of the next one always in flight ----
  struct Inc { int tri; unsigned y; };
  struct Coef { double x, y; };
  auto ld_inc = [&](const int k) {
    Inc r{-1, 0u};
    if (k < D.iw) {
      const int idx = D.ibase + k * 32 + lane;
      r.tri = __ldcs(P.inc_tri + idx);
      r.y = __ldcs(P.inc_pack + idx);
    }
    return r;
  };
  auto ld_coef = [&](const Inc& q) {
    Coef c{0.0, 0.0};
    if (q.tri >= 0) {
      c.x = __ldg(P.phx + (q.tri >> 2));
      if (STIFF == STIFF_COEF) c.y = __ldg(P.phy + (q.tri >> 2));
    }
    return c;
  };
  Inc q0 = ld_inc(0), q1 = ld_inc(1), q2 = ld_inc(2);
  int sdiag = 0, len = 0;
  double b_D = 0.0;
  if (live) {
    const unsigned ri = P.rowinfo[row];
    sdiag = ri & 31; len = ri >> 5;
    if (RHS != RHS_NONE) b_D = P.gD[row];
  }
  const int slice2 = slice + P.pf_dist;
  int4 e0 = make_int4(0, 0, 0, 0);
  if (slice2 < P.nslices) e0 = __ldg(P.sdesc + 3 * (size_t)slice2);
  Coef c0 = ld_coef(q0);
  mbar_wait(bar, 0);

  const int lown = D.lfirst + (live ? lane : 0);
  const double* own = nodes + lown * kRecD;
  const double2 o0 = *reinterpret_cast<const double2*>(own);       // {g, +-e3}: the sign carries the Dirichlet flag
  const double2 o1 = *reinterpret_cast<const double2*>(own + 2);   // {u, v}
  const double2 o2 = *reinterpret_cast<const double2*>(own + 4);   // {exp(V/2), exp(-V/2)}
  const double2 q_own = sxy[lown];
  const bool rowD = live && o0.y < 0.0;
  const double own_e3 = fabs(o0.y);
  double bsum = 0.0, lift = 0.0, dacc = 0.0;   // the diagonal accumulates in a register (its slot is fixed per row)

  auto one = [&](const unsigned y, const double phx_in, const double phy_in) {
    const int s_next = y & 31, s_prev = (y >> 5) & 31, lb = (y >> 10) & 1023, lc = (y >> 20) & 1023;
    const double* B = nodes + lb * kRecD;
    const double* C = nodes + lc * kRecD;
    const double2 b0 = *reinterpret_cast<const double2*>(B), qb = sxy[lb];
    const double2 c0 = *reinterpret_cast<const double2*>(C), qc = sxy[lc];
    // Own frame of the row: A = this row's corner, B = next, C = previous corner (cyclic).  Edge vectors opposite
    // each corner (src:243-245 up to the cyclic relabelling) and the signed area (src:246); every product is
    // written so that entry (i,j) computed for row i and entry (j,i) computed for row j use the same operands
    // in a commutative order (no 3-way selects on the corner index).
    const double eax = qb.x - qc.x, eay = qb.y - qc.y;        // e_A = q_B - q_C
    const double ebx = qc.x - q_own.x, eby = qc.y - q_own.y;  // e_B = q_C - q_A
    const double ecx = q_own.x - qb.x, ecy = q_own.y - qb.y;  // e_C = q_A - q_B
    const double ar = (eax * eby - eay * ebx) * 0.5;
    double phx = phx_in, phy = phy_in;
    if (STIFF == STIFF_EXPV) { phx = phx_in * (own_e3 * fabs(b0.y) * fabs(c0.y)); phy = phx; }   // src:469
    const double inv4 = 1.0 / fabs(ar * 4.0);  // src:257 (one division per triangle; <= 1 ulp vs x/ar4)
    // K(i,j) = (e_i.x*phx*e_j.x + e_i.y*phy*e_j.y)/|4ar| (src:260-266)
    const double k_self = (phx * (eax * eax) + phy * (eay * eay)) * inv4;
    const double k_next = (phx * (eax * ebx) + phy * (eay * eby)) * inv4;
    const double k_prev = (phx * (eax * ecx) + phy * (eay * ecy)) * inv4;
    const bool bD = b0.y < 0.0, cD = c0.y < 0.0;
    // Dirichlet elimination src:271-278: rows and columns of D are zeroed (kept in the pattern)
    double c_self = rowD ? 0.0 : k_self;
    double c_next = (rowD || bD) ? 0.0 : k_next;
    double c_prev = (rowD || cD) ? 0.0 : k_prev;
    // boundary lifting b -= K[:,D]*gD on rows outside D, src:362 (rare: rows next to a contact)
    if (!rowD) {
      if (bD) lift += k_next * P.gD[P.sn_nodes[D.nbase + lb]];
      if (cD) lift += k_prev * P.gD[P.sn_nodes[D.nbase + lc]];
    }
    if (MASS) {
      // weighted mass with nodal weights W_k = g_k*ar/30, SIGNED ar (src:375-386, Q2); J0 is never eliminated (Q3).
      const double ar30 = ar * (1.0 / 30.0);
      const double Wa = o0.x * ar30, Wb = b0.x * ar30, Wc = c0.x * ar30;
      c_self -= 3.0 * Wa + (Wb + Wc);
      c_next -= (Wa + Wb) + Wc * 0.5;
      c_prev -= (Wa + Wc) + Wb * 0.5;
    }
    dacc += c_self;
    // an off-diagonal slot receives at most two contributions (the two triangles of its edge): the first one is a plain
    // store (the column is never zeroed), only the second one reads it back
    double a_next = 0.0, a_prev = 0.0;
    if (y & 0x40000000u) a_next = acc[s_next * 32];
    if (y & 0x80000000u) a_prev = acc[s_prev * 32];
    acc[s_next * 32] = a_next + c_next;
    acc[s_prev * 32] = a_prev + c_prev;
    if (RHS == RHS_SRH) {  // src:517-519; corner a takes the midpoint between a and next (src:345-350, Q1)
      const double2 b1 = *reinterpret_cast<const double2*>(B + 2);
      const double2 b2 = *reinterpret_cast<const double2*>(B + 4);
      const double ua = (o1.x + b1.x) * 0.5, va = (o1.y + b1.y) * 0.5;
      const double eS = o2.x * b2.x, eSi = o2.y * b2.y;
      const double f = 1.0 / (P.par0 * (eS * ua + 1.0) + P.par1 * (eSi * va + 1.0));
      bsum += f * (fabs(ar) * (1.0 / 3.0));  // src:349
    }
  };
  DDPS_PRAGMA(unroll DDPS_STIFF_UNROLL)
  for (int k = 0; k < D.iw; ++k) {
    const Inc q = q0;
    const Coef c = c0;
    q0 = q1; q1 = q2; q2 = ld_inc(k + 3);
    c0 = ld_coef(q0);
    if (q.tri >= 0) one(q.y, c.x, c.y);
  }

  // ---- write the rows (coalesced per slice) ----
  if (rowD) dacc += 1.0;  // src:279-281
  for (int k = 0; k < D.w; ++k) {
    double a_k = 0.0;   // padding entries of the slice (k >= len) stay 0
    if (k < len) a_k = (k == sdiag) ? dacc : acc[k * 32];
    P.out_val[D.base + k * 32 + lane] = a_k;
  }
  if (live) {
    if (P.dinv) P.dinv[row] = 1.0 / dacc;
    if (P.lift_out) P.lift_out[row] = lift;
    if (RHS != RHS_NONE) {
      double bi = bsum - lift;                       // src:362
      if (rowD) bi = b_D;                            // src:363
      P.b[row] = bi;
    }
  }
  if (slice2 < P.nslices) prefetch_slice<false>(P, slice2, e0, lane);
}

namespace {

template <class K>
int launch_slices(Context* ctx, K kern, const AsmArgs& a, size_t warp_doubles) {
  if (a.nslices == 0) return 0;
  const size_t smem = warp_doubles * sizeof(double) * kWarps;
  if (smem > (size_t)ctx->max_smem_optin) {
    ctx->fail(DDPS_ERR_ARG, "an assembly slice needs more shared memory than the device offers (node of very high degree?)");
    return DDPS_ERR_ARG;
  }
  // the opt-in limit is a property of the FUNCTION in the context, shared by every host thread: always raise it to the device
  // maximum (idempotent) -- setting the exact size per launch lets another context's thread (in-process ranks with different
  // meshes) lower it between this call and the launch
  DDPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin));
  DDPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  kern<<<(a.nslices + kWarps - 1) / kWarps, kAsmBlock, smem, ctx->stream>>>(a);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

size_t newton_warp_doubles(const AsmArgs& a) { return 2 + (size_t)a.snmax * kRecD + 2 * (size_t)a.wmax * 32; }
size_t stiff_warp_doubles(const AsmArgs& a) { return 2 + (size_t)a.snmax * (kRecD + 2) + (size_t)a.wmax * 32; }

}  // namespace

int launch_assemble_base(Context* ctx, const AsmArgs& a) {
  return launch_slices(ctx, k_asm_stiff<STIFF_COEF, false, RHS_NONE>, a, stiff_warp_doubles(a));
}

int launch_assemble_uv(Context* ctx, const AsmArgs& a, int rhs_kind, bool mass) {
  if (rhs_kind == RHS_SRH && mass) return launch_slices(ctx, k_asm_stiff<STIFF_EXPV, true, RHS_SRH>, a, stiff_warp_doubles(a));
  if (rhs_kind == RHS_ZERO && !mass) return launch_slices(ctx, k_asm_stiff<STIFF_EXPV, false, RHS_ZERO>, a, stiff_warp_doubles(a));
  ctx->fail(DDPS_ERR_ARG, "unsupported uv assembly mode");
  return DDPS_ERR_ARG;
}

int launch_assemble_newton(Context* ctx, const AsmArgs& a, int rhs_kind) {
  switch (rhs_kind) {
    case RHS_DDP: return launch_slices(ctx, k_asm_newton<RHS_DDP>, a, newton_warp_doubles(a));
    case RHS_SINH: return launch_slices(ctx, k_asm_newton<RHS_SINH>, a, newton_warp_doubles(a));
    case RHS_POLY: return launch_slices(ctx, k_asm_newton<RHS_POLY>, a, newton_warp_doubles(a));
  }
  ctx->fail(DDPS_ERR_ARG, "unsupported Newton rhs functor");
  return DDPS_ERR_ARG;
}


// ------------------------------------------------------------------------------------------------
// Nodal prep kernels: all transcendental work of an assembly in one pass over the nodes; they write the 48-byte nodal
// records {gw, s1 | u, v | eh, ehi} the assembly warps gather (s1 = the iterate for the Newton systems, +-exp(V/3) for the
// continuity systems and +-1 for the base stiffness, the sign carrying the Dirichlet flag)
// ------------------------------------------------------------------------------------------------
namespace {

// The records of a block's 256 consecutive nodes are one contiguous 12 KB piece of memory: they go through shared memory so
// that the stores are coalesced 16-byte pieces (a thread storing its own three pieces writes 32 scattered sectors per request).
constexpr int kPrepBlock = 256;
__device__ __forceinline__ void put_rec(double* __restrict__ rec, int n, double gw, double s1, double u, double v,
                                        double eh, double ehi) {
  __shared__ double2 sh[kPrepBlock * 3];
  const int t = threadIdx.x;
  sh[3 * t] = make_double2(gw, s1);
  sh[3 * t + 1] = make_double2(u, v);
  sh[3 * t + 2] = make_double2(eh, ehi);
  __syncthreads();
  const size_t first = (size_t)blockIdx.x * kPrepBlock;               // first node of the block
  const int cnt = (int)min((size_t)kPrepBlock, (size_t)n - first) * 3;   // 16-byte pieces to store
  double2* out = reinterpret_cast<double2*>(rec + first * kRecD);
  for (int i = t; i < cnt; i += kPrepBlock) out[i] = sh[i];
}

__global__ void __launch_bounds__(kPrepBlock) k_prep_base(int n, const unsigned char* __restrict__ isD, double* __restrict__ rec) {
  const int i = min(blockIdx.x * blockDim.x + threadIdx.x, n - 1);   // (the tail threads recompute the last node; not stored)
  put_rec(rec, n, 0.0, isD[i] ? -1.0 : 1.0, 0.0, 0.0, 0.0, 0.0);
}

__global__ void __launch_bounds__(kPrepBlock) k_prep_poisson(int n, const double* __restrict__ V, const double* __restrict__ u,
                                                             const double* __restrict__ v, double delta2, double* __restrict__ rec) {
  const int i = min(blockIdx.x * blockDim.x + threadIdx.x, n - 1);
  const double Vi = V[i], ui = u[i], vi = v[i];
  const double gw = -delta2 * (ui * exp(Vi) + vi * exp(-Vi));  // src:375-377
  put_rec(rec, n, gw, Vi, ui, vi, exp(Vi / 2.0), exp(-Vi / 2.0));
}

// sign = +1: u-solve with (V, tau_p, tau_n, u0, v0); sign = -1: v-solve, the reference passes
// (-V, tau_n, tau_p, v0, u0) (src:598-599).  V here is always the true potential.
__global__ void __launch_bounds__(kPrepBlock) k_prep_uv(int n, double sign, const double* __restrict__ V, const double* __restrict__ u0,
                                                        const double* __restrict__ v0, double tau_p, double tau_n, int shockley,
                                                        const unsigned char* __restrict__ isD, double* __restrict__ rec) {
  const int i = min(blockIdx.x * blockDim.x + threadIdx.x, n - 1);
  const double Vs = sign * V[i];   // the "V" the reference's uvsolve sees
  const double a = (sign > 0) ? u0[i] : v0[i];   // callee's u0
  const double b = (sign > 0) ? v0[i] : u0[i];   // callee's v0
  const double e3 = exp(Vs / 3.0);               // exp((V1+V2+V3)/3) = product of the three nodal factors, src:469
  double gw = 0.0;
  if (shockley) {
    // src:568: g_k = -v0_k / (tau_p(e^{V}u0+1) + tau_n(e^{-V}v0+1)) in the callee's own naming
    const double tp = (sign > 0) ? tau_p : tau_n, tn = (sign > 0) ? tau_n : tau_p;
    gw = -b / (tp * (exp(Vs) * a + 1.0) + tn * (exp(-Vs) * b + 1.0));
  }
  put_rec(rec, n, gw, isD[i] ? -e3 : e3, a, b, exp(Vs / 2.0), exp(-Vs / 2.0));
}

struct PolyNodal { const double* c[kMaxPoly]; };

__global__ void __launch_bounds__(kPrepBlock) k_prep_semlin(int n, const double* __restrict__ V, int kind, double p0, double p1, PolyNodal pn,
                                                            double* __restrict__ rec) {
  const int i = min(blockIdx.x * blockDim.x + threadIdx.x, n - 1);
  const double u = V[i];
  double g;
  if (kind == DDPS_FUNCTOR_SINH) {
    g = p0 * p1 * cosh(p1 * u);   // g = df/du, test:96
  } else {
    const int deg = (int)p0;
    g = 0.0;
    for (int d = deg; d >= 1; --d) g = g * u + d * pn.c[d][i];
  }
  put_rec(rec, n, g, u, 0.0, 0.0, 0.0, 0.0);   // src:1046-1048
}

__global__ void k_permute_mid(int64_t n, const int* __restrict__ inc_tri, const double* __restrict__ mid,
                              double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = inc_tri[i];
  out[i] = t >= 0 ? mid[3 * (int64_t)(t >> 2) + (t & 3)] : 0.0;
}

}  // namespace

int launch_prep_base(Context* ctx) {
  int n = ctx->n_loc;
  if (!n) return 0;
  k_prep_base<<<(n + kPrepBlock - 1) / kPrepBlock, kPrepBlock, 0, ctx->stream>>>(n, ctx->isD, ctx->rec);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

int launch_prep_poisson(Context* ctx, const double* V, const double* u, const double* v, double delta2) {
  int n = ctx->n_loc;
  if (!n) return 0;
  k_prep_poisson<<<(n + kPrepBlock - 1) / kPrepBlock, kPrepBlock, 0, ctx->stream>>>(n, V, u, v, delta2, ctx->rec);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

int launch_prep_uv(Context* ctx, double sign, const double* V, const double* u0, const double* v0, double tau_p,
                   double tau_n, int rec_mode) {
  int n = ctx->n_loc;
  if (!n) return 0;
  k_prep_uv<<<(n + kPrepBlock - 1) / kPrepBlock, kPrepBlock, 0, ctx->stream>>>(n, sign, V, u0, v0, tau_p, tau_n, rec_mode == DDPS_REC_SHOCKLEY,
                                                    ctx->isD, ctx->rec);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

int launch_prep_semlin(Context* ctx, const double* V) {
  int n = ctx->n_loc;
  if (!n) return 0;
  PolyNodal pn;
  for (int k = 0; k < kMaxPoly; ++k) pn.c[k] = ctx->coeff[DDPS_COEFF_POLY_NODE0 + k];
  k_prep_semlin<<<(n + kPrepBlock - 1) / kPrepBlock, kPrepBlock, 0, ctx->stream>>>(n, V, ctx->functor_kind, ctx->functor_par[0],
                                                        ctx->functor_par[1], pn, ctx->rec);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

int launch_permute_mid(Context* ctx, const double* mid, double* out) {
  const int64_t n = ctx->pat.inc_entries;
  if (!n) return 0;
  k_permute_mid<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, ctx->pat.inc_tri, mid, out);
  ctx->stats.kernel_launches++;
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ddps
