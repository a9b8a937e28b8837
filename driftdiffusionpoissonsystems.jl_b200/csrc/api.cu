// C ABI (include/ddps.h) + the device-resident Gummel / Newton drivers.
//
//   ddps_ddpe_begin/step  <->  solve_ddpe's loop + G + Vsolve + uvsolve (src:630-658, 596-601, 292-425, 433-589)
//   ddps_solve_semlin     <->  solve_semlin_poisson (src:895-1097)
//
// Host code here only sequences kernels and reads back scalars (residual norms, iteration counts);
// every array stays on the device between iterations.
#include <chrono>
#include <cstdio>
#include <map>

#include "ddps_internal.cuh"

using namespace ddps;

namespace {

thread_local std::string g_create_err;

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

template <typename T>
int dev_alloc(Context* ctx, T** p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  DDPS_CUDA(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
  return 0;
}

template <typename T>
void dev_free(T*& p) { if (p) { cudaFree(p); p = nullptr; } }

void free_mesh(Context* ctx) {
  amg_free(ctx);
  p2p_teardown(ctx);
  free_pattern(ctx);
  dev_free(ctx->xy); dev_free(ctx->elem); dev_free(ctx->d_own_glob);
  dev_free(ctx->isD);
  for (int f = 0; f < 4; ++f) { dev_free(ctx->gD[f]); ctx->have_D[f] = false; }
  for (int s = 0; s < 64; ++s) { dev_free(ctx->coeff[s]); ctx->coeff_len[s] = 0; }
  dev_free(ctx->val_base); dev_free(ctx->val_op); dev_free(ctx->lift); dev_free(ctx->val_s); dev_free(ctx->sc); dev_free(ctx->bs);
  ctx->last_val = nullptr; ctx->last_dinv = nullptr;
  dev_free(ctx->V); dev_free(ctx->U); dev_free(ctx->W); dev_free(ctx->V0); dev_free(ctx->U0); dev_free(ctx->W0);
  dev_free(ctx->Unew); dev_free(ctx->rec); dev_free(ctx->cinc);
  for (int k = 0; k < kMaxPoly; ++k) dev_free(ctx->cinc_poly[k]);
  dev_free(ctx->b); dev_free(ctx->resid); dev_free(ctx->dinv); dev_free(ctx->x); dev_free(ctx->r); dev_free(ctx->q);
  dev_free(ctx->p); dev_free(ctx->partials);
  for (int k = 0; k < Context::kNewtonWarm; ++k) { dev_free(ctx->newton_x[k]); ctx->newton_x_valid[k] = false; }
  plan_free(ctx->halo);
  ctx->loc2glob.clear(); ctx->elem_glob.clear(); ctx->Dset.clear();
  ctx->have_mesh = false; ctx->base_ready = false; ctx->ddpe_active = false; ctx->local_ingest = false;
}

__global__ void k_count_negative(const int4* __restrict__ elem, const double2* __restrict__ xy, int nme, int* out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nme) return;
  int4 el = elem[e];
  double2 q1 = xy[el.x], q2 = xy[el.y], q3 = xy[el.z];
  double ar = ((q2.x - q3.x) * (q3.y - q1.y) - (q2.y - q3.y) * (q3.x - q1.x)) / 2.0;
  if (ar < 0.0) atomicAdd(out, 1);
}

// Mesh.elements (int64, 1-based, 5 x nme column-major, src:17) -> int4 {n1,n2,n3,tag} 0-based; bad[0] = 1 + index of a
// triangle that references a node outside 1..nq (0: none)
__global__ void k_convert_elements(const long long* __restrict__ el, int nme, long long nq, int4* __restrict__ out,
                                   int* __restrict__ bad) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nme) return;
  const long long* c = el + 5LL * e;
  const long long a = c[0] - 1, b = c[1] - 1, d = c[2] - 1;
  if (a < 0 || b < 0 || d < 0 || a >= nq || b >= nq || d >= nq) { atomicMax(bad, e + 1); out[e] = make_int4(0, 0, 0, 0); return; }
  out[e] = make_int4((int)a, (int)b, (int)d, (int)c[4]);
}

// Dirichlet data: flags + values scattered from the (node, value) list
__global__ void k_scatter_dirichlet(int nd, const long long* __restrict__ ids, const double* __restrict__ vals,
                                    unsigned char* __restrict__ isD, double* __restrict__ gD) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nd) return;
  isD[ids[i]] = 1;
  gD[ids[i]] = vals[i];
}

__global__ void k_check_finite(size_t n, const double* __restrict__ a, int* __restrict__ flag) {
  bool bad = false;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) bad = bad || !isfinite(a[i]);
  if (bad) *flag = 1;
}

// registered spatial functors (include/ddps.h: ddps_spatial_kind), evaluated where the reference evaluates its closures
struct SpatialFn { int kind; int ntag; double p[16]; };
__device__ __forceinline__ double eval_spatial(const SpatialFn& f, double x, double y, int tag) {
  switch (f.kind) {
    case DDPS_SPATIAL_CONST: return f.p[0];
    case DDPS_SPATIAL_AFFINE: return (f.p[0] + f.p[1] * x) + f.p[2] * y;
    case DDPS_SPATIAL_STEP_X: return x > f.p[0] ? f.p[2] : f.p[1];
    case DDPS_SPATIAL_STEP_Y: return y > f.p[0] ? f.p[2] : f.p[1];
    default: {
      double v = f.p[0];
      for (int k = 0; k < f.ntag; ++k) if ((double)tag == f.p[1 + 2 * k]) v = f.p[2 + 2 * k];
      return v;
    }
  }
}
// where = 0: centroids ((x1+x2+x3)/3, src:250-251) -> out[nme]; 1: edge midpoints m12,m23,m31 ((xa+xb)/2, src:303-308) ->
// out[3*nme]; square: store the square (lambda^2, src:254)
__global__ void k_sample_elements(int nme, const int4* __restrict__ elem, const double2* __restrict__ xy, SpatialFn f, int where,
                                  int square, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nme) return;
  const int4 el = elem[e];
  const double2 q1 = xy[el.x], q2 = xy[el.y], q3 = xy[el.z];
  if (where == 0) {
    const double v = eval_spatial(f, (q1.x + q2.x + q3.x) / 3.0, (q1.y + q2.y + q3.y) / 3.0, el.w);
    out[e] = square ? v * v : v;
  } else {
    out[3 * (size_t)e + 0] = eval_spatial(f, (q1.x + q2.x) / 2.0, (q1.y + q2.y) / 2.0, el.w);
    out[3 * (size_t)e + 1] = eval_spatial(f, (q2.x + q3.x) / 2.0, (q2.y + q3.y) / 2.0, el.w);
    out[3 * (size_t)e + 2] = eval_spatial(f, (q3.x + q1.x) / 2.0, (q3.y + q1.y) / 2.0, el.w);
  }
}
__global__ void k_sample_nodes(int n, const double2* __restrict__ xy, SpatialFn f, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = eval_spatial(f, xy[i].x, xy[i].y, 0);
}

__global__ void k_scatter_global(int n, const long long* __restrict__ glob, const double* __restrict__ src,
                                 double* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[glob[i]] = src[i];
}

// upload a host array through the stream (pageable source: the copy is staged by the driver)
template <typename T>
int upload(Context* ctx, T* dst, const T* src, size_t n) {
  if (!n) return 0;
  DDPS_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

int require_mesh(Context* ctx) {
  if (!ctx->have_mesh) { ctx->fail(DDPS_ERR_STATE, "no mesh: call ddps_mesh_set first"); return DDPS_ERR_STATE; }
  return 0;
}

// Fill a local [n_loc] array from a global [nq] host array.
int upload_nodal(Context* ctx, double* dst, const double* src_global) {
  if (ctx->loc2glob.empty()) return upload(ctx, dst, src_global, (size_t)ctx->n_loc);
  std::vector<double> tmp(ctx->n_loc);
  for (int i = 0; i < ctx->n_loc; ++i) tmp[i] = src_global[ctx->loc2glob[i]];
  int rc = upload(ctx, dst, tmp.data(), tmp.size());
  if (rc) return rc;
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int download_owned(Context* ctx, const double* src, double* dst_user);

// Copy the owned part of a local field into a global [nq] host array (all ranks get the full field).
int download_nodal(Context* ctx, const double* src_local, double* dst_global) {
  if (ctx->nranks == 1 && !ctx->loc2glob.empty()) return download_owned(ctx, src_local, dst_global);
  if (ctx->nranks == 1) {
    DDPS_CUDA(cudaMemcpyAsync(dst_global, src_local, (size_t)ctx->n_own * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  double* g = nullptr;
  DDPS_CUDA(cudaMalloc(&g, (size_t)ctx->nq_global * sizeof(double)));
  cudaMemsetAsync(g, 0, (size_t)ctx->nq_global * sizeof(double), ctx->stream);
  if (ctx->n_own) k_scatter_global<<<(ctx->n_own + 255) / 256, 256, 0, ctx->stream>>>(ctx->n_own, ctx->d_own_glob, src_local, g);
  int rc = 0;
  const int64_t chunk = 1 << 28;
  for (int64_t o = 0; o < ctx->nq_global && !rc; o += chunk)
    rc = allreduce_sum(ctx, g + o, (int)std::min<int64_t>(chunk, ctx->nq_global - o));
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(dst_global, g, (size_t)ctx->nq_global * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->fail(DDPS_ERR_CUDA, cudaGetErrorString(e)); rc = DDPS_ERR_CUDA; }
  }
  cudaFree(g);
  return rc;
}


// ---- node partition + local mesh extraction (host code; also behind ddps_partition_plan) -------
struct LocalMesh {
  int n_own = 0, n_loc = 0;
  std::vector<int64_t> loc2glob, elem_glob;
  std::vector<int4> elem;
  std::vector<int> peers, send_off, recv_off, send_idx;
};

// Rank `me` owns the nodes RCB assigns to it and keeps every triangle that touches an owned node
// (one redundantly-assembled ghost layer => no communication in assembly).  Local numbering:
// owned nodes in increasing global id, then ghosts sorted by (owner, global id); triangles in
// increasing global id (so every row sums its triangles in the same order as on one GPU).
// Send list to peer q = my owned corners of triangles that touch a q-owned node, sorted by global id
// -- exactly q's ghost block for me, in the same order, so no setup communication is needed.
int build_local_mesh(const double* nodes, int64_t nq, const int64_t* elements, int64_t nme, int nparts, int me,
                     LocalMesh& L, std::string& err) {
  std::vector<int32_t> owner(nq);
  partition_nodes_rcb(nodes, nq, nparts, owner.data());
  std::vector<int32_t> g2l(nq, -1);
  std::vector<std::pair<int, int64_t>> ghosts, sends;   // (peer, global id)
  for (int64_t e = 0; e < nme; ++e) {
    const int64_t* c = elements + 5 * e;
    int64_t g[3] = {c[0] - 1, c[1] - 1, c[2] - 1};
    for (int k = 0; k < 3; ++k)
      if (g[k] < 0 || g[k] >= nq) { err = "element " + std::to_string(e + 1) + " references a node outside 1..nq"; return -1; }
    int o[3] = {owner[g[0]], owner[g[1]], owner[g[2]]};
    if (o[0] != me && o[1] != me && o[2] != me) continue;
    L.elem_glob.push_back(e);
    for (int k = 0; k < 3; ++k) {
      if (o[k] != me) ghosts.emplace_back(o[k], g[k]);
      else
        for (int j = 0; j < 3; ++j)
          if (o[j] != me) sends.emplace_back(o[j], g[k]);
    }
  }
  auto uniq = [](std::vector<std::pair<int, int64_t>>& v) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
  };
  uniq(ghosts); uniq(sends);
  for (int64_t i = 0; i < nq; ++i) if (owner[i] == me) ++L.n_own;
  L.n_loc = L.n_own + (int)ghosts.size();
  L.loc2glob.resize(L.n_loc);
  int k = 0;
  for (int64_t i = 0; i < nq; ++i) if (owner[i] == me) { g2l[i] = k; L.loc2glob[k] = i; ++k; }
  for (auto& pr : ghosts) { g2l[pr.second] = k; L.loc2glob[k] = pr.second; ++k; }
  for (auto& pr : ghosts) L.peers.push_back(pr.first);
  for (auto& pr : sends) L.peers.push_back(pr.first);
  std::sort(L.peers.begin(), L.peers.end());
  L.peers.erase(std::unique(L.peers.begin(), L.peers.end()), L.peers.end());
  const int nn = (int)L.peers.size();
  L.send_off.assign(nn + 1, 0); L.recv_off.assign(nn + 1, 0);
  for (int j = 0; j < nn; ++j) {
    for (auto& pr : sends) if (pr.first == L.peers[j]) L.send_idx.push_back(g2l[pr.second]);
    L.send_off[j + 1] = (int)L.send_idx.size();
    int cnt = 0;
    for (auto& pr : ghosts) if (pr.first == L.peers[j]) ++cnt;
    L.recv_off[j + 1] = L.recv_off[j] + cnt;   // ghosts are sorted by (peer, gid) = local ghost order
  }
  L.elem.resize(L.elem_glob.size());
  for (size_t i = 0; i < L.elem_glob.size(); ++i) {
    const int64_t* c = elements + 5 * L.elem_glob[i];
    L.elem[i] = make_int4(g2l[c[0] - 1], g2l[c[1] - 1], g2l[c[2] - 1], (int)c[4]);
  }
  return 0;
}

// Owned-row vectors in the caller's numbering (single-GPU debug exports; identity unless the mesh was renumbered)
int upload_owned(Context* ctx, double* dst, const double* src_user) {
  if (ctx->loc2glob.empty()) return upload(ctx, dst, src_user, (size_t)ctx->n_own);
  std::vector<double> tmp(ctx->n_own);
  for (int i = 0; i < ctx->n_own; ++i) tmp[i] = src_user[ctx->loc2glob[i]];
  int rc = upload(ctx, dst, tmp.data(), tmp.size());
  if (rc) return rc;
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int download_owned(Context* ctx, const double* src, double* dst_user) {
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->loc2glob.empty()) {
    DDPS_CUDA(cudaMemcpy(dst_user, src, (size_t)ctx->n_own * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
  }
  std::vector<double> tmp(ctx->n_own);
  DDPS_CUDA(cudaMemcpy(tmp.data(), src, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int i = 0; i < ctx->n_own; ++i) dst_user[ctx->loc2glob[i]] = tmp[i];
  return 0;
}

// Space-filling-curve (Morton / Z-order) renumbering of the nodes: internal id -> user id.  gmsh numbers nodes
// in insertion order; rows that are neighbours in the mesh are then far apart in memory and the SpMV gathers
// miss L1/L2.  Sorting the nodes along a Z-curve makes index-neighbours spatial neighbours.
void morton_order(const double* nodes, int64_t nq, std::vector<int64_t>& user_of_int) {
  double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
  for (int64_t i = 0; i < nq; ++i)
    for (int d = 0; d < 2; ++d) { lo[d] = std::min(lo[d], nodes[2 * i + d]); hi[d] = std::max(hi[d], nodes[2 * i + d]); }
  const double ext = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), 1e-300);
  auto spread = [](uint32_t v) {
    v &= 0xffff; v = (v | (v << 8)) & 0x00ff00ffu; v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u; v = (v | (v << 1)) & 0x55555555u;
    return v;
  };
  std::vector<std::pair<uint32_t, int64_t>> keys(nq);
  for (int64_t i = 0; i < nq; ++i) {
    uint32_t xi = (uint32_t)std::min(65535.0, (nodes[2 * i] - lo[0]) / ext * 65535.0);
    uint32_t yi = (uint32_t)std::min(65535.0, (nodes[2 * i + 1] - lo[1]) / ext * 65535.0);
    keys[i] = {spread(xi) | (spread(yi) << 1), i};
  }
  std::sort(keys.begin(), keys.end());
  user_of_int.resize(nq);
  for (int64_t i = 0; i < nq; ++i) user_of_int[i] = keys[i].second;
}

AsmArgs base_args(Context* ctx) {
  AsmArgs a;
  memset(&a, 0, sizeof(a));
  const Pattern& P = ctx->pat;
  a.nrows = P.nrows; a.nslices = P.nslices; a.wmax = P.max_row_len; a.snmax = P.max_sn;
  static const int pf_env = getenv("DDPS_ASM_PF") ? atoi(getenv("DDPS_ASM_PF")) : -1;   // (experiments) prefetch distance in slices
  static const int pfd_env = getenv("DDPS_ASM_PFD") ? atoi(getenv("DDPS_ASM_PFD")) : -1;
  a.pf_desc = pfd_env >= 0 ? (pfd_env ? pfd_env : (1 << 30)) : ctx->num_sms * 24;
  a.pf_dist = pf_env > 0 ? pf_env : (1 << 30);   // default: off (measured: no gain once the records come by bulk copies)
  a.sdesc = P.sdesc; a.sn_nodes = P.sn_nodes; a.inc_tri = P.inc_tri; a.inc_pack = P.inc_pack; a.arinc = P.arinc;
  a.runs = P.runs; a.rowinfo = P.rowinfo; a.isD = ctx->isD; a.xy = ctx->xy; a.rec = ctx->rec;
  return a;
}

int need_coeff(Context* ctx, int slot, const char* name) {
  if (!ctx->coeff[slot]) { ctx->fail(DDPS_ERR_STATE, std::string("coefficient not set: ") + name); return DDPS_ERR_STATE; }
  return 0;
}

// MV_assemble (src:229-287) / the constant semlin stiffness (src:924-1028)
int assemble_base(Context* ctx, bool semlin) {
  int rc;
  AsmArgs a = base_args(ctx);
  int field = semlin ? DDPS_FIELD_SCALAR : DDPS_FIELD_V;
  if (!ctx->have_D[field]) { ctx->fail(DDPS_ERR_STATE, "Dirichlet data not set"); return DDPS_ERR_STATE; }
  if (semlin) {
    if ((rc = need_coeff(ctx, DDPS_COEFF_A11, "A11")) || (rc = need_coeff(ctx, DDPS_COEFF_A22, "A22"))) return rc;
    a.phx = ctx->coeff[DDPS_COEFF_A22];   // x components pair with A[4], src:926
    a.phy = ctx->coeff[DDPS_COEFF_A11];
  } else {
    if ((rc = need_coeff(ctx, DDPS_COEFF_LAMBDA2, "LAMBDA2"))) return rc;
    a.phx = a.phy = ctx->coeff[DDPS_COEFF_LAMBDA2];
  }
  a.gD = ctx->gD[field];
  a.out_val = ctx->val_base;
  a.lift_out = ctx->lift;
  if ((rc = launch_prep_base(ctx))) return rc;
  if ((rc = launch_assemble_base(ctx, a))) return rc;
  // once per solve: the midpoint samples of the Newton systems' load in incidence order (what k_asm_newton streams)
  auto permute = [&](int slot, double** out) -> int {
    if (!ctx->coeff[slot]) return 0;
    if (!*out) { int r = dev_alloc(ctx, out, (size_t)ctx->pat.inc_entries); if (r) return r; }
    return launch_permute_mid(ctx, ctx->coeff[slot], *out);
  };
  if (!semlin) {
    if ((rc = permute(DDPS_COEFF_C_MID, &ctx->cinc))) return rc;
  } else if (ctx->functor_kind == DDPS_FUNCTOR_SINH) {
    if ((rc = permute(DDPS_COEFF_H_MID, &ctx->cinc))) return rc;
  } else if (ctx->functor_kind == DDPS_FUNCTOR_POLY) {
    for (int k = 0; k < kMaxPoly; ++k)
      if ((rc = permute(DDPS_COEFF_POLY_MID0 + k, &ctx->cinc_poly[k]))) return rc;
  }
  ctx->base_ready = true;
  ctx->base_is_semlin = semlin;
  return 0;
}

// Multigrid hierarchy of the freshly assembled base operator (opt-in, single GPU): SURVEY 8f rank 1
int maybe_amg_setup(Context* ctx) {
  ctx->stats.amg_levels = 0;
  if (ctx->opts.precond != DDPS_PRECOND_AMG) { amg_free(ctx); return 0; }
  int rc = amg_setup(ctx);   // (multi-rank contexts: distributed hierarchy, collective)
  if (rc) return rc;
  int64_t nl = 0;
  amg_debug_levels(ctx, &nl, nullptr, 0);
  ctx->stats.amg_levels = (int32_t)nl;
  return 0;
}

// One Newton-system assembly at the iterate Vit: A = base - J0(V), b(V), F = base*V - b, ||F||_inf.
int assemble_newton(Context* ctx, bool semlin, double delta, const double* Vit, const double* u, const double* v) {
  int rc;
  AsmArgs a = base_args(ctx);
  a.base_val = ctx->val_base; a.lift_in = ctx->lift;
  a.out_val = ctx->val_op; a.dinv = ctx->dinv; a.b = ctx->b; a.resid = ctx->resid;
  if (!semlin) {
    if ((rc = need_coeff(ctx, DDPS_COEFF_C_MID, "C_MID"))) return rc;
    if (!ctx->cinc) { ctx->fail(DDPS_ERR_STATE, "C_MID was set after the base assembly of this solve"); return DDPS_ERR_STATE; }
    if ((rc = launch_prep_poisson(ctx, Vit, u, v, delta * delta))) return rc;
    a.gD = ctx->gD[DDPS_FIELD_V];
    a.cinc = ctx->cinc;
    a.par0 = delta * delta;
    return launch_assemble_newton(ctx, a, RHS_DDP);
  }
  if ((rc = launch_prep_semlin(ctx, Vit))) return rc;
  a.gD = ctx->gD[DDPS_FIELD_SCALAR];
  a.nodal_const = ctx->coeff[DDPS_COEFF_NEUMANN];
  if (ctx->functor_kind == DDPS_FUNCTOR_SINH) {
    if ((rc = need_coeff(ctx, DDPS_COEFF_H_MID, "H_MID"))) return rc;
    a.cinc = ctx->cinc;
    a.par0 = ctx->functor_par[0]; a.par1 = ctx->functor_par[1];
    return launch_assemble_newton(ctx, a, RHS_SINH);
  } else if (ctx->functor_kind == DDPS_FUNCTOR_POLY) {
    int deg = (int)ctx->functor_par[0];
    for (int k = 0; k <= deg; ++k) {
      if ((rc = need_coeff(ctx, DDPS_COEFF_POLY_MID0 + k, "POLY_MID")) || (rc = need_coeff(ctx, DDPS_COEFF_POLY_NODE0 + k, "POLY_NODE")))
        return rc;
      a.cinc_poly[k] = ctx->cinc_poly[k];
    }
    a.par0 = (double)deg;
    return launch_assemble_newton(ctx, a, RHS_POLY);
  }
  ctx->fail(DDPS_ERR_STATE, "no functor registered for f,g (arbitrary closures cannot run on the device; no CPU fallback)");
  return DDPS_ERR_STATE;
}

// uvsolve's system (src:433-588).  which = FIELD_U: (V, tau_p, tau_n, mu_p, u0, v0);
// which = FIELD_W: (-V, tau_n, tau_p, mu_n, v0, u0) exactly as G passes them (src:598-599).
int assemble_uv(Context* ctx, int which, int rec_mode, double tau_p, double tau_n, const double* V, const double* u0,
                const double* v0) {
  int rc;
  const bool isU = (which == DDPS_FIELD_U);
  int mu_slot = isU ? DDPS_COEFF_MU_P : DDPS_COEFF_MU_N;
  if ((rc = need_coeff(ctx, mu_slot, isU ? "MU_P" : "MU_N"))) return rc;
  if (!ctx->have_D[which]) { ctx->fail(DDPS_ERR_STATE, "Dirichlet data for u/v not set"); return DDPS_ERR_STATE; }
  if ((rc = launch_prep_uv(ctx, isU ? 1.0 : -1.0, V, u0, v0, tau_p, tau_n, rec_mode))) return rc;
  AsmArgs a = base_args(ctx);
  a.gD = ctx->gD[which];
  a.phx = ctx->coeff[mu_slot];
  a.par0 = isU ? tau_p : tau_n; a.par1 = isU ? tau_n : tau_p;   // callee's (tau_p, tau_n)
  a.out_val = ctx->val_op; a.dinv = ctx->dinv; a.b = ctx->b;
  if (rec_mode == DDPS_REC_SHOCKLEY) return launch_assemble_uv(ctx, a, RHS_SRH, true);
  return launch_assemble_uv(ctx, a, RHS_ZERO, false);
}

// Relative A-norm error every linear solve of a Newton / Gummel loop must reach (PcgScalars::eps_e2): two orders below the
// outer tolerance (the outer loops test ABSOLUTE max-norms of updates / residuals of O(1) fields, src:368,653,1040), at most
// opts.lin_etol, and never below 1e-13 (what FP64 CG can resolve).  The reference's direct solves are exact to rounding;
// a residual-only rule leaves smooth error components of size ||A^-1|| * tol ~ tol / h^2 behind.
double energy_tol(const Context* ctx, double outer_tol) {
  if (ctx->opts.lin_etol < 0.0) return 0.0;   // A/B: residual rules only
  const double cap = ctx->opts.lin_etol > 0.0 ? ctx->opts.lin_etol : 1e-10;
  static const double factor = getenv("DDPS_ETOL_FACTOR") ? atof(getenv("DDPS_ETOL_FACTOR")) : 0.01;   // (experiments)
  return std::max(1e-13, std::min(cap, factor * outer_tol));
}

struct EventTimer {
  Context* ctx;
  cudaEvent_t a, b;
  EventTimer(Context* c, cudaEvent_t a_, cudaEvent_t b_) : ctx(c), a(a_), b(b_) { cudaEventRecord(a, c->stream); }
  void stop() { cudaEventRecord(b, ctx->stream); }
  double ms() { float t = 0; cudaEventSynchronize(b); cudaEventElapsedTime(&t, a, b); return t; }
};

// The Newton loop shared by Vsolve (src:368-423) and solve_semlin_poisson (src:1040-1095):
//   assemble A = base - J0(V), b(V), F = base*V - b at the iterate;  stop when ||F||_inf <= tol;  V <- V - A^-1 F.
// The assembly at the updated iterate is both the stop test of the reference's `while` and the next step's system.
// opts.newton_damping = k > 0 (OPT-IN; the reference's update is undamped, src:394-395,1063-1064):
// the step is halved, at most k times, until ||F||_inf decreases -- backtracking on the quantity the loop tests anyway.
// on_res(res, it) sees the residual norm of every ACCEPTED iterate (it = 0: the initial one).
template <class OnRes>
int newton_loop(Context* ctx, bool semlin, double delta, double tol, const double* u, const double* v, OnRes&& on_res,
                int* count_out) {
  int rc;
  const int n = ctx->n_own;
  const int cap = ctx->opts.max_newton > 0 ? ctx->opts.max_newton : 1000;
  const int damp = std::max(0, ctx->opts.newton_damping);
  const size_t bytes = (size_t)ctx->n_loc * sizeof(double);
  auto assemble = [&](double* res) -> int {
    EventTimer ta(ctx, ctx->ev0, ctx->ev1);
    int r = assemble_newton(ctx, semlin, delta, ctx->V, u, v);
    if (r) return r;
    ta.stop();
    if ((r = fetch_resid_norm(ctx, res))) return r;
    ctx->stats.t_asm_ms += ta.ms();
    return 0;
  };
  int count = 0;
  double res = 0.0;
  // error-based stop rule of the corrections: relative to the energy of the LARGEST correction of this loop (~ the iterate's)
  DDPS_CUDA(cudaMemsetAsync(&ctx->scal->e_ref, 0, sizeof(double), ctx->stream));
  const double eps_e = energy_tol(ctx, tol);
  if ((rc = assemble(&res))) return rc;
  on_res(res, 0);
  while (res > tol) {                                                                        // src:368, 1040
    if (count >= cap) {
      if (ctx->opts.lin_cap_ok) break;   // profiling runs only
      ctx->fail(DDPS_ERR_NOCONV, semlin ? "semlin: Newton iteration cap reached" : "Vsolve: Newton iteration cap reached");
      *count_out = count;
      return DDPS_ERR_NOCONV;
    }
    EventTimer tp(ctx, ctx->ev2, ctx->ev3);
    // Newton correction: the linear residual only has to be below the Newton tolerance (so that, like the
    // reference's direct solve, a linear problem converges in one step); newton_eta is a relative floor.
    ctx->amg_kind = count == 0 ? 0 : 3;
    ctx->eps_e = eps_e;
    // Gummel loop: the k-th correction of this Vsolve starts from the k-th correction of the previous one (they converge with
    // the outer iteration; the stop rules -- absolute residual, error energy relative to the correction's -- are unchanged, so
    // the Newton iterates and counts are the reference's; only the CG iteration counts shrink).  Not for solve_semlin: one loop.
    const bool warm_slot = !semlin && ctx->ddpe_active && ctx->opts.warm_start && count < Context::kNewtonWarm;
    const bool warm = warm_slot && ctx->newton_x_valid[count];
    if (warm) DDPS_CUDA(cudaMemcpyAsync(ctx->x, ctx->newton_x[count], (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    rc = solve_system(ctx, ctx->val_op, ctx->dinv, ctx->resid, ctx->x, warm, ctx->opts.newton_eta, 0.1 * tol,
                      ctx->opts.lin_max_it, nullptr, nullptr);                               // src:394, 1063
    ctx->amg_kind = -1;
    ctx->eps_e = 0.0;
    tp.stop();
    ctx->stats.t_pcg_ms += tp.ms();
    if (rc) return rc;
    if (warm_slot) {
      if (!ctx->newton_x[count]) DDPS_CUDA(cudaMalloc(&ctx->newton_x[count], std::max<size_t>(n, 1) * sizeof(double)));
      DDPS_CUDA(cudaMemcpyAsync(ctx->newton_x[count], ctx->x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      ctx->newton_x_valid[count] = true;
    }
    if (damp) DDPS_CUDA(cudaMemcpyAsync(ctx->Unew, ctx->V, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    if ((rc = launch_axpy(ctx, -1.0, ctx->x, ctx->V, n))) return rc;                         // src:395, 1064
    if ((rc = halo_exchange(ctx, ctx->V))) return rc;
    double res_new = 0.0;
    if ((rc = assemble(&res_new))) return rc;
    double t = 1.0;
    for (int k = 0; k < damp && !(res_new < res); ++k) {
      t *= 0.5;
      DDPS_CUDA(cudaMemcpyAsync(ctx->V, ctx->Unew, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
      if ((rc = launch_axpy(ctx, -t, ctx->x, ctx->V, n))) return rc;
      if ((rc = halo_exchange(ctx, ctx->V))) return rc;
      if ((rc = assemble(&res_new))) return rc;
      ctx->stats.newton_halvings++;
    }
    res = res_new;
    ++count;
    on_res(res, count);
  }
  *count_out = count;
  return 0;
}

// Vsolve (src:292-425): Newton from V=0 on MV*V = b(V) at frozen (u0,v0).
int vsolve(Context* ctx) {
  if (!(ctx->opts.warm_start_v && ctx->stats.outer_iters > 0))   // opt-in warm start keeps the previous V
    DDPS_CUDA(cudaMemsetAsync(ctx->V, 0, (size_t)ctx->n_loc * sizeof(double), ctx->stream));   // src:311
  int newton = 0;
  int rc = newton_loop(ctx, false, ctx->delta, ctx->tol, ctx->U0, ctx->W0, [](double, int) {}, &newton);
  if (rc) return rc;
  ctx->stats.newton_total += newton;
  ctx->hist_newton.push_back((double)newton);
  return 0;
}

int uvsolve(Context* ctx, int which) {
  int rc;
  EventTimer ta(ctx, ctx->ev0, ctx->ev1);
  if ((rc = assemble_uv(ctx, which, ctx->rec_mode, ctx->tau_p, ctx->tau_n, ctx->V, ctx->U0, ctx->W0))) return rc;
  ta.stop();
  double* field = (which == DDPS_FIELD_U) ? ctx->U : ctx->W;
  EventTimer tp(ctx, ctx->ev2, ctx->ev3);
  ctx->amg_kind = (which == DDPS_FIELD_U) ? 1 : 2;
  DDPS_CUDA(cudaMemsetAsync(&ctx->scal->e_ref, 0, sizeof(double), ctx->stream));
  ctx->eps_e = energy_tol(ctx, ctx->tol);
  rc = solve_system(ctx, ctx->val_op, ctx->dinv, ctx->b, field, ctx->opts.warm_start != 0, ctx->opts.lin_rtol, 0.0,
                 ctx->opts.lin_max_it, nullptr, nullptr);                                     // src:588
  ctx->amg_kind = -1;
  ctx->eps_e = 0.0;
  tp.stop();
  ctx->stats.t_pcg_ms += tp.ms();
  ctx->stats.t_asm_ms += ta.ms();
  if (rc) return rc;
  return halo_exchange(ctx, field);
}

// One application of the Gummel map G (src:596-601) + the update norm (src:653).
int gummel_step(Context* ctx, double* control) {
  int rc;
  const size_t bytes = (size_t)ctx->n_loc * sizeof(double);
  DDPS_CUDA(cudaMemcpyAsync(ctx->U0, ctx->U, bytes, cudaMemcpyDeviceToDevice, ctx->stream));   // src:650
  DDPS_CUDA(cudaMemcpyAsync(ctx->W0, ctx->W, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  DDPS_CUDA(cudaMemcpyAsync(ctx->V0, ctx->V, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  if ((rc = vsolve(ctx))) return rc;                   // src:597
  if ((rc = uvsolve(ctx, DDPS_FIELD_U))) return rc;    // src:598
  if ((rc = uvsolve(ctx, DDPS_FIELD_W))) return rc;    // src:599 (uses the OLD u0)
  return launch_max_abs_diff3(ctx, ctx->U, ctx->U0, ctx->W, ctx->W0, ctx->V, ctx->V0, ctx->n_own, control);
}

void reset_stats(Context* ctx) {
  int neg = ctx->stats.n_negative_area;
  double ts = ctx->stats.t_setup_ms;
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  ctx->stats.n_negative_area = neg;
  ctx->stats.t_setup_ms = ts;
  ctx->hist_outer.clear(); ctx->hist_newton.clear(); ctx->hist_pcg.clear();
}

}  // namespace

// =================================================================================================
extern "C" {

int ddps_version(void) { return DDPS_VERSION; }

const char* ddps_last_error(const ddps_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

void ddps_default_opts(ddps_solver_opts* o) {
  memset(o, 0, sizeof(*o));
  o->lin_rtol = 1e-13;
  o->newton_eta = 1e-15;
  o->lin_max_it = 200000;
  o->max_newton = 0;
  o->max_gummel = 0;
  o->check_every = 32;
  o->warm_start = 1;
  o->verbose = 0;
  o->lin_etol = 1e-10;
}

int ddps_ctx_create(ddps_ctx** out, int device) {
  if (!out) return DDPS_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("no CUDA device available (this library has no CPU fallback): ") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return DDPS_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { g_create_err = "device index out of range"; return DDPS_ERR_ARG; }
  ddps_ctx* ctx = new ddps_ctx();
  ctx->device = device;
  ddps_default_opts(&ctx->opts);
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  auto bail = [&](const char* what, cudaError_t err) {
    g_create_err = std::string(what) + ": " + cudaGetErrorString(err);
    delete ctx;
    return DDPS_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
  ctx->num_sms = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  {   // setup temporaries live in the stream-ordered pool: keep freed blocks cached for the next setup
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1); cudaEventCreate(&ctx->ev2); cudaEventCreate(&ctx->ev3);
  cudaEventCreate(&ctx->ev4); cudaEventCreate(&ctx->ev5);
  if ((e = cudaMalloc(&ctx->scal, sizeof(PcgScalars))) != cudaSuccess) return bail("cudaMalloc", e);
  cudaMemset(ctx->scal, 0, sizeof(PcgScalars));
  if ((e = cudaMallocHost(&ctx->scal_host, sizeof(PcgScalars))) != cudaSuccess) return bail("cudaMallocHost", e);
  memset(ctx->scal_host, 0, sizeof(PcgScalars));
  *out = ctx;
  return DDPS_OK;
}

int ddps_ctx_destroy(ddps_ctx* ctx) {
  if (!ctx) return DDPS_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  free_mesh(ctx);
  comm_destroy(ctx);
  cudaFree(ctx->scal);
  cudaFreeHost(ctx->scal_host);
  cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaEventDestroy(ctx->ev2); cudaEventDestroy(ctx->ev3);
  cudaEventDestroy(ctx->ev4); cudaEventDestroy(ctx->ev5);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return DDPS_OK;
}

int ddps_set_opts(ddps_ctx* ctx, const ddps_solver_opts* o) {
  if (!ctx || !o) return DDPS_ERR_ARG;
  if (!(o->lin_rtol >= 0) || !(o->newton_eta >= 0)) { ctx->fail(DDPS_ERR_ARG, "negative tolerance"); return DDPS_ERR_ARG; }
  ctx->opts = *o;
  if (ctx->opts.lin_max_it <= 0) ctx->opts.lin_max_it = 200000;
  if (ctx->opts.check_every <= 0) ctx->opts.check_every = 32;
  return DDPS_OK;
}

int ddps_comm_unique_id(char id128[128]) { return comm_unique_id(id128); }

int ddps_comm_init(ddps_ctx* ctx, const char id128[128], int rank, int nranks) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->have_mesh) { ctx->fail(DDPS_ERR_STATE, "ddps_comm_init must precede ddps_mesh_set"); return DDPS_ERR_STATE; }
  return comm_init(ctx, id128, rank, nranks);
}

int ddps_partition_nodes(const double* nodes, int64_t nq, int nparts, int32_t* owner) {
  if (!nodes || !owner || nq < 0 || nparts < 1) return DDPS_ERR_ARG;
  partition_nodes_rcb(nodes, nq, nparts, owner);
  return DDPS_OK;
}

int ddps_partition_plan(const double* nodes, int64_t nq, const int64_t* elements, int64_t nme, int nparts, int rank,
                        int64_t counts[4], int64_t* loc2glob, int32_t* peers, int64_t* send_counts, int64_t* recv_counts,
                        int64_t* send_glob) {
  if (!nodes || !elements || !counts || nq <= 0 || nme <= 0 || nparts < 1 || rank < 0 || rank >= nparts) return DDPS_ERR_ARG;
  LocalMesh L;
  std::string err;
  if (build_local_mesh(nodes, nq, elements, nme, nparts, rank, L, err)) return DDPS_ERR_ARG;
  counts[0] = L.n_own; counts[1] = L.n_loc - L.n_own; counts[2] = (int64_t)L.elem_glob.size(); counts[3] = (int64_t)L.peers.size();
  if (loc2glob) for (int i = 0; i < L.n_loc; ++i) loc2glob[i] = L.loc2glob[i];
  for (size_t j = 0; j < L.peers.size(); ++j) {
    if (peers) peers[j] = L.peers[j];
    if (send_counts) send_counts[j] = L.send_off[j + 1] - L.send_off[j];
    if (recv_counts) recv_counts[j] = L.recv_off[j + 1] - L.recv_off[j];
  }
  if (send_glob) for (size_t i = 0; i < L.send_idx.size(); ++i) send_glob[i] = L.loc2glob[L.send_idx[i]];
  return DDPS_OK;
}

// common tail of ddps_mesh_set / ddps_mesh_set_local: orientation report, pattern build, work arrays, peer-memory setup
static int finish_mesh(ddps_ctx* ctx, double t0) {
  int rc;
  // orientation report (Q2: the mass weights use the SIGNED area, src:375)
  {
    int* d_neg = nullptr;
    DDPS_CUDA(cudaMalloc(&d_neg, sizeof(int)));
    cudaMemsetAsync(d_neg, 0, sizeof(int), ctx->stream);
    if (ctx->nme) k_count_negative<<<(ctx->nme + 255) / 256, 256, 0, ctx->stream>>>(ctx->elem, ctx->xy, ctx->nme, d_neg);
    int neg = 0;
    cudaMemcpyAsync(&neg, d_neg, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_neg);
    ctx->stats.n_negative_area = neg;
  }
  const double t_up = now_ms();
  if ((rc = build_pattern(ctx))) { free_mesh(ctx); return rc; }
  const double t_pat = now_ms();

  const size_t nl = ctx->n_loc, no = ctx->n_own, ne = (size_t)ctx->pat.nentries;
  if ((rc = dev_alloc(ctx, &ctx->isD, nl))) return rc;
  DDPS_CUDA(cudaMemsetAsync(ctx->isD, 0, nl, ctx->stream));
  for (int f = 0; f < 4; ++f) {
    if ((rc = dev_alloc(ctx, &ctx->gD[f], nl))) return rc;
    DDPS_CUDA(cudaMemsetAsync(ctx->gD[f], 0, nl * sizeof(double), ctx->stream));
  }
  if ((rc = dev_alloc(ctx, &ctx->rec, nl * kRecD))) return rc;
  DDPS_CUDA(cudaMemsetAsync(ctx->rec, 0, nl * kRecD * sizeof(double), ctx->stream));
  double** nodal[] = {&ctx->V, &ctx->U, &ctx->W, &ctx->V0, &ctx->U0, &ctx->W0, &ctx->Unew, &ctx->p};
  for (double** pp : nodal) {
    if ((rc = dev_alloc(ctx, pp, nl))) return rc;
    DDPS_CUDA(cudaMemsetAsync(*pp, 0, nl * sizeof(double), ctx->stream));
  }
  double** owned[] = {&ctx->b, &ctx->resid, &ctx->dinv, &ctx->x, &ctx->r, &ctx->q, &ctx->lift};
  for (double** pp : owned) {
    if ((rc = dev_alloc(ctx, pp, no))) return rc;
    DDPS_CUDA(cudaMemsetAsync(*pp, 0, no * sizeof(double), ctx->stream));
  }
  if ((rc = dev_alloc(ctx, &ctx->val_base, ne))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->val_op, ne))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->sc, nl))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->bs, no))) return rc;
  DDPS_CUDA(cudaMemsetAsync(ctx->sc, 0, std::max<size_t>(nl, 1) * sizeof(double), ctx->stream));
  DDPS_CUDA(cudaMemsetAsync(ctx->val_base, 0, std::max<size_t>(ne, 1) * sizeof(double), ctx->stream));
  DDPS_CUDA(cudaMemsetAsync(ctx->val_op, 0, std::max<size_t>(ne, 1) * sizeof(double), ctx->stream));
  int asm_blocks = (ctx->pat.nslices * 32 + kAsmBlock - 1) / kAsmBlock;
  ctx->partial_cap = std::max(asm_blocks, 4 * ctx->num_sms * 16) + 64;
  if ((rc = dev_alloc(ctx, &ctx->partials, (size_t)ctx->partial_cap))) return rc;
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  if ((rc = p2p_setup(ctx))) return rc;
  ctx->have_mesh = true;
  ctx->stats.t_setup_ms = now_ms() - t0;
  if (getenv("DDPS_AMG_TIMING"))
    fprintf(stderr, "[ddps] mesh_set %.1f ms: uploads + conversion %.1f, pattern build %.1f, allocations %.1f\n", ctx->stats.t_setup_ms,
            t_up - t0, t_pat - t_up, now_ms() - t_pat);
  return DDPS_OK;
}

int ddps_mesh_set(ddps_ctx* ctx, const double* nodes, int64_t nq, const int64_t* elements, int64_t nme) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (!nodes || !elements || nq <= 0 || nme <= 0) { ctx->fail(DDPS_ERR_ARG, "empty mesh"); return DDPS_ERR_ARG; }
  if (nq >= (1LL << 31) - 64) { ctx->fail(DDPS_ERR_ARG, "too many nodes"); return DDPS_ERR_ARG; }
  double t0 = now_ms();
  free_mesh(ctx);
  ctx->nq_global = nq; ctx->nme_global = nme;
  int rc;
  std::vector<int4> hel;
  // optional internal renumbering (the caller keeps its own numbering at the ABI)
  std::vector<int64_t> user_of_int;
  std::vector<double> pnodes;
  std::vector<int64_t> pelems;
  if (ctx->opts.renumber) {
    for (int64_t e = 0; e < nme; ++e)
      for (int k = 0; k < 3; ++k)
        if (elements[5 * e + k] < 1 || elements[5 * e + k] > nq) {
          ctx->fail(DDPS_ERR_ARG, "element " + std::to_string(e + 1) + " references a node outside 1..nq");
          return DDPS_ERR_ARG;
        }
    morton_order(nodes, nq, user_of_int);
    std::vector<int64_t> int_of_user(nq);
    for (int64_t i = 0; i < nq; ++i) int_of_user[user_of_int[i]] = i;
    pnodes.resize((size_t)2 * nq);
    for (int64_t i = 0; i < nq; ++i) { pnodes[2 * i] = nodes[2 * user_of_int[i]]; pnodes[2 * i + 1] = nodes[2 * user_of_int[i] + 1]; }
    pelems.assign(elements, elements + (size_t)5 * nme);
    for (int64_t e = 0; e < nme; ++e)
      for (int k = 0; k < 3; ++k) pelems[5 * e + k] = int_of_user[elements[5 * e + k] - 1] + 1;
    nodes = pnodes.data();
    elements = pelems.data();
  }
  if (ctx->nranks == 1) {
    ctx->n_own = ctx->n_loc = (int)nq;
    if (nme >= (1LL << 29)) { ctx->fail(DDPS_ERR_ARG, "too many triangles for one GPU"); return DDPS_ERR_ARG; }
    ctx->nme = (int)nme;
    if (user_of_int.empty()) {
      // the caller's arrays go to the device as they are (full PCIe speed when they are pinned); the conversion to the
      // packed 0-based int4 records and the range check run on the device
      long long* d_el = nullptr;
      int* d_bad = nullptr;
      DDPS_CUDA(cudaMallocAsync(&d_el, (size_t)5 * nme * sizeof(long long), ctx->stream));   // (pool: see setup.cu)
      if ((rc = dev_alloc(ctx, &ctx->elem, (size_t)nme)) || (rc = dev_alloc(ctx, &ctx->xy, (size_t)nq))) { cudaFreeAsync(d_el, ctx->stream); return rc; }
      cudaMalloc(&d_bad, sizeof(int));
      cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream);
      cudaMemcpyAsync(d_el, elements, (size_t)5 * nme * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream);
      cudaMemcpyAsync(ctx->xy, nodes, (size_t)2 * nq * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
      k_convert_elements<<<(unsigned)((nme + 255) / 256), 256, 0, ctx->stream>>>(d_el, (int)nme, (long long)nq, ctx->elem, d_bad);
      int bad = 0;
      cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
      cudaFreeAsync(d_el, ctx->stream);
      cudaError_t ce = cudaStreamSynchronize(ctx->stream);
      cudaFree(d_bad);
      if (ce != cudaSuccess) { ctx->fail(DDPS_ERR_CUDA, cudaGetErrorString(ce)); return DDPS_ERR_CUDA; }
      if (bad) {
        ctx->fail(DDPS_ERR_ARG, "element " + std::to_string(bad) + " references a node outside 1..nq");
        return DDPS_ERR_ARG;
      }
    } else {
    hel.resize(nme);
    for (int64_t e = 0; e < nme; ++e) {
      const int64_t* c = elements + 5 * e;
      int64_t a = c[0] - 1, b = c[1] - 1, d = c[2] - 1;
      if (a < 0 || b < 0 || d < 0 || a >= nq || b >= nq || d >= nq) {
        ctx->fail(DDPS_ERR_ARG, "element " + std::to_string(e + 1) + " references a node outside 1..nq");
        return DDPS_ERR_ARG;
      }
      hel[e] = make_int4((int)a, (int)b, (int)d, (int)c[4]);
    }
    if ((rc = dev_alloc(ctx, &ctx->xy, (size_t)nq))) return rc;
    if ((rc = upload(ctx, (double*)ctx->xy, nodes, (size_t)2 * nq))) return rc;
    }
    if (!user_of_int.empty()) {
      ctx->loc2glob = user_of_int;
      if ((rc = dev_alloc(ctx, &ctx->d_own_glob, (size_t)nq))) return rc;
      std::vector<long long> og(user_of_int.begin(), user_of_int.end());
      if ((rc = upload(ctx, ctx->d_own_glob, og.data(), og.size()))) return rc;
      DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
  } else {
    LocalMesh L;
    std::string perr;
    if (build_local_mesh(nodes, nq, elements, nme, ctx->nranks, ctx->rank, L, perr)) { ctx->fail(DDPS_ERR_ARG, perr); return DDPS_ERR_ARG; }
    ctx->n_own = L.n_own;
    ctx->n_loc = L.n_loc;
    ctx->loc2glob = L.loc2glob;
    HaloPlan& H = ctx->halo;
    H.nneigh = (int)L.peers.size(); H.peer = L.peers; H.send_off = L.send_off; H.recv_off = L.recv_off;
    H.nsend = (int)L.send_idx.size();
    H.n_own = L.n_own; H.n_ghost = L.n_loc - L.n_own;
    if ((rc = dev_alloc(ctx, &H.send_idx, L.send_idx.size()))) return rc;
    if ((rc = dev_alloc(ctx, &H.send_buf, L.send_idx.size()))) return rc;
    if ((rc = upload(ctx, H.send_idx, L.send_idx.data(), L.send_idx.size()))) return rc;
    ctx->nme = (int)L.elem_glob.size();
    ctx->elem_glob = L.elem_glob;
    hel = L.elem;
    std::vector<double> hxy((size_t)2 * ctx->n_loc);
    for (int i = 0; i < ctx->n_loc; ++i) { hxy[2 * i] = nodes[2 * ctx->loc2glob[i]]; hxy[2 * i + 1] = nodes[2 * ctx->loc2glob[i] + 1]; }
    if ((rc = dev_alloc(ctx, &ctx->xy, (size_t)ctx->n_loc))) return rc;
    if ((rc = upload(ctx, (double*)ctx->xy, hxy.data(), hxy.size()))) return rc;
    if (!user_of_int.empty())   // compose with the renumbering: local -> internal global -> caller's numbering
      for (int i = 0; i < ctx->n_loc; ++i) ctx->loc2glob[i] = user_of_int[ctx->loc2glob[i]];
    if ((rc = dev_alloc(ctx, &ctx->d_own_glob, (size_t)ctx->n_own))) return rc;
    std::vector<long long> og(ctx->loc2glob.begin(), ctx->loc2glob.begin() + ctx->n_own);
    if ((rc = upload(ctx, ctx->d_own_glob, og.data(), og.size()))) return rc;
    DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (!hel.empty() || !ctx->elem) {
    if ((rc = dev_alloc(ctx, &ctx->elem, (size_t)ctx->nme))) return rc;
    if ((rc = upload(ctx, ctx->elem, hel.data(), hel.size()))) return rc;
  }
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  hel.clear(); hel.shrink_to_fit();

  return finish_mesh(ctx, t0);
}

// Per-rank mesh ingest: nobody holds the global mesh.  The halo plan follows from the local
// triangles alone: my send list to peer q = my owned corners of triangles that also have a q-owned corner, in increasing
// global id -- exactly q's ghost block for me (q holds every triangle touching one of its nodes, so it sees the same ones).
int ddps_mesh_set_local(ddps_ctx* ctx, const double* nodes, int64_t n_own, int64_t n_loc, const int64_t* node_glob,
                        const int32_t* ghost_owner, const int64_t* elements, int64_t nme_loc, const int64_t* elem_glob,
                        int64_t nq_global, int64_t nme_global) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (!nodes || !node_glob || !elements || n_own < 0 || n_loc < n_own || n_loc <= 0 || nme_loc <= 0 || (n_loc > n_own && !ghost_owner)) {
    ctx->fail(DDPS_ERR_ARG, "empty or inconsistent local mesh");
    return DDPS_ERR_ARG;
  }
  if (n_loc >= (1LL << 31) - 64 || nme_loc >= (1LL << 29) || nq_global >= (1LL << 31) - 64) { ctx->fail(DDPS_ERR_ARG, "local mesh too large for one GPU"); return DDPS_ERR_ARG; }
  if (ctx->opts.renumber) { ctx->fail(DDPS_ERR_ARG, "opts.renumber is not available with ddps_mesh_set_local (number the local nodes as you like)"); return DDPS_ERR_ARG; }
  const double t0 = now_ms();
  free_mesh(ctx);
  ctx->nq_global = nq_global; ctx->nme_global = nme_global;
  const int me = ctx->rank, nr = ctx->nranks;
  int rc;
  for (int64_t i = 1; i < n_own; ++i)
    if (node_glob[i] <= node_glob[i - 1]) { ctx->fail(DDPS_ERR_ARG, "owned nodes must come in increasing global id"); return DDPS_ERR_ARG; }
  for (int64_t g = 0; g < n_loc - n_own; ++g) {
    const int o = ghost_owner[g];
    if (o < 0 || o >= nr || o == me) { ctx->fail(DDPS_ERR_ARG, "bad ghost owner"); return DDPS_ERR_ARG; }
    if (g > 0 && (o < ghost_owner[g - 1] || (o == ghost_owner[g - 1] && node_glob[n_own + g] <= node_glob[n_own + g - 1]))) {
      ctx->fail(DDPS_ERR_ARG, "ghost nodes must be sorted by (owner, global id)");
      return DDPS_ERR_ARG;
    }
  }
  ctx->n_own = (int)n_own; ctx->n_loc = (int)n_loc; ctx->nme = (int)nme_loc;
  ctx->loc2glob.resize(n_loc);
  for (int64_t i = 0; i < n_loc; ++i) {
    ctx->loc2glob[i] = node_glob[i] - 1;
    if (ctx->loc2glob[i] < 0 || ctx->loc2glob[i] >= nq_global) { ctx->fail(DDPS_ERR_ARG, "global node id outside 1..nq_global"); return DDPS_ERR_ARG; }
  }
  if (elem_glob) {
    ctx->elem_glob.resize(nme_loc);
    for (int64_t e = 0; e < nme_loc; ++e) ctx->elem_glob[e] = elem_glob[e] - 1;
  }
  // triangles: local 1-based ids -> packed records (device), validated on the device like ddps_mesh_set does
  {
    long long* d_el = nullptr;
    int* d_bad = nullptr;
    DDPS_CUDA(cudaMallocAsync(&d_el, (size_t)5 * nme_loc * sizeof(long long), ctx->stream));
    if ((rc = dev_alloc(ctx, &ctx->elem, (size_t)nme_loc)) || (rc = dev_alloc(ctx, &ctx->xy, (size_t)n_loc))) { cudaFreeAsync(d_el, ctx->stream); return rc; }
    cudaMalloc(&d_bad, sizeof(int));
    cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream);
    cudaMemcpyAsync(d_el, elements, (size_t)5 * nme_loc * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(ctx->xy, nodes, (size_t)2 * n_loc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    k_convert_elements<<<(unsigned)((nme_loc + 255) / 256), 256, 0, ctx->stream>>>(d_el, (int)nme_loc, (long long)n_loc, ctx->elem, d_bad);
    int bad = 0;
    cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaFreeAsync(d_el, ctx->stream);
    cudaError_t ce = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_bad);
    if (ce != cudaSuccess) { ctx->fail(DDPS_ERR_CUDA, cudaGetErrorString(ce)); return DDPS_ERR_CUDA; }
    if (bad) { ctx->fail(DDPS_ERR_ARG, "local element " + std::to_string(bad) + " references a node outside 1..n_loc"); return DDPS_ERR_ARG; }
  }
  // halo plan
  HaloPlan& H = ctx->halo;
  H.n_own = (int)n_own; H.n_ghost = (int)(n_loc - n_own);
  if (nr > 1) {
    std::vector<std::pair<int, int>> sends;   // (peer, local owned id); owned ids ascend with the global id
    for (int64_t e = 0; e < nme_loc; ++e) {
      const int64_t* c = elements + 5 * e;
      int o[3];
      for (int k = 0; k < 3; ++k) o[k] = c[k] - 1 < n_own ? me : ghost_owner[c[k] - 1 - n_own];
      if (o[0] == me && o[1] == me && o[2] == me) continue;
      for (int k = 0; k < 3; ++k)
        if (o[k] == me)
          for (int j = 0; j < 3; ++j)
            if (o[j] != me) sends.emplace_back(o[j], (int)(c[k] - 1));
    }
    std::sort(sends.begin(), sends.end());
    sends.erase(std::unique(sends.begin(), sends.end()), sends.end());
    std::vector<int> peers;
    for (int64_t g = 0; g < n_loc - n_own; ++g) peers.push_back(ghost_owner[g]);
    for (auto& pr : sends) peers.push_back(pr.first);
    std::sort(peers.begin(), peers.end());
    peers.erase(std::unique(peers.begin(), peers.end()), peers.end());
    const int nn = (int)peers.size();
    std::vector<int> send_idx;
    H.nneigh = nn; H.peer = peers;
    H.send_off.assign(nn + 1, 0); H.recv_off.assign(nn + 1, 0);
    for (int j = 0; j < nn; ++j) {
      for (auto& pr : sends) if (pr.first == peers[j]) send_idx.push_back(pr.second);
      H.send_off[j + 1] = (int)send_idx.size();
      int cnt = 0;
      for (int64_t g = 0; g < n_loc - n_own; ++g) if (ghost_owner[g] == peers[j]) ++cnt;
      H.recv_off[j + 1] = H.recv_off[j] + cnt;
    }
    H.nsend = (int)send_idx.size();
    if ((rc = dev_alloc(ctx, &H.send_idx, send_idx.size()))) return rc;
    if ((rc = dev_alloc(ctx, &H.send_buf, send_idx.size()))) return rc;
    if ((rc = upload(ctx, H.send_idx, send_idx.data(), send_idx.size()))) return rc;
    DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if ((rc = dev_alloc(ctx, &ctx->d_own_glob, (size_t)n_own))) return rc;
  {
    std::vector<long long> og(ctx->loc2glob.begin(), ctx->loc2glob.begin() + n_own);
    if ((rc = upload(ctx, ctx->d_own_glob, og.data(), og.size()))) return rc;
    DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  ctx->local_ingest = true;
  return finish_mesh(ctx, t0);
}

// V, u, v of the nodes this rank OWNS (n_own each, in the order of ddps_mesh_set_local's owned nodes / increasing global
// id): no global gather -- the natural output of a partitioned solve
int ddps_ddpe_fetch_owned(ddps_ctx* ctx, double* V, double* u, double* v, int64_t cap) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if (cap < ctx->n_own) { ctx->fail(DDPS_ERR_ARG, "output buffers too small"); return DDPS_ERR_ARG; }
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  const size_t b = (size_t)ctx->n_own * sizeof(double);
  if (V) DDPS_CUDA(cudaMemcpy(V, ctx->V, b, cudaMemcpyDeviceToHost));
  if (u) DDPS_CUDA(cudaMemcpy(u, ctx->U, b, cudaMemcpyDeviceToHost));
  if (v) DDPS_CUDA(cudaMemcpy(v, ctx->W, b, cudaMemcpyDeviceToHost));
  return DDPS_OK;
}

int ddps_dirichlet_set(ddps_ctx* ctx, int field, const int64_t* idx1, const double* val, int64_t nd) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if (field < 0 || field > 3 || nd < 0 || (nd > 0 && (!idx1 || !val))) { ctx->fail(DDPS_ERR_ARG, "bad Dirichlet arguments"); return DDPS_ERR_ARG; }
  std::vector<int64_t> ids(nd);
  for (int64_t i = 0; i < nd; ++i) {
    ids[i] = idx1[i] - 1;
    if (ids[i] < 0 || ids[i] >= ctx->nq_global) { ctx->fail(DDPS_ERR_ARG, "Dirichlet node outside 1..nq"); return DDPS_ERR_ARG; }
  }
  std::vector<int64_t> sorted = ids;
  std::sort(sorted.begin(), sorted.end());
  if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end()) {
    // Q5: the reference would put 2 on the diagonal and double-subtract the lifting; not supported.
    ctx->fail(DDPS_ERR_ARG, "a Dirichlet node is listed twice (conflicting boundary values at a shared node are not supported)");
    return DDPS_ERR_ARG;
  }
  // solve_ddpe eliminates the node set of the V bddata for all three equations (src:640 -> 598-599);
  // the u/v value lists must cover exactly that set (the reference throws DimensionMismatch at src:563).
  if (field == DDPS_FIELD_SCALAR) {
    ctx->have_D[0] = ctx->have_D[1] = ctx->have_D[2] = false;   // isD is shared: a new scalar set invalidates the DDP sets
  } else {
    ctx->have_D[DDPS_FIELD_SCALAR] = false;
    bool any = false;
    for (int f = 0; f < 3; ++f) any = any || (f != field && ctx->have_D[f]);
    if (any && sorted != ctx->Dset) {
      ctx->fail(DDPS_ERR_ARG, "Dirichlet node sets of V, u and v differ (reference: DimensionMismatch at src:563)");
      return DDPS_ERR_ARG;
    }
  }
  ctx->Dset = sorted;
  if (ctx->loc2glob.empty()) {   // single GPU in the caller's numbering: scatter the list on the device
    long long* d_ids = nullptr;
    double* d_vals = nullptr;
    DDPS_CUDA(cudaMalloc(&d_ids, std::max<size_t>(nd, 1) * sizeof(long long)));
    DDPS_CUDA(cudaMalloc(&d_vals, std::max<size_t>(nd, 1) * sizeof(double)));
    cudaMemsetAsync(ctx->isD, 0, (size_t)ctx->n_loc, ctx->stream);
    cudaMemsetAsync(ctx->gD[field], 0, (size_t)ctx->n_loc * sizeof(double), ctx->stream);
    if (nd > 0) {
      cudaMemcpyAsync(d_ids, ids.data(), (size_t)nd * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream);
      cudaMemcpyAsync(d_vals, val, (size_t)nd * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
      k_scatter_dirichlet<<<(unsigned)((nd + 255) / 256), 256, 0, ctx->stream>>>((int)nd, d_ids, d_vals, ctx->isD, ctx->gD[field]);
    }
    cudaError_t ce = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_ids); cudaFree(d_vals);
    if (ce != cudaSuccess) { ctx->fail(DDPS_ERR_CUDA, cudaGetErrorString(ce)); return DDPS_ERR_CUDA; }
    ctx->have_D[field] = true;
    ctx->base_ready = false;
    return DDPS_OK;
  }
  // local copies of the flags / values: every listed node this rank holds (owned or ghost), found by binary search in
  // the sorted runs of loc2glob (owned nodes ascending, then one ascending run per ghost owner) -- no global array
  std::vector<unsigned char> lfl(ctx->n_loc, 0);
  std::vector<double> lg(ctx->n_loc, 0.0);
  {
    const std::vector<int64_t>& G = ctx->loc2glob;
    std::vector<std::pair<int, int>> runs;   // [begin, end) ascending runs
    int b0 = 0;
    for (int i = 1; i <= ctx->n_loc; ++i)
      if (i == ctx->n_loc || i == ctx->n_own || G[i] < G[i - 1]) { if (i > b0) runs.emplace_back(b0, i); b0 = i; }
    std::vector<std::pair<int64_t, int>> lookup;   // internally renumbered meshes have no long sorted runs: sort once
    if (runs.size() > 64) {
      lookup.resize(ctx->n_loc);
      for (int i = 0; i < ctx->n_loc; ++i) lookup[i] = {G[i], i};
      std::sort(lookup.begin(), lookup.end());
    }
    for (int64_t k = 0; k < nd; ++k) {
      if (!lookup.empty()) {
        auto it = std::lower_bound(lookup.begin(), lookup.end(), std::make_pair(ids[k], -1));
        for (; it != lookup.end() && it->first == ids[k]; ++it) { lfl[it->second] = 1; lg[it->second] = val[k]; }
        continue;
      }
      for (auto& rn : runs) {
        auto it = std::lower_bound(G.begin() + rn.first, G.begin() + rn.second, ids[k]);
        if (it != G.begin() + rn.second && *it == ids[k]) { const int i = (int)(it - G.begin()); lfl[i] = 1; lg[i] = val[k]; }
      }
    }
  }
  if ((rc = upload(ctx, ctx->isD, lfl.data(), lfl.size()))) return rc;
  if ((rc = upload(ctx, ctx->gD[field], lg.data(), lg.size()))) return rc;
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->have_D[field] = true;
  ctx->base_ready = false;
  return DDPS_OK;
}

static int coeff_slot_kind(int slot) {   // 0: [nme], 1: [3,nme], 2: [nq], -1: unknown
  if (slot == DDPS_COEFF_C_MID || slot == DDPS_COEFF_H_MID || (slot >= DDPS_COEFF_POLY_MID0 && slot < DDPS_COEFF_POLY_MID0 + kMaxPoly)) return 1;
  if (slot == DDPS_COEFF_NEUMANN || (slot >= DDPS_COEFF_POLY_NODE0 && slot < DDPS_COEFF_POLY_NODE0 + kMaxPoly)) return 2;
  if (slot >= 0 && slot <= DDPS_COEFF_A22) return 0;
  return -1;
}

int ddps_coeff_set(ddps_ctx* ctx, int slot, const double* data, int64_t len) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if (slot < 0 || slot >= 64 || !data) { ctx->fail(DDPS_ERR_ARG, "bad coefficient slot"); return DDPS_ERR_ARG; }
  const int kind = coeff_slot_kind(slot);  // 0: [nme], 1: [3,nme], 2: [nq]
  if (kind < 0) { ctx->fail(DDPS_ERR_ARG, "unknown coefficient slot"); return DDPS_ERR_ARG; }
  int64_t want = kind == 0 ? ctx->nme_global : kind == 1 ? 3 * ctx->nme_global : ctx->nq_global;
  if (len != want) { ctx->fail(DDPS_ERR_ARG, "coefficient array has the wrong length"); return DDPS_ERR_ARG; }
  const bool direct = (kind == 2) ? ctx->loc2glob.empty() : ctx->elem_glob.empty();
  if (!direct)
    for (int64_t i = 0; i < len; ++i)
      if (!std::isfinite(data[i])) { ctx->fail(DDPS_ERR_ARG, "coefficient array contains a non-finite value"); return DDPS_ERR_ARG; }
  size_t nloc = kind == 0 ? (size_t)ctx->nme : kind == 1 ? (size_t)3 * ctx->nme : (size_t)ctx->n_loc;
  if ((rc = dev_alloc(ctx, &ctx->coeff[slot], nloc))) return rc;
  ctx->coeff_len[slot] = (int64_t)nloc;
  if (direct) {
    // straight H2D (full PCIe speed from pinned memory); the finiteness check runs on the device
    if ((rc = upload(ctx, ctx->coeff[slot], data, nloc))) return rc;
    int* d_flag = nullptr;
    DDPS_CUDA(cudaMalloc(&d_flag, sizeof(int)));
    cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream);
    if (nloc) k_check_finite<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(nloc, ctx->coeff[slot], d_flag);
    int flag = 0;
    cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t ce = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_flag);
    if (ce != cudaSuccess) { ctx->fail(DDPS_ERR_CUDA, cudaGetErrorString(ce)); return DDPS_ERR_CUDA; }
    if (flag) {
      dev_free(ctx->coeff[slot]);
      ctx->fail(DDPS_ERR_ARG, "coefficient array contains a non-finite value");
      return DDPS_ERR_ARG;
    }
  } else {
    std::vector<double> tmp(nloc);
    if (kind == 0) for (int e = 0; e < ctx->nme; ++e) tmp[e] = data[ctx->elem_glob[e]];
    else if (kind == 1) for (int e = 0; e < ctx->nme; ++e) for (int k = 0; k < 3; ++k) tmp[3 * (size_t)e + k] = data[3 * ctx->elem_glob[e] + k];
    else for (int i = 0; i < ctx->n_loc; ++i) tmp[i] = data[ctx->loc2glob[i]];
    if ((rc = upload(ctx, ctx->coeff[slot], tmp.data(), nloc))) return rc;
    DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (slot == DDPS_COEFF_LAMBDA2 || slot == DDPS_COEFF_A11 || slot == DDPS_COEFF_A22) ctx->base_ready = false;
  return DDPS_OK;
}

int ddps_coeff_functor(ddps_ctx* ctx, int slot, int kind, const double* params, int nparams) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  const int sk = coeff_slot_kind(slot);
  if (sk < 0 || slot == DDPS_COEFF_NEUMANN) { ctx->fail(DDPS_ERR_ARG, "slot cannot be filled from a spatial functor"); return DDPS_ERR_ARG; }
  static const int want[] = {0, 1, 3, 3, 3};
  const bool by_tag = kind == DDPS_SPATIAL_BY_TAG;
  if (kind < DDPS_SPATIAL_CONST || kind > DDPS_SPATIAL_BY_TAG || !params || (!by_tag && nparams != want[kind]) ||
      (by_tag && (nparams < 1 || nparams > 15 || nparams % 2 == 0))) {
    ctx->fail(DDPS_ERR_ARG, "unknown spatial functor or wrong parameter count (host-sample the closure and use ddps_coeff_set)");
    return DDPS_ERR_ARG;
  }
  if (by_tag && sk == 2) { ctx->fail(DDPS_ERR_ARG, "BY_TAG functors exist per triangle only"); return DDPS_ERR_ARG; }
  SpatialFn f;
  memset(&f, 0, sizeof(f));
  f.kind = kind;
  f.ntag = by_tag ? (nparams - 1) / 2 : 0;
  for (int i = 0; i < nparams; ++i) {
    if (!std::isfinite(params[i])) { ctx->fail(DDPS_ERR_ARG, "spatial functor parameter is not finite"); return DDPS_ERR_ARG; }
    f.p[i] = params[i];
  }
  const size_t nloc = sk == 0 ? (size_t)ctx->nme : sk == 1 ? (size_t)3 * ctx->nme : (size_t)ctx->n_loc;
  if ((rc = dev_alloc(ctx, &ctx->coeff[slot], nloc))) return rc;
  ctx->coeff_len[slot] = (int64_t)nloc;
  if (sk == 2) {
    if (ctx->n_loc) k_sample_nodes<<<(ctx->n_loc + 255) / 256, 256, 0, ctx->stream>>>(ctx->n_loc, ctx->xy, f, ctx->coeff[slot]);
  } else if (ctx->nme) {
    k_sample_elements<<<(ctx->nme + 255) / 256, 256, 0, ctx->stream>>>(ctx->nme, ctx->elem, ctx->xy, f, sk, slot == DDPS_COEFF_LAMBDA2,
                                                                      ctx->coeff[slot]);
  }
  DDPS_CUDA(cudaGetLastError());
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (slot == DDPS_COEFF_LAMBDA2 || slot == DDPS_COEFF_A11 || slot == DDPS_COEFF_A22) ctx->base_ready = false;
  return DDPS_OK;
}

int ddps_functor_set(ddps_ctx* ctx, int kind, const double* params, int nparams) {
  if (!ctx) return DDPS_ERR_ARG;
  if (kind == DDPS_FUNCTOR_SINH) {
    if (nparams != 2 || !params) { ctx->fail(DDPS_ERR_ARG, "SINH functor takes {alpha,beta}"); return DDPS_ERR_ARG; }
  } else if (kind == DDPS_FUNCTOR_POLY) {
    if (nparams != 1 || !params || params[0] < 0 || params[0] >= kMaxPoly || params[0] != floor(params[0])) {
      ctx->fail(DDPS_ERR_ARG, "POLY functor takes {degree}, degree in 0..4");
      return DDPS_ERR_ARG;
    }
  } else {
    ctx->fail(DDPS_ERR_ARG, "unknown functor kind (arbitrary closures cannot run on the device; no CPU fallback)");
    return DDPS_ERR_ARG;
  }
  ctx->functor_kind = kind;
  for (int i = 0; i < nparams && i < 8; ++i) ctx->functor_par[i] = params[i];
  return DDPS_OK;
}

// ---- solve_ddpe ---------------------------------------------------------------------------------
int ddps_ddpe_begin(ddps_ctx* ctx, int rec_mode, double delta, double tau_p, double tau_n, double tol) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if (rec_mode != DDPS_REC_ZERO && rec_mode != DDPS_REC_SHOCKLEY) { ctx->fail(DDPS_ERR_ARG, "rec_mode must be :shockley or :zero"); return DDPS_ERR_ARG; }
  for (int f = 0; f < 3; ++f)
    if (!ctx->have_D[f]) { ctx->fail(DDPS_ERR_STATE, "Dirichlet data for V, u and v must be set"); return DDPS_ERR_STATE; }
  if ((rc = need_coeff(ctx, DDPS_COEFF_MU_P, "MU_P")) || (rc = need_coeff(ctx, DDPS_COEFF_MU_N, "MU_N")) ||
      (rc = need_coeff(ctx, DDPS_COEFF_C_MID, "C_MID")))
    return rc;
  ctx->rec_mode = rec_mode; ctx->delta = delta; ctx->tau_p = tau_p; ctx->tau_n = tau_n; ctx->tol = tol;
  reset_stats(ctx);
  if ((rc = assemble_base(ctx, false))) return rc;                              // src:643, once per solve
  {
    const double ta = now_ms();
    if ((rc = maybe_amg_setup(ctx))) return rc;
    ctx->stats.t_amg_setup_ms = now_ms() - ta;
  }
  const size_t bytes = (size_t)ctx->n_loc * sizeof(double);
  DDPS_CUDA(cudaMemsetAsync(ctx->U, 0, bytes, ctx->stream));                     // src:635-637
  DDPS_CUDA(cudaMemsetAsync(ctx->W, 0, bytes, ctx->stream));
  DDPS_CUDA(cudaMemsetAsync(ctx->V, 0, bytes, ctx->stream));
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < Context::kNewtonWarm; ++k) ctx->newton_x_valid[k] = false;   // a new solve: no earlier corrections
  ctx->ddpe_active = true;
  return DDPS_OK;
}

int ddps_ddpe_step(ddps_ctx* ctx, int nsteps, int stop_at_tol, double* control, int* steps_done) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (!ctx->ddpe_active) { ctx->fail(DDPS_ERR_STATE, "call ddps_ddpe_begin first"); return DDPS_ERR_STATE; }
  int done = 0, rc = 0;
  double t0 = now_ms();
  cudaEventRecord(ctx->ev4, ctx->stream);
  for (int k = 0; k < nsteps; ++k) {
    double c = 0;
    if ((rc = gummel_step(ctx, &c))) break;
    ++done;
    ctx->stats.outer_iters++;
    ctx->stats.control = c;
    ctx->hist_outer.push_back(c);
    if (control) control[k] = c;
    if (ctx->opts.verbose) printf("iteration %d, tolerance: %.17g\n", ctx->stats.outer_iters, c);   // src:655
    if (stop_at_tol && !(c > ctx->tol)) break;                                                     // src:649
  }
  cudaEventRecord(ctx->ev5, ctx->stream);
  cudaEventSynchronize(ctx->ev5);
  float dev_ms = 0;
  cudaEventElapsedTime(&dev_ms, ctx->ev4, ctx->ev5);
  ctx->stats.t_solve_ms += dev_ms;                 // CUDA events on the library stream
  ctx->stats.t_wall_ms += now_ms() - t0;           // host wall clock around the same region
  if (steps_done) *steps_done = done;
  if (!rc) rc = p2p_check_error(ctx);
  ctx->stats.status = rc;
  return rc;
}

int ddps_ddpe_fetch(ddps_ctx* ctx, double* V, double* u, double* v) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if (V && (rc = download_nodal(ctx, ctx->V, V))) return rc;
  if (u && (rc = download_nodal(ctx, ctx->U, u))) return rc;
  if (v && (rc = download_nodal(ctx, ctx->W, v))) return rc;
  return DDPS_OK;
}

int ddps_solve_ddpe(ddps_ctx* ctx, int rec_mode, double delta, double tau_p, double tau_n, double tol, double* V,
                    double* u, double* v, ddps_stats* stats) {
  if (!ctx) return DDPS_ERR_ARG;
  int rc = ddps_ddpe_begin(ctx, rec_mode, delta, tau_p, tau_n, tol);
  if (rc) return rc;
  const int cap = ctx->opts.max_gummel > 0 ? ctx->opts.max_gummel : 1000;
  double control = 1.0;   // src:645
  int total = 0;
  while (control > tol) {
    if (total >= cap) { ctx->fail(DDPS_ERR_NOCONV, "Gummel iteration cap reached"); rc = DDPS_ERR_NOCONV; break; }
    int done = 0;
    rc = ddps_ddpe_step(ctx, 1, 1, &control, &done);
    if (rc) break;
    ++total;
  }
  ctx->stats.status = rc;
  int rc2 = ddps_ddpe_fetch(ctx, V, u, v);
  if (stats) *stats = ctx->stats;
  return rc ? rc : rc2;
}

// ---- solve_semlin_poisson ---------------------------------------------------------------------
int ddps_solve_semlin(ddps_ctx* ctx, double tol, double* Vout, ddps_stats* stats) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  reset_stats(ctx);
  ctx->ddpe_active = false;
  double t0 = now_ms();
  if ((rc = assemble_base(ctx, true))) return rc;                                            // src:924-1028
  {
    const double ta = now_ms();
    if ((rc = maybe_amg_setup(ctx))) return rc;
    ctx->stats.t_amg_setup_ms = now_ms() - ta;
  }
  DDPS_CUDA(cudaMemsetAsync(ctx->V, 0, (size_t)ctx->n_loc * sizeof(double), ctx->stream));   // src:948
  int count = 0;
  rc = newton_loop(ctx, true, 0.0, tol, nullptr, nullptr,
                   [&](double res, int it) {
                     ctx->hist_outer.push_back(res);
                     ctx->stats.control = res;
                     if (it > 0 && ctx->opts.verbose) printf("iteration %d, tolerance: %.17g\n", it, res);   // src:1093
                   },
                   &count);
  ctx->stats.outer_iters = count;
  ctx->stats.newton_total = count;
  ctx->stats.status = rc;
  ctx->stats.t_solve_ms = now_ms() - t0;
  int rc2 = Vout ? download_nodal(ctx, ctx->V, Vout) : 0;
  if (stats) *stats = ctx->stats;
  return rc ? rc : rc2;
}

int ddps_get_history(ddps_ctx* ctx, int kind, double* out, int64_t cap, int64_t* n) {
  if (!ctx || !n) return DDPS_ERR_ARG;
  const std::vector<double>& h = kind == 0 ? ctx->hist_outer : kind == 1 ? ctx->hist_newton : ctx->hist_pcg;
  *n = (int64_t)h.size();
  if (out) for (int64_t i = 0; i < cap && i < (int64_t)h.size(); ++i) out[i] = h[i];
  return DDPS_OK;
}

int ddps_get_stats(ddps_ctx* ctx, ddps_stats* stats) {
  if (!ctx || !stats) return DDPS_ERR_ARG;
  *stats = ctx->stats;
  return DDPS_OK;
}

// ---- parity / debug exports -------------------------------------------------------------------
int ddps_get_csr_size(ddps_ctx* ctx, int64_t* nrows, int64_t* nnz) {
  if (!ctx) return DDPS_ERR_ARG;
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if (nrows) *nrows = ctx->pat.nrows;
  if (nnz) *nnz = ctx->pat.nnz;
  return DDPS_OK;
}

int ddps_get_csr(ddps_ctx* ctx, int which, int64_t* rowptr, int64_t* col, double* val) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  const Pattern& P = ctx->pat;
  std::vector<int> rp(P.nrows + 1), cc(P.nnz), sp(P.nslices + 1);
  DDPS_CUDA(cudaMemcpy(rp.data(), P.rowptr, rp.size() * sizeof(int), cudaMemcpyDeviceToHost));
  DDPS_CUDA(cudaMemcpy(cc.data(), P.csrcol, cc.size() * sizeof(int), cudaMemcpyDeviceToHost));
  DDPS_CUDA(cudaMemcpy(sp.data(), P.slice_ptr, sp.size() * sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<double> vals;
  if (val) {
    const double* src = which == DDPS_MAT_BASE ? ctx->val_base : ctx->val_op;
    std::vector<double> sv(P.nentries);
    DDPS_CUDA(cudaMemcpy(sv.data(), src, sv.size() * sizeof(double), cudaMemcpyDeviceToHost));
    vals.resize(P.nnz);
    for (int r = 0; r < P.nrows; ++r) {
      int base = sp[r >> 5], lane = r & 31;
      for (int k = 0; k < rp[r + 1] - rp[r]; ++k) vals[rp[r] + k] = sv[(size_t)base + 32 * k + lane];
    }
  }
  if (ctx->nranks == 1 && !ctx->loc2glob.empty()) {
    // renumbered mesh: hand the matrix back in the CALLER's numbering (rows and columns permuted, columns re-sorted)
    const int n = P.nrows;
    std::vector<int> int_of_user(n);
    for (int i = 0; i < n; ++i) int_of_user[ctx->loc2glob[i]] = i;
    int64_t pos = 0;
    std::vector<std::pair<int64_t, double>> rowbuf;
    for (int ur = 0; ur < n; ++ur) {
      const int r = int_of_user[ur];
      if (rowptr) rowptr[ur] = pos;
      rowbuf.clear();
      for (int k = rp[r]; k < rp[r + 1]; ++k) rowbuf.emplace_back(ctx->loc2glob[cc[k]], val ? vals[k] : 0.0);
      std::sort(rowbuf.begin(), rowbuf.end());
      for (auto& e : rowbuf) { if (col) col[pos] = e.first; if (val) val[pos] = e.second; ++pos; }
    }
    if (rowptr) rowptr[n] = pos;
    return DDPS_OK;
  }
  if (rowptr) for (size_t i = 0; i < rp.size(); ++i) rowptr[i] = rp[i];
  if (col) for (size_t i = 0; i < cc.size(); ++i) col[i] = ctx->loc2glob.empty() ? cc[i] : ctx->loc2glob[cc[i]];
  if (val) for (int64_t i = 0; i < P.nnz; ++i) val[i] = vals[i];
  return DDPS_OK;
}

int ddps_get_vector(ddps_ctx* ctx, int which, double* out, int64_t len) {
  if (!ctx || !out) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  const double* src = nullptr;
  switch (which) {
    case DDPS_VEC_RHS: src = ctx->b; break;
    case DDPS_VEC_RESID: src = ctx->resid; break;
    case DDPS_VEC_V: src = ctx->V; break;
    case DDPS_VEC_U: src = ctx->U; break;
    case DDPS_VEC_W: src = ctx->W; break;
    case DDPS_VEC_LIFT: src = ctx->lift; break;
    default: ctx->fail(DDPS_ERR_ARG, "unknown vector"); return DDPS_ERR_ARG;
  }
  if (len != ctx->n_own) { ctx->fail(DDPS_ERR_ARG, "vector length must equal the number of owned rows"); return DDPS_ERR_ARG; }
  if (ctx->nranks == 1) return download_owned(ctx, src, out);
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  DDPS_CUDA(cudaMemcpy(out, src, (size_t)len * sizeof(double), cudaMemcpyDeviceToHost));
  return DDPS_OK;
}

int ddps_get_coeff(ddps_ctx* ctx, int slot, double* out, int64_t cap, int64_t* len) {
  if (!ctx || !len) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if (slot < 0 || slot >= 64 || !ctx->coeff[slot]) { ctx->fail(DDPS_ERR_STATE, "coefficient slot is empty"); return DDPS_ERR_STATE; }
  if (ctx->nranks != 1 || !ctx->loc2glob.empty()) { ctx->fail(DDPS_ERR_STATE, "debug exports are single-GPU, caller's numbering"); return DDPS_ERR_STATE; }
  *len = ctx->coeff_len[slot];
  if (!out) return DDPS_OK;
  if (cap < *len) { ctx->fail(DDPS_ERR_ARG, "output buffer too small"); return DDPS_ERR_ARG; }
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  DDPS_CUDA(cudaMemcpy(out, ctx->coeff[slot], (size_t)*len * sizeof(double), cudaMemcpyDeviceToHost));
  return DDPS_OK;
}

int ddps_debug_assemble_base(ddps_ctx* ctx, int use_semlin) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if ((rc = assemble_base(ctx, use_semlin != 0))) return rc;
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return maybe_amg_setup(ctx);
}

int ddps_debug_assemble_newton(ddps_ctx* ctx, int use_semlin, double delta, const double* V, const double* u,
                               const double* v, double* resid_inf) {
  if (!ctx || !V) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if (!ctx->base_ready || ctx->base_is_semlin != (use_semlin != 0)) { ctx->fail(DDPS_ERR_STATE, "assemble the matching base matrix first"); return DDPS_ERR_STATE; }
  if ((rc = upload_nodal(ctx, ctx->V, V))) return rc;
  if (!use_semlin) {
    if (!u || !v) return DDPS_ERR_ARG;
    if ((rc = upload_nodal(ctx, ctx->U0, u)) || (rc = upload_nodal(ctx, ctx->W0, v))) return rc;
  }
  if ((rc = assemble_newton(ctx, use_semlin != 0, delta, ctx->V, ctx->U0, ctx->W0))) return rc;
  double res;
  if ((rc = fetch_resid_norm(ctx, &res))) return rc;
  if (resid_inf) *resid_inf = res;
  return DDPS_OK;
}

int ddps_debug_assemble_uv(ddps_ctx* ctx, int which, int rec_mode, double tau_p, double tau_n, const double* V,
                           const double* u0, const double* v0) {
  if (!ctx || !V || !u0 || !v0) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if (which != DDPS_FIELD_U && which != DDPS_FIELD_W) { ctx->fail(DDPS_ERR_ARG, "which must be U or W"); return DDPS_ERR_ARG; }
  if ((rc = upload_nodal(ctx, ctx->V, V)) || (rc = upload_nodal(ctx, ctx->U0, u0)) || (rc = upload_nodal(ctx, ctx->W0, v0))) return rc;
  if ((rc = assemble_uv(ctx, which, rec_mode, tau_p, tau_n, ctx->V, ctx->U0, ctx->W0))) return rc;
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return DDPS_OK;
}

int ddps_debug_spmv(ddps_ctx* ctx, int which, const double* x, double* y) {
  if (!ctx || !x || !y) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if ((rc = upload_nodal(ctx, ctx->p, x))) return rc;
  if ((rc = launch_spmv(ctx, which == DDPS_MAT_BASE ? ctx->val_base : ctx->val_op, ctx->p, ctx->q))) return rc;
  if (ctx->nranks == 1) return download_owned(ctx, ctx->q, y);
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  DDPS_CUDA(cudaMemcpy(y, ctx->q, (size_t)ctx->n_own * sizeof(double), cudaMemcpyDeviceToHost));
  return DDPS_OK;
}

int ddps_debug_pcg(ddps_ctx* ctx, const double* rhs, double* x, double rtol, int64_t max_it, int64_t* iters, double* relres) {
  if (!ctx || !rhs || !x) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if (ctx->nranks != 1) { ctx->fail(DDPS_ERR_STATE, "debug exports are single-GPU"); return DDPS_ERR_STATE; }
  if ((rc = upload_owned(ctx, ctx->b, rhs)) || (rc = upload_owned(ctx, ctx->x, x))) return rc;
  rc = solve_system(ctx, ctx->val_op, ctx->dinv, ctx->b, ctx->x, true, rtol, 0.0, max_it, iters, relres);
  int rc3 = download_owned(ctx, ctx->x, x);
  return rc ? rc : rc3;
}

// ---- multigrid debug exports -----------------------------------------------------------------------
int ddps_amg_levels(ddps_ctx* ctx, int64_t* nlevels, int64_t* sizes) {
  if (!ctx || !nlevels) return DDPS_ERR_ARG;
  return amg_debug_levels(ctx, nlevels, sizes, 16);
}

int ddps_amg_numeric(ddps_ctx* ctx, int which) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if (!amg_ready(ctx)) { ctx->fail(DDPS_ERR_STATE, "no multigrid hierarchy (set opts.precond = DDPS_PRECOND_AMG before the base assembly)"); return DDPS_ERR_STATE; }
  if (which == DDPS_MAT_BASE) {
    // inverse diagonal of the base operator: the assembled dinv belongs to DDPS_MAT_OP, so take it from a unit solve
    ctx->fail(DDPS_ERR_ARG, "numeric phase runs on DDPS_MAT_OP (assemble an operator first)");
    return DDPS_ERR_ARG;
  }
  if ((rc = amg_numeric(ctx, ctx->val_op, ctx->dinv, -1))) return rc;
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return DDPS_OK;
}

int ddps_amg_get_values(ddps_ctx* ctx, int level, double* val, double* dinv) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  return amg_debug_get_values(ctx, level, val, dinv);
}

int ddps_amg_get_level(ddps_ctx* ctx, int level, int64_t sizes[4], int64_t* rowptr, int64_t* col, int64_t* Pptr, int64_t* Pcol,
                       double* Pval) {
  if (!ctx) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  return amg_debug_get_level(ctx, level, sizes, rowptr, col, Pptr, Pcol, Pval);
}

int ddps_amg_apply(ddps_ctx* ctx, const double* r, double* z) {
  if (!ctx || !r || !z) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  if (!amg_ready(ctx)) { ctx->fail(DDPS_ERR_STATE, "no multigrid hierarchy"); return DDPS_ERR_STATE; }
  if (ctx->nranks != 1) { ctx->fail(DDPS_ERR_STATE, "debug exports are single-GPU"); return DDPS_ERR_STATE; }
  if ((rc = upload_owned(ctx, ctx->r, r))) return rc;
  DDPS_CUDA(cudaMemsetAsync(&ctx->scal->done, 0, sizeof(int), ctx->stream));
  if ((rc = amg_vcycle(ctx, ctx->r, ctx->q, false))) return rc;
  return download_owned(ctx, ctx->q, z);
}

// ---- kernel timing on the library stream --------------------------------------------------------
int ddps_time_kernel(ddps_ctx* ctx, int kernel, int warm, int reps, double* ms_per_launch, double* algo_bytes) {
  if (!ctx || reps <= 0) return DDPS_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc;
  if ((rc = require_mesh(ctx))) return rc;
  const Pattern& P = ctx->pat;
  const double nnz = (double)P.nnz, nq = (double)ctx->n_own, nme = (double)ctx->nme;
  double bytes = 0;
  auto run = [&](int it) -> int {
    switch (kernel) {
      case 0: return launch_spmv(ctx, ctx->last_val ? ctx->last_val : ctx->val_op, ctx->p, ctx->q);
      case 1: return assemble_uv(ctx, DDPS_FIELD_U, ctx->rec_mode, ctx->tau_p, ctx->tau_n, ctx->V, ctx->U0, ctx->W0);
      case 2: return pcg_iteration_launch(ctx, ctx->last_val ? ctx->last_val : ctx->val_op, ctx->last_val ? ctx->last_dinv : ctx->dinv, ctx->x, it, 1 << 30, it == 1);
      case 3: return assemble_newton(ctx, false, ctx->delta, ctx->V, ctx->U0, ctx->W0);
      case 4: return assemble_base(ctx, ctx->base_is_semlin);
      case 5: return launch_spmv_dot(ctx, ctx->last_val ? ctx->last_val : ctx->val_op, ctx->p, ctx->q);
      case 6: return amg_pcg_iteration_launch(ctx, ctx->x, it, 1 << 30);
      case 7: return amg_numeric(ctx, ctx->val_op, ctx->dinv, -1);
      case 8: return amg_vcycle(ctx, ctx->r, ctx->q, false);
      case 9: case 10: case 11: case 12: case 13: return amg_time_launch(ctx, kernel, it, nullptr);
      case 14: return launch_prep_poisson(ctx, ctx->V, ctx->U0, ctx->W0, ctx->delta * ctx->delta);   // the nodal passes alone
      case 15: return launch_prep_uv(ctx, 1.0, ctx->V, ctx->U0, ctx->W0, ctx->tau_p, ctx->tau_n, ctx->rec_mode);
    }
    ctx->fail(DDPS_ERR_ARG, "unknown kernel id");
    return DDPS_ERR_ARG;
  };
  // algorithmic bytes per launch: every input read once, every output written once,
  // FP64 values, int32 indices
  switch (kernel) {
    case 0: bytes = 12.0 * nnz + 4.0 * (nq + 1) + 16.0 * nq; break;                     // 14.86 B/nnz
    case 1: bytes = 12.0 * nme + 16.0 * nq + 3 * 8.0 * nq + 8.0 * nme + 8.0 * nnz + 8.0 * nq; break;   // 72 B/triangle
    case 2: bytes = 12.0 * nnz + 4.0 * (nq + 1) + 16.0 * nq + 10 * 8.0 * nq; break;     // SURVEY 8d PCG iteration
    case 3: bytes = 12.0 * nme + 16.0 * nq + 3 * 8.0 * nq + 24.0 * nme + 12.0 * nnz + 8.0 * nnz + 16.0 * nq; break;  // 134 B/triangle
    case 4: bytes = 12.0 * nme + 16.0 * nq + 8.0 * nme + 8.0 * nnz + 8.0 * nq; break;
    case 5: bytes = 12.0 * nnz + 4.0 * (nq + 1) + 16.0 * nq; break;
    case 14: case 15: bytes = (3 * 8.0 + 8.0 * kRecD) * (double)ctx->n_loc; break;
  }
  if (kernel >= 6 && kernel <= 13) {
    if (!amg_ready(ctx)) { ctx->fail(DDPS_ERR_STATE, "no multigrid hierarchy (opts.precond = DDPS_PRECOND_AMG and a base assembly first)"); return DDPS_ERR_STATE; }
    // fresh state on the last assembled system: one capped solve leaves r, p, z and the scalars consistent
    const int32_t keep = ctx->opts.lin_cap_ok;
    ctx->opts.lin_cap_ok = 1;
    rc = amg_pcg_solve(ctx, ctx->val_op, ctx->dinv, ctx->b, ctx->x, false, 0.0, -1.0, 2, nullptr, nullptr);
    ctx->opts.lin_cap_ok = keep;
    if (rc) return rc;
    DDPS_CUDA(cudaMemsetAsync(&ctx->scal->done, 0, sizeof(int), ctx->stream));
    if (kernel == 6) bytes = amg_iteration_bytes(ctx);
    if (kernel >= 9) {
      if ((rc = amg_time_launch(ctx, kernel, 0, &bytes))) { ctx->fail(DDPS_ERR_ARG, "unknown kernel id"); return rc; }
    }
  }
  if (kernel == 5) DDPS_CUDA(cudaMemsetAsync(&ctx->scal->done, 0, sizeof(int), ctx->stream));
  if (kernel == 2) {
    // fresh PCG state on the last assembled system; rtol 0 keeps the done flag down
    if ((rc = halo_exchange(ctx, ctx->p))) return rc;
    rc = pcg_solve(ctx, ctx->last_val ? ctx->last_val : ctx->val_op, ctx->last_val ? ctx->last_dinv : ctx->dinv,
                   (ctx->val_s && ctx->last_val == ctx->val_s) ? ctx->bs : ctx->b, ctx->x, false, 0.0, -1.0, 1, nullptr, nullptr);
    if (rc != DDPS_ERR_NOCONV && rc != 0) return rc;
    DDPS_CUDA(cudaMemsetAsync(&ctx->scal->done, 0, sizeof(int), ctx->stream));
  }
  int it = 1;
  for (int i = 0; i < warm; ++i, ++it) if ((rc = run(it))) return rc;
  DDPS_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int i = 0; i < reps; ++i, ++it) if ((rc = run(it))) return rc;
  DDPS_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  DDPS_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0;
  DDPS_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  if (ms_per_launch) *ms_per_launch = ms / reps;
  if (algo_bytes) *algo_bytes = bytes;
  return DDPS_OK;
}

}  // extern "C"
