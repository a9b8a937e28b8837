// Internal structures of libddps_b200.so (not part of the C ABI).
//
// Layout decisions:
//  * node coordinates: double2 per node (x,y interleaved exactly like Mesh.nodes, src:11) -> one
//    16-byte load per corner.
//  * triangles: int4 {n1,n2,n3,_} with 0-based LOCAL node ids -> one 16-byte load per triangle.
//  * sparsity pattern: CSR with sorted columns, stored slice-interleaved ("SELL-32"): rows are
//    grouped in slices of 32 (one warp), every slice is padded to its longest row and stored
//    column-major inside the slice, so lane l of a warp reads entry k of row 32*s+l at
//    slice_ptr[s] + 32*k + l  -> perfectly coalesced value/column streams for both the row-major
//    assembly (writes) and the SpMV (reads).  Padding entries carry value 0 and column = own row.
//  * row -> incident triangle lists ("incidence"), also slice-interleaved, as three parallel streams: the
//    triangle id + the row's local corner (0..2); the pre-computed slots of that triangle's off-diagonal
//    row entries + the positions of its two other corners in the SLICE's node list (4 bytes); the signed
//    area in the row's own frame.  The scatter needs no search and no atomics.
//  * per slice of 32 rows: the sorted list of distinct nodes its rows touch, staged once per slice in
//    shared memory by the warp that assembles the slice (cp.async gathers of 48-byte nodal records).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ddps.h"

#ifdef DDPS_WITH_NCCL
#include <nccl.h>
#endif

namespace ddps {

constexpr int kSlice = 32;        // rows per slice = warp size
constexpr int kAsmBlock = 128;    // threads per block of the assembly kernels: 4 independent warps, one slice each
constexpr int kMaxRowLen = 40;    // shared-memory accumulator columns available per row
constexpr int kVecBlock = 256;    // threads per block of SpMV / vector kernels
constexpr int kMaxPoly = 5;       // polynomial functor: degree 0..4

#define DDPS_CUDA(call)                                                                         \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      ctx->fail(DDPS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));            \
      return DDPS_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

struct Pattern {
  int nrows = 0;        // owned rows
  int ncols = 0;        // local nodes (owned + ghost)
  int nslices = 0;
  int64_t nnz = 0;      // true nonzeros of the owned rows
  int64_t nentries = 0; // stored entries incl. padding (= slice_ptr[nslices])
  int max_row_len = 0;
  int max_inc = 0;      // largest number of triangles around a node
  int* slice_ptr = nullptr;  // [nslices+1]
  int* col = nullptr;        // [nentries] slice-interleaved
  int* rowptr = nullptr;     // [nrows+1] plain CSR (export + slot search)
  int* csrcol = nullptr;     // [nnz]
  // incidence streams (setup.cu: k_fill_incidence)
  int64_t inc_entries = 0;
  int* inc_slice_ptr = nullptr;  // [nslices+1]
  int* inc_tri = nullptr;        // [inc_entries] tri*4+corner | -1
  unsigned* inc_pack = nullptr;  // [inc_entries] slot_next | slot_prev<<5 | l_next<<10 | l_prev<<20 | again<<30,31 | 0xffffffff
  double* arinc = nullptr;       // [inc_entries] signed area in the row's own frame
  unsigned short* rowinfo = nullptr;   // [nrows] slot of the diagonal entry | row length << 5
  // per slice: sorted distinct nodes the slice's rows touch + the packed descriptor the assembling warp starts from
  int max_sn = 0, sn_total = 0;
  int* sn_ptr = nullptr;         // [nslices+1]
  int* sn_nodes = nullptr;       // [sn_total]
  int64_t nruns = 0;
  int2* runs = nullptr;          // runs of consecutive node ids of every slice list: {first node, position | length<<16}
  int4* sdesc = nullptr;         // [nslices][3] {base, ibase, nbase, w | iw<<6 | nn<<12 | l_first<<22}, {run base, #runs, run0}, {run1, run2}
};

// device-side scalar block of one PCG solve (all control flow stays on the device)
struct PcgScalars {
  double rz[2];     // r.z, ping-pong by iteration parity
  double pq;        // p.Ap
  double rr;        // r.r
  double bb;        // rhs.rhs
  double thresh2;   // stop when rr <= thresh2
  double sums[4];   // staging for allreduce: {pq} or {rz_new, rr}
  int done;
  int iters;
  int breakdown;
  int pad0;
  unsigned cnt[4];  // last-block tickets
  double norm_inf;  // generic max-norm result
  double rmax;      // ||r||_inf of the PCG residual
  double thresh_inf;
  double norm_aux2;
  // multigrid-preconditioned CG (amg.cu): energy-norm stop rule of warm-started solves
  double thresh_e;  // stop only when r.Br <= thresh_e
  double xb;        // x0.rhs ~ ||x||_A^2
  int warm;
  int conv;         // 1 once every stop rule is met (0 after an exhausted iteration cap)
  // error-based stop rule (both preconditioners): the A-norm of the error must fall below eps_e * sqrt(E_ref), E_ref = the
  // largest solution energy seen (x0.rhs of a warm start, the energy of the largest Newton correction of the running
  // Newton loop).  Multigrid path: r.Br ~ e.Ae directly.  Jacobi path: Hestenes-Stiefel estimate, the energy the last
  // hs_d updates added, sum alpha_j r_j.z_j (= ||x_k - x_{k-d}||_A^2 <= ||x - x_{k-d}||_A^2).
  double eps_e2;    // eps_e^2 (0: rule off)
  double e_ref;     // reference energy carried across the solves of one Newton loop
  double hs_total;  // energy added so far in this solve
  double hs_ring[32];
  int hs_d;
  int pad2;
};
constexpr int kHsDelay = 32;

// ---- peer-memory (NVLink P2P via CUDA IPC) windows ---------------------------------------------------------------------
// Two users: (a) the Jacobi-PCG iteration pushes the halo of p and exchanges its dot products from INSIDE its three kernels
// (halo_flag / redA / redB; kernels.cu); (b) the generic staged exchange every other vector of every multigrid level goes
// through (xflag + the staging buffers) and the generic scalar allreduce (rflag / rval) -- one tiny kernel each, so no full
// grid ever spins on a peer (comm.cu).
constexpr int kMaxRanks = 8;
struct P2PWin {                                  // one per rank, written by the peers
  unsigned long long halo_flag[kMaxRanks];       // [src] tag of the last halo src pushed into my ghost block (Jacobi path)
  unsigned long long redA_flag[kMaxRanks];       // [src] tag of src's p.Ap partial
  unsigned long long redB_flag[kMaxRanks];       // [src] tag of src's {r.z, r.r, max|r|} partials
  double redA[kMaxRanks][4];
  double redB[kMaxRanks][4];
  unsigned long long xflag[kMaxRanks];           // [src] tag of the last staged halo block src wrote into my staging buffer
  unsigned long long rflag[2][kMaxRanks];        // generic scalar allreduce, double-buffered by tag parity
  double rval[2][kMaxRanks][4];
  unsigned xticket[kMaxRanks];                   // slice tickets of the staged exchange (local use)
  int error;
  int pad;
};
struct P2PDev {                                  // by-value kernel argument
  int nranks, rank, nneigh;
  int neigh[kMaxRanks];
  P2PWin* win[kMaxRanks];                        // win[rank] is local, the others are IPC mappings
  unsigned long long tag_halo, tagA, tagB;
};
struct PushArgs {
  int nneigh;
  int off[kMaxRanks + 1];
  double* dst[kMaxRanks];
  unsigned long long* flag[kMaxRanks];
  unsigned long long tag;
};
struct P2PHost {
  bool enabled = false;                          // (a) fused Jacobi-PCG path
  bool xenabled = false;                         // (b) staged exchange + scalar allreduce
  P2PWin* win_local = nullptr;
  P2PWin* win_peer[kMaxRanks] = {nullptr};
  double* p_peer[kMaxRanks] = {nullptr};
  double* stage_local = nullptr;                 // [2][stage_cap] receive staging, double-buffered by tag parity
  double* stage_peer[kMaxRanks] = {nullptr};
  size_t stage_cap = 0;
  bool ipc = false;                              // peer pointers are IPC mappings (to be closed), not in-process pointers
  PushArgs push;
  P2PDev dev;
  unsigned long long tag = 0, xtag = 0, rtag = 0;
};

// halo plan of one level: owned entries first, ghosts grouped by owner; the send list to peer q is exactly q's ghost block
// for me, in the same order
struct HaloPlan {
  int n_own = 0, n_ghost = 0;
  int nneigh = 0;
  std::vector<int> peer;        // neighbour ranks
  std::vector<int> send_off;    // [nneigh+1] offsets into send_idx
  std::vector<int> recv_off;    // [nneigh+1] offsets into the ghost block (relative to n_own)
  std::vector<int> remote_off;  // [nneigh] offset of my block inside the peer's ghost block (= its staging offset)
  int* send_idx = nullptr;      // device: local owned ids to pack, grouped by neighbour
  double* send_buf = nullptr;   // device
  int nsend = 0;
  bool p2p = false;             // staged peer-memory exchange usable for this plan (every rank agrees)
};

// communication backends: NCCL (one process per GPU) or an in-process group (ranks = host threads sharing one device; lets
// the whole distributed logic run, and be tested, on a single GPU)
struct LocalGroup;
enum { COMM_SUM = 0, COMM_MAX = 1, COMM_MIN = 2 };
struct CommMsg { int peer; const void* send; size_t send_bytes; void* recv; size_t recv_bytes; };

struct Amg;   // amg.cu: device-side smoothed-aggregation hierarchy

struct Context {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev4 = nullptr, ev5 = nullptr;
  int num_sms = 148;
  int max_smem_optin = 227 * 1024;
  std::string err;
  int err_code = 0;
  ddps_solver_opts opts;
  ddps_stats stats;

  // communicator
  int rank = 0, nranks = 1;
#ifdef DDPS_WITH_NCCL
  ncclComm_t comm = nullptr;
#endif
  std::shared_ptr<LocalGroup> lgroup;   // in-process communicator (ddps_comm_init with a "LOCAL:<name>" id)
  HaloPlan halo;
  P2PHost p2p;

  // mesh (local part)
  bool have_mesh = false;
  bool local_ingest = false;          // mesh came through ddps_mesh_set_local (no rank holds the global mesh)
  int64_t nq_global = 0, nme_global = 0;
  int n_own = 0, n_loc = 0, nme = 0;
  double2* xy = nullptr;
  int4* elem = nullptr;
  std::vector<int64_t> loc2glob;      // [n_loc] global 0-based node id of every local node
  std::vector<int64_t> elem_glob;     // [nme] global element id of every local element
  long long* d_own_glob = nullptr;    // device copy of loc2glob[0:n_own] (gather of results across ranks)
  Pattern pat;

  // boundary data
  unsigned char* isD = nullptr;       // [n_loc]
  double* gD[4] = {nullptr, nullptr, nullptr, nullptr};  // [n_loc] per ddps_field
  bool have_D[4] = {false, false, false, false};
  std::vector<int64_t> Dset;          // sorted global Dirichlet ids (from the first dirichlet_set)

  // sampled coefficients: slot -> device array (element slots hold local elements only)
  double* coeff[64] = {nullptr};
  int64_t coeff_len[64] = {0};
  int functor_kind = 0;
  double functor_par[8] = {0};

  // matrices (values in SELL order)
  double* val_base = nullptr;  // MV / semlin M
  double* val_op = nullptr;    // operator handed to the solver
  double* val_s = nullptr;     // symmetrically Jacobi-scaled copy D^-1/2 A D^-1/2 the CG actually iterates on
  double* sc = nullptr;        // [n_loc] scale factors D^-1/2
  double* bs = nullptr;        // [n_own] scaled right-hand side
  const double* last_val = nullptr;   // matrix / inverse diagonal of the last solve (kernel timing)
  const double* last_dinv = nullptr;
  double* lift = nullptr;      // [n_own] K[:,D]*gD of the base stiffness
  bool base_ready = false;
  bool base_is_semlin = false;

  // nodal work arrays [n_loc]
  double *V = nullptr, *U = nullptr, *W = nullptr;          // fields
  double *V0 = nullptr, *U0 = nullptr, *W0 = nullptr;       // previous Gummel iterate
  double *Unew = nullptr;                                   // scratch field
  double* rec = nullptr;                                    // [n_loc][kRecD] nodal records of the running assembly (prep kernels)
  double* cinc = nullptr;                                   // [inc_entries] C_MID / H_MID in incidence order
  double* cinc_poly[kMaxPoly] = {nullptr};
  // solver vectors
  double *b = nullptr, *resid = nullptr, *dinv = nullptr;   // [n_own]
  double *x = nullptr, *r = nullptr, *q = nullptr;          // [n_own]
  double *p = nullptr;                                      // [n_loc] (gathered by SpMV -> needs ghosts)
  double* partials = nullptr;                               // block partial sums
  int partial_cap = 0;
  PcgScalars* scal = nullptr;                               // device
  PcgScalars* scal_host = nullptr;                          // pinned

  // Newton corrections of the previous Vsolve, by Newton step: initial guesses of the same step's linear solve in the next Gummel
  // iteration (opts.warm_start; the Newton iterates themselves still restart from V = 0, src:311)
  static constexpr int kNewtonWarm = 8;
  double* newton_x[kNewtonWarm] = {nullptr};
  bool newton_x_valid[kNewtonWarm] = {false};

  // ddpe state
  bool ddpe_active = false;
  int rec_mode = 0;
  double delta = 1, tau_p = 1, tau_n = 1, tol = 1e-9;

  std::vector<double> hist_outer, hist_newton, hist_pcg;

  // optional multigrid preconditioner (opts.precond == DDPS_PRECOND_AMG, single GPU)
  double eps_e = 0.0;  // energy-norm stop tolerance of the NEXT solve_system call (0: residual rules only)
  Amg* amg = nullptr;
  int amg_kind = -1;   // which kind of operator the next solve_system call handles (amg_numeric: lagged coarse operators)

  void fail(int code, const std::string& msg) {
    err_code = code;
    err = msg;
  }
};

// ---- setup.cu ----
int build_pattern(Context* ctx);
void free_pattern(Context* ctx);
// ---- partition.cpp-like host code (in setup.cu) ----
void partition_nodes_rcb(const double* nodes, int64_t nq, int nparts, int32_t* owner);

// ---- kernels.cu ----
constexpr int kRecD = 6;          // doubles per nodal record: {gw, s1 | u, v | eh, ehi}
struct AsmArgs {
  int nrows, nslices, wmax, snmax;
  int pf_dist;            // L2 prefetch distance in slices (about one wave of resident warps)
  int pf_desc;            // descriptor prefetch distance in slices
  const int4* sdesc;
  const int* sn_nodes;
  const int* inc_tri;
  const unsigned* inc_pack;
  const double* arinc;
  const int2* runs;
  const unsigned short* rowinfo;
  const unsigned char* isD;
  const double2* xy;
  const double* rec;      // [n_loc][kRecD] nodal records written by the prep kernels
  const double* gD;
  const double* phx;      // per-element coefficient (x pairing) or mu
  const double* phy;      // per-element coefficient (y pairing)
  const double* base_val; // base matrix values (Newton systems)
  const double* cinc;     // midpoint samples (C or H) in incidence order
  const double* cinc_poly[kMaxPoly];
  const double* nodal_const;  // neumann loads (may be null)
  const double* lift_in;      // precomputed lifting of the base stiffness (Newton systems)
  double par0, par1;          // delta^2 | tau_p,tau_n | alpha,beta | degree
  // outputs
  double* out_val;
  double* dinv;
  double* b;
  double* lift_out;
  double* resid;
};

enum { STIFF_BASE = 0, STIFF_COEF = 1, STIFF_EXPV = 2 };
enum { RHS_NONE = 0, RHS_DDP = 1, RHS_SRH = 2, RHS_ZERO = 3, RHS_SINH = 4, RHS_POLY = 5 };

int launch_assemble_base(Context* ctx, const AsmArgs& a);
int launch_assemble_uv(Context* ctx, const AsmArgs& a, int rhs_kind, bool mass);
int launch_assemble_newton(Context* ctx, const AsmArgs& a, int rhs_kind);

int launch_prep_base(Context* ctx);
int launch_prep_poisson(Context* ctx, const double* V, const double* u, const double* v, double delta2);
int launch_prep_uv(Context* ctx, double sign, const double* V, const double* u0, const double* v0, double tau_p,
                   double tau_n, int rec_mode);
int launch_prep_semlin(Context* ctx, const double* V);
// element-midpoint samples [3*nme] -> incidence order [inc_entries] (what the Newton-system kernels stream)
int launch_permute_mid(Context* ctx, const double* mid, double* out);

int launch_spmv(Context* ctx, const double* val, const double* x, double* y);
int launch_spmv_dot(Context* ctx, const double* val, const double* x, double* y);  // the PCG variant (fused p.Ap)
int launch_axpy(Context* ctx, double a, const double* x, double* y, int n);          // y += a*x
int launch_negate_copy(Context* ctx, const double* x, double* y, double s, int n);   // y = s*x
int launch_max_abs_diff3(Context* ctx, const double* a0, const double* a1, const double* b0, const double* b1,
                         const double* c0, const double* c1, int n, double* host_out);
int launch_max_abs(Context* ctx, const double* a, int n, double* host_out);

int fetch_resid_norm(Context* ctx, double* host_out);  // norm_inf written by the last RESID assembly
int pcg_iteration_launch(Context* ctx, const double* val, const double* dinv, double* x, int it, int max_it, bool first);
// Jacobi-PCG: solve val*x = rhs; x holds the initial guess if x0_nonzero (else it is zeroed).
int solve_system(Context* ctx, const double* val, const double* dinv, const double* rhs, double* x, bool x0_nonzero,
                 double rtol, double atol_inf, int64_t max_it, int64_t* iters_out, double* relres_out);
int pcg_solve(Context* ctx, const double* val, const double* dinv, const double* rhs, double* x, bool x0_nonzero,
              double rtol, double atol_inf, int64_t max_it, int64_t* iters_out, double* relres_out);

// ---- amg.cu ----
int amg_setup(Context* ctx);      // build the hierarchy from the base operator (after assemble_base); host + uploads
void amg_free(Context* ctx);
bool amg_ready(const Context* ctx);
// per-operator numeric phase: Galerkin products of `val` on the frozen pattern + coarsest-level inverse
int amg_numeric(Context* ctx, const double* val, const double* dinv, int kind);
int amg_vcycle(Context* ctx, const double* r, double* z, bool dot_rz);
int amg_pcg_solve(Context* ctx, const double* val, const double* dinv, const double* rhs, double* x, bool x0_nonzero,
                  double rtol, double atol_inf, int64_t max_it, int64_t* iters_out, double* relres_out);
int amg_pcg_iteration_launch(Context* ctx, double* x, int it, int max_it);
int amg_time_launch(Context* ctx, int which, int it, double* bytes);   // single kernels of the iteration (ddps_time_kernel 9..13)
double amg_iteration_bytes(Context* ctx);
int amg_debug_levels(Context* ctx, int64_t* nlevels, int64_t* sizes, int cap);
int amg_debug_get_values(Context* ctx, int level, double* val_csr, double* dinv);
int amg_debug_get_level(Context* ctx, int level, int64_t* sizes, int64_t* rowptr, int64_t* col, int64_t* Pptr, int64_t* Pcol,
                        double* Pval);

// ---- comm.cu ----
int halo_exchange(Context* ctx, double* vec);            // fill ghost part of a [n_loc] vector (finest level's plan)
int allreduce_sum(Context* ctx, double* dev, int count); // in place on the library stream
int allreduce_max(Context* ctx, double* dev, int count);
// backend primitives (NCCL or in-process): grouped point-to-point exchange of device buffers on the library stream, allreduce of
// device doubles, allgather of small host structs (blocking)
int comm_exchange(Context* ctx, const CommMsg* msgs, int n);
int comm_allreduce(Context* ctx, double* dev, size_t count, int op);
int comm_allgather_host(Context* ctx, const void* mine, size_t bytes, void* all);
// halo exchange on any level's plan: staged peer-memory kernel when plan.p2p, else pack + comm_exchange.  check_done: the
// kernel early-exits once the running solve's `done` flag is up (identical on all ranks)
int halo_xchg(Context* ctx, HaloPlan& plan, double* vec, bool check_done);
int halo_exchange_bytes(Context* ctx, const HaloPlan& plan, void* vec, size_t elem);   // setup: any element size, backend exchange
// in-place allreduce of nsum sums followed by nmax maxima (nsum + nmax <= 4) living at `dev_vals`: one tiny peer-memory
// kernel, or backend allreduces
int scalar_allreduce(Context* ctx, double* dev_vals, int nsum, int nmax, bool check_done);
int plan_finish(Context* ctx, HaloPlan& plan);   // collective: remote offsets + agree on peer-memory use
void plan_free(HaloPlan& plan);
int p2p_check_error(Context* ctx);               // raise DDPS_ERR_COMM if a peer-memory wait timed out since the last check
int comm_init(Context* ctx, const char* id128, int rank, int nranks);
int comm_unique_id(char* id128);
void comm_destroy(Context* ctx);
int p2p_setup(Context* ctx);      // after the vectors are allocated (needs ctx->p)
int p2p_teardown(Context* ctx);   // before they are freed
int comm_barrier(Context* ctx);

}  // namespace ddps

struct ddps_ctx : public ddps::Context {};
