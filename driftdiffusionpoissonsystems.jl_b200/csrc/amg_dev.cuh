// Device-side structures of the multigrid hierarchy, shared by amg.cu (kernels, cycle, CG) and amg_setup.cu
// (device-side hierarchy construction).
#pragma once

#include <vector>

#include "amg.h"
#include "ddps_internal.cuh"

namespace ddps {

constexpr int kAmgMaxLevels = 12;
constexpr int kAmgDenseMax = 512;

struct AmgLevelDev {
  int n = 0, nslices = 0, nc = 0, wmax = 0;   // wmax: longest row (= widest slice) of this level's matrix
  int64_t nnz = 0, nentries = 0;
  bool owns = false;                 // level 0 aliases the fine pattern of the context
  int* slice_ptr = nullptr;
  int* col = nullptr;
  int* rowptr = nullptr;
  double* val = nullptr;             // level >= 1: Galerkin values of the current operator (SELL order)
  double* dinv = nullptr;            // level >= 1
  float* val32 = nullptr;            // large levels: fp32 copy of the current values for the cycle's SpMVs
  const float* cur32 = nullptr;      // the fp32 copy the cycle uses (level-owned or of the active operator set)
  const double* cur_val = nullptr;   // matrix / inverse diagonal used by the cycle (level 0: set per operator)
  const double* cur_dinv = nullptr;
  // transfers to the next level, CSR with packed 8-byte entries {index, float value}: the prolongator values are
  // rounded to fp32 at setup (P, R = P^T and the Galerkin products all use the same rounded numbers)
  int *Pptr = nullptr, *Rptr = nullptr;
  int2 *Pent = nullptr, *Rent = nullptr;
  // two-stage Galerkin product (device-built levels): pattern of A*P row by row (sorted distinct coarse columns; rows with
  // an empty prolongator row are left out, nothing reads them) and the value workspace stage 1 fills for stage 2
  int *APptr = nullptr, *APcol = nullptr;
  double* APval = nullptr;
  int64_t nAP = 0;
  double *b = nullptr, *x = nullptr, *x2 = nullptr, *t = nullptr;
  // ---- distributed levels (multi-rank contexts).  Rows = the level's rows this rank owns (n of them, global id goff + i),
  // columns = local ids: owned first, then ghosts grouped by owner (n_loc in all); vectors x, t, x2 have n_loc entries, their
  // ghost parts are refreshed over `plan`.
  //   dist == 1: the next level is distributed too (rank-local aggregates, ghost prolongator rows and A*P rows of the
  //              neighbours' boundary rows come by one exchange each);
  //   dist == 2: "transition": the next level (and everything below) is REPLICATED on every rank -- restriction and the
  //              Galerkin product give partial sums over the owned fine rows that one allreduce completes;
  //   dist == 0: single-rank context or replicated level.
  int dist = 0;
  int n_loc = 0;
  long long goff = 0;
  HaloPlan* plan = nullptr;
  bool owns_plan = false;
  int* glob = nullptr;               // [n_loc] global row id of every local column
  // A*P values of the boundary rows travel to the neighbours once per operator (dist == 1)
  int ap_nsend = 0;
  int* ap_send_pos = nullptr;        // positions in APval of the entries to pack, grouped by neighbour
  double* ap_send_buf = nullptr;
  std::vector<int> ap_send_off, ap_recv_off;   // [nneigh+1] entry offsets per neighbour (send buffer / ghost part of APval)
  int ap_ghost_base = 0;             // APptr[n]: the ghost rows' entries start here
};

// Galerkin values of one KIND of operator (first Newton system of a Vsolve / later Newton systems / u- / v-continuity
// system): consecutive operators of the same kind differ little once the outer iteration settles, so their coarse
// operators are reused while the fine operator stays within a tolerance of the snapshot they were computed from.
constexpr int kAmgKinds = 4;
struct AmgOpSet {
  bool allocated = false, valid = false;
  double* snap = nullptr;                      // fine values the coarse operators were computed from (SELL order)
  double* val[kAmgMaxLevels] = {nullptr};      // levels >= 1
  double* dinv[kAmgMaxLevels] = {nullptr};
  float* val32[kAmgMaxLevels] = {nullptr};
  double* dense = nullptr;
};

// DDPS_PROFILE_ITER=1: CUDA-event timeline of ONE CG iteration per solve (the 4th), accumulated per phase and printed when the
// hierarchy is freed -- where the iteration's time goes, including the exchanges (no profiler needed, works on every rank)
struct AmgProf {
  bool on = false, armed = false;
  int n = 0, samples = 0;
  cudaEvent_t ev[96];
  const char* name[96];
  double acc[96] = {0};
  bool have_ev = false;
};

struct Amg {
  bool ready = false;
  AmgProf prof;
  int nlev = 0;
  AmgLevelDev lv[kAmgMaxLevels];
  double* dense = nullptr;   // [n_coarsest^2] inverse of the coarsest operator
  const double* cur_dense = nullptr;
  AmgOpSet sets[kAmgKinds];
  long long reuses = 0, numerics = 0;
  double* z = nullptr;       // [n0] preconditioned residual of the CG
  double omega = 2.0 / 3.0;
  double setup_ms = 0.0;
  int tail_from = 0, tail_grid = 0;   // fused tail of the cycle (k_amg_tail): first level it covers (0: off), CTAs
  unsigned* tail_bar = nullptr;       // its grid-barrier counters
  // transition level of a distributed hierarchy: the partial right-hand sides of the first replicated level travel by ONE staged
  // peer-memory exchange (every rank to every rank) and are summed in rank order -- instead of a library allreduce
  HaloPlan* tail_plan = nullptr;
  double* tail_parts = nullptr;       // [nranks][n of the first replicated level]: own part first, then the peers' in rank order
};


// amg.cu: Galerkin product of the fine matrix values `f_val` (SELL order of level F) into level C (values + inverse diagonal)
int amg_launch_rap(Context* ctx, const AmgLevelDev& F, const double* f_val, const AmgLevelDev& C, double* out_val, double* out_dinv);
// amg_setup.cu: build transfers of level l and the pattern + base values of level l+1 on the device.
// Returns 0 on success with *nc_out = rows of the next level (0: coarsening stalled, nothing installed), <0 on error,
// >0 when the device path does not apply (caller falls back to the host path).
int amg_dev_coarsen(Context* ctx, Amg* A, int l, const unsigned char* d_fixed, const AmgParams& prm, int* nc_out);
// amg_setup.cu, multi-rank contexts.  Distributed coarsening of level l (collective): *ncg_out = global rows of level l+1
// (0: coarsening stalled on the whole, nothing installed).  Transition: gather level l, build the replicated rest of the
// hierarchy on the host, install it; *nlev_out = total number of levels (0: no coarsening possible).
int amg_dist_coarsen(Context* ctx, Amg* A, int l, const AmgParams& prm, long long* ncg_out);
int amg_dist_transition(Context* ctx, Amg* A, int l, const AmgParams& prm, int* nlev_out);
int amg_dist_ap_halo(Context* ctx, AmgLevelDev& F);   // numeric phase: A*P values of the neighbours' boundary rows
// amg.cu: upload helpers of host-built levels
int amg_install_pattern(Context* ctx, AmgLevelDev& L, const AmgHostLevel& HL);
int amg_install_transfers(Context* ctx, AmgLevelDev& L, const AmgHostLevel& HL);

}  // namespace ddps
