// Multi-GPU plumbing.  The reference has no communication at all ("Distributed communication backend:
// none exists"); the path has two real exchange steps -- halos of nodal vectors (fields, CG search direction, the cycle
// vectors of every distributed multigrid level) and scalar reductions (CG dots, stop tests) -- plus one-time setup exchanges.
//
// Backends (same primitives: grouped point-to-point exchange, allreduce, allgather of small host structs):
//   * NCCL over NVLink: one process per GPU (ddps_comm_init with an ncclUniqueId);
//   * in-process group: ranks are host threads of ONE process sharing one device (id "LOCAL:<name>").  Exchanges are
//     device-to-device copies ordered by host barriers.  It exists so that the whole distributed logic (partition, ghost
//     layers, distributed multigrid hierarchy) runs -- and is tested by `pytest -m gpu` -- on a single GPU.
// On top of either backend the hot exchanges use peer memory directly (CUDA IPC mappings over NVLink, or plain pointers in
// the in-process group): a staged halo exchange and a scalar allreduce, ONE tiny kernel each, no library call and no host
// involvement inside the iteration (k_halo_xchg, k_p2p_allreduce below).  The Jacobi-PCG iteration additionally has its
// exchanges fused into its three kernels (kernels.cu).
#include <chrono>
#include <cstdio>
#include <condition_variable>
#include <map>
#include <mutex>

#include "ddps_internal.cuh"

namespace ddps {

#ifdef DDPS_WITH_NCCL
#define DDPS_NCCL(call)                                                                          \
  do {                                                                                           \
    ncclResult_t r__ = (call);                                                                   \
    if (r__ != ncclSuccess) {                                                                    \
      ctx->fail(DDPS_ERR_COMM, std::string(#call) + ": " + ncclGetErrorString(r__));             \
      return DDPS_ERR_COMM;                                                                      \
    }                                                                                            \
  } while (0)
#endif

// ------------------------------------------------------------------------------------------------------------
// in-process group
// ------------------------------------------------------------------------------------------------------------
struct LocalGroup {
  int nranks = 0;
  std::mutex mu;
  std::condition_variable cv;
  int waiting = 0;
  unsigned long gen = 0;
  bool broken = false;
  const CommMsg* msgs[kMaxRanks] = {nullptr};
  int nmsg[kMaxRanks] = {0};
  const void* slot[kMaxRanks] = {nullptr};
  std::vector<double> red[kMaxRanks];
  // all ranks arrive or the group is declared broken (a rank died: its thread raised instead of calling the collective)
  bool barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (broken) return false;
    const unsigned long g = gen;
    if (++waiting == nranks) {
      waiting = 0;
      ++gen;
      cv.notify_all();
      return true;
    }
    if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return gen != g || broken; })) broken = true;
    if (broken) cv.notify_all();
    return !broken;
  }
};

namespace {
std::mutex g_groups_mu;
std::map<std::string, std::weak_ptr<LocalGroup>> g_groups;

int local_fail(Context* ctx) {
  ctx->fail(DDPS_ERR_COMM, "in-process communicator: a rank stopped participating (barrier timed out)");
  return DDPS_ERR_COMM;
}
}  // namespace

// ------------------------------------------------------------------------------------------------------------
// backend primitives
// ------------------------------------------------------------------------------------------------------------
int comm_exchange(Context* ctx, const CommMsg* msgs, int n) {
  if (ctx->nranks == 1 || n == 0) {
    if (ctx->lgroup) {   // collective even when this rank has nothing to exchange
      if (!ctx->lgroup->barrier() || !ctx->lgroup->barrier()) return local_fail(ctx);
    }
    return 0;
  }
  if (ctx->lgroup) {
    LocalGroup& G = *ctx->lgroup;
    G.msgs[ctx->rank] = msgs;
    G.nmsg[ctx->rank] = n;
    DDPS_CUDA(cudaStreamSynchronize(ctx->stream));   // my send buffers are complete
    if (!G.barrier()) return local_fail(ctx);
    for (int k = 0; k < n; ++k) {
      const CommMsg& m = msgs[k];
      if (!m.recv_bytes) continue;
      const CommMsg* src = nullptr;
      for (int q = 0; q < G.nmsg[m.peer]; ++q)
        if (G.msgs[m.peer][q].peer == ctx->rank && G.msgs[m.peer][q].send_bytes) src = &G.msgs[m.peer][q];
      if (!src || src->send_bytes != m.recv_bytes) { ctx->fail(DDPS_ERR_COMM, "in-process exchange: message sizes do not match"); return DDPS_ERR_COMM; }
      DDPS_CUDA(cudaMemcpyAsync(m.recv, src->send, m.recv_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!G.barrier()) return local_fail(ctx);          // the senders may reuse their buffers
    return 0;
  }
#ifdef DDPS_WITH_NCCL
  DDPS_NCCL(ncclGroupStart());
  for (int k = 0; k < n; ++k) {
    const CommMsg& m = msgs[k];
    if (m.send_bytes) DDPS_NCCL(ncclSend(m.send, m.send_bytes, ncclInt8, m.peer, ctx->comm, ctx->stream));
    if (m.recv_bytes) DDPS_NCCL(ncclRecv(m.recv, m.recv_bytes, ncclInt8, m.peer, ctx->comm, ctx->stream));
  }
  DDPS_NCCL(ncclGroupEnd());
  return 0;
#else
  ctx->fail(DDPS_ERR_COMM, "library built without NCCL");
  return DDPS_ERR_COMM;
#endif
}

int comm_allreduce(Context* ctx, double* dev, size_t count, int op) {
  if (ctx->nranks == 1 || count == 0) return 0;
  if (ctx->lgroup) {
    LocalGroup& G = *ctx->lgroup;
    std::vector<double>& mine = G.red[ctx->rank];
    mine.resize(count);
    DDPS_CUDA(cudaMemcpyAsync(mine.data(), dev, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!G.barrier()) return local_fail(ctx);
    std::vector<double> out(G.red[0]);   // combined in rank order: identical on every rank
    for (int r = 1; r < G.nranks; ++r) {
      if (G.red[r].size() != count) { ctx->fail(DDPS_ERR_COMM, "in-process allreduce: counts differ"); return DDPS_ERR_COMM; }
      for (size_t i = 0; i < count; ++i)
        out[i] = op == COMM_SUM ? out[i] + G.red[r][i] : op == COMM_MAX ? std::max(out[i], G.red[r][i]) : std::min(out[i], G.red[r][i]);
    }
    if (!G.barrier()) return local_fail(ctx);            // everybody has read every slot
    DDPS_CUDA(cudaMemcpyAsync(dev, out.data(), count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
#ifdef DDPS_WITH_NCCL
  DDPS_NCCL(ncclAllReduce(dev, dev, count, ncclDouble, op == COMM_SUM ? ncclSum : op == COMM_MAX ? ncclMax : ncclMin, ctx->comm, ctx->stream));
  return 0;
#else
  return DDPS_ERR_COMM;
#endif
}

int comm_allgather_host(Context* ctx, const void* mine, size_t bytes, void* all) {
  if (ctx->nranks == 1) { memcpy(all, mine, bytes); return 0; }
  if (ctx->lgroup) {
    LocalGroup& G = *ctx->lgroup;
    G.slot[ctx->rank] = mine;
    if (!G.barrier()) return local_fail(ctx);
    for (int r = 0; r < G.nranks; ++r) memcpy((char*)all + (size_t)r * bytes, G.slot[r], bytes);
    if (!G.barrier()) return local_fail(ctx);
    return 0;
  }
#ifdef DDPS_WITH_NCCL
  char *d_send = nullptr, *d_recv = nullptr;
  DDPS_CUDA(cudaMalloc(&d_send, std::max<size_t>(bytes, 1)));
  DDPS_CUDA(cudaMalloc(&d_recv, std::max<size_t>(bytes, 1) * ctx->nranks));
  DDPS_CUDA(cudaMemcpyAsync(d_send, mine, bytes, cudaMemcpyHostToDevice, ctx->stream));
  DDPS_NCCL(ncclAllGather(d_send, d_recv, bytes, ncclInt8, ctx->comm, ctx->stream));
  DDPS_CUDA(cudaMemcpyAsync(all, d_recv, bytes * ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_send);
  cudaFree(d_recv);
  return 0;
#else
  return DDPS_ERR_COMM;
#endif
}

int allreduce_sum(Context* ctx, double* dev, int count) { return comm_allreduce(ctx, dev, (size_t)count, COMM_SUM); }
int allreduce_max(Context* ctx, double* dev, int count) { return comm_allreduce(ctx, dev, (size_t)count, COMM_MAX); }

int comm_init(Context* ctx, const char* id128, int rank, int nranks) {
  if (nranks < 1 || rank < 0 || rank >= nranks || nranks > kMaxRanks) { ctx->fail(DDPS_ERR_ARG, "bad rank/nranks (at most 8 ranks)"); return DDPS_ERR_ARG; }
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks == 1) return 0;
  if (strncmp(id128, "LOCAL:", 6) == 0) {
    const char* ml = getenv("CUDA_MODULE_LOADING");
    if (!ml || strcmp(ml, "EAGER") != 0) {
      static bool warned = false;
      if (!warned) fprintf(stderr, "[ddps] in-process communicator: set CUDA_MODULE_LOADING=EAGER before the first CUDA call -- with lazy module "
                                   "loading the first launch of a kernel can block behind a kernel that waits for it (ranks share one context)\n");
      warned = true;
    }
    const std::string key(id128, strnlen(id128, 128));
    std::lock_guard<std::mutex> lk(g_groups_mu);
    std::shared_ptr<LocalGroup> g = g_groups[key].lock();
    if (!g) {
      g = std::make_shared<LocalGroup>();
      g->nranks = nranks;
      g_groups[key] = g;
    }
    if (g->nranks != nranks) { ctx->fail(DDPS_ERR_ARG, "in-process communicator: nranks differs between the ranks"); return DDPS_ERR_ARG; }
    ctx->lgroup = g;
    return 0;
  }
#ifdef DDPS_WITH_NCCL
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(&id, id128, 128);
  DDPS_NCCL(ncclCommInitRank(&ctx->comm, nranks, id, rank));
  return 0;
#else
  ctx->fail(DDPS_ERR_COMM, "library built without NCCL");
  return DDPS_ERR_COMM;
#endif
}

int comm_unique_id(char* id128) {
#ifdef DDPS_WITH_NCCL
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return DDPS_ERR_COMM;
  memcpy(id128, &id, 128);
  return 0;
#else
  return DDPS_ERR_COMM;
#endif
}

int comm_barrier(Context* ctx) {
  if (ctx->nranks == 1) return 0;
  int rc = comm_allreduce(ctx, &ctx->scal->norm_aux2, 1, COMM_MAX);
  if (rc) return rc;
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// halo exchange through the backend (setup exchanges, fallback of the hot exchanges)
// ------------------------------------------------------------------------------------------------------------
__global__ void k_pack(int n, const int* __restrict__ idx, const double* __restrict__ vec, double* __restrict__ buf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = vec[idx[i]];
}
__global__ void k_pack_bytes(int n, const int* __restrict__ idx, const char* __restrict__ vec, char* __restrict__ buf, int elem) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const char* s = vec + (size_t)idx[i] * elem;
  char* d = buf + (size_t)i * elem;
  for (int b = 0; b < elem; ++b) d[b] = s[b];
}

static int halo_backend(Context* ctx, const HaloPlan& H, void* vec, size_t elem, void* send_buf) {
  std::vector<CommMsg> msgs;
  for (int j = 0; j < H.nneigh; ++j) {
    CommMsg m;
    m.peer = H.peer[j];
    m.send = (char*)send_buf + (size_t)H.send_off[j] * elem;
    m.send_bytes = (size_t)(H.send_off[j + 1] - H.send_off[j]) * elem;
    m.recv = (char*)vec + ((size_t)H.n_own + H.recv_off[j]) * elem;
    m.recv_bytes = (size_t)(H.recv_off[j + 1] - H.recv_off[j]) * elem;
    msgs.push_back(m);
  }
  return comm_exchange(ctx, msgs.data(), (int)msgs.size());
}

int halo_exchange_bytes(Context* ctx, const HaloPlan& H, void* vec, size_t elem) {
  if (ctx->nranks == 1) return 0;
  char* buf = nullptr;
  DDPS_CUDA(cudaMalloc(&buf, std::max<size_t>((size_t)H.nsend * elem, 1)));
  if (H.nsend > 0) k_pack_bytes<<<(H.nsend + 255) / 256, 256, 0, ctx->stream>>>(H.nsend, H.send_idx, (const char*)vec, buf, (int)elem);
  int rc = halo_backend(ctx, H, vec, elem, buf);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(buf);
  return rc;
}

// ------------------------------------------------------------------------------------------------------------
// peer-memory exchanges: one tiny kernel each
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool xspin(const volatile unsigned long long* flag, unsigned long long tag, volatile int* err) {
  const long long t0 = clock64();
  while (*flag < tag) {
    if (*err) return false;
    if (clock64() - t0 > 20000000000LL) { *err = 1; return false; }
  }
  return true;
}

struct XchgArgs {
  int nneigh, n_own, nslice;
  int send_off[kMaxRanks + 1], recv_off[kMaxRanks + 1];
  double* rstage[kMaxRanks];                     // the peer's staging buffer (this tag's half) at my block's offset
  unsigned long long* rflag[kMaxRanks];          // &peer window->xflag[me]
  const unsigned long long* lflag[kMaxRanks];    // &my window->xflag[peer]
  const double* lstage;                          // my staging buffer (this tag's half)
  unsigned* ticket;                              // [kMaxRanks] my window's slice tickets
  unsigned long long tag;
  int* err;
};

// nslice blocks serve neighbour j (blockIdx.x = j * nslice + s): each stores its slice of my boundary values into the
// neighbour's staging buffer over NVLink; the block that finishes last raises the neighbour's flag; then every block waits
// for the neighbour's block to land in MY staging buffer and copies its slice into the ghost part of `vec`.  (One block of
// 256 threads per neighbour made the two gather loops -- 8193 entries per neighbour at 268M triangles on 8 GPUs -- a chain
// of 32 dependent-latency steps each: ~20 us per exchange; sliced, an exchange is two short steps around the NVLink round
// trip.)  When the kernel has finished the ghosts are final, so every consumer reads them through the normal (read-only)
// path.  Staging is double-buffered by tag parity: a peer can be at most one exchange ahead of me (its exchange k+1 completes
// only after my push k+1, which follows my copy-out k).
__global__ void __launch_bounds__(512) k_halo_xchg(const XchgArgs A, const int* __restrict__ send_idx, double* vec, PcgScalars* scal,
                                                  int check_done) {
  if (check_done && scal->done) return;
  __shared__ bool s_last;
  const int j = blockIdx.x / A.nslice, sl = blockIdx.x - j * A.nslice;
  const int s0 = A.send_off[j], ns = A.send_off[j + 1] - s0;
  double* dst = A.rstage[j];
  for (int i = sl * blockDim.x + threadIdx.x; i < ns; i += A.nslice * blockDim.x) dst[i] = vec[send_idx[s0 + i]];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    s_last = atomicInc(&A.ticket[j], A.nslice - 1) == (unsigned)(A.nslice - 1);
    if (s_last) {
      __threadfence_system();
      *(volatile unsigned long long*)A.rflag[j] = A.tag;
    }
    if (!xspin(A.lflag[j], A.tag, A.err)) {
      if (sl == 0)
        printf("[ddps] staged halo exchange timed out: tag %llu, neighbour %d of %d, flag seen %llu, n_own %d\n", A.tag, j, A.nneigh,
               *(volatile const unsigned long long*)A.lflag[j], A.n_own);
      scal->breakdown = 2; scal->done = 1;
    }
    __threadfence_system();
  }
  __syncthreads();
  const int r0 = A.recv_off[j], nr = A.recv_off[j + 1] - r0;
  for (int i = sl * blockDim.x + threadIdx.x; i < nr; i += A.nslice * blockDim.x) vec[A.n_own + r0 + i] = __ldcg(A.lstage + r0 + i);
}

struct RedArgs {
  int nranks, rank, nv;
  unsigned maxmask;
  P2PWin* win[kMaxRanks];
  unsigned long long tag;
};

// vals[0..nv) <- allreduce over the ranks (component i: max if bit i of maxmask, else sum), combined in rank order: bitwise
// identical on every rank.  One warp: lane r publishes my values into rank r's window and later waits for rank r's values
// in mine.  Slots are double-buffered by tag parity (same argument as the staging buffers).
__global__ void k_p2p_allreduce(const RedArgs A, double* vals, PcgScalars* scal, int check_done) {
  if (check_done && scal->done) return;
  const int lane = threadIdx.x;
  const int par = (int)(A.tag & 1);
  P2PWin* me = A.win[A.rank];
  double mine[4] = {0.0, 0.0, 0.0, 0.0};
  bool ok = true;
  if (lane < A.nranks) {
    P2PWin* W = A.win[lane];
    for (int i = 0; i < A.nv; ++i) ((volatile double*)W->rval[par][A.rank])[i] = vals[i];
    __threadfence_system();
    *(volatile unsigned long long*)&W->rflag[par][A.rank] = A.tag;
    ok = xspin(&me->rflag[par][lane], A.tag, &me->error);
    __threadfence_system();
    for (int i = 0; i < A.nv; ++i) mine[i] = ((const volatile double*)me->rval[par][lane])[i];
  }
  ok = __all_sync(0xffffffffu, ok);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int q = 0; q < A.nranks; ++q)
    for (int i = 0; i < 4; ++i) {
      const double t = __shfl_sync(0xffffffffu, mine[i], q);
      if (q == 0) acc[i] = t;
      else acc[i] = ((A.maxmask >> i) & 1) ? fmax(acc[i], t) : acc[i] + t;
    }
  if (lane < A.nranks && !ok)
    printf("[ddps] scalar allreduce timed out: rank %d waits for rank %d, tag %llu, flag seen %llu\n", A.rank, lane, A.tag,
           *(volatile unsigned long long*)&me->rflag[par][lane]);
  if (lane == 0) {
    for (int i = 0; i < A.nv; ++i) vals[i] = acc[i];
    if (!ok) { scal->breakdown = 2; scal->done = 1; }
  }
}

int halo_xchg(Context* ctx, HaloPlan& H, double* vec, bool check_done) {
  if (ctx->nranks == 1) return 0;
  if (H.p2p && ctx->p2p.xenabled) {
    // In-process group: the ranks share one device AND one driver context.  A rank whose host runs ahead can enter a driver
    // call that blocks on its own stream (a copy to pageable memory, cudaFree) while holding the context lock -- behind a kernel
    // that spins on a peer whose host then cannot even launch the kernel it waits for.  Host barriers around the launch of
    // every spinning kernel make sure both kernels sit in their streams before anybody proceeds (test transport only).
    if (ctx->lgroup && !ctx->lgroup->barrier()) return local_fail(ctx);
    if (H.nneigh == 0) {
      if (ctx->lgroup && !ctx->lgroup->barrier()) return local_fail(ctx);
      return 0;
    }
    P2PHost& P = ctx->p2p;
    XchgArgs A;
    memset(&A, 0, sizeof(A));
    A.nneigh = H.nneigh;
    A.n_own = H.n_own;
    A.tag = ++P.xtag;
    const size_t half = (size_t)(A.tag & 1) * P.stage_cap;
    for (int j = 0; j <= H.nneigh; ++j) { A.send_off[j] = H.send_off[j]; A.recv_off[j] = H.recv_off[j]; }
    for (int j = 0; j < H.nneigh; ++j) {
      const int q = H.peer[j];
      A.rstage[j] = P.stage_peer[q] + half + H.remote_off[j];
      A.rflag[j] = &P.win_peer[q]->xflag[ctx->rank];
      A.lflag[j] = &P.win_local->xflag[q];
    }
    A.lstage = P.stage_local + half;
    A.err = &P.win_local->error;
    A.ticket = P.win_local->xticket;
    int widest = 1;
    for (int j = 0; j < H.nneigh; ++j) widest = std::max(widest, std::max(H.send_off[j + 1] - H.send_off[j], H.recv_off[j + 1] - H.recv_off[j]));
    A.nslice = std::max(1, std::min(16, (widest + 511) / 512));
    k_halo_xchg<<<H.nneigh * A.nslice, 512, 0, ctx->stream>>>(A, H.send_idx, vec, ctx->scal, check_done ? 1 : 0);
    ctx->stats.kernel_launches++;
    DDPS_CUDA(cudaGetLastError());
    if (ctx->lgroup && !ctx->lgroup->barrier()) return local_fail(ctx);
    return 0;
  }
  // every rank takes part in the backend exchange, also one without neighbours (the in-process backend is collective)
  if (H.nsend > 0) {
    k_pack<<<(H.nsend + 255) / 256, 256, 0, ctx->stream>>>(H.nsend, H.send_idx, vec, H.send_buf);
    ctx->stats.kernel_launches++;
  }
  return halo_backend(ctx, H, vec, sizeof(double), H.send_buf);
}

int halo_exchange(Context* ctx, double* vec) { return halo_xchg(ctx, ctx->halo, vec, false); }

int scalar_allreduce(Context* ctx, double* dev_vals, int nsum, int nmax, bool check_done) {
  if (ctx->nranks == 1) return 0;
  if (ctx->p2p.xenabled) {
    if (ctx->lgroup && !ctx->lgroup->barrier()) return local_fail(ctx);   // (see halo_xchg)
    P2PHost& P = ctx->p2p;
    RedArgs A;
    memset(&A, 0, sizeof(A));
    A.nranks = ctx->nranks;
    A.rank = ctx->rank;
    A.nv = nsum + nmax;
    A.maxmask = ((1u << nmax) - 1u) << nsum;
    for (int r = 0; r < ctx->nranks; ++r) A.win[r] = P.win_peer[r];
    A.tag = ++P.rtag;
    k_p2p_allreduce<<<1, 32, 0, ctx->stream>>>(A, dev_vals, ctx->scal, check_done ? 1 : 0);
    ctx->stats.kernel_launches++;
    DDPS_CUDA(cudaGetLastError());
    if (ctx->lgroup && !ctx->lgroup->barrier()) return local_fail(ctx);
    return 0;
  }
  int rc;
  if (nsum && (rc = comm_allreduce(ctx, dev_vals, (size_t)nsum, COMM_SUM))) return rc;
  if (nmax && (rc = comm_allreduce(ctx, dev_vals + nsum, (size_t)nmax, COMM_MAX))) return rc;
  return 0;
}

int p2p_check_error(Context* ctx) {
  if (!ctx->p2p.win_local) return 0;
  int err = 0;
  DDPS_CUDA(cudaMemcpyAsync(&err, &ctx->p2p.win_local->error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (err) {
    cudaMemsetAsync(&ctx->p2p.win_local->error, 0, sizeof(int), ctx->stream);
    ctx->fail(DDPS_ERR_COMM, "peer-memory exchange timed out (a rank stopped participating)");
    return DDPS_ERR_COMM;
  }
  return 0;
}

// remote offsets of a plan (where my block starts inside each neighbour's ghost region) + agreement on peer-memory use
int plan_finish(Context* ctx, HaloPlan& H) {
  H.p2p = false;
  H.remote_off.assign(H.nneigh, 0);
  if (ctx->nranks == 1) return 0;
  struct Info { int off[kMaxRanks]; int n_ghost; int pad; };
  Info mine;
  memset(&mine, 0, sizeof(mine));
  for (int q = 0; q < kMaxRanks; ++q) mine.off[q] = -1;
  for (int j = 0; j < H.nneigh; ++j) mine.off[H.peer[j]] = H.recv_off[j];
  mine.n_ghost = H.n_ghost;
  std::vector<Info> all(ctx->nranks);
  int rc = comm_allgather_host(ctx, &mine, sizeof(Info), all.data());
  if (rc) return rc;
  bool fits = ctx->p2p.xenabled;
  for (int r = 0; r < ctx->nranks; ++r) fits = fits && (size_t)all[r].n_ghost <= ctx->p2p.stage_cap;
  for (int j = 0; j < H.nneigh; ++j) {
    const int off = all[H.peer[j]].off[ctx->rank];
    if (off < 0) { ctx->fail(DDPS_ERR_COMM, "halo plans of two ranks do not match"); return DDPS_ERR_COMM; }
    H.remote_off[j] = off;
  }
  H.p2p = fits;   // (xenabled and stage_cap are identical on all ranks by construction)
  return 0;
}

void plan_free(HaloPlan& H) {
  if (H.send_idx) cudaFree(H.send_idx);
  if (H.send_buf) cudaFree(H.send_buf);
  H = HaloPlan();
}

// ------------------------------------------------------------------------------------------------------------
// peer-memory setup (at ddps_mesh_set, after the vectors exist): windows, staging buffers, and -- NCCL backend only -- the
// IPC mapping of every peer's search-direction vector p for the fused Jacobi-PCG path
// ------------------------------------------------------------------------------------------------------------
int p2p_setup(Context* ctx) {
  P2PHost& H = ctx->p2p;
  H.enabled = false;
  H.xenabled = false;
  if (ctx->nranks == 1) return 0;
  const char* env = getenv("DDPS_P2P");
  const bool local = (bool)ctx->lgroup;
  // (in-process group: the staged peer-memory kernels run too, with plain pointers instead of IPC mappings and host
  //  barriers around their launches, see halo_xchg; DDPS_P2P=0 selects the host-ordered backend exchange for everything)
  const bool off = env && env[0] == '0';
  const int me = ctx->rank, nr = ctx->nranks;
  int rc;
  struct Xchg { cudaIpcMemHandle_t hp, hw, hs; void *pp, *pw, *ps; int off[kMaxRanks]; int device; int ok; long long cap; };
  Xchg mine;
  memset(&mine, 0, sizeof(mine));
  mine.device = ctx->device;
  mine.ok = off ? 0 : 1;
  // staging capacity: identical on all ranks (largest ghost block of the finest level, doubled: the coarse levels' ghost
  // sets reach two aggregate rings deep but shrink with the level size)
  // ... and room for the replicated right-hand side of the multigrid transition level, which every rank receives from every
  // other rank in one exchange (amg.cu: up to kAmgSmallLevel = 32768 doubles per rank)
  long long cap = std::max(2LL * ctx->halo.n_ghost + 1024, (long long)(nr - 1) * 32768);
  {
    double c = (double)cap;
    DDPS_CUDA(cudaMemcpyAsync(&ctx->scal->norm_aux2, &c, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = comm_allreduce(ctx, &ctx->scal->norm_aux2, 1, COMM_MAX))) return rc;
    DDPS_CUDA(cudaMemcpyAsync(&c, &ctx->scal->norm_aux2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
    cap = (long long)c;
  }
  if (!off) {
    DDPS_CUDA(cudaMalloc(&H.win_local, sizeof(P2PWin)));
    DDPS_CUDA(cudaMemset(H.win_local, 0, sizeof(P2PWin)));
    DDPS_CUDA(cudaMalloc(&H.stage_local, 2 * (size_t)cap * sizeof(double)));
    DDPS_CUDA(cudaMemset(H.stage_local, 0, 2 * (size_t)cap * sizeof(double)));
    H.stage_cap = (size_t)cap;
    mine.cap = cap;
    mine.pp = ctx->p; mine.pw = H.win_local; mine.ps = H.stage_local;
    if (!local) {
      if (cudaIpcGetMemHandle(&mine.hp, ctx->p) != cudaSuccess || cudaIpcGetMemHandle(&mine.hw, H.win_local) != cudaSuccess ||
          cudaIpcGetMemHandle(&mine.hs, H.stage_local) != cudaSuccess) {
        cudaGetLastError();
        mine.ok = 0;
      }
    }
  }
  for (int q = 0; q < kMaxRanks; ++q) mine.off[q] = -1;
  for (int j = 0; j < ctx->halo.nneigh; ++j) mine.off[ctx->halo.peer[j]] = ctx->n_own + ctx->halo.recv_off[j];
  std::vector<Xchg> all(nr);
  if ((rc = comm_allgather_host(ctx, &mine, sizeof(Xchg), all.data()))) return rc;
  bool ok = true;
  for (int r = 0; r < nr; ++r) ok = ok && all[r].ok;
  if (ok) {
    H.ipc = !local;
    for (int r = 0; r < nr && ok; ++r) {
      if (r == me) { H.win_peer[r] = H.win_local; H.p_peer[r] = ctx->p; H.stage_peer[r] = H.stage_local; continue; }
      if (local) {
        H.win_peer[r] = (P2PWin*)all[r].pw; H.p_peer[r] = (double*)all[r].pp; H.stage_peer[r] = (double*)all[r].ps;
        continue;
      }
      int can = 0;
      cudaDeviceCanAccessPeer(&can, ctx->device, all[r].device);
      if (!can) { ok = false; break; }
      if (cudaIpcOpenMemHandle((void**)&H.win_peer[r], all[r].hw, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
      if (ok && cudaIpcOpenMemHandle((void**)&H.p_peer[r], all[r].hp, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
      if (ok && cudaIpcOpenMemHandle((void**)&H.stage_peer[r], all[r].hs, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
    }
    if (!ok) cudaGetLastError();
  }
  // every rank must take the same decision: reduce the flag
  double flag = ok ? 1.0 : 0.0;
  DDPS_CUDA(cudaMemcpyAsync(&ctx->scal->norm_aux2, &flag, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = comm_allreduce(ctx, &ctx->scal->norm_aux2, 1, COMM_MIN))) return rc;
  DDPS_CUDA(cudaMemcpyAsync(&flag, &ctx->scal->norm_aux2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (flag < 0.5) {   // fall back to the backend path everywhere
    for (int r = 0; r < nr; ++r) if (r != me) {
      if (H.ipc) {
        if (H.win_peer[r]) cudaIpcCloseMemHandle(H.win_peer[r]);
        if (H.p_peer[r]) cudaIpcCloseMemHandle(H.p_peer[r]);
        if (H.stage_peer[r]) cudaIpcCloseMemHandle(H.stage_peer[r]);
      }
      H.win_peer[r] = nullptr; H.p_peer[r] = nullptr; H.stage_peer[r] = nullptr;
    }
    if (H.win_local) cudaFree(H.win_local);
    if (H.stage_local) cudaFree(H.stage_local);
    H.win_local = nullptr; H.stage_local = nullptr; H.stage_cap = 0;
    return plan_finish(ctx, ctx->halo);
  }
  memset(&H.push, 0, sizeof(H.push));
  memset(&H.dev, 0, sizeof(H.dev));
  H.push.nneigh = ctx->halo.nneigh;
  for (int j = 0; j < ctx->halo.nneigh; ++j) {
    const int q = ctx->halo.peer[j];
    H.push.off[j] = ctx->halo.send_off[j];
    H.push.dst[j] = H.p_peer[q] + all[q].off[me];
    H.push.flag[j] = &H.win_peer[q]->halo_flag[me];
    H.dev.neigh[j] = q;
  }
  H.push.off[ctx->halo.nneigh] = ctx->halo.send_off[ctx->halo.nneigh];
  H.dev.nranks = nr; H.dev.rank = me; H.dev.nneigh = ctx->halo.nneigh;
  for (int r = 0; r < nr; ++r) H.dev.win[r] = H.win_peer[r];
  H.tag = H.xtag = H.rtag = 0;
  H.xenabled = true;
  // the fused Jacobi path spins with FULL grids on its peers: fine with one GPU per rank, a deadlock when the ranks share a
  // device (a spinning grid can keep the kernel it waits for from ever being scheduled) -> in-process group: staged path only
  H.enabled = !local;
  return plan_finish(ctx, ctx->halo);
}

int p2p_teardown(Context* ctx) {
  P2PHost& H = ctx->p2p;
  if (!H.enabled && !H.xenabled) return 0;
  cudaStreamSynchronize(ctx->stream);
  comm_barrier(ctx);                       // nobody writes into a window any more
  for (int r = 0; r < ctx->nranks; ++r) if (r != ctx->rank) {
    if (H.ipc) {
      if (H.win_peer[r]) cudaIpcCloseMemHandle(H.win_peer[r]);
      if (H.p_peer[r]) cudaIpcCloseMemHandle(H.p_peer[r]);
      if (H.stage_peer[r]) cudaIpcCloseMemHandle(H.stage_peer[r]);
    }
    H.win_peer[r] = nullptr; H.p_peer[r] = nullptr; H.stage_peer[r] = nullptr;
  }
  comm_barrier(ctx);                       // every mapping is closed before the owners free their memory
  cudaFree(H.win_local);
  cudaFree(H.stage_local);
  H.win_local = nullptr; H.stage_local = nullptr; H.stage_cap = 0;
  H.enabled = false;
  H.xenabled = false;
  return 0;
}

void comm_destroy(Context* ctx) {
#ifdef DDPS_WITH_NCCL
  if (ctx->comm) { ncclCommDestroy(ctx->comm); ctx->comm = nullptr; }
#endif
  ctx->lgroup.reset();
  plan_free(ctx->halo);
}

}  // namespace ddps
