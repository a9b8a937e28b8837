// Multi-GPU plumbing: one process per GPU, NCCL over NVLink for the two exchange steps the path
// has -- the SpMV halo (ghost values of the search direction / of the fields) and the scalar
// allreduces of the CG dot products and stop tests.  The reference has no communication at all
// ("Distributed communication backend: none exists").
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

__global__ void k_pack(int n, const int* __restrict__ idx, const double* __restrict__ vec, double* __restrict__ buf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = vec[idx[i]];
}

int halo_exchange(Context* ctx, double* vec) {
  if (ctx->nranks == 1) return 0;
#ifdef DDPS_WITH_NCCL
  HaloPlan& H = ctx->halo;
  if (H.nneigh == 0) return 0;
  if (H.nsend > 0) {
    k_pack<<<(H.nsend + 255) / 256, 256, 0, ctx->stream>>>(H.nsend, H.send_idx, vec, H.send_buf);
    ctx->stats.kernel_launches++;
  }
  DDPS_NCCL(ncclGroupStart());
  for (int j = 0; j < H.nneigh; ++j) {
    int ns = H.send_off[j + 1] - H.send_off[j], nr = H.recv_off[j + 1] - H.recv_off[j];
    if (ns > 0) DDPS_NCCL(ncclSend(H.send_buf + H.send_off[j], ns, ncclDouble, H.peer[j], ctx->comm, ctx->stream));
    if (nr > 0) DDPS_NCCL(ncclRecv(vec + ctx->n_own + H.recv_off[j], nr, ncclDouble, H.peer[j], ctx->comm, ctx->stream));
  }
  DDPS_NCCL(ncclGroupEnd());
  return 0;
#else
  ctx->fail(DDPS_ERR_COMM, "library built without NCCL");
  return DDPS_ERR_COMM;
#endif
}

int allreduce_sum(Context* ctx, double* dev, int count) {
  if (ctx->nranks == 1) return 0;
#ifdef DDPS_WITH_NCCL
  DDPS_NCCL(ncclAllReduce(dev, dev, count, ncclDouble, ncclSum, ctx->comm, ctx->stream));
  return 0;
#else
  return DDPS_ERR_COMM;
#endif
}

int allreduce_max(Context* ctx, double* dev, int count) {
  if (ctx->nranks == 1) return 0;
#ifdef DDPS_WITH_NCCL
  DDPS_NCCL(ncclAllReduce(dev, dev, count, ncclDouble, ncclMax, ctx->comm, ctx->stream));
  return 0;
#else
  return DDPS_ERR_COMM;
#endif
}

int comm_init(Context* ctx, const char* id128, int rank, int nranks) {
#ifdef DDPS_WITH_NCCL
  if (nranks < 1 || rank < 0 || rank >= nranks) { ctx->fail(DDPS_ERR_ARG, "bad rank/nranks"); return DDPS_ERR_ARG; }
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks == 1) return 0;
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
#ifdef DDPS_WITH_NCCL
  DDPS_NCCL(ncclAllReduce(&ctx->scal->norm_inf, &ctx->scal->norm_inf, 1, ncclDouble, ncclMax, ctx->comm, ctx->stream));
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
#endif
  return 0;
}

// Map every peer's search-direction vector p and its P2PWin into this process (CUDA IPC over NVLink).
// The handles travel through one ncclAllGather; afterwards the PCG inner loop makes no NCCL call at all:
// the halo is pushed straight into the neighbours' ghost blocks and the dot products are exchanged
// through the windows by the kernels that produce / consume them (kernels.cu).
int p2p_setup(Context* ctx) {
  P2PHost& H = ctx->p2p;
  H.enabled = false;
  if (ctx->nranks == 1) return 0;
#ifdef DDPS_WITH_NCCL
  const char* env = getenv("DDPS_P2P");
  if (env && env[0] == '0') return 0;
  if (ctx->nranks > kMaxRanks) return 0;
  const int me = ctx->rank, nr = ctx->nranks;
  struct Xchg { cudaIpcMemHandle_t hp, hw; int off[kMaxRanks]; int device; int ok; };
  DDPS_CUDA(cudaMalloc(&H.win_local, sizeof(P2PWin)));
  DDPS_CUDA(cudaMemset(H.win_local, 0, sizeof(P2PWin)));
  Xchg mine;
  memset(&mine, 0, sizeof(mine));
  mine.device = ctx->device;
  mine.ok = 1;
  if (cudaIpcGetMemHandle(&mine.hp, ctx->p) != cudaSuccess || cudaIpcGetMemHandle(&mine.hw, H.win_local) != cudaSuccess) {
    cudaGetLastError();
    mine.ok = 0;
  }
  for (int q = 0; q < kMaxRanks; ++q) mine.off[q] = -1;
  for (int j = 0; j < ctx->halo.nneigh; ++j) mine.off[ctx->halo.peer[j]] = ctx->n_own + ctx->halo.recv_off[j];
  Xchg *d_send = nullptr, *d_recv = nullptr;
  DDPS_CUDA(cudaMalloc(&d_send, sizeof(Xchg)));
  DDPS_CUDA(cudaMalloc(&d_recv, sizeof(Xchg) * nr));
  DDPS_CUDA(cudaMemcpyAsync(d_send, &mine, sizeof(Xchg), cudaMemcpyHostToDevice, ctx->stream));
  DDPS_NCCL(ncclAllGather(d_send, d_recv, sizeof(Xchg), ncclUint8, ctx->comm, ctx->stream));
  std::vector<Xchg> all(nr);
  DDPS_CUDA(cudaMemcpyAsync(all.data(), d_recv, sizeof(Xchg) * nr, cudaMemcpyDeviceToHost, ctx->stream));
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_send); cudaFree(d_recv);
  bool ok = true;
  for (int r = 0; r < nr; ++r) ok = ok && all[r].ok;
  if (ok) {
    for (int r = 0; r < nr && ok; ++r) {
      if (r == me) { H.win_peer[r] = H.win_local; H.p_peer[r] = ctx->p; continue; }
      int can = 0;
      cudaDeviceCanAccessPeer(&can, ctx->device, all[r].device);
      if (!can) { ok = false; break; }
      if (cudaIpcOpenMemHandle((void**)&H.win_peer[r], all[r].hw, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
      if (ok && cudaIpcOpenMemHandle((void**)&H.p_peer[r], all[r].hp, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
    }
    if (!ok) cudaGetLastError();
  }
  // every rank must take the same decision: reduce the flag
  double flag = ok ? 1.0 : 0.0;
  DDPS_CUDA(cudaMemcpyAsync(&ctx->scal->norm_aux2, &flag, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  DDPS_NCCL(ncclAllReduce(&ctx->scal->norm_aux2, &ctx->scal->norm_aux2, 1, ncclDouble, ncclMin, ctx->comm, ctx->stream));
  DDPS_CUDA(cudaMemcpyAsync(&flag, &ctx->scal->norm_aux2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DDPS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (flag < 0.5) {   // fall back to the NCCL path everywhere
    for (int r = 0; r < nr; ++r) if (r != me) {
      if (H.win_peer[r]) cudaIpcCloseMemHandle(H.win_peer[r]);
      if (H.p_peer[r]) cudaIpcCloseMemHandle(H.p_peer[r]);
      H.win_peer[r] = nullptr; H.p_peer[r] = nullptr;
    }
    cudaFree(H.win_local); H.win_local = nullptr;
    return 0;
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
  H.tag = 0;
  H.enabled = true;
#endif
  return 0;
}

int p2p_teardown(Context* ctx) {
  P2PHost& H = ctx->p2p;
  if (!H.enabled) return 0;
  cudaStreamSynchronize(ctx->stream);
  comm_barrier(ctx);                       // nobody writes into a window any more
  for (int r = 0; r < ctx->nranks; ++r) if (r != ctx->rank) {
    if (H.win_peer[r]) cudaIpcCloseMemHandle(H.win_peer[r]);
    if (H.p_peer[r]) cudaIpcCloseMemHandle(H.p_peer[r]);
    H.win_peer[r] = nullptr; H.p_peer[r] = nullptr;
  }
  comm_barrier(ctx);                       // every mapping is closed before the owners free their memory
  cudaFree(H.win_local);
  H.win_local = nullptr;
  H.enabled = false;
  return 0;
}

void comm_destroy(Context* ctx) {
#ifdef DDPS_WITH_NCCL
  if (ctx->comm) { ncclCommDestroy(ctx->comm); ctx->comm = nullptr; }
#endif
  if (ctx->halo.send_idx) cudaFree(ctx->halo.send_idx);
  if (ctx->halo.send_buf) cudaFree(ctx->halo.send_buf);
  ctx->halo = HaloPlan();
}

}  // namespace ddps
