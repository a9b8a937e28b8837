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

void comm_destroy(Context* ctx) {
#ifdef DDPS_WITH_NCCL
  if (ctx->comm) { ncclCommDestroy(ctx->comm); ctx->comm = nullptr; }
#endif
  if (ctx->halo.send_idx) cudaFree(ctx->halo.send_idx);
  if (ctx->halo.send_buf) cudaFree(ctx->halo.send_buf);
  ctx->halo = HaloPlan();
}

}  // namespace ddps
