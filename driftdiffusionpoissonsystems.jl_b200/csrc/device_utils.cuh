// Device helpers shared by kernels.cu and amg.cu: deterministic warp/block/grid reductions and streaming loads.
#pragma once

#include "ddps_internal.cuh"

namespace ddps {

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch: the kernels of the CG iteration are launched with the programmatic-stream-serialization
// attribute, so the NEXT kernel's CTAs are scheduled while the tail of the current one is still running and find their launch
// latency hidden; each such kernel starts with pdl_enter(): wait until the preceding kernel has completed and its writes are
// visible (nothing before this touches memory), then allow its own dependent to be scheduled.  Both instructions are no-ops for
// a kernel launched the ordinary way.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
}

inline bool pdl_enabled() {
  static const bool on = getenv("DDPS_NO_PDL") == nullptr;
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

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

// Same with a per-component operation: bit i of MAXMASK set -> component i is a max, else a sum.
template <int NV, unsigned MAXMASK>
__device__ __forceinline__ bool grid_reduce_mixed(double (&v)[NV], double* partials, unsigned* ticket, double* sh,
                                                  double (&total)[NV]) {
  __shared__ bool s_last;
  double bv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) bv[i] = ((MAXMASK >> i) & 1) ? block_reduce<true>(v[i], sh) : block_reduce<false>(v[i], sh);
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
    const bool mx = (MAXMASK >> i) & 1;
    double s = 0.0;
    const volatile double* pp = partials + (size_t)i * gridDim.x;
    for (unsigned j = threadIdx.x; j < gridDim.x; j += blockDim.x) s = mx ? fmax(s, pp[j]) : s + pp[j];
    total[i] = mx ? block_reduce<true>(s, sh) : block_reduce<false>(s, sh);
  }
  return true;
}

__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }

}  // namespace ddps
