// Gradient all-reduce inside the NVSwitch (row 8e): the step engine's flat fp32 gradient buffer lives in symmetric
// memory with an NVLS multicast mapping (one virtual address that reaches the same offset on every GPU of the box).
// Rank r owns 1/world of the range: `multimem.ld_reduce.add.v4.f32` returns the SUM over all GPUs computed by the switch,
// the kernel scales it (1/world: the mean the reference's DDP wrapper produces, mtl/apis/train.py:37-46) and
// `multimem.st.v4.f32` writes it back to every GPU -- each gradient element crosses each NVLink once per direction, no
// ring / tree steps, no staging buffers, and the CTAs carry no shared memory, so they co-reside with the persistent GEMM
// CTAs of the backward pass that this exchange overlaps.  The cross-GPU rendezvous before (all ranks have written the
// range) and after (all slices are stored) are the symmetric-memory barriers of the caller (mtl/engine/step.py).
#include "common.cuh"

namespace rsc {
namespace nvls {

__device__ __forceinline__ float4 mc_ld_reduce(const float *mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float *mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

constexpr int UNROLL = 4;

// v4 elements [v0, v1) of the multicast view
__global__ void __launch_bounds__(512) allreduce_mean_kernel(float *__restrict__ mc, int64_t v0, int64_t v1, float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = v0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < v1; i += UNROLL * stride) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = mc_ld_reduce(mc + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      v[u].x *= scale, v[u].y *= scale, v[u].z *= scale, v[u].w *= scale;
      mc_st(mc + 4 * (i + u * stride), v[u]);
    }
  }
  for (; i < v1; i += stride) {
    float4 v = mc_ld_reduce(mc + 4 * i);
    v.x *= scale, v.y *= scale, v.z *= scale, v.w *= scale;
    mc_st(mc + 4 * i, v);
  }
}

}  // namespace nvls
}  // namespace rsc

using namespace rsc;

// mc: multicast address of element 0 of the symmetric buffer; [lo, hi) float elements (multiples of 4, 16-byte aligned);
// this rank reduces its 1/world share of the range.  Caller: barrier -> this -> barrier.
extern "C" int rsc_nvls_allreduce_mean(void *mc, int64_t lo, int64_t hi, int rank, int world, float scale, int ctas, void *stream) {
  RSC_CHECK_ARG(mc && world > 0 && rank >= 0 && rank < world && hi > lo, "rsc_nvls_allreduce_mean: bad arguments");
  RSC_CHECK_ARG(lo % 4 == 0 && hi % 4 == 0 && ((uintptr_t)mc & 15) == 0, "rsc_nvls_allreduce_mean: range must be 16-byte aligned");
  const int64_t nv = (hi - lo) / 4, per = (nv + world - 1) / world;
  const int64_t v0 = lo / 4 + (int64_t)rank * per, v1 = lo / 4 + (((int64_t)rank + 1) * per < nv ? ((int64_t)rank + 1) * per : nv);
  if (v1 <= v0) return RSC_OK;
  if (ctas <= 0) ctas = 64;
  int64_t want = (v1 - v0 + 512 * nvls::UNROLL - 1) / (512 * nvls::UNROLL);
  const int grid = (int)(want < ctas ? want : ctas);
  nvls::allreduce_mean_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>((float *)mc, v0, v1, scale);
  RSC_CHECK_LAUNCH("rsc_nvls_allreduce_mean");
  return RSC_OK;
}
