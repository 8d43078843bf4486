// TMA throughput of the window-tile loads alone (no compute): 148 x CTAS_PER_SM CTAs of 64 threads stream every
// (window, head) unit of a (B,200,200,288) bf16 tensor through a two-stage shared-memory ring.
//   variant 0: 6 ops / unit  -- 4-D box (32 ch, 4 tokens, 7 rows), SWIZZLE_64B          (what wmsa_tma.cu does)
//   variant 1: 3 ops / unit  -- 4-D box (32 ch, 8 tokens, 7 rows), SWIZZLE_64B
//   variant 2: 1 op  / unit  -- 5-D box (32 ch, 8 tokens, 7 rows, 3 parts, 1 head), SWIZZLE_64B
//   variant 3: 3 ops / unit  -- like 1 with SWIZZLE_NONE
// Prints GB/s of useful (49-token) bytes and of fetched bytes.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../rscotr_b200/csrc/tc_common.cuh"
using namespace rsc::tc;

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
               "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma5(uint32_t dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
               "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

constexpr int H = 200, W = 200, C = 96, HEADS = 3, NW = 29;

__global__ void __launch_bounds__(64) stream(const __grid_constant__ CUtensorMap map, int variant, int B, unsigned long long *sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[2];
  const uint32_t sb = smem_u32(smem);
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
  __syncthreads();
  const int num_items = B * NW * NW * HEADS;
  const uint32_t bytes = variant == 0 ? 6 * 1792 : 3 * 3584;
  auto issue = [&](int item, int s) {
    const int head = item % HEADS, win = item / HEADS;
    const int ww = win % NW, wh = (win / NW) % NW, b = win / (NW * NW);
    const uint32_t dst = sb + s * 12288;
    mbar_expect_tx(&full[s], bytes);
    if (variant == 0) {
      for (int p = 0; p < 3; ++p) {
        tma4(dst + p * 4096, &map, &full[s], p * C + head * 32, ww * 7, wh * 7, b);
        tma4(dst + p * 4096 + 1792, &map, &full[s], p * C + head * 32, ww * 7 + 4, wh * 7, b);
      }
    } else if (variant == 1 || variant == 3) {
      for (int p = 0; p < 3; ++p) tma4(dst + p * 3584, &map, &full[s], p * C + head * 32, ww * 7, wh * 7, b);
    } else {
      tma5(dst, &map, &full[s], 0, ww * 7, b * H + wh * 7, 0, head);
    }
  };
  int item = blockIdx.x;
  if (tid == 0) {
    if (item < num_items) issue(item, 0);
    if (item + (int)gridDim.x < num_items) issue(item + gridDim.x, 1);
  }
  unsigned long long acc = 0;
  for (int it = 0; item < num_items; ++it, item += gridDim.x) {
    const int s = it & 1;
    mbar_wait(&full[s], (it >> 1) & 1);
    acc += *reinterpret_cast<volatile uint32_t *>(smem + s * 12288 + tid * 64);
    __syncthreads();
    if (tid == 0 && item + 2 * (int)gridDim.x < num_items) issue(item + 2 * gridDim.x, s);
  }
  if (acc == 0x123456789ull) *sink = acc;
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
  const int B = 16;
  const int ctas = argc > 1 ? atoi(argv[1]) : 8;
  __nv_bfloat16 *x;
  const size_t n = (size_t)B * H * W * 3 * C;
  CHECK(cudaMalloc(&x, n * 2));
  CHECK(cudaMemset(x, 0, n * 2));
  unsigned long long *sink;
  CHECK(cudaMalloc(&sink, 8));
  EncodeTiled enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &q));
  CHECK(cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576));
  for (int variant = 0; variant < 4; ++variant) {
    CUtensorMap map;
    CUresult r;
    const CUtensorMapSwizzle sw = variant == 3 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_64B;
    if (variant != 2) {
      const cuuint64_t dims[4] = {3 * C, W, H, (cuuint64_t)B};
      const cuuint64_t strides[3] = {3 * C * 2, (cuuint64_t)W * 3 * C * 2, (cuuint64_t)H * W * 3 * C * 2};
      const cuuint32_t box[4] = {32, (cuuint32_t)(variant == 0 ? 4 : 8), 7, 1}, es[4] = {1, 1, 1, 1};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {   // (32 ch, W, B*H, 3 parts, heads)
      const cuuint64_t dims[5] = {32, W, (cuuint64_t)B * H, 3, HEADS};
      const cuuint64_t strides[4] = {3 * C * 2, (cuuint64_t)W * 3 * C * 2, C * 2, 32 * 2};
      const cuuint32_t box[5] = {32, 8, 7, 3, 1}, es[5] = {1, 1, 1, 1, 1};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("variant %d: encode failed %d\n", variant, (int)r); continue; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      stream<<<148 * ctas, 64, 24576>>>(map, variant, B, sink);
      cudaEventRecord(e1);
      CHECK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double items = (double)B * NW * NW * HEADS;
    printf("variant %d (%d CTAs/SM): %.1f us  useful %.0f GB/s  fetched %.0f GB/s\n", variant, ctas, best * 1e3,
           items * 49 * 64 * 3 / best / 1e6, items * (variant == 0 ? 6 * 1792 : 3 * 3584) / best / 1e6);
  }
  return 0;
}
