// Probe for the TMA-staged window loader the next window-attention kernel wants (run on a B200):
//   * a 4-D tensor map over a (B, H, W, 3C) bf16 activation, box = (1, 7, 7, 32): ONE bulk-tensor copy fetches the
//     q (or k, or v) slice of one head for one 7x7 window as 49 consecutive 64-byte rows, out-of-bounds tokens
//     (right / bottom padding of the map) arrive as zeros;
//   * CU_TENSOR_MAP_SWIZZLE_64B on the TMA side + SWIZZLE_64B K-major shared-memory descriptors on the UMMA side:
//     S = Q K^T (M=128 over a 64-row tile, N=64, K=32) straight from the TMA-written tiles.
// Prints where every 16-byte chunk of the tile landed (the swizzle pattern) and checks S against the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe_tma tools/tc_probe_tma.cu
// Usage: tc_probe_tma [h0 w0]      (window origin; default 14 14 on a 20x20 map -> one padded row and column)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../rscotr_b200/csrc/tc_common.cuh"

using namespace rsc::tc;

#define CHECK(x)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                       \
    }                                                                                 \
  } while (0)

constexpr int H = 20, W = 20, C3 = 96, HD = 32, WS = 7, NT = WS * WS;
constexpr uint32_t TILE_BYTES = 64 * 64;          // 64 rows x 32 bf16
constexpr uint32_t BOX_BYTES = NT * HD * 2;        // what one TMA box delivers

// shared-memory matrix descriptor, K-major, SWIZZLE_64B: 8-row groups 512 bytes apart (rows are 64 bytes)
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(1) << 16;                        // LBO: unused for swizzled K-major (canonical value 1)
  d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;      // SBO
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                          // layout_type = SWIZZLE_64B
  return d;
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map, int h0, int w0, float *S_out, uint8_t *raw) {
  __shared__ __align__(1024) uint8_t sQ[TILE_BYTES];
  __shared__ __align__(1024) uint8_t sK[TILE_BYTES];
  __shared__ __align__(1024) uint8_t guard[TILE_BYTES];   // rows 64..127 of the M=128 A operand read into here / sK
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 64);
  for (int i = tid; i < (int)TILE_BYTES / 16; i += 128) {
    reinterpret_cast<uint4 *>(sQ)[i] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4 *>(sK)[i] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4 *>(guard)[i] = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (tid == 0) {
    mbar_expect_tx(&bar_tma, 2 * BOX_BYTES);
    tma_load_4d(sQ, &map, &bar_tma, 0 * HD, w0, h0, 0);     // q channels of head 0
    tma_load_4d(sK, &map, &bar_tma, 1 * HD, w0, h0, 0);     // (the probe tensor packs q | k | v as 3 x 32 channels)
  }
  mbar_wait(&bar_tma, 0);
  for (int i = tid; i < (int)TILE_BYTES; i += 128) raw[i] = sQ[i];
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    fence_after_sync();
    const uint32_t idesc = make_idesc_bf16(128, 64, false, false);
#pragma unroll
    for (int k = 0; k < 2; ++k)      // K = 32 = 2 x 16 elements; the second step starts 32 bytes into the 64-byte rows
      mma_bf16_ss(tm, make_desc_sw64(smem_u32(sQ) + k * 32), make_desc_sw64(smem_u32(sK) + k * 32), idesc, k > 0);
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  fence_after_sync();
  if (warp < 2) {
    uint32_t r[32];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 64; c += 32) {
      tmem_ld32(taddr + c, r);
      tmem_ld_wait();
      for (int k = 0; k < 32; ++k) S_out[tid * 64 + c + k] = __uint_as_float(r[k]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
  const int h0 = argc > 2 ? atoi(argv[1]) : 14, w0 = argc > 2 ? atoi(argv[2]) : 14;
  std::vector<__nv_bfloat16> x((size_t)H * W * C3);
  std::vector<float> xf(x.size());
  srand(1);
  for (size_t i = 0; i < x.size(); ++i) {
    xf[i] = (float)(rand() % 7 - 3);               // small integers: bf16 products and fp32 sums are exact
    x[i] = __float2bfloat16(xf[i]);
  }
  __nv_bfloat16 *dx;
  float *dS;
  uint8_t *draw;
  CHECK(cudaMalloc(&dx, x.size() * 2));
  CHECK(cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice));
  CHECK(cudaMalloc(&dS, 64 * 64 * 4));
  CHECK(cudaMalloc(&draw, TILE_BYTES));
  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
  if (!encode || qres != cudaDriverEntryPointSuccess) {
    printf("cuTensorMapEncodeTiled not available\n");
    return 2;
  }
  CUtensorMap map;
  const cuuint64_t dims[4] = {(cuuint64_t)C3, (cuuint64_t)W, (cuuint64_t)H, 1};
  const cuuint64_t strides[3] = {(cuuint64_t)C3 * 2, (cuuint64_t)W * C3 * 2, (cuuint64_t)H * W * C3 * 2};
  const cuuint32_t box[4] = {HD, WS, WS, 1}, estr[4] = {1, 1, 1, 1};
  const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return 2;
  }
  probe<<<1, 128>>>(map, h0, w0, dS, draw);
  CHECK(cudaGetLastError());
  CHECK(cudaDeviceSynchronize());
  std::vector<float> S(64 * 64);
  std::vector<uint8_t> raw(TILE_BYTES);
  CHECK(cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(raw.data(), draw, raw.size(), cudaMemcpyDeviceToHost));
  // token t of the window = (h0 + t / 7, w0 + t % 7); out-of-bounds tokens are zero rows
  auto val = [&](int t, int part, int c) -> float {
    const int hh = h0 + t / WS, ww = w0 + t % WS;
    if (t >= NT || hh < 0 || hh >= H || ww < 0 || ww >= W) return 0.f;
    return xf[((size_t)hh * W + ww) * C3 + part * HD + c];
  };
  // 1. where did each 16-byte chunk (8 channels) of each token row land?  (signature = the 8 bf16 values)
  int moved = 0, lost = 0;
  printf("swizzle pattern of the TMA-written q tile (row r, chunk c -> byte offset), first 16 rows:\n");
  for (int t = 0; t < NT; ++t)
    for (int c = 0; c < 4; ++c) {
      __nv_bfloat16 want[8];
      bool allzero = true;
      for (int e = 0; e < 8; ++e) {
        want[e] = __float2bfloat16(val(t, 0, c * 8 + e));
        allzero &= val(t, 0, c * 8 + e) == 0.f;
      }
      int found = -1;
      for (int cc = 0; cc < 4 && found < 0; ++cc)
        if (!memcmp(want, raw.data() + t * 64 + cc * 16, 16)) found = cc;
      if (found < 0) ++lost;
      else if (found != c && !allzero) ++moved;
      if (t < 16) printf("  r%2d c%d -> %s%d%s", t, c, found < 0 ? "?" : "chunk ", found, c == 3 ? "\n" : "");
    }
  printf("chunks displaced within their row: %d, chunks not found in their row: %d (of %d)\n", moved, lost, NT * 4);
  // 2. S = Q K^T
  double maxerr = 0;
  for (int i = 0; i < NT; ++i)
    for (int j = 0; j < NT; ++j) {
      float ref = 0.f;
      for (int c = 0; c < HD; ++c) ref += val(i, 0, c) * val(j, 1, c);
      maxerr = fmax(maxerr, fabs((double)S[i * 64 + j] - ref));
    }
  printf("S = Q K^T from the TMA tiles (window origin %d,%d): max |err| = %g  -> %s\n", h0, w0, maxerr, maxerr == 0 ? "MATCH" : "MISMATCH");
  return maxerr == 0 ? 0 : 1;
}
