// Stand-alone probe of the no-swizzle tcgen05 building blocks (round 1; the P / dS tiles of wmsa_tma.cu still use them):
//   S = [Q_A;Q_B] (128x32, K-major) x [K_A;K_B]^T (128x32, K-major) -> TMEM 128x128 fp32
//   O = P (128x64, K-major)  x  V (64 keys x 32 dims, MN-major B)   -> TMEM 128x32 fp32
// with the no-swizzle core-matrix smem layouts, checked against a host fp32
// reference.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe tc_probe.cu
// Exit code 0 = both MMAs match.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../rscotr_b200/csrc/tc_common.cuh"

using namespace rsc::tc;

// token-major tile layout shared by Q, K, V: 16-byte chunk c (8 bf16) of row r at
//   (r/8)*ROWGRP + c*128 + (r%8)*16      (ROWGRP = 128 * chunks per row)
__host__ __device__ inline int tile_off(int r, int c, int chunks) { return (r / 8) * (128 * chunks) + c * 128 + (r % 8) * 16; }
// P tile (128 rows x 64 keys): chunk kc of row r at kc*2048 + (r/8)*128 + (r%8)*16
__host__ __device__ inline int p_off(int r, int kc) { return kc * 2048 + (r / 8) * 128 + (r % 8) * 16; }

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16 *Q, const __nv_bfloat16 *K, const __nv_bfloat16 *V,
                                             const __nv_bfloat16 *P, float *S_out, float *O_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sQ = smem;            // 128 x 32 bf16 = 8 KB
  uint8_t *sK = sQ + 8192;       // 8 KB
  uint8_t *sV = sK + 8192;       // 64 x 32 = 4 KB
  uint8_t *sP = sV + 4096;       // 128 x 64 = 16 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  // fill smem tiles (16-byte chunks)
  for (int i = tid; i < 128 * 4; i += 128) {
    int r = i / 4, c = i % 4;
    *reinterpret_cast<uint4 *>(sQ + tile_off(r, c, 4)) = *reinterpret_cast<const uint4 *>(Q + r * 32 + c * 8);
    *reinterpret_cast<uint4 *>(sK + tile_off(r, c, 4)) = *reinterpret_cast<const uint4 *>(K + r * 32 + c * 8);
    if (r < 64) *reinterpret_cast<uint4 *>(sV + tile_off(r, c, 4)) = *reinterpret_cast<const uint4 *>(V + r * 32 + c * 8);
  }
  for (int i = tid; i < 128 * 8; i += 128) {
    int r = i / 8, kc = i % 8;
    *reinterpret_cast<uint4 *>(sP + p_off(r, kc)) = *reinterpret_cast<const uint4 *>(P + r * 64 + kc * 8);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base;

  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 128, false, false);
#pragma unroll
    for (int k = 0; k < 2; ++k) {  // K = 32 = 2 x 16; one K step = 2 chunks = 256 B
      uint64_t a = make_smem_desc(smem_u32(sQ) + k * 256, 128, 512);
      uint64_t b = make_smem_desc(smem_u32(sK) + k * 256, 128, 512);
      mma_bf16_ss(tm, a, b, idesc, k > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  {
    uint32_t r[32];
    for (int c0 = 0; c0 < 128; c0 += 32) {
      tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) S_out[tid * 128 + c0 + j] = __uint_as_float(r[j]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    fence_after_sync();
    const uint32_t idesc = make_idesc_bf16(128, 32, false, true);
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // K = 64 keys = 4 x 16
      uint64_t a = make_smem_desc(smem_u32(sP) + k * 4096, 2048, 128);   // K-major: LBO = K-chunk stride, SBO = row-group stride
      uint64_t b = make_smem_desc(smem_u32(sV) + k * 1024, 512, 128);    // MN-major: LBO = key-group stride, SBO = dim-chunk stride
      mma_bf16_ss(tm + 128, a, b, idesc, k > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 1);
  fence_after_sync();
  {
    uint32_t r[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + 128, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) O_out[tid * 32 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      return 2;                                                                \
    }                                                                          \
  } while (0)

int main() {
  std::vector<__nv_bfloat16> Q(128 * 32), K(128 * 32), V(64 * 32), P(128 * 64);
  srand(1);
  auto rnd = []() { return (rand() % 2001 - 1000) / 1000.0f; };
  for (auto &x : Q) x = __float2bfloat16(rnd());
  for (auto &x : K) x = __float2bfloat16(rnd());
  for (auto &x : V) x = __float2bfloat16(rnd());
  for (auto &x : P) x = __float2bfloat16(rnd());
  __nv_bfloat16 *dQ, *dK, *dV, *dP;
  float *dS, *dO;
  CK(cudaMalloc(&dQ, Q.size() * 2));
  CK(cudaMalloc(&dK, K.size() * 2));
  CK(cudaMalloc(&dV, V.size() * 2));
  CK(cudaMalloc(&dP, P.size() * 2));
  CK(cudaMalloc(&dS, 128 * 128 * 4));
  CK(cudaMalloc(&dO, 128 * 32 * 4));
  CK(cudaMemcpy(dQ, Q.data(), Q.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dK, K.data(), K.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dV, V.data(), V.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dP, P.data(), P.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
  probe<<<1, 128, 36864>>>(dQ, dK, dV, dP, dS, dO);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> S(128 * 128), O(128 * 32);
  CK(cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  double es = 0, eo = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j) {
      float ref = 0;
      for (int k = 0; k < 32; ++k) ref += __bfloat162float(Q[i * 32 + k]) * __bfloat162float(K[j * 32 + k]);
      es = fmax(es, fabs(ref - S[i * 128 + j]));
    }
  for (int i = 0; i < 128; ++i)
    for (int d = 0; d < 32; ++d) {
      float ref = 0;
      for (int k = 0; k < 64; ++k) ref += __bfloat162float(P[i * 64 + k]) * __bfloat162float(V[k * 32 + d]);
      eo = fmax(eo, fabs(ref - O[i * 32 + d]));
    }
  printf("tc_probe: max|S err| = %.3e  max|O err| = %.3e\n", es, eo);
  printf("S[0][0..3] = %f %f %f %f   O[0][0..3] = %f %f %f %f\n", S[0], S[1], S[2], S[3], O[0], O[1], O[2], O[3]);
  bool ok = es < 1e-3 && eo < 1e-3;
  printf(ok ? "tc_probe PASS\n" : "tc_probe FAIL\n");
  return ok ? 0 : 1;
}
