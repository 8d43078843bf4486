// Hardware probe for the TMA-staged window-attention kernels (csrc/wmsa_tma.cu).  Run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/wmsa_probe tools/wmsa_probe.cu && /tmp/wmsa_probe <mode>
// What it pins (each against a host fp32 reference on small-integer data, so every result is exact):
//   * the 56-slot window tile: a 7x7 window is staged as [half][7 rows][4 tokens] x 32 channels (64-byte rows),
//     half 0 = window columns 0..3, half 1 = columns 4..7 (column 7 is a neighbour's token: never a valid key);
//     two TMA boxes (32 ch, 4 tokens, 7 rows) per operand, CU_TENSOR_MAP_SWIZZLE_64B, destination offsets 0 and
//     1792 bytes; windows that wrap around the rolled map use (4 rows) + (3 rows) boxes from two more tensor maps;
//     out-of-bounds tokens (the zero padding of the map, negative coordinates) arrive as zeros;
//   * S = Q K^T  with both operands K-major SWIZZLE_64B descriptors on the TMA-written tiles (M=128, N=64, K=32);
//   * O = P V    with P stored by threads in the no-swizzle core-matrix layout (K-major A) and V read as an
//     MN-major SWIZZLE_64B B operand straight from its TMA tile (K = 64 key slots);
//   * T = P^T Q  with P read MN-major (no swizzle) and Q as MN-major SWIZZLE_64B B operand (the dV / dK shape).
// Modes: "plain h0 w0" (one window at origin h0,w0; default 7 7), "wrap" (last window of a shifted 20x20 map:
// wraps in h and w, one padded row / column).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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

constexpr int H = 20, W = 20, C3 = 96, HD = 32, WS = 7;
constexpr uint32_t TILE = 4096;   // 64 slots x 64 bytes
constexpr uint32_t HALF = 1792;   // 28 slots

__host__ __device__ inline bool slot_valid(int s) { return s < 56 && (s / 28) * 4 + s % 4 < WS; }
__host__ __device__ inline int slot_i(int s) { return (s % 28) / 4; }
__host__ __device__ inline int slot_j(int s) { return (s / 28) * 4 + s % 4; }
__host__ __device__ inline int p_off(int r, int kc) { return kc * 1024 + (r >> 3) * 128 + (r & 7) * 16; }

// K-major SWIZZLE_64B: rows of 64 bytes, 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;   // SWIZZLE_64B
  return d;
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct Plan {   // up to 4 boxes per operand
  int n;
  int which[4];   // 0: 7 rows, 1: 4 rows, 2: 3 rows
  int w[4], h[4];
  uint32_t dst[4], bytes[4];
};

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap m7, const __grid_constant__ CUtensorMap m4,
                                             const __grid_constant__ CUtensorMap m3, Plan plan, const __nv_bfloat16 *P,
                                             float *S_out, float *O_out, float *T_out, uint8_t *raw, long long *clk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // [Q | K | V | P 8 KB | guard 8 KB]
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sb = smem_u32(smem);
  const uint32_t sQ = sb, sK = sb + TILE, sV = sb + 2 * TILE, sP = sb + 3 * TILE;
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  for (int i = tid; i < (int)(3 * TILE + 16384) / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  __syncthreads();
  // P tile (64 rows x 64 key slots) in the no-swizzle core-matrix layout
  for (int i = tid; i < 64 * 8; i += 128) {
    const int r = i / 8, kc = i % 8;
    *reinterpret_cast<uint4 *>(smem + 3 * TILE + p_off(r, kc)) = *reinterpret_cast<const uint4 *>(P + r * 64 + kc * 8);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  long long t0 = clock64();
  if (tid == 0) {
    uint32_t total = 0;
    for (int k = 0; k < plan.n; ++k) total += 3 * plan.bytes[k];
    mbar_expect_tx(&bar_tma, total);
    for (int part = 0; part < 3; ++part)
      for (int k = 0; k < plan.n; ++k) {
        const CUtensorMap *m = plan.which[k] == 0 ? &m7 : (plan.which[k] == 1 ? &m4 : &m3);
        tma_load_4d(sb + part * TILE + plan.dst[k], m, &bar_tma, part * HD, plan.w[k], plan.h[k], 0);
      }
  }
  mbar_wait(&bar_tma, 0);
  long long t1 = clock64();
  for (int i = tid; i < (int)TILE; i += 128) raw[i] = smem[i];
  fence_before_sync();
  __syncthreads();
  const uint32_t tm = tmem_base;
  // latency of a minimal MMA batch (2 x M128 N64 K16 + commit -> mbarrier wake-up), measured 3 times
  __shared__ uint64_t bar_lat;
  if (tid == 0) {
    mbar_init(&bar_lat, 1);
    mbar_fence_init();
  }
  __syncthreads();
  for (int rep = 0; rep < 3; ++rep) {
    long long a = clock64();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
      for (int k = 0; k < 2; ++k)
        mma_bf16_ss(tm, desc_sw64(sQ + k * 32, 16, 512), desc_sw64(sK + k * 32, 16, 512), idesc_s, k > 0);
      mma_commit(&bar_lat);
    }
    mbar_wait(&bar_lat, rep & 1);
    long long b = clock64();
    fence_after_sync();
    uint32_t r[32];
    if (warp < 4) {
      tmem_ld32(tm + ((uint32_t)((warp & 3) * 32) << 16), r);
      tmem_ld_wait();
    }
    long long c = clock64();
    if (tid == 0) clk[2 + 2 * rep] = b - a, clk[3 + 2 * rep] = c - b;
    fence_before_sync();
    __syncthreads();
  }
  if (tid == 0) clk[0] = t1 - t0;
  if (tid == 0) {
    fence_after_sync();
    const uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
#pragma unroll
    for (int k = 0; k < 2; ++k)   // K = 32 channels = 2 x 16: +32 bytes inside the 64-byte rows
      mma_bf16_ss(tm, desc_sw64(sQ + k * 32, 16, 512), desc_sw64(sK + k * 32, 16, 512), idesc_s, k > 0);
    const uint32_t idesc_o = make_idesc_bf16(128, 32, false, true);
#pragma unroll
    for (int k = 0; k < 4; ++k)   // K = 64 key slots = 4 x 16: V advances 16 rows = 1024 bytes
      mma_bf16_ss(tm + 64, make_smem_desc(sP + k * 2048, 1024, 128), desc_sw64(sV + k * 1024, 16, 512), idesc_o, k > 0);
    const uint32_t idesc_t = make_idesc_bf16(128, 32, true, true);
#pragma unroll
    for (int k = 0; k < 4; ++k)   // K = 64 query rows
      mma_bf16_ss(tm + 96, make_smem_desc(sP + k * 256, 128, 1024), desc_sw64(sQ + k * 1024, 16, 512), idesc_t, k > 0);
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  fence_after_sync();
  if (warp < 2) {
    uint32_t r[32];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 128; c += 32) {
      tmem_ld32(taddr + c, r);
      tmem_ld_wait();
      for (int k = 0; k < 32; ++k) {
        const float v = __uint_as_float(r[k]);
        if (c < 64) S_out[tid * 64 + c + k] = v;
        else if (c < 96) O_out[tid * 32 + k] = v;
        else T_out[tid * 32 + k] = v;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
  const bool wrap = argc > 1 && !strcmp(argv[1], "wrap");
  const int h0 = (!wrap && argc > 3) ? atoi(argv[2]) : 7, w0 = (!wrap && argc > 3) ? atoi(argv[3]) : 7;
  const int Hp = 21, Wp = 21, shift = 3;
  std::vector<__nv_bfloat16> x((size_t)H * W * C3), P(64 * 64);
  std::vector<float> xf(x.size()), Pf(P.size());
  srand(1);
  for (size_t i = 0; i < x.size(); ++i) {
    xf[i] = (float)(rand() % 7 - 3);
    x[i] = __float2bfloat16(xf[i]);
  }
  for (int r = 0; r < 64; ++r)
    for (int k = 0; k < 64; ++k) {
      Pf[r * 64 + k] = (slot_valid(r) && slot_valid(k)) ? (float)(rand() % 5 - 2) : 0.f;
      P[r * 64 + k] = __float2bfloat16(Pf[r * 64 + k]);
    }
  __nv_bfloat16 *dx, *dP;
  float *dS, *dO, *dT;
  uint8_t *draw;
  CHECK(cudaMalloc(&dx, x.size() * 2));
  CHECK(cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice));
  CHECK(cudaMalloc(&dP, P.size() * 2));
  CHECK(cudaMemcpy(dP, P.data(), P.size() * 2, cudaMemcpyHostToDevice));
  CHECK(cudaMalloc(&dS, 64 * 64 * 4));
  CHECK(cudaMalloc(&dO, 64 * 32 * 4));
  CHECK(cudaMalloc(&dT, 64 * 32 * 4));
  CHECK(cudaMalloc(&draw, TILE));
  long long *dclk;
  CHECK(cudaMalloc(&dclk, 16 * sizeof(long long)));
  CHECK(cudaMemset(dclk, 0, 16 * sizeof(long long)));
  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
  if (!encode || qres != cudaDriverEntryPointSuccess) {
    printf("cuTensorMapEncodeTiled not available\n");
    return 2;
  }
  CUtensorMap maps[3];
  const int rows[3] = {7, 4, 3};
  for (int m = 0; m < 3; ++m) {
    const cuuint64_t dims[4] = {(cuuint64_t)C3, (cuuint64_t)W, (cuuint64_t)H, 1};
    const cuuint64_t strides[3] = {(cuuint64_t)C3 * 2, (cuuint64_t)W * C3 * 2, (cuuint64_t)H * W * C3 * 2};
    const cuuint32_t box[4] = {HD, 4, (cuuint32_t)rows[m], 1}, estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(&maps[m], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
      return 2;
    }
  }
  // source token of slot s: (h, w) or out of bounds
  auto src = [&](int s, int &h, int &w) -> bool {
    if (!slot_valid(s) && !(s < 56)) return false;
    const int i = slot_i(s), j = slot_j(s);   // j == 7: the neighbour token (loaded, never valid)
    if (wrap) {
      h = Hp - WS + i + shift;
      if (h >= Hp) h -= Hp;
      w = Wp - WS + j + shift;
      if (j < 4) { if (w >= Wp) w -= Wp; } else w = j - 4;   // second half box starts at w = 0
    } else {
      h = h0 + i;
      w = w0 + j;
    }
    return h >= 0 && h < H && w >= 0 && w < W;
  };
  auto val = [&](int s, int part, int c) -> float {
    int h, w;
    if (s >= 56 || !src(s, h, w)) return 0.f;
    return xf[((size_t)h * W + w) * C3 + part * HD + c];
  };
  Plan plan;
  memset(&plan, 0, sizeof(plan));
  if (!wrap) {
    plan.n = 2;
    for (int k = 0; k < 2; ++k) {
      plan.which[k] = 0, plan.w[k] = w0 + 4 * k, plan.h[k] = h0, plan.dst[k] = k * HALF, plan.bytes[k] = 7 * 4 * 64;
    }
  } else {
    plan.n = 4;
    for (int k = 0; k < 4; ++k) {
      const int half = k & 1, lower = k >> 1;   // lower: window rows 4..6 <- map rows 0..2
      plan.which[k] = lower ? 2 : 1;
      plan.w[k] = half ? 0 : Wp - 4;
      plan.h[k] = lower ? 0 : Hp - 4;
      plan.dst[k] = half * HALF + (lower ? 16 * 64 : 0);
      plan.bytes[k] = (lower ? 3 : 4) * 4 * 64;
    }
  }
  CHECK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * TILE + 16384));
  probe<<<1, 128, 3 * TILE + 16384>>>(maps[0], maps[1], maps[2], plan, dP, dS, dO, dT, draw, dclk);
  CHECK(cudaGetLastError());
  CHECK(cudaDeviceSynchronize());
  std::vector<float> S(64 * 64), O(64 * 32), T(64 * 32);
  std::vector<uint8_t> raw(TILE);
  long long clk[16];
  CHECK(cudaMemcpy(clk, dclk, sizeof(clk), cudaMemcpyDeviceToHost));
  printf("CLOCKS : TMA issue->landed %lld | MMA batch issue->wake-up %lld %lld %lld | tcgen05.ld+wait %lld %lld %lld\n", clk[0], clk[2],
         clk[4], clk[6], clk[3], clk[5], clk[7]);
  CHECK(cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(T.data(), dT, T.size() * 4, cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(raw.data(), draw, raw.size(), cudaMemcpyDeviceToHost));
  // 1. swizzle pattern: chunk c of slot s expected at s*64 + ((c ^ ((s >> 1) & 3)) * 16)
  int bad = 0;
  for (int s = 0; s < 56; ++s)
    for (int c = 0; c < 4; ++c) {
      __nv_bfloat16 want[8];
      for (int e = 0; e < 8; ++e) want[e] = __float2bfloat16(val(s, 0, c * 8 + e));
      if (memcmp(want, raw.data() + s * 64 + ((c ^ ((s >> 1) & 3)) * 16), 16)) {
        if (bad < 8) {
          int found = -1;
          for (int cc = 0; cc < 4; ++cc)
            if (!memcmp(want, raw.data() + s * 64 + cc * 16, 16)) found = cc;
          printf("  slot %d chunk %d: not at the expected swizzled position (found in chunk %d)\n", s, c, found);
        }
        ++bad;
      }
    }
  int nz = 0;
  for (int b = 56 * 64; b < 64 * 64; ++b) nz += raw[b] != 0;
  printf("TILE   : %d of %d chunks misplaced, %d non-zero bytes in slots 56..63 -> %s\n", bad, 56 * 4, nz,
         (bad == 0 && nz == 0) ? "MATCH" : "MISMATCH");
  double es = 0, eo = 0, et = 0;
  for (int i = 0; i < 56; ++i)
    for (int j = 0; j < 56; ++j) {
      float ref = 0.f;
      for (int c = 0; c < HD; ++c) ref += val(i, 0, c) * val(j, 1, c);
      es = fmax(es, fabs((double)S[i * 64 + j] - ref));
    }
  for (int i = 0; i < 64; ++i)
    for (int d = 0; d < HD; ++d) {
      float ro = 0.f, rt = 0.f;
      for (int k = 0; k < 64; ++k) {
        ro += Pf[i * 64 + k] * val(k, 2, d);   // O[i] = sum_k P[i][k] V[k]
        rt += Pf[k * 64 + i] * val(k, 0, d);   // T[i] = sum_k P[k][i] Q[k]
      }
      eo = fmax(eo, fabs((double)O[i * 32 + d] - ro));
      et = fmax(et, fabs((double)T[i * 32 + d] - rt));
    }
  printf("S=QK^T : max |err| = %g -> %s\n", es, es == 0 ? "MATCH" : "MISMATCH");
  printf("O=PV   : max |err| = %g -> %s\n", eo, eo == 0 ? "MATCH" : "MISMATCH");
  printf("T=P^TQ : max |err| = %g -> %s\n", et, et == 0 ? "MATCH" : "MISMATCH");
  const bool ok = bad == 0 && nz == 0 && es == 0 && eo == 0 && et == 0;
  printf("wmsa_probe %s (%s)\n", ok ? "PASS" : "FAIL", wrap ? "wrap" : "plain");
  return ok ? 0 : 1;
}
