// Dense bf16 GEMMs of the Linear layers on the 5th-gen tensor cores, with the surrounding element-wise passes fused
// into the epilogue (Swin qkv / proj / MLP, encoder FFN; reference sites
// configs/multi/MTL_slvlcls_swin-t-p4-w7_1x1_resisc&dior&potsdam.py:9-25, :34-50 through mmcv FFN / nn.Linear).
//
//   rsc_linear_fwd : Y = act(X W^T + b)            X (M,K), W (N,K) both K-major            [+ H = X W^T + b kept for GELU]
//   rsc_linear_dx  : dX = (dY W) * act'(H)          dY (M,N) K-major, W (N,K) read MN-major  [act' from the saved H / Y]
//   rsc_linear_dw  : dW += dY^T X, db += colsum(dY) both operands MN-major, split over the token dimension, fp32 atomics
//
// One persistent, warp-specialised kernel shape (the canonical sm_100 structure):
//   warp 0  : TMA producer -- A tile (128 rows x 64 k) and B tile (BN rows x 64 k) per stage, SWIZZLE_128B
//   warp 1  : TMEM allocation + MMA issue (tcgen05.mma kind::f16, M = 128, N = BN, 4 x K=16 per stage), accumulator
//             double-buffered in TMEM so that the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2-9: epilogue -- two warps per TMEM lane quadrant, each owning half of the columns: tcgen05.ld -> bias /
//             activation in packed fp32x2 arithmetic -> bf16 -> swizzled staging tile -> TMA store (32-column boxes)
// The GELU is torch's erf form evaluated as x * (0.5 + 0.5 tanh(g(x))) with g an odd polynomial fitted to the normal
// CDF (|GELU error| < 3e-4 absolute incl. tanh.approx, below bf16 resolution): ONE MUFU per element instead of two --
// the fused epilogue is MUFU-bound otherwise -- and its backward is the exact derivative of that function.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace rsc {
namespace gemm {

using namespace tc;

// SMs the persistent kernels size their single wave for.  A data-parallel step overlaps its gradient exchange with
// backward: while the collective's CTAs hold SMs, a persistent CTA that cannot become resident turns one wave into two.
// rsc_set_gemm_sms(sms, below_rows) (called by the step engine when world > 1; RSC_GEMM_SMS / RSC_GEMM_SMS_BELOW override)
// sizes the wave of the GEMMs with fewer than `below_rows` token rows for `sms` SMs.  Measured on 2 x B200: inside the
// box-to-box spread for the det / seg steps, a loss for the tensor-bound cls GEMMs -> a tuning knob, off by default.
static int g_sms_small = 0;
static int64_t g_sms_below = 0;
static int gemm_sms(int64_t rows) {
  static const int env_sms = []() {
    const char *e = getenv("RSC_GEMM_SMS");
    const int n = e ? atoi(e) : 0;
    return n > 0 && n <= kNumSMs ? n : 0;
  }();
  static const int64_t env_below = []() {
    const char *e = getenv("RSC_GEMM_SMS_BELOW");
    return e ? (int64_t)atoll(e) : (int64_t)-1;
  }();
  const int sms = env_sms ? env_sms : g_sms_small;
  const int64_t below = env_below >= 0 ? env_below : g_sms_below;
  return (sms > 0 && (below == 0 || rows < below)) ? sms : kNumSMs;
}

constexpr int BM = 128, BK = 64;
constexpr int EPI_WARPS = 8, THREADS = 64 + EPI_WARPS * 32;
enum { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RELU = 2, EPI_DGELU = 3, EPI_DRELU = 4, EPI_ADD_LN = 5 };

// ---- small PTX helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void st_shared16(uint32_t dst, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared16(uint32_t src) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  uint64_t ra = *reinterpret_cast<uint64_t *>(&a), rb = *reinterpret_cast<uint64_t *>(&b), rc = *reinterpret_cast<uint64_t *>(&c), rd;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  uint64_t ra = *reinterpret_cast<uint64_t *>(&a), rb = *reinterpret_cast<uint64_t *>(&b), rd;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  uint64_t ra = *reinterpret_cast<uint64_t *>(&a), rb = *reinterpret_cast<uint64_t *>(&b), rd;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- GELU (erf form) through one tanh ---------------------------------------------------------------------------
// Phi(x) = 0.5 + 0.5 tanh(g(x)),  g(x) = x (G0 + G1 x^2 + G2 x^4) with x^2 clamped at 49: the logistic fit of
// fused_ew.cu (ACT_GELU_SIG: 1 / (1 + 2^(-x q(x^2)))) rewritten with tanh(z) = 2 / (1 + e^(-2z)) - 1, so
// g = (ln 2 / 2) x q(x^2).
constexpr float G0 = 0.5f * 0.6931471805599453f * 2.3011213394566354f;
constexpr float G1 = 0.5f * 0.6931471805599453f * 0.10677572400272597f;
constexpr float G2 = 0.5f * 0.6931471805599453f * -0.0010142630552444944f;

__device__ __forceinline__ float2 gelu2(float2 x) {
  float2 x2 = mul2(x, x);
  x2.x = fminf(x2.x, 49.f), x2.y = fminf(x2.y, 49.f);
  float2 q = fma2(x2, make_float2(G2, G2), make_float2(G1, G1));
  q = fma2(x2, q, make_float2(G0, G0));
  const float2 g = mul2(x, q);
  const float2 t = make_float2(tanh_approx(g.x), tanh_approx(g.y));
  const float2 hx = mul2(x, make_float2(0.5f, 0.5f));
  return fma2(hx, t, hx);                                     // x (0.5 + 0.5 t)
}
// d/dx [x Phi(x)] = Phi + x Phi',  Phi' = 0.5 (1 - t^2) g',  g' = G0 + 3 G1 x^2 + 5 G2 x^4 (0 beyond the clamp: there
// 1 - t^2 is 0 to fp32 precision anyway)
__device__ __forceinline__ float2 gelu_grad2(float2 x) {
  float2 x2 = mul2(x, x);
  x2.x = fminf(x2.x, 49.f), x2.y = fminf(x2.y, 49.f);
  float2 q = fma2(x2, make_float2(G2, G2), make_float2(G1, G1));
  q = fma2(x2, q, make_float2(G0, G0));
  const float2 g = mul2(x, q);
  float2 dq = fma2(x2, make_float2(5.f * G2, 5.f * G2), make_float2(3.f * G1, 3.f * G1));
  dq = fma2(x2, dq, make_float2(G0, G0));                      // g'(x)
  const float2 t = make_float2(tanh_approx(g.x), tanh_approx(g.y));
  const float2 phi = fma2(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
  const float2 omt2 = fma2(t, make_float2(-t.x, -t.y), make_float2(1.f, 1.f));    // 1 - t^2
  const float2 xd = mul2(mul2(x, make_float2(0.5f, 0.5f)), dq);                   // 0.5 x g'
  return fma2(xd, omt2, phi);
}

// ---- descriptors ---------------------------------------------------------------------------------------------------
// SWIZZLE_128B tiles: rows of 128 bytes, 8-row groups 1024 bytes apart.  K-major: start + 32 bytes per K=16 step.
// MN-major (64-element atoms along the non-contraction dim, `lbo` bytes apart): start + 2048 bytes per K=16 step.
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

struct Params {
  int M, N, K;             // D (M,N) = sum_k A(m,k) B(n,k)
  const float *bias;       // (N) or null
  int n_store;             // outputs: 1 (D) or 2 (D and D2 = the pre-activation H of EPI_BIAS_GELU)
  // EPI_ADD_LN: r = identity + (acc + bias) * scale[row / rows_per_sample];  n = LayerNorm(r) * gamma + beta
  const float *scale, *gamma, *beta;
  float *mean, *rstd;
  int rows_per_sample;
  float eps;
};

// smem carve-up (all tile bases 1024-byte aligned).  Every epilogue warp owns a private staging area for its
// (32 rows x BN/2 columns) part of the output tile: [BN/64 blocks of 32 columns][32 rows][64 B], SWIZZLE_64B, stored
// with its own TMA ops -- the eight warps never meet at a barrier.  One-output epilogues double-buffer it (the store
// of tile i is still reading while tile i+1 is written); the two-tile epilogues (GELU forward: Y and H; activation
// backward: aux in, dX out) keep one buffer per tile.
template <int BN, int STAGES, int EPI>
struct Smem {
  static constexpr bool TWO = EPI == EPI_BIAS_GELU || EPI == EPI_DGELU || EPI == EPI_DRELU || EPI == EPI_ADD_LN;
  static constexpr bool DOUBLE = !TWO && BN < 256;            // BN = 256 spends the second buffer on a third pipeline stage
  static constexpr uint32_t A_BYTES = BM * BK * 2;            // 16 KB
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE = A_BYTES + B_BYTES;
  static constexpr uint32_t WARP_STAGING = 32 * (BN / 2) * 2; // one warp, one buffer
  static constexpr uint32_t STAGING = EPI_WARPS * WARP_STAGING;       // = one bf16 output tile
  static constexpr uint32_t OFF_STAGING = STAGES * STAGE;
  static constexpr uint32_t OFF_STAGING2 = OFF_STAGING + STAGING;     // second buffer / second output / aux input
  static constexpr uint32_t OFF_BIAS = OFF_STAGING2 + ((TWO || DOUBLE) ? STAGING : 0);   // [2 tile parities][BN] floats
  static constexpr uint32_t OFF_BAR = OFF_BIAS + (2 * BN * 4 > 2048 ? 2 * BN * 4 : 2048);   // (EPI_ADD_LN: the 2 KB row-statistics exchange)
  static constexpr uint32_t TOTAL = OFF_BAR + 256;
};

// B_MN: B tile is read MN-major (the weight (N_w, K_w) as the (K_w out, N_w contraction) operand of dX = dY W)
template <int BN, int STAGES, bool B_MN, int EPI>
__global__ void __launch_bounds__(THREADS, 1)
    gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmD2,
                const __grid_constant__ CUtensorMap tmD3, Params p) {
  using S = Smem<BN, STAGES, EPI>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  if (sb & 1023u) __trap();
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + S::OFF_BAR);
  uint64_t *empty = full + STAGES;
  uint64_t *tfull = empty + STAGES;      // [2] accumulator ready
  uint64_t *tempty = tfull + 2;          // [2] accumulator drained
  uint64_t *auxbar = tempty + 2;         // [EPI_WARPS] aux part landed (EPI_DGELU / EPI_DRELU), one per warp
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(auxbar + EPI_WARPS);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512));
  constexpr bool AUX_IN = EPI == EPI_DGELU || EPI == EPI_DRELU || EPI == EPI_ADD_LN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    for (int a = 0; a < 2; ++a) mbar_init(&tfull[a], 1), mbar_init(&tempty[a], EPI_WARPS);
    for (int e = 0; e < EPI_WARPS; ++e) mbar_init(&auxbar[e], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = *tmem_slot;

  const int m_blks = (p.M + BM - 1) / BM, n_blks = (p.N + BN - 1) / BN;
  const int num_tiles = m_blks * n_blks;
  const int k_blks = (p.K + BK - 1) / BK;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_blks) * BM, n0 = (tile % n_blks) * BN;
        for (int kb = 0; kb < k_blks; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          const uint32_t a = sb + s * S::STAGE, b = a + S::A_BYTES;
          mbar_expect_tx(&full[s], S::STAGE);
          tma_load_2d(a, &tmA, &full[s], kb * BK, m0);
          if (!B_MN) {
            tma_load_2d(b, &tmB, &full[s], kb * BK, n0);
          } else {   // 64-wide output atoms, each (64 contraction rows x 128 bytes)
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(b + j * (BK * 128), &tmB, &full[s], n0 + j * 64, kb * BK);
          }
          if (++s == STAGES) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(BM, BN, false, B_MN);
      int s = 0;
      uint32_t ph = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        mbar_wait(&tempty[acc], ((local >> 1) & 1) ^ 1);
        fence_after_sync();
        for (int kb = 0; kb < k_blks; ++kb) {
          mbar_wait(&full[s], ph);
          fence_after_sync();
          const uint32_t a = sb + s * S::STAGE, b = a + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = desc_sw128(a + k * 32, 16);
            const uint64_t db = B_MN ? desc_sw128(b + k * 2048, BK * 128) : desc_sw128(b + k * 32, 16);
            mma_bf16_ss(tm + acc * BN, da, db, idesc, (kb | k) != 0);
          }
          mma_commit(&empty[s]);                       // frees the stage when these MMAs have read it
          if (kb == k_blks - 1) mma_commit(&tfull[acc]);
          if (++s == STAGES) s = 0, ph ^= 1;
        }
      }
    }
  } else {
    // ===== epilogue: warp e owns TMEM lanes 32 (e % 4) .. +31 (its hardware quadrant) and column half e / 4;
    //       it stages and stores its (32 rows x BN/2 columns) part on its own: no barrier between the warps =====
    const int q = warp & 3, half = (warp - 2) >> 2, e = warp - 2;
    constexpr int HC = BN / 2;                                  // columns per warp
    static_assert(HC % 32 == 0, "BN must be a multiple of 64");
    const uint32_t stg0 = sb + S::OFF_STAGING + e * S::WARP_STAGING, stg1 = sb + S::OFF_STAGING2 + e * S::WARP_STAGING;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const int m0 = (tile / n_blks) * BM + q * 32, n0 = (tile % n_blks) * BN + half * HC;   // this warp's part
      // which staging buffer; it is free once the stores that last read it are done reading
      const uint32_t out = (S::DOUBLE && (local & 1)) ? stg1 : stg0;
      if (lane == 0) {
        if (S::DOUBLE) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
      if (AUX_IN) {                                             // the saved H (or Y) values of this part
        if (lane == 0) {
          mbar_expect_tx(&auxbar[e], S::WARP_STAGING);
#pragma unroll
          for (int j = 0; j < HC / 32; ++j)
            tma_load_2d(stg1 + j * 2048, EPI == EPI_ADD_LN ? &tmD3 : &tmD2, &auxbar[e], n0 + j * 32, m0);
        }
      }
      mbar_wait(&tfull[acc], (local >> 1) & 1);
      fence_after_sync();
      // bias slice of this warp's columns.  The four warps of a column half write the SAME values; the slot alternates
      // with the tile parity: once this tile's accumulator is ready, every warp has drained the tile before last (the
      // MMA of this tile waited for that), i.e. nobody still reads this slot
      float *sbias = reinterpret_cast<float *>(smem + S::OFF_BIAS) + (local & 1) * BN + half * HC;
      if (EPI != EPI_ADD_LN) {
        for (int c = lane; c < HC; c += 32) sbias[c] = (p.bias && n0 + c < p.N) ? __ldg(p.bias + n0 + c) : 0.f;
      }
      __syncwarp();
      if (AUX_IN) mbar_wait(&auxbar[e], local & 1);
      const uint32_t taddr = tm + ((uint32_t)(q * 32) << 16) + acc * BN + half * HC;
      if (EPI == EPI_ADD_LN) {
        // ---- residual add + LayerNorm over the full row (the row is split between this warp and its partner in the other
        //      column half: per-half two-pass statistics, combined exactly (Chan) through a 64-thread named barrier) ----
        const int row_g = m0 + lane;
        const float sc = (p.scale && row_g < p.M) ? __ldg(p.scale + row_g / p.rows_per_sample) : 1.0f;
        const int nh = max(0, min(HC, p.N - n0));               // valid columns of this half
        float sum = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < nh; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t soff = (c0 >> 5) * 2048 + lane * 64 + ((j ^ ((lane >> 1) & 3)) * 16);
            const uint4 iq = ld_shared16(stg1 + soff);
            const uint32_t iw[4] = {iq.x, iq.y, iq.z, iq.w};
            uint32_t ow[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int col = n0 + c0 + j * 8 + 2 * k;
              const float2 b2 = p.bias ? __ldg(reinterpret_cast<const float2 *>(p.bias + col)) : make_float2(0.f, 0.f);
              float2 v = add2(make_float2(__uint_as_float(r[j * 8 + 2 * k]), __uint_as_float(r[j * 8 + 2 * k + 1])), b2);
              v = fma2(v, make_float2(sc, sc), unpack_bf16(iw[k]));
              ow[k] = pack_bf16(v.x, v.y);
              // the stored residual is the bf16 value the rest of the network sees: normalise exactly that value
              const float2 vr = unpack_bf16(ow[k]);
              sum += vr.x + vr.y;
            }
            st_shared16(stg1 + soff, make_uint4(ow[0], ow[1], ow[2], ow[3]));
          }
        }
        // the accumulator is drained: hand it back before the statistics
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        const float mean_h = nh > 0 ? sum / nh : 0.f;
        float m2 = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < nh; c0 += 32) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 rq = ld_shared16(stg1 + (c0 >> 5) * 2048 + lane * 64 + ((j ^ ((lane >> 1) & 3)) * 16));
            const uint32_t rw[4] = {rq.x, rq.y, rq.z, rq.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 v = unpack_bf16(rw[k]);
              m2 = fmaf(v.x - mean_h, v.x - mean_h, m2);
              m2 = fmaf(v.y - mean_h, v.y - mean_h, m2);
            }
          }
        }
        float2 *xch = reinterpret_cast<float2 *>(smem + S::OFF_BIAS);     // [quadrant][half][lane] (mean, M2)
        xch[(q * 2 + half) * 32 + lane] = make_float2(mean_h, m2);
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        const float2 other = xch[(q * 2 + (half ^ 1)) * 32 + lane];
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");          // (both have read before the next tile writes)
        const int no = p.N - nh;                                           // the partner's column count
        const float delta = other.x - mean_h;
        const float mean = mean_h + delta * ((float)no / (float)p.N);
        const float var = (m2 + other.y + delta * delta * ((float)nh * (float)no / (float)p.N)) / (float)p.N;
        const float rstd = rsqrtf(var + p.eps);
        if (half == 0 && row_g < p.M) p.mean[row_g] = mean, p.rstd[row_g] = rstd;
#pragma unroll 1
        for (int c0 = 0; c0 < nh; c0 += 32) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t soff = (c0 >> 5) * 2048 + lane * 64 + ((j ^ ((lane >> 1) & 3)) * 16);
            const uint4 rq = ld_shared16(stg1 + soff);
            const uint32_t rw[4] = {rq.x, rq.y, rq.z, rq.w};
            uint32_t ow[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int col = n0 + c0 + j * 8 + 2 * k;
              const float2 g2 = __ldg(reinterpret_cast<const float2 *>(p.gamma + col));
              const float2 be2 = __ldg(reinterpret_cast<const float2 *>(p.beta + col));
              const float2 v = unpack_bf16(rw[k]);
              ow[k] = pack_bf16(fmaf((v.x - mean) * rstd, g2.x, be2.x), fmaf((v.y - mean) * rstd, g2.y, be2.y));
            }
            st_shared16(stg0 + soff, make_uint4(ow[0], ow[1], ow[2], ow[3]));
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (m0 < p.M) {
#pragma unroll
            for (int j = 0; j < HC / 32; ++j) {
              if (n0 + j * 32 < p.N) {
                tma_store_2d(&tmD, stg1 + j * 2048, n0 + j * 32, m0);      // r
                tma_store_2d(&tmD2, stg0 + j * 2048, n0 + j * 32, m0);     // n
              }
            }
          }
          tma_store_commit();
        }
        continue;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < HC; c0 += 32) {
        if (n0 + c0 >= p.N) break;                              // ragged last tile: nothing to store from here on
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {                           // 8-column groups = 16-byte chunks of the bf16 row
          const int col = c0 + j * 8;                           // column inside this warp's part
          // 32-column store block (2 KB: 32 rows x 64 B, SWIZZLE_64B), 16-byte chunk j of this lane's row
          const uint32_t soff = (c0 >> 5) * 2048 + lane * 64 + ((j ^ ((lane >> 1) & 3)) * 16);
          float2 v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 b2 = *reinterpret_cast<const float2 *>(sbias + col + 2 * k);
            v[k] = add2(make_float2(__uint_as_float(r[j * 8 + 2 * k]), __uint_as_float(r[j * 8 + 2 * k + 1])), b2);
          }
          if (EPI == EPI_BIAS_GELU) {
            uint4 hq;
            hq.x = pack_bf16(v[0].x, v[0].y), hq.y = pack_bf16(v[1].x, v[1].y), hq.z = pack_bf16(v[2].x, v[2].y),
            hq.w = pack_bf16(v[3].x, v[3].y);
            st_shared16(stg1 + soff, hq);                       // H (pre-activation, bf16) is kept for the backward
            // the activation sees the ROUNDED pre-activation: forward and backward differentiate the same function
            v[0] = gelu2(unpack_bf16(hq.x)), v[1] = gelu2(unpack_bf16(hq.y)), v[2] = gelu2(unpack_bf16(hq.z)),
            v[3] = gelu2(unpack_bf16(hq.w));
          } else if (EPI == EPI_BIAS_RELU) {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k].x = fmaxf(v[k].x, 0.f), v[k].y = fmaxf(v[k].y, 0.f);
          } else if (EPI == EPI_DGELU || EPI == EPI_DRELU) {
            const uint4 aq = ld_shared16(stg1 + soff);
            const uint32_t aw[4] = {aq.x, aq.y, aq.z, aq.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 a = unpack_bf16(aw[k]);
              if (EPI == EPI_DGELU) v[k] = mul2(v[k], gelu_grad2(a));
              else v[k].x = a.x > 0.f ? v[k].x : 0.f, v[k].y = a.y > 0.f ? v[k].y : 0.f;
            }
          }
          uint4 o;
          o.x = pack_bf16(v[0].x, v[0].y), o.y = pack_bf16(v[1].x, v[1].y), o.z = pack_bf16(v[2].x, v[2].y),
          o.w = pack_bf16(v[3].x, v[3].y);
          st_shared16(out + soff, o);
        }
      }
      // this warp has drained its part of the accumulator; its staged rows go out through TMA
      fence_before_sync();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tempty[acc]);
        if (m0 < p.M) {
#pragma unroll
          for (int j = 0; j < HC / 32; ++j) {
            if (n0 + j * 32 < p.N) {
              tma_store_2d(&tmD, out + j * 2048, n0 + j * 32, m0);
              if (EPI == EPI_BIAS_GELU) tma_store_2d(&tmD2, stg1 + j * 2048, n0 + j * 32, m0);
            }
          }
        }
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_all();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, TMEM_COLS);
}

// ---- host side -------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled get_encode() {
  static EncodeTiled fn = []() -> EncodeTiled {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (EncodeTiled)p;
  }();
  return fn;
}
// row-major (rows, cols) bf16 matrix with leading dimension ld (elements); box (box_cols, box_rows)
static bool map2d(CUtensorMap *m, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                  CUtensorMapSwizzle sw) {
  EncodeTiled enc = get_encode();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int STAGES, bool B_MN, int EPI>
static int launch(const CUtensorMap &tA, const CUtensorMap &tB, const CUtensorMap &tD, const CUtensorMap &tD2,
                  const CUtensorMap &tD3, const Params &p, cudaStream_t st) {
  using S = Smem<BN, STAGES, EPI>;
  static_assert(S::TOTAL <= 227 * 1024, "shared memory");
  auto kern = gemm_kernel<BN, STAGES, B_MN, EPI>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int sms = gemm_sms(p.M);
  const int grid = tiles < sms ? tiles : sms;
  kern<<<grid, THREADS, S::TOTAL, st>>>(tA, tB, tD, tD2, tD3, p);
  return 0;
}

template <bool B_MN, int EPI>
static int dispatch_bn(int BN, const CUtensorMap &tA, const CUtensorMap &tB, const CUtensorMap &tD, const CUtensorMap &tD2,
                       const CUtensorMap &tD3, const Params &p, cudaStream_t st) {
  switch (BN) {
    case 64: return launch<64, 6, B_MN, EPI>(tA, tB, tD, tD2, tD3, p, st);
    case 128: return launch<128, 4, B_MN, EPI>(tA, tB, tD, tD2, tD3, p, st);
    case 192: return launch<192, 3, B_MN, EPI>(tA, tB, tD, tD2, tD3, p, st);
    case 256: return launch<256, Smem<256, 2, EPI>::TWO ? 2 : 3, B_MN, EPI>(tA, tB, tD, tD2, tD3, p, st);
  }
  return -1;
}

// tile width: the candidate (256, 192, 128; 64 for narrow outputs) with the least padded columns, the wider one on a
// tie.  A ragged last tile costs nothing but idle MMA columns: TMA zero-fills the loads and the epilogue skips the
// column blocks beyond N.
static int pick_bn(int N) {
  if (N <= 64) return 64;
  int best = 128, waste = 1 << 30;
  for (int bn : {128, 192, 256}) {
    const int w = (N + bn - 1) / bn * bn - N;
    if (w <= waste) best = bn, waste = w;
  }
  return best;
}

}  // namespace gemm
}  // namespace rsc

using namespace rsc;

// Y (M,N) = act(X (M,K) W(N,K)^T + bias).  act: 0 none, 1 GELU (erf form; H = X W^T + bias is also written to `h`), 2 ReLU.
// bf16 operands / outputs, fp32 accumulate, fp32 bias (may be NULL).  ldx / ldw / ldy in elements.
extern "C" int rsc_linear_fwd(const void *x, const void *w, const float *bias, void *y, void *h, int64_t M, int N, int K,
                              int64_t ldx, int64_t ldw, int64_t ldy, int act, void *stream) {
  RSC_CHECK_ARG(x && w && y && M > 0 && N > 0 && K > 0, "rsc_linear_fwd: null pointer / empty shape");
  RSC_CHECK_ARG(act >= 0 && act <= 2 && (act != 1 || h), "rsc_linear_fwd: act must be 0, 1 (needs h) or 2");
  RSC_CHECK_ARG(N % 8 == 0 && K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0 && ldy % 8 == 0 && M < (1ll << 31),
                "rsc_linear_fwd: N, K and the leading dimensions must be multiples of 8 (16-byte TMA strides)");
  RSC_CHECK_ARG((((uintptr_t)x | (uintptr_t)w | (uintptr_t)y | (uintptr_t)h) & 15) == 0, "rsc_linear_fwd: 16-byte alignment");
  const int BN = gemm::pick_bn(N);
  CUtensorMap tA, tB, tD, tD2;
  bool ok = gemm::map2d(&tA, x, M, K, ldx, gemm::BK, gemm::BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
            gemm::map2d(&tB, w, N, K, ldw, gemm::BK, BN, CU_TENSOR_MAP_SWIZZLE_128B) &&
            gemm::map2d(&tD, y, M, N, ldy, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B) &&
            gemm::map2d(&tD2, act == 1 ? h : y, M, N, ldy, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  RSC_CHECK_ARG(ok, "rsc_linear_fwd: cuTensorMapEncodeTiled failed");
  gemm::Params p{(int)M, N, K, bias, act == 1 ? 2 : 1, nullptr, nullptr, nullptr, nullptr, nullptr, 1, 0.f};
  int rc;
  if (act == 1) rc = gemm::dispatch_bn<false, gemm::EPI_BIAS_GELU>(BN, tA, tB, tD, tD2, tD2, p, (cudaStream_t)stream);
  else if (act == 2) rc = gemm::dispatch_bn<false, gemm::EPI_BIAS_RELU>(BN, tA, tB, tD, tD2, tD2, p, (cudaStream_t)stream);
  else rc = gemm::dispatch_bn<false, gemm::EPI_BIAS>(BN, tA, tB, tD, tD2, tD2, p, (cudaStream_t)stream);
  RSC_CHECK_ARG(rc == 0, "rsc_linear_fwd: no tile shape for N = %d", N);
  RSC_CHECK_LAUNCH("rsc_linear_fwd");
  return RSC_OK;
}

// (r, n) = (identity + (X W^T + bias) * scale[row / rows_per_sample], LayerNorm(r) * gamma + beta): the Linear, the residual
// add (with the DropPath scale of the row's sample) and the NEXT LayerNorm in one kernel; N (= the normalised width) must
// fit one tile: N % 32 == 0, N <= 256.  mean / rstd (M) float are written for the backward (rsc_add_ln_bwd).
extern "C" int rsc_linear_add_ln_fwd(const void *x, const void *w, const float *bias, const void *identity, const float *scale,
                                     const float *gamma, const float *beta, void *r_out, void *n_out, float *mean, float *rstd,
                                     int64_t M, int N, int K, int64_t ldx, int64_t ldw, int64_t rows_per_sample, float eps,
                                     void *stream) {
  RSC_CHECK_ARG(x && w && identity && gamma && beta && r_out && n_out && mean && rstd && M > 0 && K > 0,
                "rsc_linear_add_ln_fwd: null pointer / empty shape");
  RSC_CHECK_ARG(N % 32 == 0 && N >= 32 && N <= 256, "rsc_linear_add_ln_fwd: N = %d must be a multiple of 32 in [32, 256]", N);
  RSC_CHECK_ARG(K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0 && M < (1ll << 31) && rows_per_sample > 0,
                "rsc_linear_add_ln_fwd: K and the leading dimensions must be multiples of 8");
  RSC_CHECK_ARG((((uintptr_t)x | (uintptr_t)w | (uintptr_t)identity | (uintptr_t)r_out | (uintptr_t)n_out) & 15) == 0 &&
                    (((uintptr_t)bias | (uintptr_t)gamma | (uintptr_t)beta) & 7) == 0,
                "rsc_linear_add_ln_fwd: alignment");
  const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : (N <= 192 ? 192 : 256));
  CUtensorMap tA, tB, tR, tN, tI;
  bool ok = gemm::map2d(&tA, x, M, K, ldx, gemm::BK, gemm::BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
            gemm::map2d(&tB, w, N, K, ldw, gemm::BK, BN, CU_TENSOR_MAP_SWIZZLE_128B) &&
            gemm::map2d(&tR, r_out, M, N, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B) &&
            gemm::map2d(&tN, n_out, M, N, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B) &&
            gemm::map2d(&tI, identity, M, N, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  RSC_CHECK_ARG(ok, "rsc_linear_add_ln_fwd: cuTensorMapEncodeTiled failed");
  gemm::Params p{(int)M, N, K, bias, 2, scale, gamma, beta, mean, rstd, (int)rows_per_sample, eps};
  const int rc = gemm::dispatch_bn<false, gemm::EPI_ADD_LN>(BN, tA, tB, tR, tN, tI, p, (cudaStream_t)stream);
  RSC_CHECK_ARG(rc == 0, "rsc_linear_add_ln_fwd: no tile shape for N = %d", N);
  RSC_CHECK_LAUNCH("rsc_linear_add_ln_fwd");
  return RSC_OK;
}

// dX (M,K) = (dY (M,N) W (N,K)) * act'(aux).  act: 0 none, 1 GELU (aux = the saved H), 2 ReLU (aux = the saved Y; only
// its sign is used).  aux has the shape and leading dimension of dX.
extern "C" int rsc_linear_dx(const void *dy, const void *w, const void *aux, void *dx, int64_t M, int N, int K, int64_t lddy,
                             int64_t ldw, int64_t lddx, int act, void *stream) {
  RSC_CHECK_ARG(dy && w && dx && M > 0 && N > 0 && K > 0, "rsc_linear_dx: null pointer / empty shape");
  RSC_CHECK_ARG(act >= 0 && act <= 2 && (act == 0 || aux), "rsc_linear_dx: act must be 0, 1 or 2 (1 / 2 need aux)");
  RSC_CHECK_ARG(N % 8 == 0 && K % 8 == 0 && lddy % 8 == 0 && ldw % 8 == 0 && lddx % 8 == 0 && M < (1ll << 31),
                "rsc_linear_dx: N, K and the leading dimensions must be multiples of 8 (16-byte TMA strides)");
  RSC_CHECK_ARG((((uintptr_t)dy | (uintptr_t)w | (uintptr_t)dx | (uintptr_t)aux) & 15) == 0, "rsc_linear_dx: 16-byte alignment");
  // GEMM view: D (M, K) = sum_n dY(m, n) W(n, k): contraction length N, output width K, B read MN-major
  const int BN = gemm::pick_bn(K);
  CUtensorMap tA, tB, tD, tD2;
  bool ok = gemm::map2d(&tA, dy, M, N, lddy, gemm::BK, gemm::BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
            gemm::map2d(&tB, w, N, K, ldw, 64, gemm::BK, CU_TENSOR_MAP_SWIZZLE_128B) &&
            gemm::map2d(&tD, dx, M, K, lddx, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B) &&
            gemm::map2d(&tD2, act ? aux : dx, M, K, lddx, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  RSC_CHECK_ARG(ok, "rsc_linear_dx: cuTensorMapEncodeTiled failed");
  gemm::Params p{(int)M, K, N, nullptr, 1, nullptr, nullptr, nullptr, nullptr, nullptr, 1, 0.f};
  int rc;
  if (act == 1) rc = gemm::dispatch_bn<true, gemm::EPI_DGELU>(BN, tA, tB, tD, tD2, tD2, p, (cudaStream_t)stream);
  else if (act == 2) rc = gemm::dispatch_bn<true, gemm::EPI_DRELU>(BN, tA, tB, tD, tD2, tD2, p, (cudaStream_t)stream);
  else rc = gemm::dispatch_bn<true, gemm::EPI_BIAS>(BN, tA, tB, tD, tD2, tD2, p, (cudaStream_t)stream);
  RSC_CHECK_ARG(rc == 0, "rsc_linear_dx: no tile shape for K = %d", K);
  RSC_CHECK_LAUNCH("rsc_linear_dx");
  return RSC_OK;
}

// =====================================================================================================================
// dW (N,K) += dY (M,N)^T X (M,K),  db (N) += colsum(dY): both operands MN-major (the contraction runs over the token rows),
// the token dimension split across CTAs, fp32 results added with vector reductions (red.global.add.v4.f32).
// The bias gradient costs nothing: a constant "ones" atom behind the B tile makes accumulator column BN the column sum
// of the dY tile (the tensor core adds it up).
// =====================================================================================================================
namespace rsc {
namespace gemm {

constexpr int DW_THREADS = 64 + 4 * 32;

template <int BN, int STAGES>
struct SmemDW {
  static constexpr uint32_t A_BYTES = BM * BK * 2;                 // 2 atoms of (64 tokens x 64 n)
  static constexpr uint32_t X_BYTES = BN * BK * 2;                 // BN / 64 atoms of (64 tokens x 64 k)
  static constexpr uint32_t STAGE = A_BYTES + X_BYTES + BK * 128;  // + the ones atom
  static constexpr uint32_t OFF_BAR = STAGES * STAGE;
  static constexpr uint32_t TOTAL = OFF_BAR + 256;
};

__device__ __forceinline__ void red_add_v4(float *dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// grid = n_tiles * k_tiles * chunks; CTA = (output tile (128 n x BN k), chunk of the token range)
template <int BN, int STAGES>
__global__ void __launch_bounds__(DW_THREADS, 1)
    gemm_dw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float *__restrict__ dw,
                   float *__restrict__ db, int M, int N, int K, int64_t lddw, int k_tiles, int chunks) {
  using S = SmemDW<BN, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  if (sb & 1023u) __trap();
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + S::OFF_BAR);
  uint64_t *empty = full + STAGES;
  uint64_t *done = empty + STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  constexpr int NMMA = BN + 16;                                    // + the column-sum column (and 15 zero columns)
  constexpr uint32_t TMEM_COLS = NMMA <= 128 ? 128 : 256;

  const int chunk = blockIdx.x % chunks, tile = blockIdx.x / chunks;
  const int n0 = (tile / k_tiles) * BM, k0 = (tile % k_tiles) * BN;
  const int blocks_total = (M + BK - 1) / BK;
  const int per = (blocks_total + chunks - 1) / chunks;
  const int kb0 = chunk * per, kb1 = min(blocks_total, kb0 + per);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    mbar_init(done, 1);
    mbar_fence_init();
  }
  // the constant ones atom of every stage: element 0 of each token row = 1.0 (SWIZZLE_128B: chunk 0 of row r sits at
  // chunk position r & 7), the rest 0
  for (int i = threadIdx.x; i < STAGES * (BK * 128 / 16); i += DW_THREADS) {
    const int s = i / (BK * 128 / 16), r = (i % (BK * 128 / 16)) / 8, c = i % 8;
    const uint4 v = make_uint4(c == (r & 7) ? 0x00003F80u : 0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4 *>(smem + s * S::STAGE + S::A_BYTES + S::X_BYTES + r * 128 + c * 16) = v;
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        const uint32_t a = sb + s * S::STAGE, b = a + S::A_BYTES;
        mbar_expect_tx(&full[s], S::A_BYTES + S::X_BYTES);
#pragma unroll
        for (int j = 0; j < BM / 64; ++j) tma_load_2d(a + j * (BK * 128), &tmA, &full[s], n0 + j * 64, kb * BK);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_2d(b + j * (BK * 128), &tmB, &full[s], k0 + j * 64, kb * BK);
        if (++s == STAGES) s = 0, ph ^= 1;
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(BM, NMMA, true, true);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[s], ph);
        fence_after_sync();
        const uint32_t a = sb + s * S::STAGE, b = a + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          mma_bf16_ss(tm, desc_sw128(a + k * 2048, BK * 128), desc_sw128(b + k * 2048, BK * 128), idesc, (kb > kb0) || k > 0);
        mma_commit(&empty[s]);
        if (++s == STAGES) s = 0, ph ^= 1;
      }
      mma_commit(done);
    }
  } else if (kb1 > kb0) {
    // ===== epilogue: thread = output row n (TMEM lane), BN (+1) fp32 columns added to dW (and db) =====
    const int q = warp & 3, row = q * 32 + lane;
    mbar_wait(done, 0);
    fence_after_sync();
    const int n = n0 + row;
    const uint32_t taddr = tm + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + c0, r);
      tmem_ld_wait();
      if (n < N) {
        float *dst = dw + (int64_t)n * lddw + k0 + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (k0 + c0 + j < K)
            red_add_v4(dst + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    }
    if (db != nullptr && k0 == 0) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr + BN)
          : "memory");
      tmem_ld_wait();
      if (n < N) atomicAdd(db + n, __uint_as_float(r[0]));
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, TMEM_COLS);
}

template <int BN, int STAGES>
static int launch_dw(const CUtensorMap &tA, const CUtensorMap &tB, float *dw, float *db, int M, int N, int K, int64_t lddw,
                     cudaStream_t st) {
  using S = SmemDW<BN, STAGES>;
  static_assert(S::TOTAL <= 227 * 1024, "shared memory");
  auto kern = gemm_dw_kernel<BN, STAGES>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
  const int n_tiles = (N + BM - 1) / BM, k_tiles = (K + BN - 1) / BN, tiles = n_tiles * k_tiles;
  const int blocks_total = (M + BK - 1) / BK;
  const int sms = gemm_sms(M);
  int chunks = tiles >= sms ? 1 : sms / tiles;                    // one wave of CTAs
  if (chunks > blocks_total) chunks = blocks_total;
  kern<<<tiles * chunks, DW_THREADS, S::TOTAL, st>>>(tA, tB, dw, db, M, N, K, lddw, k_tiles, chunks);
  return 0;
}

}  // namespace gemm
}  // namespace rsc

extern "C" int rsc_linear_dw(const void *dy, const void *x, float *dw, float *db, int64_t M, int N, int K, int64_t lddy,
                             int64_t ldx, int64_t lddw, void *stream) {
  RSC_CHECK_ARG(dy && x && dw && M > 0 && N > 0 && K > 0, "rsc_linear_dw: null pointer / empty shape");
  RSC_CHECK_ARG(N % 8 == 0 && K % 8 == 0 && lddy % 8 == 0 && ldx % 8 == 0 && lddw % 4 == 0 && M < (1ll << 31),
                "rsc_linear_dw: N, K, lddy, ldx must be multiples of 8 and lddw of 4");
  RSC_CHECK_ARG((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dw) & 15) == 0, "rsc_linear_dw: 16-byte alignment");
  CUtensorMap tA, tB;
  bool ok = gemm::map2d(&tA, dy, M, N, lddy, 64, gemm::BK, CU_TENSOR_MAP_SWIZZLE_128B) &&
            gemm::map2d(&tB, x, M, K, ldx, 64, gemm::BK, CU_TENSOR_MAP_SWIZZLE_128B);
  RSC_CHECK_ARG(ok, "rsc_linear_dw: cuTensorMapEncodeTiled failed");
  const int BN = K <= 64 ? 64 : (K <= 128 || (K % 192 != 0 && K % 128 == 0) ? 128 : 192);
  int rc;
  if (BN == 64) rc = gemm::launch_dw<64, 6>(tA, tB, dw, db, (int)M, N, K, lddw, (cudaStream_t)stream);
  else if (BN == 128) rc = gemm::launch_dw<128, 5>(tA, tB, dw, db, (int)M, N, K, lddw, (cudaStream_t)stream);
  else rc = gemm::launch_dw<192, 4>(tA, tB, dw, db, (int)M, N, K, lddw, (cudaStream_t)stream);
  (void)rc;
  RSC_CHECK_LAUNCH("rsc_linear_dw");
  return RSC_OK;
}

extern "C" int rsc_set_gemm_sms(int sms, int64_t below_rows) {
  RSC_CHECK_ARG(sms >= 0 && sms <= kNumSMs && below_rows >= 0, "rsc_set_gemm_sms: sms must be in [0, %d] (0 = all)", kNumSMs);
  rsc::gemm::g_sms_small = sms;
  rsc::gemm::g_sms_below = below_rows;
  return RSC_OK;
}
