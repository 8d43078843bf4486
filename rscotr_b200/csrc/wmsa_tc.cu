// Fused (shifted-)window attention forward on the 5th-gen tensor cores (bf16).
//
//   S = Q K^T  and  O = P V  are tcgen05.mma tiles (kind::f16, bf16 x bf16 -> fp32 in TMEM);
//   two (window, head) units are stacked along M (M = 128 = 2 x 64 rows, 49 valid each):
//     S[128 x 128] = [Q_a;Q_b] x [K_a;K_b]^T          (only the two diagonal 64x64 blocks are used)
//     O_a[128 x 32] = P x V_a ,  O_b[128 x 32] = P x V_b   (rows 0-63 of O_a, rows 64-127 of O_b are used)
//   The softmax (scale, relative-position bias, shift mask -100, exp, row sum) runs on the
//   SIMT lanes between the two MMAs: thread r owns row r of S (tcgen05.ld 32x32b).
//   qkv is gathered ONCE from its natural (B,H,W,3C) layout through the padded / cyclically
//   shifted window coordinates with 16-byte cp.async into the no-swizzle core-matrix smem
//   layout (tools/tc_probe.cu pins these descriptors); padded tokens are synthesised from the
//   qkv bias.  The kernel is HBM-bound (24.5 flop/B): the tensor cores only keep the
//   math off the critical path; 4 CTAs/SM (128 TMEM columns each) overlap load / MMA / softmax.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace rsc {
namespace wtc {

using namespace tc;

constexpr int WS = 7, NT = 49, HD = 32, TBL = 169;
constexpr int THREADS = 128;
constexpr int SM_Q = 0, SM_K = 8192, SM_V = 16384, SM_P = 24576;  // byte offsets
constexpr int SM_TBL = SM_P + 16384;                               // 169 floats
constexpr int SM_TOTAL = 50 * 1024;                                // sized so that exactly 4 CTAs fit per SM
constexpr int TMEM_COLS = 128;

// 16-byte chunk c (8 bf16) of row r of a token-major [rows][32] tile
__device__ __forceinline__ int tile_off(int r, int c) { return (r >> 3) * 512 + c * 128 + (r & 7) * 16; }
// P tile (128 rows x 64 keys): chunk kc of row r
__device__ __forceinline__ int p_off(int r, int kc) { return kc * 2048 + (r >> 3) * 128 + (r & 7) * 16; }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}

constexpr float LOG2E = 1.4426950408889634f;

// softmax numerator row in the exp2 domain: sv[j] = S[j]*scale*log2e + bias2[j] (+ mask), returns the row max
template <bool MASK>
__device__ __forceinline__ float score_row(const uint32_t (&s0)[32], const uint32_t (&s1)[32], const float *tb,
                                           float scale2, uint32_t rowbits, uint32_t colbits, float (&sv)[NT]) {
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int jr = j / WS, jc = j % WS;
    float s = fmaf(__uint_as_float(j < 32 ? s0[j] : s1[j - 32]), scale2, tb[-(jr * (2 * WS - 1) + jc)]);
    if (MASK) {
      if (((rowbits >> jr) | (colbits >> jc)) & 1u) s += -100.0f * LOG2E;
    }
    sv[j] = s;
    m = fmaxf(m, s);
  }
  return m;
}

// region bits of the shift mask for query slot (ri, ci) of window (wh, ww): bit j of rowbits / colbits is set
// when key row / column j lies in a different region (only the last window row / column has two regions)
__device__ __forceinline__ void mask_bits(const WinGeom &g, int wh, int ww, int ri, int ci, uint32_t &rowbits,
                                          uint32_t &colbits) {
  rowbits = colbits = 0;
  const bool lr = wh == g.nWh - 1, lc = ww == g.nWw - 1;
  const int rh_i = lr ? (ri < WS - g.shift ? 1 : 2) : 0, rw_i = lc ? (ci < WS - g.shift ? 1 : 2) : 0;
#pragma unroll
  for (int j = 0; j < WS; ++j) {
    const int rj = j < WS - g.shift ? 1 : 2;
    rowbits |= (uint32_t)((lr ? rj : 0) != rh_i) << j;
    colbits |= (uint32_t)((lc ? rj : 0) != rw_i) << j;
  }
}

__global__ void __launch_bounds__(THREADS, 4)
    wmsa_fwd_tc_kernel(const __nv_bfloat16 *__restrict__ qkv, const float *__restrict__ qkv_bias,
                       const float *__restrict__ table, __nv_bfloat16 *__restrict__ out, WinGeom g, int C, int heads,
                       float scale, int num_windows, int num_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float *tbl = reinterpret_cast<float *>(smem + SM_TBL);
  const int head = blockIdx.x % heads;   // gridDim.x is a multiple of heads: the head is fixed per CTA

  if (warp == 0) tmem_alloc(&tmem_base_s, TMEM_COLS);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  // zero the operand tiles once: rows 49..63 of every unit stay zero for the whole kernel
  for (int i = tid; i < SM_P / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int k = tid; k < TBL; k += THREADS) tbl[k] = __ldg(table + k * heads + head) * LOG2E;   // exp2 domain
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);
  const uint32_t idesc_o = make_idesc_bf16(128, 32, false, true);
  uint32_t phase = 0;
  const float scale2 = scale * LOG2E;

  const int unit = tid >> 6;   // which of the two stacked units this thread's row belongs to
  const int i = tid & 63;      // row inside the unit (query token), valid if < 49
  const int ri = i / WS, ci = i % WS;
  const float *tb = tbl + (ri + WS - 1) * (2 * WS - 1) + (ci + WS - 1);
  const uint32_t my_row = smem_base + tile_off(tid, 0);   // this thread's token row in the q|k|v tiles

  for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
    const int pair = item / heads;
    // ---------------- phase 1: every thread gathers the q|k|v head slices of ITS token ----------------
    const int win = 2 * pair + unit;
    const bool row_ok = win < num_windows && i < NT;
    int b = 0, wh = 0, ww = 0, h = 0, w = 0;
    bool tok_ok = false;
    if (row_ok) {
      ww = win % g.nWw;
      wh = (win / g.nWw) % g.nWh;
      b = win / (g.nWw * g.nWh);
      tok_ok = g.source(wh, ww, ri, ci, h, w);
      if (tok_ok) {
        const __nv_bfloat16 *src = qkv + (((int64_t)b * g.H + h) * g.W + w) * (3 * C) + head * HD;
#pragma unroll
        for (int part = 0; part < 3; ++part)
#pragma unroll
          for (int c = 0; c < 4; ++c) cp_async16(my_row + part * 8192 + c * 128, src + part * C + c * 8);
      } else {   // zero-padded token: its qkv row is the qkv bias
#pragma unroll
        for (int part = 0; part < 3; ++part)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (qkv_bias) {
              const int col = part * C + head * HD + c * 8;
              const float4 f0 = __ldg(reinterpret_cast<const float4 *>(qkv_bias + col));
              const float4 f1 = __ldg(reinterpret_cast<const float4 *>(qkv_bias + col + 4));
              v = make_uint4(pack_bf16(f0.x, f0.y), pack_bf16(f0.z, f0.w), pack_bf16(f1.x, f1.y), pack_bf16(f1.z, f1.w));
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + part * 8192 + c * 128), "r"(v.x),
                         "r"(v.y), "r"(v.z), "r"(v.w)
                         : "memory");
          }
      }
    }
    cp_async_wait_all();
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---------------- phase 2: S = Q K^T ----------------
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int k = 0; k < 2; ++k)
        mma_bf16_ss(tm, make_smem_desc(smem_base + SM_Q + k * 256, 128, 512),
                    make_smem_desc(smem_base + SM_K + k * 256, 128, 512), idesc_s, k > 0);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    // ---------------- phase 3: softmax on this thread's row ----------------
    float inv_l = 0.f;
    {
      uint32_t s0[32], s1[32];
      const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + unit * 64;
      tmem_ld32(taddr, s0);
      tmem_ld32(taddr + 32, s1);
      tmem_ld_wait();
      uint32_t pk[32];   // 64 bf16 probabilities, packed
      if (row_ok) {
        float sv[NT];
        float m;
        // the shift mask only exists in the last window row / column (warp-uniform branch: a unit = 2 warps)
        if (g.shift > 0 && (wh == g.nWh - 1 || ww == g.nWw - 1)) {
          uint32_t rowbits, colbits;
          mask_bits(g, wh, ww, ri, ci, rowbits, colbits);
          m = score_row<true>(s0, s1, tb, scale2, rowbits, colbits, sv);
        } else {
          m = score_row<false>(s0, s1, tb, scale2, 0u, 0u, sv);
        }
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float p = exp2f(sv[j] - m);
          l += p;
          sv[j] = p;
        }
        inv_l = 1.0f / l;
#pragma unroll
        for (int j = 0; j < 24; ++j) pk[j] = pack_bf16(sv[2 * j], sv[2 * j + 1]);
        pk[24] = pack_bf16(sv[48], 0.f);
#pragma unroll
        for (int j = 25; j < 32; ++j) pk[j] = 0u;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) pk[j] = 0u;
      }
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        const uint32_t dst = smem_base + SM_P + p_off(tid, kc);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * kc]), "r"(pk[4 * kc + 1]),
                     "r"(pk[4 * kc + 2]), "r"(pk[4 * kc + 3])
                     : "memory");
      }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---------------- phase 4: O_a = P V_a, O_b = P V_b (overwrites S columns 0..63) ----------------
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_bf16_ss(tm + u * 32, make_smem_desc(smem_base + SM_P + k * 4096, 2048, 128),
                      make_smem_desc(smem_base + SM_V + u * 4096 + k * 1024, 512, 128), idesc_o, k > 0);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    // ---------------- phase 5: normalise and store this thread's output row ----------------
    {
      uint32_t o[32];
      tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + unit * 32, o);
      tmem_ld_wait();
      if (tok_ok) {
        __nv_bfloat16 *dst = out + (((int64_t)b * g.H + h) * g.W + w) * C + head * HD;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(o[8 * c + 0]) * inv_l, __uint_as_float(o[8 * c + 1]) * inv_l);
          v.y = pack_bf16(__uint_as_float(o[8 * c + 2]) * inv_l, __uint_as_float(o[8 * c + 3]) * inv_l);
          v.z = pack_bf16(__uint_as_float(o[8 * c + 4]) * inv_l, __uint_as_float(o[8 * c + 5]) * inv_l);
          v.w = pack_bf16(__uint_as_float(o[8 * c + 6]) * inv_l, __uint_as_float(o[8 * c + 7]) * inv_l);
          *reinterpret_cast<uint4 *>(dst + 8 * c) = v;
        }
      }
    }
    fence_before_sync();
    __syncthreads();  // smem tiles and TMEM are free for the next item
    fence_after_sync();
  }
  if (warp == 0) tmem_dealloc(tm, TMEM_COLS);
}

}  // namespace wtc
}  // namespace rsc

using namespace rsc;

extern "C" int rsc_wmsa_fwd_simt(const void *qkv, const float *qkv_bias, const float *bias_table, void *out, int B,
                                 int H, int W, int C, int heads, int ws, int shift, float scale, int dtype,
                                 void *stream);

extern "C" int rsc_wmsa_fwd(const void *qkv, const float *qkv_bias, const float *bias_table, void *out, int B, int H,
                            int W, int C, int heads, int ws, int shift, float scale, int dtype, void *stream) {
  static const bool force_simt = getenv("RSC_WMSA_SIMT") != nullptr;
  const bool tc_ok = dtype == RSC_BF16 && ws == 7 && (shift == 0 || shift == 3) && heads > 0 && C == heads * 32 &&
                     B > 0 && H > 0 && W > 0 && qkv && bias_table && out && heads <= 4 * kNumSMs;
  if (!tc_ok || force_simt)  // fp32 (exact-arithmetic parity path) and argument errors go through the SIMT entry
    return rsc_wmsa_fwd_simt(qkv, qkv_bias, bias_table, out, B, H, W, C, heads, ws, shift, scale, dtype, stream);
  WinGeom g(B, H, W, ws, shift);
  const int num_windows = B * g.nWh * g.nWw;
  const int num_items = ((num_windows + 1) / 2) * heads;
  auto kern = wtc::wmsa_fwd_tc_kernel;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, wtc::SM_TOTAL);
  int grid = (kNumSMs * 4) / heads * heads;    // a multiple of heads: every CTA keeps one head (bias table loaded once)
  if (grid > num_items) grid = num_items;      // num_items is a multiple of heads
  kern<<<grid, wtc::THREADS, wtc::SM_TOTAL, (cudaStream_t)stream>>>(
      (const __nv_bfloat16 *)qkv, qkv_bias, bias_table, (__nv_bfloat16 *)out, g, C, heads, scale, num_windows,
      num_items);
  RSC_CHECK_LAUNCH("rsc_wmsa_fwd");
  return RSC_OK;
}

// =====================================================================================
// Backward on the tensor cores (bf16).  Per pair of (window, head) units, five MMA groups:
//   S  = [Qa;Qb][Ka;Kb]^T            dP = [dOa;dOb][Va;Vb]^T          (recompute; TMEM 2 x 128 columns)
//   -- SIMT: P = softmax(scale*S + bias + mask), D = sum_j P*dP, dS = P*(dP - D), both written to smem
//      as block-diagonal [128 x 128] bf16 tiles (off-diagonal blocks stay zero) --
//   dV = P^T dO ,  dK = dS^T Q        (A = the same smem tile read MN-major, K = 128 query rows)
//   dQ = dS K                          (A = dS K-major, K = 128 key columns)
// d(bias table) is accumulated per thread in registers across the persistent loop (fixed head per
// CTA) and folded once at the end; gradients of padded rows go to the qkv-bias gradient.
// =====================================================================================
namespace rsc {
namespace wtc {

constexpr int B_Q = 0, B_K = 8192, B_V = 16384, B_DO = 24576, B_P = 32768, B_DS = 65536;
constexpr int B_TBL = 98304;             // 169 floats (+pad)
constexpr int B_PAD = B_TBL + 704;       // 3 x 32 floats: qkv-bias gradient of padded rows
constexpr int B_TOTAL = B_PAD + 384;     // ~97.4 KB -> 2 CTAs / SM
constexpr int B_TMEM = 256;

// block-diagonal [128 rows][128 cols] bf16 tile: 16-byte chunk kc (8 columns) of row r
__device__ __forceinline__ int bd_off(int r, int kc) { return kc * 2048 + (r >> 3) * 128 + (r & 7) * 16; }

__global__ void __launch_bounds__(THREADS, 2)
    wmsa_bwd_tc_kernel(const __nv_bfloat16 *__restrict__ qkv, const float *__restrict__ qkv_bias,
                       const float *__restrict__ table, const __nv_bfloat16 *__restrict__ dout,
                       __nv_bfloat16 *__restrict__ dqkv, float *__restrict__ dtable, float *__restrict__ dqkv_bias,
                       WinGeom g, int C, int heads, float scale, int num_windows, int num_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float *tbl = reinterpret_cast<float *>(smem + B_TBL);
  float *padacc = reinterpret_cast<float *>(smem + B_PAD);
  const int head = blockIdx.x % heads;   // gridDim.x is a multiple of heads

  if (warp == 0) tmem_alloc(&tmem_base_s, B_TMEM);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  for (int k = tid; k < B_TBL / 16; k += THREADS) reinterpret_cast<uint4 *>(smem)[k] = make_uint4(0, 0, 0, 0);
  for (int k = tid; k < TBL; k += THREADS) tbl[k] = __ldg(table + k * heads + head) * LOG2E;   // exp2 domain
  for (int k = tid; k < 96; k += THREADS) padacc[k] = 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  const uint32_t sb = smem_u32(smem);
  const uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);
  const uint32_t idesc_t = make_idesc_bf16(128, 32, true, true);    // A^T (MN-major) x MN-major B
  const uint32_t idesc_q = make_idesc_bf16(128, 32, false, true);   // K-major A x MN-major B
  uint32_t phase = 0;

  const int unit = tid >> 6, i = tid & 63;
  const int ri = i / WS, ci = i % WS;
  const float *tb = tbl + (ri + WS - 1) * (2 * WS - 1) + (ci + WS - 1);
  const uint32_t my_row = sb + tile_off(tid, 0);   // this thread's token row in the q|k|v|dO tiles
  const float scale2 = scale * LOG2E;
  float dbacc[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) dbacc[j] = 0.f;

  for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
    const int pair = item / heads;
    // ---------------- every thread gathers the q|k|v|dO head slices of ITS token ----------------
    const int win = 2 * pair + unit;
    const bool unit_ok = win < num_windows;
    const bool row_ok = unit_ok && i < NT;
    int b = 0, wh = 0, ww = 0, h = 0, w = 0;
    bool tok_ok = false;
    if (unit_ok) {
      ww = win % g.nWw;
      wh = (win / g.nWw) % g.nWh;
      b = win / (g.nWw * g.nWh);
    }
    if (row_ok) {
      tok_ok = g.source(wh, ww, ri, ci, h, w);
      if (tok_ok) {
        const int64_t tok = ((int64_t)b * g.H + h) * g.W + w;
        const __nv_bfloat16 *src = qkv + tok * (3 * C) + head * HD;
        const __nv_bfloat16 *dsrc = dout + tok * C + head * HD;
#pragma unroll
        for (int part = 0; part < 3; ++part)
#pragma unroll
          for (int c = 0; c < 4; ++c) cp_async16(my_row + part * 8192 + c * 128, src + part * C + c * 8);
#pragma unroll
        for (int c = 0; c < 4; ++c) cp_async16(my_row + 3 * 8192 + c * 128, dsrc + c * 8);
      } else {   // zero-padded token: qkv row = qkv bias, its output row is cropped (dO = 0)
#pragma unroll
        for (int part = 0; part < 4; ++part)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (part < 3 && qkv_bias) {
              const int col = part * C + head * HD + c * 8;
              const float4 f0 = __ldg(reinterpret_cast<const float4 *>(qkv_bias + col));
              const float4 f1 = __ldg(reinterpret_cast<const float4 *>(qkv_bias + col + 4));
              v = make_uint4(pack_bf16(f0.x, f0.y), pack_bf16(f0.z, f0.w), pack_bf16(f1.x, f1.y), pack_bf16(f1.z, f1.w));
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + part * 8192 + c * 128), "r"(v.x),
                         "r"(v.y), "r"(v.z), "r"(v.w)
                         : "memory");
          }
      }
    }
    cp_async_wait_all();
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---------------- S = Q K^T (cols 0..127), dP = dO V^T (cols 128..255) ----------------
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int k = 0; k < 2; ++k)
        mma_bf16_ss(tm, make_smem_desc(sb + B_Q + k * 256, 128, 512), make_smem_desc(sb + B_K + k * 256, 128, 512),
                    idesc_s, k > 0);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        mma_bf16_ss(tm + 128, make_smem_desc(sb + B_DO + k * 256, 128, 512),
                    make_smem_desc(sb + B_V + k * 256, 128, 512), idesc_s, k > 0);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    // ---------------- softmax backward on this thread's row ----------------
    {
      uint32_t s0[32], s1[32];
      const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + unit * 64;
      tmem_ld32(taddr, s0);
      tmem_ld32(taddr + 32, s1);
      tmem_ld_wait();
      uint32_t pp[32], pd[32];   // packed bf16 P row and dS row (64 columns each)
      // The TMEM loads are warp-aligned instructions and stay outside the branches; the row arithmetic runs
      // only on real query rows (the others store zeros), with warp-uniform fast paths: the shift mask exists
      // only in the last window row / column, padded keys only where the map is not a multiple of 7.
      float p[NT];
      float wp = 0.f, wds = 0.f;
      uint32_t rowpad = 0, colpad = 0;
      if (row_ok) {
        float m;
        if (g.shift > 0 && (wh == g.nWh - 1 || ww == g.nWw - 1)) {
          uint32_t rowbits, colbits;
          mask_bits(g, wh, ww, ri, ci, rowbits, colbits);
          m = score_row<true>(s0, s1, tb, scale2, rowbits, colbits, p);
        } else {
          m = score_row<false>(s0, s1, tb, scale2, 0u, 0u, p);
        }
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          p[j] = exp2f(p[j] - m);
          l += p[j];
        }
        const float inv_l = 1.0f / l;
#pragma unroll
        for (int j = 0; j < NT; ++j) p[j] *= inv_l;
        // key tokens of this window that are zero-padding (their dk/dv flow to the qkv-bias gradient)
        if (g.Hp != g.H || g.Wp != g.W) {
#pragma unroll
          for (int j = 0; j < WS; ++j) {
            int hh = wh * WS + j + g.shift, wc = ww * WS + j + g.shift;
            if (hh >= g.Hp) hh -= g.Hp;
            if (wc >= g.Wp) wc -= g.Wp;
            rowpad |= (uint32_t)(hh >= g.H) << j;
            colpad |= (uint32_t)(wc >= g.W) << j;
          }
        }
      }
      tmem_ld32(taddr + 128, s0);   // dP row (reuses the S registers)
      tmem_ld32(taddr + 160, s1);
      tmem_ld_wait();
      if (row_ok) {
        // Column 49 of the P / dS tiles (a zero-padding column of the MMA) carries the row's sum over the
        // PADDED keys: row 49 of dV = P^T dO / dK = dS^T Q then IS the padded-row gradient sum, for free.
        float D = 0.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) D = fmaf(p[j], __uint_as_float(j < 32 ? s0[j] : s1[j - 32]), D);
        if (rowpad | colpad) {
#pragma unroll
          for (int j = 0; j < NT; ++j)
            if (((rowpad >> (j / WS)) | (colpad >> (j % WS))) & 1u) wp += p[j];
        }
#pragma unroll
        for (int j = 0; j < 24; ++j) pp[j] = pack_bf16(p[2 * j], p[2 * j + 1]);
        pp[24] = pack_bf16(p[48], wp);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float ds = p[j] * (__uint_as_float(j < 32 ? s0[j] : s1[j - 32]) - D);
          dbacc[j] += ds;
          p[j] = ds;
        }
        if (rowpad | colpad) {
#pragma unroll
          for (int j = 0; j < NT; ++j)
            if (((rowpad >> (j / WS)) | (colpad >> (j % WS))) & 1u) wds += p[j];
        }
#pragma unroll
        for (int j = 0; j < 24; ++j) pd[j] = pack_bf16(p[2 * j], p[2 * j + 1]);
        pd[24] = pack_bf16(p[48], wds);
#pragma unroll
        for (int j = 25; j < 32; ++j) pp[j] = 0u, pd[j] = 0u;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) pp[j] = 0u, pd[j] = 0u;
      }
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        const uint32_t o = bd_off(tid, unit * 8 + kc);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + B_P + o), "r"(pp[4 * kc]),
                     "r"(pp[4 * kc + 1]), "r"(pp[4 * kc + 2]), "r"(pp[4 * kc + 3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + B_DS + o), "r"(pd[4 * kc]),
                     "r"(pd[4 * kc + 1]), "r"(pd[4 * kc + 2]), "r"(pd[4 * kc + 3])
                     : "memory");
      }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---------------- dV = P^T dO (cols 0..31), dK = dS^T Q (32..63), dQ = dS K (64..95) ----------------
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int k = 0; k < 8; ++k)   // K = 128 query rows, 16 per step
        mma_bf16_ss(tm, make_smem_desc(sb + B_P + k * 256, 128, 2048), make_smem_desc(sb + B_DO + k * 1024, 512, 128),
                    idesc_t, k > 0);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        mma_bf16_ss(tm + 32, make_smem_desc(sb + B_DS + k * 256, 128, 2048),
                    make_smem_desc(sb + B_Q + k * 1024, 512, 128), idesc_t, k > 0);
#pragma unroll
      for (int k = 0; k < 8; ++k)   // K = 128 key columns, 16 per step
        mma_bf16_ss(tm + 64, make_smem_desc(sb + B_DS + k * 4096, 2048, 128),
                    make_smem_desc(sb + B_K + k * 1024, 512, 128), idesc_q, k > 0);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    // ---------------- store dq | dk | dv of this thread's token ----------------
    {
      uint32_t o[32];
      const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
      __nv_bfloat16 *dst = dqkv + (((int64_t)b * g.H + h) * g.W + w) * (3 * C) + head * HD;
#pragma unroll
      for (int part = 0; part < 3; ++part) {   // TMEM columns: dV 0, dK 32, dQ 64 -> dqkv parts 2, 1, 0
        tmem_ld32(taddr + part * 32, o);
        tmem_ld_wait();
        const float sc = part == 0 ? 1.0f : scale;
        const int qpart = 2 - part;
        if (tok_ok) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(o[8 * c + 0]) * sc, __uint_as_float(o[8 * c + 1]) * sc);
            v.y = pack_bf16(__uint_as_float(o[8 * c + 2]) * sc, __uint_as_float(o[8 * c + 3]) * sc);
            v.z = pack_bf16(__uint_as_float(o[8 * c + 4]) * sc, __uint_as_float(o[8 * c + 5]) * sc);
            v.w = pack_bf16(__uint_as_float(o[8 * c + 6]) * sc, __uint_as_float(o[8 * c + 7]) * sc);
            *reinterpret_cast<uint4 *>(dst + qpart * C + 8 * c) = v;
          }
        } else if (unit_ok && i == NT && part < 2 && dqkv_bias) {
          // row 49 = sum over this window's padded keys (zero when the window has none); dq of padded
          // queries is identically zero (their output rows are cropped)
#pragma unroll
          for (int d = 0; d < 32; ++d) {
            const float v = __uint_as_float(o[d]) * sc;
            if (v != 0.f) atomicAdd(padacc + qpart * 32 + d, v);
          }
        }
      }
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  // ---------------- fold the bias-table gradient: registers -> smem table -> global ----------------
  float *fold = reinterpret_cast<float *>(smem + B_P);   // P tile is free now
  for (int k = tid; k < TBL; k += THREADS) fold[k] = 0.f;
  __syncthreads();
  if (i < NT) {
#pragma unroll
    for (int j = 0; j < NT; ++j) atomicAdd(fold + (ri - j / WS + WS - 1) * (2 * WS - 1) + (ci - j % WS + WS - 1), dbacc[j]);
  }
  __syncthreads();
  for (int k = tid; k < TBL; k += THREADS) atomicAdd(dtable + k * heads + head, fold[k]);
  if (dqkv_bias)
    for (int k = tid; k < 96; k += THREADS) {
      const float v = padacc[k];
      if (v != 0.f) atomicAdd(dqkv_bias + (k / 32) * C + head * HD + (k % 32), v);
    }
  if (warp == 0) tmem_dealloc(tm, B_TMEM);
}

}  // namespace wtc
}  // namespace rsc

extern "C" int rsc_wmsa_bwd_simt(const void *qkv, const float *qkv_bias, const float *bias_table, const void *dout,
                                 void *dqkv, float *dbias_table, float *dqkv_bias, int B, int H, int W, int C,
                                 int heads, int ws, int shift, float scale, int dtype, void *stream);

extern "C" int rsc_wmsa_bwd(const void *qkv, const float *qkv_bias, const float *bias_table, const void *dout,
                            void *dqkv, float *dbias_table, float *dqkv_bias, int B, int H, int W, int C, int heads,
                            int ws, int shift, float scale, int dtype, void *stream) {
  static const bool force_simt = getenv("RSC_WMSA_SIMT") != nullptr;
  const bool tc_ok = dtype == RSC_BF16 && ws == 7 && (shift == 0 || shift == 3) && heads > 0 && C == heads * 32 &&
                     B > 0 && H > 0 && W > 0 && qkv && bias_table && dout && dqkv && dbias_table &&
                     !(dqkv_bias && !qkv_bias) && heads <= 2 * kNumSMs;
  if (!tc_ok || force_simt)
    return rsc_wmsa_bwd_simt(qkv, qkv_bias, bias_table, dout, dqkv, dbias_table, dqkv_bias, B, H, W, C, heads, ws,
                             shift, scale, dtype, stream);
  WinGeom g(B, H, W, ws, shift);
  const int num_windows = B * g.nWh * g.nWw;
  const int num_items = ((num_windows + 1) / 2) * heads;
  auto kern = wtc::wmsa_bwd_tc_kernel;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, wtc::B_TOTAL);
  int grid = (kNumSMs * 2) / heads * heads;
  if (grid > num_items) grid = num_items;   // num_items is a multiple of heads
  kern<<<grid, wtc::THREADS, wtc::B_TOTAL, (cudaStream_t)stream>>>(
      (const __nv_bfloat16 *)qkv, qkv_bias, bias_table, (const __nv_bfloat16 *)dout, (__nv_bfloat16 *)dqkv,
      dbias_table, dqkv_bias, g, C, heads, scale, num_windows, num_items);
  RSC_CHECK_LAUNCH("rsc_wmsa_bwd");
  return RSC_OK;
}
