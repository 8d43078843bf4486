// Fused (shifted-)window attention on the 5th-gen tensor cores (bf16), forward and backward.
//
// One item = one (window, head) UNIT: 49 tokens x 32 dims, padded to 64 rows.  A CTA is 2 warps (64 threads,
// thread = token row) and is persistent over the items of ONE head; several CTAs per SM (6 forward, 4 backward:
// TMEM columns and shared memory are sized for that) give independent load -> MMA -> softmax -> MMA -> store
// pipelines, and inside a CTA the gather of item i+1 (16-byte cp.async into the second input buffer) overlaps
// the math of item i.  All contractions are tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM) with M = 128:
// only accumulator rows 0..63 are real; the A descriptors simply run on into the following shared-memory bytes
// for rows 64..127 (their results are never read -- rows of D are independent), so no zero padding is staged.
//   forward : S = Q K^T (N=64) ; softmax on the row-owner threads (tcgen05.ld 32x32b) ; O = P V (N=32)
//   backward: S = Q K^T, dP = dO V^T ; P, D = rowsum(P*dP), dS = P*(dP-D) ; dV = P^T dO, dK = dS^T Q, dQ = dS K
// qkv is gathered ONCE from its natural (B,H,W,3C) layout through the padded / cyclically shifted window
// coordinates into the no-swizzle core-matrix layout (tools/tc_probe.cu pins these descriptors); padded tokens
// are synthesised from the qkv bias.  HBM-bound (24.5 flop/B forward).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace rsc {
namespace wtc {

using namespace tc;

constexpr int WS = 7, NT = 49, HD = 32, TBL = 169;
constexpr int THREADS = 64;
constexpr int TILE = 4096;   // one operand tile: 64 rows x 32 dims bf16

// 16-byte chunk c (8 bf16) of row r of a token-major [64][32] tile: 8-row groups of 512 B
__device__ __forceinline__ int tile_off(int r, int c) { return (r >> 3) * 512 + c * 128 + (r & 7) * 16; }
// P / dS tile (64 query rows x 64 keys): 16-byte chunk kc (8 keys) of row r: [kc][row group][row in group]
__device__ __forceinline__ int p_off(int r, int kc) { return kc * 1024 + (r >> 3) * 128 + (r & 7) * 16; }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void st_shared16(uint32_t dst, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

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

// Per-item geometry of this thread's token row.
struct RowGeo {
  int b, wh, ww, h, w;
  bool row_ok, tok_ok;   // row_ok: a real window slot (i < 49); tok_ok: it maps to a real (un-padded) token
};

__device__ __forceinline__ RowGeo row_geo(const WinGeom &g, int win, int i, int ri, int ci) {
  RowGeo r;
  r.ww = win % g.nWw;
  r.wh = (win / g.nWw) % g.nWh;
  r.b = win / (g.nWw * g.nWh);
  r.row_ok = i < NT;
  r.h = r.w = 0;
  r.tok_ok = r.row_ok && g.source(r.wh, r.ww, ri, ci, r.h, r.w);
  return r;
}

// This thread's token: PARTS head slices (q|k|v [|dO]) -> its row of the operand tiles at `row` (cp.async).
template <int PARTS, int Q0, int DO0>
__device__ __forceinline__ void gather_row(const RowGeo &r, const WinGeom &g, const __nv_bfloat16 *qkv,
                                           const __nv_bfloat16 *dout, const float *qkv_bias, int C, int head,
                                           uint32_t row) {
  if (!r.row_ok) return;   // rows 49..63 stay zero for the whole kernel
  if (r.tok_ok) {
    const int64_t tok = ((int64_t)r.b * g.H + r.h) * g.W + r.w;
    const __nv_bfloat16 *src = qkv + tok * (3 * C) + head * HD;
#pragma unroll
    for (int part = 0; part < 3; ++part)
#pragma unroll
      for (int c = 0; c < 4; ++c) cp_async16(row + Q0 + part * TILE + c * 128, src + part * C + c * 8);
    if (PARTS == 4) {
      const __nv_bfloat16 *dsrc = dout + tok * C + head * HD;
#pragma unroll
      for (int c = 0; c < 4; ++c) cp_async16(row + DO0 + c * 128, dsrc + c * 8);
    }
  } else {   // zero-padded token: its qkv row is the qkv bias, its output row is cropped (dO = 0)
#pragma unroll
    for (int part = 0; part < PARTS; ++part)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (part < 3 && qkv_bias) {
          const int col = part * C + head * HD + c * 8;
          const float4 f0 = __ldg(reinterpret_cast<const float4 *>(qkv_bias + col));
          const float4 f1 = __ldg(reinterpret_cast<const float4 *>(qkv_bias + col + 4));
          v = make_uint4(pack_bf16(f0.x, f0.y), pack_bf16(f0.z, f0.w), pack_bf16(f1.x, f1.y), pack_bf16(f1.z, f1.w));
        }
        st_shared16(row + (part < 3 ? Q0 + part * TILE : DO0) + c * 128, v);
      }
  }
}

// ---- forward ------------------------------------------------------------------------------------------------
// smem: [P 8 KB | in0: q k v | in1: q k v | bias table].  The M = 128 descriptors of the A operands (P, q) read up
// to one tile past their own 64 rows: P runs into in0, q into k -- always allocated bytes.
constexpr int F_P = 0;
constexpr int F_IN0 = 8192;
constexpr int F_IN = 3 * TILE;                   // q | k | v of one item
constexpr int F_TBL = F_IN0 + 2 * F_IN;
constexpr int F_TOTAL = F_TBL + 704;             // ~32.7 KB -> 6 CTAs / SM
constexpr int F_TMEM = 64;

__global__ void __launch_bounds__(THREADS, 6)
    wmsa_fwd_tc_kernel(const __nv_bfloat16 *__restrict__ qkv, const float *__restrict__ qkv_bias,
                       const float *__restrict__ table, __nv_bfloat16 *__restrict__ out, WinGeom g, int C, int heads,
                       float scale, int num_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float *tbl = reinterpret_cast<float *>(smem + F_TBL);
  const int head = blockIdx.x % heads;   // gridDim.x is a multiple of heads: the head is fixed per CTA

  if (warp == 0) tmem_alloc(&tmem_base_s, F_TMEM);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < F_TBL / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int k = tid; k < TBL; k += THREADS) tbl[k] = __ldg(table + k * heads + head) * LOG2E;   // exp2 domain
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  const uint32_t sb = smem_u32(smem);
  const uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
  const uint32_t idesc_o = make_idesc_bf16(128, 32, false, true);
  uint32_t phase = 0;
  const float scale2 = scale * LOG2E;
  const int i = tid, ri = i / WS, ci = i % WS;   // this thread's row = window slot i (valid if < 49)
  const float *tb = tbl + (ri + WS - 1) * (2 * WS - 1) + (ci + WS - 1);
  const uint32_t row_off = tile_off(tid, 0);
  const int step = gridDim.x;

  int item = blockIdx.x, buf = 0;
  RowGeo cur = row_geo(g, item < num_items ? item / heads : 0, i, ri, ci);
  if (item < num_items) gather_row<3, 0, 0>(cur, g, qkv, nullptr, qkv_bias, C, head, sb + F_IN0 + row_off);
  cp_async_commit();
  for (; item < num_items; item += step, buf ^= 1) {
    // ---- prefetch the next item into the other buffer, then wait for this one ----
    const int nxt = item + step;
    RowGeo nx = cur;
    if (nxt < num_items) {
      nx = row_geo(g, nxt / heads, i, ri, ci);
      gather_row<3, 0, 0>(nx, g, qkv, nullptr, qkv_bias, C, head, sb + F_IN0 + (buf ^ 1) * F_IN + row_off);
    }
    cp_async_commit();
    cp_async_wait_group<1>();
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    const uint32_t in = sb + F_IN0 + buf * F_IN;
    // ---- S = Q K^T (rows 0..63 real) ----
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int k = 0; k < 2; ++k)
        mma_bf16_ss(tm, make_smem_desc(in + k * 256, 128, 512), make_smem_desc(in + TILE + k * 256, 128, 512), idesc_s,
                    k > 0);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    // ---- softmax on this thread's row ----
    float inv_l = 0.f;
    {
      uint32_t s0[32], s1[32];
      const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
      tmem_ld32(taddr, s0);
      tmem_ld32(taddr + 32, s1);
      tmem_ld_wait();
      uint32_t pk[32];   // 64 bf16 probabilities, packed
      if (cur.row_ok) {
        float sv[NT];
        float m;
        // the shift mask only exists in the last window row / column (CTA-uniform branch)
        if (g.shift > 0 && (cur.wh == g.nWh - 1 || cur.ww == g.nWw - 1)) {
          uint32_t rowbits, colbits;
          mask_bits(g, cur.wh, cur.ww, ri, ci, rowbits, colbits);
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
      for (int kc = 0; kc < 8; ++kc)
        st_shared16(sb + F_P + p_off(tid, kc), make_uint4(pk[4 * kc], pk[4 * kc + 1], pk[4 * kc + 2], pk[4 * kc + 3]));
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---- O = P V (overwrites S columns 0..31) ----
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int k = 0; k < 4; ++k)   // K = 64 keys, 16 per step
        mma_bf16_ss(tm, make_smem_desc(sb + F_P + k * 2048, 1024, 128),
                    make_smem_desc(in + 2 * TILE + k * 1024, 512, 128), idesc_o, k > 0);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    // ---- normalise and store this thread's output row ----
    {
      uint32_t o[32];
      tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), o);
      tmem_ld_wait();
      if (cur.tok_ok) {
        __nv_bfloat16 *dst = out + (((int64_t)cur.b * g.H + cur.h) * g.W + cur.w) * C + head * HD;
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
    cur = nx;
    fence_before_sync();   // (the __syncthreads at the top of the next trip orders these TMEM reads before its MMAs)
  }
  cp_async_wait_group<0>();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, F_TMEM);
}

}  // namespace wtc
}  // namespace rsc

using namespace rsc;

extern "C" int rsc_wmsa_fwd_simt(const void *qkv, const float *qkv_bias, const float *bias_table, void *out, int B,
                                 int H, int W, int C, int heads, int ws, int shift, float scale, int dtype,
                                 void *stream);

int rsc_wmsa_fwd_tma(const void *qkv, const float *qkv_bias, const float *bias_table, void *out, int B, int H, int W, int C,
                     int heads, int shift, float scale, void *stream);

extern "C" int rsc_wmsa_fwd(const void *qkv, const float *qkv_bias, const float *bias_table, void *out, int B, int H,
                            int W, int C, int heads, int ws, int shift, float scale, int dtype, void *stream) {
  static const bool force_simt = getenv("RSC_WMSA_SIMT") != nullptr;
  const bool tc_ok = dtype == RSC_BF16 && ws == 7 && (shift == 0 || shift == 3) && heads > 0 && C == heads * 32 &&
                     B > 0 && H > 0 && W > 0 && qkv && bias_table && out && heads <= 6 * kNumSMs;
  if (!tc_ok || force_simt)  // fp32 (exact-arithmetic parity path) and argument errors go through the SIMT entry
    return rsc_wmsa_fwd_simt(qkv, qkv_bias, bias_table, out, B, H, W, C, heads, ws, shift, scale, dtype, stream);
  static const bool v3 = getenv("RSC_WMSA_V3") != nullptr;   // the round-1 cp.async kernel (kept for A/B runs)
  if (!v3) {
    const int rc = rsc_wmsa_fwd_tma(qkv, qkv_bias, bias_table, out, B, H, W, C, heads, shift, scale, stream);
    if (rc >= 0) return rc;
  }
  WinGeom g(B, H, W, ws, shift);
  const int num_items = B * g.nWh * g.nWw * heads;
  auto kern = wtc::wmsa_fwd_tc_kernel;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, wtc::F_TOTAL);
  int grid = (kNumSMs * 6) / heads * heads;    // a multiple of heads: every CTA keeps one head (bias table loaded once)
  if (grid > num_items) grid = num_items;      // num_items is a multiple of heads
  kern<<<grid, wtc::THREADS, wtc::F_TOTAL, (cudaStream_t)stream>>>(
      (const __nv_bfloat16 *)qkv, qkv_bias, bias_table, (__nv_bfloat16 *)out, g, C, heads, scale, num_items);
  RSC_CHECK_LAUNCH("rsc_wmsa_fwd");
  return RSC_OK;
}

// =====================================================================================
// Backward (bf16).  Per (window, head) unit, two MMA groups around the SIMT softmax backward:
//   S  = Q K^T (TMEM cols 0..63)        dP = dO V^T (cols 64..127)            (S is recomputed)
//   -- SIMT: P = softmax(scale*S + bias + mask), D = sum_j P*dP, dS = P*(dP - D); P and dS rows -> smem --
//   dV = P^T dO (cols 0..31) , dK = dS^T Q (32..63)   (A = the smem tile read MN-major, K = 64 query rows)
//   dQ = dS K   (cols 64..95)                          (A = dS K-major, K = 64 key columns)
// d(bias table) is accumulated per thread in registers across the persistent loop (fixed head per
// CTA) and folded once at the end; gradients of padded rows go to the qkv-bias gradient.
// =====================================================================================
namespace rsc {
namespace wtc {

// smem: [P 8 KB | dS 8 KB | in0: dO q k v | in1: dO q k v | bias table | padded-row sums].  The M = 128 descriptors
// of the A operands read up to 8 KB past their own 64 rows / keys: P runs into dS, dS into in0, dO into q, q into
// k -- always allocated bytes (k and v are only B operands with exactly 64 rows).
constexpr int B_P = 0;
constexpr int B_DS = 8192;
constexpr int B_IN0 = 16384;
constexpr int B_IN = 4 * TILE;                  // dO | q | k | v of one item
constexpr int B_TBL = B_IN0 + 2 * B_IN;
constexpr int B_PAD = B_TBL + 704;              // 3 x 32 floats: qkv-bias gradient of padded rows
constexpr int B_TOTAL = B_PAD + 384;            // ~49.1 KB -> 4 CTAs / SM (4 x 128 TMEM columns = all 512)
constexpr int B_TMEM = 128;

__global__ void __launch_bounds__(THREADS, 4)
    wmsa_bwd_tc_kernel(const __nv_bfloat16 *__restrict__ qkv, const float *__restrict__ qkv_bias,
                       const float *__restrict__ table, const __nv_bfloat16 *__restrict__ dout,
                       __nv_bfloat16 *__restrict__ dqkv, float *__restrict__ dtable, float *__restrict__ dqkv_bias,
                       WinGeom g, int C, int heads, float scale, int num_items) {
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
  const uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
  const uint32_t idesc_t = make_idesc_bf16(128, 32, true, true);    // A^T (MN-major) x MN-major B
  const uint32_t idesc_q = make_idesc_bf16(128, 32, false, true);   // K-major A x MN-major B
  uint32_t phase = 0;
  const float scale2 = scale * LOG2E;
  const int i = tid, ri = i / WS, ci = i % WS;
  const float *tb = tbl + (ri + WS - 1) * (2 * WS - 1) + (ci + WS - 1);
  const uint32_t row_off = tile_off(tid, 0);
  const int step = gridDim.x;
  float dbacc[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) dbacc[j] = 0.f;

  int item = blockIdx.x, buf = 0;
  RowGeo cur = row_geo(g, item < num_items ? item / heads : 0, i, ri, ci);
  if (item < num_items) gather_row<4, TILE, 0>(cur, g, qkv, dout, qkv_bias, C, head, sb + B_IN0 + row_off);
  cp_async_commit();
  for (; item < num_items; item += step, buf ^= 1) {
    // ---- prefetch the next item into the other buffer, then wait for this one ----
    const int nxt = item + step;
    RowGeo nx = cur;
    if (nxt < num_items) {
      nx = row_geo(g, nxt / heads, i, ri, ci);
      gather_row<4, TILE, 0>(nx, g, qkv, dout, qkv_bias, C, head, sb + B_IN0 + (buf ^ 1) * B_IN + row_off);
    }
    cp_async_commit();
    cp_async_wait_group<1>();
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    const uint32_t in = sb + B_IN0 + buf * B_IN;
    const uint32_t inDO = in, inQ = in + TILE, inK = in + 2 * TILE, inV = in + 3 * TILE;
    // ---------------- S = Q K^T (cols 0..63), dP = dO V^T (cols 64..127) ----------------
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int k = 0; k < 2; ++k)
        mma_bf16_ss(tm, make_smem_desc(inQ + k * 256, 128, 512), make_smem_desc(inK + k * 256, 128, 512), idesc_s, k > 0);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        mma_bf16_ss(tm + 64, make_smem_desc(inDO + k * 256, 128, 512), make_smem_desc(inV + k * 256, 128, 512), idesc_s,
                    k > 0);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    // ---------------- softmax backward on this thread's row ----------------
    {
      uint32_t s0[32], s1[32];
      const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
      tmem_ld32(taddr, s0);
      tmem_ld32(taddr + 32, s1);
      tmem_ld_wait();
      uint32_t pp[32], pd[32];   // packed bf16 P row and dS row (64 columns each)
      // The TMEM loads are warp-aligned instructions and stay outside the branches; the row arithmetic runs
      // only on real query rows (the others store zeros), with CTA-uniform fast paths: the shift mask exists
      // only in the last window row / column, padded keys only where the map is not a multiple of 7.
      float p[NT];
      float wp = 0.f, wds = 0.f;
      uint32_t rowpad = 0, colpad = 0;
      if (cur.row_ok) {
        float m;
        if (g.shift > 0 && (cur.wh == g.nWh - 1 || cur.ww == g.nWw - 1)) {
          uint32_t rowbits, colbits;
          mask_bits(g, cur.wh, cur.ww, ri, ci, rowbits, colbits);
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
            int hh = cur.wh * WS + j + g.shift, wc = cur.ww * WS + j + g.shift;
            if (hh >= g.Hp) hh -= g.Hp;
            if (wc >= g.Wp) wc -= g.Wp;
            rowpad |= (uint32_t)(hh >= g.H) << j;
            colpad |= (uint32_t)(wc >= g.W) << j;
          }
        }
      }
      tmem_ld32(taddr + 64, s0);   // dP row (reuses the S registers)
      tmem_ld32(taddr + 96, s1);
      tmem_ld_wait();
      if (cur.row_ok) {
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
        const uint32_t o = p_off(tid, kc);
        st_shared16(sb + B_P + o, make_uint4(pp[4 * kc], pp[4 * kc + 1], pp[4 * kc + 2], pp[4 * kc + 3]));
        st_shared16(sb + B_DS + o, make_uint4(pd[4 * kc], pd[4 * kc + 1], pd[4 * kc + 2], pd[4 * kc + 3]));
      }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---------------- dV = P^T dO (cols 0..31), dK = dS^T Q (32..63), dQ = dS K (64..95) ----------------
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int k = 0; k < 4; ++k)   // K = 64 query rows, 16 per step
        mma_bf16_ss(tm, make_smem_desc(sb + B_P + k * 256, 128, 1024), make_smem_desc(inDO + k * 1024, 512, 128), idesc_t,
                    k > 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_bf16_ss(tm + 32, make_smem_desc(sb + B_DS + k * 256, 128, 1024), make_smem_desc(inQ + k * 1024, 512, 128),
                    idesc_t, k > 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)   // K = 64 key columns, 16 per step
        mma_bf16_ss(tm + 64, make_smem_desc(sb + B_DS + k * 2048, 1024, 128), make_smem_desc(inK + k * 1024, 512, 128),
                    idesc_q, k > 0);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    // ---------------- store dq | dk | dv of this thread's token ----------------
    {
      uint32_t o[32];
      const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
      __nv_bfloat16 *dst = dqkv + (((int64_t)cur.b * g.H + cur.h) * g.W + cur.w) * (3 * C) + head * HD;
#pragma unroll
      for (int part = 0; part < 3; ++part) {   // TMEM columns: dV 0, dK 32, dQ 64 -> dqkv parts 2, 1, 0
        tmem_ld32(taddr + part * 32, o);
        tmem_ld_wait();
        const float sc = part == 0 ? 1.0f : scale;
        const int qpart = 2 - part;
        if (cur.tok_ok) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(o[8 * c + 0]) * sc, __uint_as_float(o[8 * c + 1]) * sc);
            v.y = pack_bf16(__uint_as_float(o[8 * c + 2]) * sc, __uint_as_float(o[8 * c + 3]) * sc);
            v.z = pack_bf16(__uint_as_float(o[8 * c + 4]) * sc, __uint_as_float(o[8 * c + 5]) * sc);
            v.w = pack_bf16(__uint_as_float(o[8 * c + 6]) * sc, __uint_as_float(o[8 * c + 7]) * sc);
            *reinterpret_cast<uint4 *>(dst + qpart * C + 8 * c) = v;
          }
        } else if (i == NT && part < 2 && dqkv_bias) {
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
    cur = nx;
    fence_before_sync();   // (the __syncthreads at the top of the next trip orders these TMEM reads before its MMAs)
  }
  cp_async_wait_group<0>();
  __syncthreads();
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

int rsc_wmsa_bwd_tma(const void *qkv, const float *qkv_bias, const float *bias_table, const void *dout, void *dqkv,
                     float *dbias_table, float *dqkv_bias, int B, int H, int W, int C, int heads, int shift, float scale,
                     void *stream);

extern "C" int rsc_wmsa_bwd(const void *qkv, const float *qkv_bias, const float *bias_table, const void *dout,
                            void *dqkv, float *dbias_table, float *dqkv_bias, int B, int H, int W, int C, int heads,
                            int ws, int shift, float scale, int dtype, void *stream) {
  static const bool force_simt = getenv("RSC_WMSA_SIMT") != nullptr;
  const bool tc_ok = dtype == RSC_BF16 && ws == 7 && (shift == 0 || shift == 3) && heads > 0 && C == heads * 32 &&
                     B > 0 && H > 0 && W > 0 && qkv && bias_table && dout && dqkv && dbias_table &&
                     !(dqkv_bias && !qkv_bias) && heads <= 4 * kNumSMs;
  if (!tc_ok || force_simt)
    return rsc_wmsa_bwd_simt(qkv, qkv_bias, bias_table, dout, dqkv, dbias_table, dqkv_bias, B, H, W, C, heads, ws,
                             shift, scale, dtype, stream);
  static const bool v3 = getenv("RSC_WMSA_V3") != nullptr;   // the round-1 cp.async kernel (kept for A/B runs)
  if (!v3) {
    const int rc = rsc_wmsa_bwd_tma(qkv, qkv_bias, bias_table, dout, dqkv, dbias_table, dqkv_bias, B, H, W, C, heads, shift,
                                    scale, stream);
    if (rc >= 0) return rc;
  }
  WinGeom g(B, H, W, ws, shift);
  const int num_items = B * g.nWh * g.nWw * heads;
  auto kern = wtc::wmsa_bwd_tc_kernel;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, wtc::B_TOTAL);
  int grid = (kNumSMs * 4) / heads * heads;
  if (grid > num_items) grid = num_items;   // num_items is a multiple of heads
  kern<<<grid, wtc::THREADS, wtc::B_TOTAL, (cudaStream_t)stream>>>(
      (const __nv_bfloat16 *)qkv, qkv_bias, bias_table, (const __nv_bfloat16 *)dout, (__nv_bfloat16 *)dqkv,
      dbias_table, dqkv_bias, g, C, heads, scale, num_items);
  RSC_CHECK_LAUNCH("rsc_wmsa_bwd");
  return RSC_OK;
}
