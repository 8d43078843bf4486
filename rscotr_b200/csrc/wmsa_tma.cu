// Fused (shifted-)window attention on the 5th-gen tensor cores (bf16), forward and backward, TMA-staged.
// tools/wmsa_probe.cu pins every shared-memory layout / descriptor used here against a host reference.
//
// Unit = one (window, head): 49 tokens x 32 dims.  The window is staged as a 56-SLOT tile
//     slot s = half*28 + i*4 + jj      (i = window row 0..6, window column j = half*4 + jj, jj = 0..3)
// of 64-byte rows (32 bf16), so that TWO TMA boxes of (32 channels, 4 tokens, 7 rows) fetch one operand of one unit
// straight from the natural (B,H,W,3C) activation: no per-thread gathers, no address arithmetic, out-of-bounds
// tokens (the zero padding mmdet applies after norm1) arrive as zeros.  Column j = 7 (slot jj = 3 of half 1) is a
// neighbour's token: it is loaded but never a valid key / query.  Windows of a shifted block that wrap around the
// rolled map use (4 rows) + (3 rows) boxes per half (two more tensor maps).  TMA writes the tiles with
// SWIZZLE_64B; the same bytes are read by tcgen05.mma as K-major operands (S = Q K^T, dP = dO V^T) and as
// MN-major B operands (O = P V, dV = P^T dO, dK = dS^T Q, dQ = dS K).  All contractions are kind::f16 MMAs with
// M = 128 whose accumulator rows 64..127 are never read (the A descriptors run on into the following tile).
//
// Zero-padded tokens: mmdet pads AFTER norm1, so the k / v row of a padded token is the qkv bias.  In windows that
// hold padding, the threads that own padded slots overwrite the zeros TMA delivered with the bias rows before the
// MMAs read the tiles; in the backward pass column 56 of the P / dS tiles carries the row sums over the padded keys,
// so row 56 of dV / dK (summed by the tensor core) is the gradient that reaches the qkv bias through them.
//
// CTA = 64 threads (thread = slot = query row = TMEM lane), persistent over the units of ONE head, 8 (forward) /
// 4 (backward) CTAs per SM.  One elected lane of warp 0 drives TMA (two-stage ring, prefetch distance one unit) and
// issues the MMAs on the uniform datapath.  The relative-position biases of the thread's row live in registers as
// slot pairs; softmax and its backward run in packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2).  HBM-bound
// (24.5 flop/B forward): SURVEY 8d.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace rsc {
namespace wtm {

using namespace tc;

constexpr int WS = 7, NT = 49, HD = 32;   // NT: valid keys per window
constexpr int THREADS = 64;
constexpr uint32_t TILE = 4096;   // 64 slots x 64 bytes
constexpr uint32_t HALF = 1792;   // 28 slots
constexpr int PADSLOT = 56;       // (backward) column of P / dS that carries the row sums over the padded keys
constexpr float LOG2E = 1.4426950408889634f;

// n-th valid key (0..48) -> slot, and slot -> (window row, window column)
__host__ __device__ constexpr int kslot(int n) { return n < 28 ? n : 28 + ((n - 28) / 3) * 4 + (n - 28) % 3; }
__host__ __device__ constexpr int slot_r(int s) { return (s % 28) / 4; }
__host__ __device__ constexpr int slot_c(int s) { return (s / 28) * 4 + s % 4; }
__host__ __device__ constexpr bool slot_ok(int s) { return s < 56 && slot_c(s) < WS; }
// slot -> index among the valid keys (only for valid slots)
__host__ __device__ constexpr int kidx(int s) { return s < 28 ? s : 28 + ((s - 28) / 4) * 3 + (s - 28) % 4; }

// P / dS tile (64 query slots x 64 key slots), no-swizzle core matrices: 16-byte chunk kc (8 keys) of row r
__device__ __forceinline__ uint32_t p_off(int r, int kc) { return kc * 1024 + (r >> 3) * 128 + (r & 7) * 16; }

__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {   // rows of 64 B, 8-row groups 512 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
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
__device__ __forceinline__ void st_shared16(uint32_t dst, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// window walk: (b, wh, ww) advanced by a constant number of windows per trip, no divisions in the loop
struct WinPos {
  int b, wh, ww;
};
struct WinStep {
  int db, dwh, dww;
};
__device__ __forceinline__ void advance(WinPos &p, const WinStep &s, const WinGeom &g) {
  p.ww += s.dww;
  if (p.ww >= g.nWw) p.ww -= g.nWw, ++p.wh;
  p.wh += s.dwh;
  if (p.wh >= g.nWh) p.wh -= g.nWh, ++p.b;
  p.b += s.db;
}

// one operand tile of one unit: channels [c0, c0+32) of window (b, wh, ww)
__device__ __forceinline__ void tma_window(uint32_t dst, uint64_t *bar, const CUtensorMap *m7, const CUtensorMap *m4,
                                           const CUtensorMap *m3, const WinGeom &g, const WinPos &p, int c0) {
  const int hs = p.wh * WS + g.shift, ws0 = p.ww * WS + g.shift;
  const int w1 = ws0 + WS > g.Wp ? 0 : ws0 + 4;   // wrap in w: columns 4..6 come from w = 0..2
  if (hs + WS > g.Hp) {                            // wrap in h: rows 0..3 <- Hp-4.., rows 4..6 <- 0..2
    tma_load_4d(dst, m4, bar, c0, ws0, hs, p.b);
    tma_load_4d(dst + HALF, m4, bar, c0, w1, hs, p.b);
    tma_load_4d(dst + 1024, m3, bar, c0, ws0, 0, p.b);
    tma_load_4d(dst + HALF + 1024, m3, bar, c0, w1, 0, p.b);
  } else {
    tma_load_4d(dst, m7, bar, c0, ws0, hs, p.b);
    tma_load_4d(dst + HALF, m7, bar, c0, w1, hs, p.b);
  }
}

// shift-mask region bits of query (ri, ci) in window p: bit j of rowbits / colbits = key row / column j lies in
// another region (only the last window row / column of a shifted block has two regions)
__device__ __forceinline__ void mask_bits(const WinGeom &g, const WinPos &p, int ri, int ci, uint32_t &rowbits,
                                          uint32_t &colbits) {
  rowbits = colbits = 0;
  if (g.shift == 0) return;
  const bool lr = p.wh == g.nWh - 1, lc = p.ww == g.nWw - 1;
  const int rh_i = lr ? (ri < WS - g.shift ? 1 : 2) : 0, rw_i = lc ? (ci < WS - g.shift ? 1 : 2) : 0;
#pragma unroll
  for (int j = 0; j < WS; ++j) {
    const int rj = j < WS - g.shift ? 1 : 2;
    rowbits |= (uint32_t)((lr ? rj : 0) != rh_i) << j;
    colbits |= (uint32_t)((lc ? rj : 0) != rw_i) << j;
  }
}
// bit j of rowpad / colpad = window row / column j is zero padding
__device__ __forceinline__ void pad_bits(const WinGeom &g, const WinPos &p, uint32_t &rowpad, uint32_t &colpad) {
  rowpad = colpad = 0;
  if (g.Hp == g.H && g.Wp == g.W) return;
#pragma unroll
  for (int j = 0; j < WS; ++j) {
    int hh = p.wh * WS + j + g.shift, wc = p.ww * WS + j + g.shift;
    if (hh >= g.Hp) hh -= g.Hp;
    if (wc >= g.Wp) wc -= g.Wp;
    rowpad |= (uint32_t)(hh >= g.H) << j;
    colpad |= (uint32_t)(wc >= g.W) << j;
  }
}

// 49-bit key mask from the 7-bit row / column masks
__device__ __forceinline__ uint64_t key_mask(uint32_t rowbits, uint32_t colbits) {
  uint64_t m = 0;
#pragma unroll
  for (int n = 0; n < NT; ++n)
    m |= (uint64_t)(((rowbits >> slot_r(kslot(n))) | (colbits >> slot_c(kslot(n)))) & 1u) << n;
  return m;
}
// ---- forward softmax: packed fp32x2 arithmetic (FFMA2 / FADD2) on key-slot PAIRS (2w, 2w+1) --------------------
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  uint64_t ra = *reinterpret_cast<uint64_t *>(&a), rb = *reinterpret_cast<uint64_t *>(&b), rc = *reinterpret_cast<uint64_t *>(&c), rd;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  uint64_t ra = *reinterpret_cast<uint64_t *>(&a), rb = *reinterpret_cast<uint64_t *>(&b), rd;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2 *>(&rd);
}
constexpr int NP = 28;   // slot pairs 0..27 cover slots 0..55; the odd slot of a pair may be the never-valid column 7

// scores (exp2 domain) of slot pair P: S * scale2 + bias (+ shift mask)
template <bool MASK>
__device__ __forceinline__ float2 score2(int P, uint32_t s0, uint32_t s1, const float2 (&bp)[NP], float scale2, uint64_t msk) {
  float2 t = fma2(make_float2(__uint_as_float(s0), __uint_as_float(s1)), make_float2(scale2, scale2), bp[P]);
  if (MASK) {
    if ((msk >> kidx(2 * P)) & 1ull) t.x += -100.0f * LOG2E;
    if (slot_ok(2 * P + 1)) {
      if ((msk >> kidx(2 * P + 1)) & 1ull) t.y += -100.0f * LOG2E;
    }
  }
  return t;
}
// running maximum over one 32-column half (HALF_ID 0: slots 0..31, 1: slots 32..63)
template <bool MASK, int HALF_ID>
__device__ __forceinline__ float half_max(const uint32_t (&x)[32], const float2 (&bp)[NP], float scale2, uint64_t msk, float m) {
#pragma unroll
  for (int w = 0; w < 16; ++w) {
    const int P = HALF_ID * 16 + w;
    if (P < NP) {
      const float2 t = score2<MASK>(P, x[2 * w], x[2 * w + 1], bp, scale2, msk);
      m = slot_ok(2 * P + 1) ? fmaxf(m, fmaxf(t.x, t.y)) : fmaxf(m, t.x);
    }
  }
  return m;
}
// softmax numerators of one half as 16 packed bf16 words, straight into the P tile row; accumulates the row sum
template <bool MASK, int HALF_ID>
__device__ __forceinline__ void half_probs(const uint32_t (&x)[32], const float2 (&bp)[NP], float scale2, uint64_t msk, float m,
                                           float2 &l, uint32_t prow) {
  uint32_t pk[16];
  const float2 nm = make_float2(-m, -m);
#pragma unroll
  for (int w = 0; w < 16; ++w) {
    const int P = HALF_ID * 16 + w;
    if (P < NP) {
      const float2 d = add2(score2<MASK>(P, x[2 * w], x[2 * w + 1], bp, scale2, msk), nm);
      float2 p;
      p.x = ex2(d.x);
      p.y = slot_ok(2 * P + 1) ? ex2(d.y) : 0.f;
      l = add2(l, p);
      pk[w] = pack_bf16(p.x, p.y);
    } else {
      pk[w] = 0u;
    }
  }
#pragma unroll
  for (int kc = 0; kc < 4; ++kc)
    st_shared16(prow + (HALF_ID * 4 + kc) * 1024, make_uint4(pk[4 * kc], pk[4 * kc + 1], pk[4 * kc + 2], pk[4 * kc + 3]));
}

// Softmax of this thread's row: accumulator row at TMEM address taddr (64 columns) -> bf16 numerators in the P tile
// row at shared address prow; returns the row sum.  Executed by all 32 lanes of the warp (tcgen05.ld is
// warp-collective).  The row is read from TMEM half a row at a time, once for the maximum and once more for the
// exponentials: 32 accumulator + 56 bias registers live instead of 64 + 56 + 49.  Rows that are not real queries
// (slots 56..63, window column 7) produce finite junk: in the forward pass such rows of P only reach rows of O
// that are never stored.
template <bool MASK>
__device__ __forceinline__ float softmax_row(uint32_t taddr, const float2 (&bp)[NP], float scale2, uint64_t msk, uint32_t prow) {
  uint32_t x[32];
  tmem_ld32(taddr + 32, x);
  tmem_ld_wait();
  float m = half_max<MASK, 1>(x, bp, scale2, msk, -INFINITY);
  tmem_ld32(taddr, x);
  tmem_ld_wait();
  m = half_max<MASK, 0>(x, bp, scale2, msk, m);
  float2 l = make_float2(0.f, 0.f);
  tmem_ld32(taddr, x);   // (a fresh copy: keeps the first-pass scores from being held in registers)
  tmem_ld_wait();
  half_probs<MASK, 0>(x, bp, scale2, msk, m, l, prow);
  tmem_ld32(taddr + 32, x);
  tmem_ld_wait();
  half_probs<MASK, 1>(x, bp, scale2, msk, m, l, prow);
  return l.x + l.y;
}

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

// ---- forward ------------------------------------------------------------------------------------------------
// smem: [Q0 | K0 | Q1 | K1 | V0 | V1 | k,v bias rows | barriers].  The P tile (8 KB) overwrites Q | K of its own stage
// once S is in TMEM.  Every M = 128 over-read (8 KB behind Q, 16 KB behind P) stays inside the tiles; rows 64..127
// of D are never read.
constexpr uint32_t F_QK = 2 * TILE;                 // Q | K of one stage (= its P tile)
constexpr uint32_t F_V0 = 2 * F_QK;
constexpr uint32_t F_BIAS = F_V0 + 2 * TILE;        // 24 KB of tiles, then 2 x 64 bytes: k / v bias of this head (bf16)
constexpr uint32_t F_BAR = F_BIAS + 128;
constexpr uint32_t F_TOTAL = F_BAR + 64;            // -> 8 CTAs / SM (8 x 64 TMEM columns = all 512)
constexpr int F_TMEM = 64;

// A zero-padded token's k / v row is the qkv bias (mmdet pads after norm1): the thread that owns a padded slot
// overwrites the zeros TMA delivered (swizzled 16-byte chunks of row `slot`).
__device__ __forceinline__ void fix_padded_row(uint32_t tile, uint32_t bias_row, int slot) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(bias_row + c * 16));
    st_shared16(tile + slot * 64 + ((c ^ ((slot >> 1) & 3)) * 16), v);
  }
}

template <int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS)
    wmsa_fwd_tma_kernel(const __grid_constant__ CUtensorMap m7, const __grid_constant__ CUtensorMap m4,
                        const __grid_constant__ CUtensorMap m3, const float *__restrict__ qkv_bias,
                        const float *__restrict__ table, __nv_bfloat16 *__restrict__ out, WinGeom g, int C, int heads,
                        float scale, int num_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + F_BAR);   // full[0], full[1]
  uint64_t &mbar = full[2];
  uint32_t &tmem_base_s = *reinterpret_cast<uint32_t *>(smem + F_BAR + 32);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction: the control code of
                                                            // warp 0 (TMA, MMA issue) runs on the uniform datapath
  const uint32_t sb = smem_u32(smem);
  if (sb & 1023u) __trap();   // the swizzled tiles need the 1024-byte alignment the declaration asks for
  const int head = blockIdx.x % heads;   // gridDim.x is a multiple of heads: the head is fixed per CTA

  if (warp == 0) tmem_alloc(&tmem_base_s, F_TMEM);
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&mbar, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < (int)F_BAR / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (qkv_bias && tid < 8) {   // k (tid 0..3) and v (4..7) bias slices of this head, bf16
    const float *src = qkv_bias + (1 + (tid >> 2)) * C + head * HD + (tid & 3) * 8;
    const float4 f0 = __ldg(reinterpret_cast<const float4 *>(src)), f1 = __ldg(reinterpret_cast<const float4 *>(src + 4));
    st_shared16(sb + F_BIAS + tid * 16,
                make_uint4(pack_bf16(f0.x, f0.y), pack_bf16(f0.z, f0.w), pack_bf16(f1.x, f1.y), pack_bf16(f1.z, f1.w)));
  }
  // this thread's query slot and the relative-position biases of its row (exp2 domain), as slot pairs
  const bool row_ok = slot_ok(tid);
  const int ri = slot_r(tid), ci = slot_c(tid);
  float2 bp[NP];
#pragma unroll
  for (int P = 0; P < NP; ++P) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = 2 * P + e;
      const int idx = (ri - slot_r(c) + WS - 1) * (2 * WS - 1) + (ci - slot_c(c) + WS - 1);
      v[e] = (row_ok && slot_ok(c)) ? __ldg(table + idx * heads + head) * LOG2E : 0.f;
    }
    bp[P] = make_float2(v[0], v[1]);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  const uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
  const uint32_t idesc_o = make_idesc_bf16(128, 32, false, true);
  const float scale2 = scale * LOG2E;
  const bool any_pad = g.Hp != g.H || g.Wp != g.W;
  uint32_t phase = 0;

  // window walk
  const int wstep = gridDim.x / heads;
  WinStep st;
  st.dww = wstep % g.nWw;
  st.dwh = (wstep / g.nWw) % g.nWh;
  st.db = wstep / (g.nWw * g.nWh);
  WinPos cur;
  {
    const int win = blockIdx.x / heads;
    cur.ww = win % g.nWw;
    cur.wh = (win / g.nWw) % g.nWh;
    cur.b = win / (g.nWw * g.nWh);
  }
  const int step = gridDim.x;
  int item = blockIdx.x;
  WinPos pf = cur;   // window of the next unit to prefetch (tracked by every lane: all of it is warp-uniform)
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (warp == 0 && item + s * step < num_items) {
      if (elect_one()) {
        mbar_expect_tx(&full[s], 6 * HALF);
#pragma unroll
        for (int part = 0; part < 3; ++part)
          tma_window(sb + (part < 2 ? s * F_QK + part * TILE : F_V0 + s * TILE), &full[s], &m7, &m4, &m3, g, pf,
                     part * C + head * HD);
      }
      __syncwarp();
    }
    advance(pf, st, g);
  }

  for (int it = 0; item < num_items; ++it, item += step) {
    const int buf = it & 1;
    const uint32_t in = sb + buf * F_QK, inV = sb + F_V0 + buf * TILE;   // Q | K (later P) and V of this stage
    // geometry of this thread's token
    int h = cur.wh * WS + ri + g.shift, w = cur.ww * WS + ci + g.shift;
    if (h >= g.Hp) h -= g.Hp;
    if (w >= g.Wp) w -= g.Wp;
    const bool tok_ok = row_ok && h < g.H && w < g.W;
    const bool rim = cur.wh == g.nWh - 1 || cur.ww == g.nWw - 1;   // CTA-uniform
    // does the window hold zero-padded tokens?  (its last un-wrapped source row / column reaches the padding; with
    // pad + shift > 7 that already happens in the second-to-last window row / column)
    const bool has_pad = any_pad && (min(cur.wh * WS + WS - 1 + g.shift, g.Hp - 1) >= g.H ||
                                     min(cur.ww * WS + WS - 1 + g.shift, g.Wp - 1) >= g.W);
    if (has_pad) {   // padded keys of this window: zeros -> bias rows, before the MMAs read the tiles
      mbar_wait(&full[buf], (it >> 1) & 1);
      if (row_ok && !tok_ok) {
        fix_padded_row(in + TILE, sb + F_BIAS, tid);
        fix_padded_row(inV, sb + F_BIAS + 64, tid);
      }
      fence_async_smem();
    }
    // (the TMEM reads of the previous unit are ordered before this unit's MMAs)
    fence_before_sync();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        fence_after_sync();
        mbar_wait(&full[buf], (it >> 1) & 1);
#pragma unroll
        for (int k = 0; k < 2; ++k)   // K = 32 channels: +32 bytes inside the swizzled 64-byte rows
          mma_bf16_ss(tm, desc_sw64(in + k * 32), desc_sw64(in + TILE + k * 32), idesc_s, k > 0);
        mma_commit(&mbar);
      }
      __syncwarp();
    }
    uint64_t msk = 0;
    const bool masked = rim && g.shift > 0;
    if (masked) {
      uint32_t rowbits, colbits;
      mask_bits(g, cur, ri, ci, rowbits, colbits);
      msk = key_mask(rowbits, colbits);
    }
    mbar_wait(&mbar, phase);
    phase ^= 1;
    fence_after_sync();
    // ---- softmax on this thread's row; P goes where Q | K were ----
    float inv_l;
    {
      const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16), prow = in + p_off(tid, 0);
      const float l = masked ? softmax_row<true>(taddr, bp, scale2, msk, prow) : softmax_row<false>(taddr, bp, scale2, 0ull, prow);
      inv_l = 1.0f / l;
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---- O = P V (overwrites S columns 0..31) ----
    if (warp == 0) {
      if (elect_one()) {
        fence_after_sync();
#pragma unroll
        for (int k = 0; k < 4; ++k)   // K = 64 key slots, 16 per step
          mma_bf16_ss(tm, make_smem_desc(in + k * 2048, 1024, 128), desc_sw64(inV + k * 1024), idesc_o, k > 0);
        mma_commit(&mbar);
      }
      __syncwarp();
    }
    mbar_wait(&mbar, phase);
    phase ^= 1;
    fence_after_sync();
    uint32_t o[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), o);
    // the stage is free again: refill it with the unit after next
    if (warp == 0 && item + 2 * step < num_items) {
      if (elect_one()) {
        mbar_expect_tx(&full[buf], 6 * HALF);
#pragma unroll
        for (int part = 0; part < 3; ++part)
          tma_window(part < 2 ? in + part * TILE : inV, &full[buf], &m7, &m4, &m3, g, pf, part * C + head * HD);
      }
      __syncwarp();
    }
    advance(pf, st, g);
    tmem_ld_wait();
    // ---- normalise and store this thread's output row ----
    if (tok_ok) {
      __nv_bfloat16 *dst = out + (((int64_t)cur.b * g.H + h) * g.W + w) * C + head * HD;
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
    advance(cur, st, g);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, F_TMEM);
}

// ---- backward -----------------------------------------------------------------------------------------------
// Per unit, two MMA groups around the SIMT softmax backward (S is recomputed):
//   S  = Q K^T (TMEM cols 0..63)          dP = dO V^T (cols 64..127)
//   -- SIMT: P = softmax(scale*S + bias + mask), D = sum_j P*dP, dS' = scale * P*(dP - D); P and dS' rows -> smem --
//   dV = P^T dO (cols 0..31), dK = dS'^T Q (32..63)   (A = the P / dS' tile read MN-major, K = 64 query slots)
//   dQ = dS' K  (cols 64..95)                          (A = dS' K-major, K = 64 key slots)
// d(bias table) accumulates per thread in 49 registers across the persistent loop (in units of `scale`) and is
// folded once per CTA.  Windows with zero-padded tokens: the padded key rows of K / V are set to the qkv bias before
// the MMAs, column 56 of P / dS' carries the row sums over the padded keys, and row 56 of dV / dK is then the
// gradient that reaches the k / v bias through the padded rows (summed by the tensor core).
// smem: [P 8 KB | dS 8 KB | stage 0: dO Q K V | stage 1 | k,v bias rows | padded-row sums | barriers]; every
// M = 128 over-read (8 KB behind P / dS / dO / Q) lands in the tile that follows.
constexpr uint32_t B_P = 0;
constexpr uint32_t B_DS = 8192;
constexpr uint32_t B_IN0 = 16384;
constexpr uint32_t B_STAGE = 4 * TILE;              // dO | Q | K | V of one unit
constexpr uint32_t B_BIAS = B_IN0 + 2 * B_STAGE;    // 48 KB of tiles, then 2 x 64 bytes of bias rows
constexpr uint32_t B_PADACC = B_BIAS + 128;         // 64 floats: d(k bias), d(v bias) of this head from padded rows
constexpr uint32_t B_BAR = B_PADACC + 256;
constexpr uint32_t B_TOTAL = B_BAR + 64;            // ~48.5 KB -> 4 CTAs / SM (4 x 128 TMEM columns = all 512)
constexpr int B_TMEM = 128;

__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  uint64_t ra = *reinterpret_cast<uint64_t *>(&a), rb = *reinterpret_cast<uint64_t *>(&b), rd;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2 *>(&rd);
}

// one row of a P / dS tile from its slot pairs (pair w = packed bf16 word w); `pad` goes to slot 56
__device__ __forceinline__ void store_row(uint32_t row_addr, const float2 (&v)[NP], float pad) {
#pragma unroll
  for (int kc = 0; kc < 8; ++kc) {
    uint32_t w4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int w = kc * 4 + j;
      w4[j] = w < NP ? pack_bf16(v[w].x, v[w].y) : (w == NP ? pack_bf16(pad, 0.f) : 0u);
    }
    st_shared16(row_addr + kc * 1024, make_uint4(w4[0], w4[1], w4[2], w4[3]));
  }
}

// Softmax backward of this thread's row, packed fp32x2 arithmetic on slot pairs.  S at TMEM taddr (64 cols), dP at
// taddr + 64.  Writes the P and dS' = scale * P * (dP - D) rows and accumulates the bias-table gradient (in units of
// scale).  pad = 49-bit mask of the zero-padded keys, msk = of the shift-masked keys (RIM only).  Rows that are not
// real queries (row_ok false) write zeros: dV / dK sum over ALL 64 query slots.
template <bool RIM>
__device__ __forceinline__ void softmax_bwd_row(uint32_t taddr, const float2 (&bp)[NP], float scale2, float scale, uint64_t msk,
                                                uint64_t pad, bool row_ok, uint32_t prow, uint32_t dsrow, float2 (&dbacc)[NP]) {
  float2 p[NP];
  uint32_t x[32];
  float m = -INFINITY;
  tmem_ld32(taddr, x);
  tmem_ld_wait();
#pragma unroll
  for (int w = 0; w < 16; ++w) {
    p[w] = score2<RIM>(w, x[2 * w], x[2 * w + 1], bp, scale2, msk);
    m = slot_ok(2 * w + 1) ? fmaxf(m, fmaxf(p[w].x, p[w].y)) : fmaxf(m, p[w].x);
  }
  tmem_ld32(taddr + 32, x);
  tmem_ld_wait();
#pragma unroll
  for (int w = 16; w < NP; ++w) {
    p[w] = score2<RIM>(w, x[2 * w - 32], x[2 * w - 31], bp, scale2, msk);
    m = slot_ok(2 * w + 1) ? fmaxf(m, fmaxf(p[w].x, p[w].y)) : fmaxf(m, p[w].x);
  }
  const float2 nm = make_float2(-m, -m);
  float2 l = make_float2(0.f, 0.f);
#pragma unroll
  for (int w = 0; w < NP; ++w) {
    const float2 d = add2(p[w], nm);
    p[w].x = ex2(d.x);
    p[w].y = slot_ok(2 * w + 1) ? ex2(d.y) : 0.f;
    l = add2(l, p[w]);
  }
  const float inv_l = row_ok ? 1.0f / (l.x + l.y) : 0.f;   // (rows that are no queries become zero rows)
  const float2 il = make_float2(inv_l, inv_l);
  float wp = 0.f;
#pragma unroll
  for (int w = 0; w < NP; ++w) {
    p[w] = mul2(p[w], il);
    if (RIM) {
      if ((pad >> kidx(2 * w)) & 1ull) wp += p[w].x;
      if (slot_ok(2 * w + 1)) {
        if ((pad >> kidx(2 * w + 1)) & 1ull) wp += p[w].y;
      }
    }
  }
  store_row(prow, p, wp);
  // D = sum_j P * dP
  float2 D2 = make_float2(0.f, 0.f);
  tmem_ld32(taddr + 64, x);
  tmem_ld_wait();
#pragma unroll
  for (int w = 0; w < 16; ++w) D2 = fma2(p[w], make_float2(__uint_as_float(x[2 * w]), __uint_as_float(x[2 * w + 1])), D2);
  uint32_t y[32];
  tmem_ld32(taddr + 96, y);
  tmem_ld_wait();
#pragma unroll
  for (int w = 16; w < NP; ++w)
    D2 = fma2(p[w], make_float2(__uint_as_float(y[2 * w - 32]), __uint_as_float(y[2 * w - 31])), D2);
  const float nDs = -(D2.x + D2.y) * scale;
  const float2 nD = make_float2(nDs, nDs), sc = make_float2(scale, scale);
  // dS' = P * (scale * dP - scale * D), in place of p
  float wds = 0.f;
#pragma unroll
  for (int w = 0; w < NP; ++w) {
    const float2 dp = w < 16 ? make_float2(__uint_as_float(x[2 * w]), __uint_as_float(x[2 * w + 1]))
                             : make_float2(__uint_as_float(y[2 * w - 32]), __uint_as_float(y[2 * w - 31]));
    p[w] = mul2(p[w], fma2(dp, sc, nD));
    dbacc[w] = add2(dbacc[w], p[w]);
    if (RIM) {
      if ((pad >> kidx(2 * w)) & 1ull) wds += p[w].x;
      if (slot_ok(2 * w + 1)) {
        if ((pad >> kidx(2 * w + 1)) & 1ull) wds += p[w].y;
      }
    }
  }
  store_row(dsrow, p, wds);
}

__global__ void __launch_bounds__(THREADS, 4)
    wmsa_bwd_tma_kernel(const __grid_constant__ CUtensorMap m7, const __grid_constant__ CUtensorMap m4,
                        const __grid_constant__ CUtensorMap m3, const __grid_constant__ CUtensorMap d7,
                        const __grid_constant__ CUtensorMap d4, const __grid_constant__ CUtensorMap d3,
                        const float *__restrict__ qkv_bias, const float *__restrict__ table, __nv_bfloat16 *__restrict__ dqkv,
                        float *__restrict__ dtable, float *__restrict__ dqkv_bias, WinGeom g, int C, int heads, float scale,
                        int num_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + B_BAR);   // full[0], full[1]
  uint64_t &mbar = full[2];
  uint32_t &tmem_base_s = *reinterpret_cast<uint32_t *>(smem + B_BAR + 32);
  float *padacc = reinterpret_cast<float *>(smem + B_PADACC);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction (control code of warp 0)
  const uint32_t sb = smem_u32(smem);
  if (sb & 1023u) __trap();
  const int head = blockIdx.x % heads;   // gridDim.x is a multiple of heads

  if (warp == 0) tmem_alloc(&tmem_base_s, B_TMEM);
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&mbar, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < (int)B_BAR / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (qkv_bias && tid < 8) {   // k (tid 0..3) and v (4..7) bias slices of this head, bf16
    const float *src = qkv_bias + (1 + (tid >> 2)) * C + head * HD + (tid & 3) * 8;
    const float4 f0 = __ldg(reinterpret_cast<const float4 *>(src)), f1 = __ldg(reinterpret_cast<const float4 *>(src + 4));
    st_shared16(sb + B_BIAS + tid * 16,
                make_uint4(pack_bf16(f0.x, f0.y), pack_bf16(f0.z, f0.w), pack_bf16(f1.x, f1.y), pack_bf16(f1.z, f1.w)));
  }
  const bool row_ok = slot_ok(tid);
  const int ri = slot_r(tid), ci = slot_c(tid);
  float2 bp[NP], dbacc[NP];
#pragma unroll
  for (int P = 0; P < NP; ++P) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = 2 * P + e;
      const int idx = (ri - slot_r(c) + WS - 1) * (2 * WS - 1) + (ci - slot_c(c) + WS - 1);
      v[e] = (row_ok && slot_ok(c)) ? __ldg(table + idx * heads + head) * LOG2E : 0.f;
    }
    bp[P] = make_float2(v[0], v[1]);
    dbacc[P] = make_float2(0.f, 0.f);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  const uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
  const uint32_t idesc_t = make_idesc_bf16(128, 32, true, true);    // A^T (MN-major) x MN-major B
  const uint32_t idesc_q = make_idesc_bf16(128, 32, false, true);   // K-major A x MN-major B
  const float scale2 = scale * LOG2E;
  const bool any_pad = g.Hp != g.H || g.Wp != g.W;
  uint32_t phase = 0;

  const int wstep = gridDim.x / heads;
  WinStep st;
  st.dww = wstep % g.nWw;
  st.dwh = (wstep / g.nWw) % g.nWh;
  st.db = wstep / (g.nWw * g.nWh);
  WinPos cur;
  {
    const int win = blockIdx.x / heads;
    cur.ww = win % g.nWw;
    cur.wh = (win / g.nWw) % g.nWh;
    cur.b = win / (g.nWw * g.nWh);
  }
  const int step = gridDim.x;
  int item = blockIdx.x;
  WinPos pf = cur;
  auto load_unit = [&](int s) {   // one elected lane: the four operand tiles of the unit at `pf` into stage s
    const uint32_t base = sb + B_IN0 + s * B_STAGE;
    mbar_expect_tx(&full[s], 8 * HALF);
    tma_window(base, &full[s], &d7, &d4, &d3, g, pf, head * HD);
#pragma unroll
    for (int part = 0; part < 3; ++part)
      tma_window(base + (1 + part) * TILE, &full[s], &m7, &m4, &m3, g, pf, part * C + head * HD);
  };
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (warp == 0 && item + s * step < num_items) {
      if (elect_one()) load_unit(s);
      __syncwarp();
    }
    advance(pf, st, g);
  }

  for (int it = 0; item < num_items; ++it, item += step) {
    const int buf = it & 1;
    const uint32_t in = sb + B_IN0 + buf * B_STAGE;
    const uint32_t inDO = in, inQ = in + TILE, inK = in + 2 * TILE, inV = in + 3 * TILE;
    int h = cur.wh * WS + ri + g.shift, w = cur.ww * WS + ci + g.shift;
    if (h >= g.Hp) h -= g.Hp;
    if (w >= g.Wp) w -= g.Wp;
    const bool tok_ok = row_ok && h < g.H && w < g.W;
    const bool rim = cur.wh == g.nWh - 1 || cur.ww == g.nWw - 1;   // CTA-uniform
    const bool has_pad = any_pad && (min(cur.wh * WS + WS - 1 + g.shift, g.Hp - 1) >= g.H ||
                                     min(cur.ww * WS + WS - 1 + g.shift, g.Wp - 1) >= g.W);
    if (has_pad) {   // padded keys: zeros -> bias rows
      mbar_wait(&full[buf], (it >> 1) & 1);
      if (row_ok && !tok_ok) {
        fix_padded_row(inK, sb + B_BIAS, tid);
        fix_padded_row(inV, sb + B_BIAS + 64, tid);
      }
      fence_async_smem();
    }
    fence_before_sync();
    __syncthreads();
    // ---------------- S = Q K^T (cols 0..63), dP = dO V^T (cols 64..127) ----------------
    if (warp == 0) {
      if (elect_one()) {
        fence_after_sync();
        mbar_wait(&full[buf], (it >> 1) & 1);
#pragma unroll
        for (int k = 0; k < 2; ++k) mma_bf16_ss(tm, desc_sw64(inQ + k * 32), desc_sw64(inK + k * 32), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 2; ++k) mma_bf16_ss(tm + 64, desc_sw64(inDO + k * 32), desc_sw64(inV + k * 32), idesc_s, k > 0);
        mma_commit(&mbar);
      }
      __syncwarp();
    }
    uint64_t msk = 0, pad = 0;
    const bool slow = has_pad || (rim && g.shift > 0);
    if (slow) {
      uint32_t rowbits, colbits;
      mask_bits(g, cur, ri, ci, rowbits, colbits);
      if (rim) msk = key_mask(rowbits, colbits);
      pad_bits(g, cur, rowbits, colbits);
      pad = key_mask(rowbits, colbits);
    }
    mbar_wait(&mbar, phase);
    phase ^= 1;
    fence_after_sync();
    // ---------------- softmax backward on this thread's row ----------------
    {
      const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
      const uint32_t prow = sb + B_P + p_off(tid, 0), dsrow = sb + B_DS + p_off(tid, 0);
      if (slow) softmax_bwd_row<true>(taddr, bp, scale2, scale, msk, pad, row_ok, prow, dsrow, dbacc);
      else softmax_bwd_row<false>(taddr, bp, scale2, scale, 0ull, 0ull, row_ok, prow, dsrow, dbacc);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---------------- dV = P^T dO (cols 0..31), dK = dS'^T Q (32..63), dQ = dS' K (64..95) ----------------
    if (warp == 0) {
      if (elect_one()) {
        fence_after_sync();
#pragma unroll
        for (int k = 0; k < 4; ++k)   // K = 64 query slots, 16 per step
          mma_bf16_ss(tm, make_smem_desc(sb + B_P + k * 256, 128, 1024), desc_sw64(inDO + k * 1024), idesc_t, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_bf16_ss(tm + 32, make_smem_desc(sb + B_DS + k * 256, 128, 1024), desc_sw64(inQ + k * 1024), idesc_t, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // K = 64 key slots, 16 per step
          mma_bf16_ss(tm + 64, make_smem_desc(sb + B_DS + k * 2048, 1024, 128), desc_sw64(inK + k * 1024), idesc_q, k > 0);
        mma_commit(&mbar);
      }
      __syncwarp();
    }
    mbar_wait(&mbar, phase);
    phase ^= 1;
    fence_after_sync();
    // the stage is free again: refill it with the unit after next
    if (warp == 0 && item + 2 * step < num_items) {
      if (elect_one()) load_unit(buf);
      __syncwarp();
    }
    advance(pf, st, g);
    // ---------------- store dv | dk | dq of this thread's token ----------------
    {
      const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
      uint32_t o[3][32];   // three register sets: a store still reading its registers never blocks the next TMEM load
#pragma unroll
      for (int part = 0; part < 3; ++part) tmem_ld32(taddr + part * 32, o[part]);   // TMEM columns: dV 0, dK 32, dQ 64
      tmem_ld_wait();
      if (tok_ok) {
        __nv_bfloat16 *dst = dqkv + (((int64_t)cur.b * g.H + h) * g.W + w) * (3 * C) + head * HD;
#pragma unroll
        for (int part = 0; part < 3; ++part) {   // -> dqkv parts 2, 1, 0
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(o[part][8 * c + 0]), __uint_as_float(o[part][8 * c + 1]));
            v.y = pack_bf16(__uint_as_float(o[part][8 * c + 2]), __uint_as_float(o[part][8 * c + 3]));
            v.z = pack_bf16(__uint_as_float(o[part][8 * c + 4]), __uint_as_float(o[part][8 * c + 5]));
            v.w = pack_bf16(__uint_as_float(o[part][8 * c + 6]), __uint_as_float(o[part][8 * c + 7]));
            *reinterpret_cast<uint4 *>(dst + (2 - part) * C + 8 * c) = v;
          }
        }
      } else if (tid == PADSLOT && has_pad) {
        // row 56 = sum over this window's padded keys (only this thread touches padacc: no atomics)
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int d = 0; d < 32; ++d) padacc[(1 - part) * 32 + d] += __uint_as_float(o[part][d]);
      }
    }
    advance(cur, st, g);
  }
  fence_before_sync();
  __syncthreads();
  // ---------------- fold the bias-table gradient: registers -> [slot][key] matrix -> 169 table entries ----------------
  float *mat = reinterpret_cast<float *>(smem);   // 64 x 49 floats over the (now idle) P / dS tiles
  if (row_ok) {
#pragma unroll
    for (int P = 0; P < NP; ++P) {
      mat[tid * NT + kidx(2 * P)] = dbacc[P].x;
      if (slot_ok(2 * P + 1)) mat[tid * NT + kidx(2 * P + 1)] = dbacc[P].y;
    }
  }
  __syncthreads();
  const float inv_scale = 1.0f / scale;
  for (int k = tid; k < (2 * WS - 1) * (2 * WS - 1); k += THREADS) {
    const int dr = k / (2 * WS - 1) - (WS - 1), dc = k % (2 * WS - 1) - (WS - 1);   // query - key offsets
    float acc = 0.f;
    for (int qr = max(0, dr); qr < min(WS, WS + dr); ++qr)
      for (int qc = max(0, dc); qc < min(WS, WS + dc); ++qc) {
        const int kr = qr - dr, kc = qc - dc;
        const int qs = (qc >> 2) * 28 + qr * 4 + (qc & 3), ks = (kc >> 2) * 28 + kr * 4 + (kc & 3);
        acc += mat[qs * NT + kidx(ks)];
      }
    atomicAdd(dtable + k * heads + head, acc * inv_scale);
  }
  if (dqkv_bias && any_pad && tid < 64) {
    const float v = padacc[tid];   // [0,32): k part, [32,64): v part
    if (v != 0.f) atomicAdd(dqkv_bias + (1 + (tid >> 5)) * C + head * HD + (tid & 31), v);
  }
  if (warp == 0) tmem_dealloc(tm, B_TMEM);
}

// ---- host side: tensor maps ----------------------------------------------------------------------------------
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

// maps over a (B, H, W, ch) bf16 tensor with boxes (32 channels, 4 tokens, rows, 1)
static bool window_maps(const void *base, int B, int H, int W, int ch, CUtensorMap *m7, CUtensorMap *m4, CUtensorMap *m3) {
  EncodeTiled enc = get_encode();
  if (!enc) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)ch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)ch * 2, (cuuint64_t)W * ch * 2, (cuuint64_t)H * W * ch * 2};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap *maps[3] = {m7, m4, m3};
  const cuuint32_t rows[3] = {7, 4, 3};
  for (int i = 0; i < 3; ++i) {
    const cuuint32_t box[4] = {HD, 4, rows[i], 1};
    if (enc(maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  return true;
}

}  // namespace wtm
}  // namespace rsc

using namespace rsc;

extern "C" int rsc_wmsa_fwd_simt(const void *qkv, const float *qkv_bias, const float *bias_table, void *out, int B, int H,
                                 int W, int C, int heads, int ws, int shift, float scale, int dtype, void *stream);
extern "C" int rsc_wmsa_bwd_simt(const void *qkv, const float *qkv_bias, const float *bias_table, const void *dout, void *dqkv,
                                 float *dbias_table, float *dqkv_bias, int B, int H, int W, int C, int heads, int ws, int shift,
                                 float scale, int dtype, void *stream);

static bool tc_shape_ok(int dtype, int ws, int shift, int heads, int C, int B, int H, int W) {
  return dtype == RSC_BF16 && ws == 7 && (shift == 0 || shift == 3) && heads > 0 && C == heads * 32 && B > 0 && H > 0 && W > 0;
}

extern "C" int rsc_wmsa_fwd(const void *qkv, const float *qkv_bias, const float *bias_table, void *out, int B, int H, int W,
                            int C, int heads, int ws, int shift, float scale, int dtype, void *stream) {
  static const bool force_simt = getenv("RSC_WMSA_SIMT") != nullptr;
  if (force_simt || !tc_shape_ok(dtype, ws, shift, heads, C, B, H, W) || !qkv || !bias_table || !out || heads > 8 * kNumSMs)
    // fp32 (the exact-arithmetic parity path) and argument errors go through the SIMT entry, which validates
    return rsc_wmsa_fwd_simt(qkv, qkv_bias, bias_table, out, B, H, W, C, heads, ws, shift, scale, dtype, stream);
  RSC_CHECK_ARG(((uintptr_t)qkv & 15) == 0, "rsc_wmsa_fwd: qkv must be 16-byte aligned (TMA)");
  WinGeom g(B, H, W, wtm::WS, shift);
  CUtensorMap m7, m4, m3;
  RSC_CHECK_ARG(wtm::window_maps(qkv, B, H, W, 3 * C, &m7, &m4, &m3), "rsc_wmsa_fwd: cuTensorMapEncodeTiled failed");
  const int num_items = B * g.nWh * g.nWw * heads;
  auto kern = wtm::wmsa_fwd_tma_kernel<8>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, wtm::F_TOTAL);
  int grid = (kNumSMs * 8) / heads * heads;   // a multiple of heads: every CTA keeps one head
  if (grid > num_items) grid = num_items;     // num_items is a multiple of heads
  kern<<<grid, wtm::THREADS, wtm::F_TOTAL, (cudaStream_t)stream>>>(m7, m4, m3, qkv_bias, bias_table, (__nv_bfloat16 *)out, g, C,
                                                                   heads, scale, num_items);
  RSC_CHECK_LAUNCH("rsc_wmsa_fwd");
  return RSC_OK;
}

extern "C" int rsc_wmsa_bwd(const void *qkv, const float *qkv_bias, const float *bias_table, const void *dout, void *dqkv,
                            float *dbias_table, float *dqkv_bias, int B, int H, int W, int C, int heads, int ws, int shift,
                            float scale, int dtype, void *stream) {
  static const bool force_simt = getenv("RSC_WMSA_SIMT") != nullptr;
  if (force_simt || !tc_shape_ok(dtype, ws, shift, heads, C, B, H, W) || !qkv || !bias_table || !dout || !dqkv || !dbias_table ||
      (dqkv_bias && !qkv_bias) || heads > 4 * kNumSMs)
    return rsc_wmsa_bwd_simt(qkv, qkv_bias, bias_table, dout, dqkv, dbias_table, dqkv_bias, B, H, W, C, heads, ws, shift, scale,
                             dtype, stream);
  RSC_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)dout) & 15) == 0, "rsc_wmsa_bwd: qkv / dout must be 16-byte aligned (TMA)");
  WinGeom g(B, H, W, wtm::WS, shift);
  CUtensorMap m7, m4, m3, d7, d4, d3;
  RSC_CHECK_ARG(wtm::window_maps(qkv, B, H, W, 3 * C, &m7, &m4, &m3) && wtm::window_maps(dout, B, H, W, C, &d7, &d4, &d3),
                "rsc_wmsa_bwd: cuTensorMapEncodeTiled failed");
  const int num_items = B * g.nWh * g.nWw * heads;
  auto kern = wtm::wmsa_bwd_tma_kernel;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, wtm::B_TOTAL);
  int grid = (kNumSMs * 4) / heads * heads;
  if (grid > num_items) grid = num_items;
  kern<<<grid, wtm::THREADS, wtm::B_TOTAL, (cudaStream_t)stream>>>(m7, m4, m3, d7, d4, d3, qkv_bias, bias_table,
                                                                   (__nv_bfloat16 *)dqkv, dbias_table, dqkv_bias, g, C, heads,
                                                                   scale, num_items);
  RSC_CHECK_LAUNCH("rsc_wmsa_bwd");
  return RSC_OK;
}
