// Decoder attention: the softmax(Q K^T / sqrt(d) + mask) V core of
// nn.MultiheadAttention as the reference's transformer bricks call it
//   * DINO decoder self-attention: ~1100 queries x ~1100 keys, one constant boolean denoising mask (Lq, Lk)
//     (models/multi/bbox_head/dino_head.py -> mmcv MultiheadAttention, query_denoising.py:167-190 builds the mask)
//   * Mask2Former-style seg decoder: 100 queries x (100^2 | 50^2 | 25^2) keys, mask = sigmoid(resize(mask_pred)) < 0.5 per
//     (image, query, key), "all-masked row -> unmask"  (models/multi/seg_head/mask2former_head.py:111-139,174-197)
// head dim 32 (256 / 8), bf16 in, fp32 accumulate.  The mask never exists as more than ONE BIT per (image | 1, query, key):
//   rsc_m2f_mask_bits   bilinear resize of mask_pred + threshold + the all-masked rule -> bit words (the reference
//                       materialises resized logits, an 8x head-repeated boolean tensor and the SDPA additive bias)
//   rsc_pack_mask_bits  any boolean mask -> bit words (the constant DINO mask: once per ground-truth signature)
//   rsc_attn_fwd        flash-style forward; keys split over CTAs when there are few queries (100 queries x 10^4 keys
//                       would otherwise be 32 CTAs on 148 SMs), partial (m, l, O) merged by a second tiny kernel
//   rsc_attn_bwd        key-stationary backward on TRANSPOSED score tiles (S^T = K Q^T): P^T and dS^T come out of the MMA
//                       in exactly the register layout the dV += P^T dO and dK += dS^T Q MMAs want as their A operand,
//                       so only dS takes one trip through shared memory (for dQ += dS K, red.global.add.v2.f32)
// These are launch-latency-sized problems (<= 2 GFLOP, 16..150 CTAs): warp-level mma.sync m16n8k16 with register-resident
// accumulators -- no TMEM allocation / descriptor set-up per launch; BASELINE north_star keeps tcgen05 for the window
// attention and the GEMMs.
#include "common.cuh"

namespace rsc {
namespace attn {

constexpr int D = 32;         // head dim
constexpr int ROWB = 80;      // bytes per shared-memory tile row: 64 B of data + 16 B pad (ldmatrix conflict-free)
constexpr int TILE = 64 * ROWB;
constexpr int DSB = 144;      // bytes per row of the dS^T tile (64 bf16 + pad)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(uint32_t dst, const void *src, bool pred) {
  const int sz = pred ? 16 : 0;      // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}

// 64 rows x 32 bf16 of a (L, B, H*32)-strided tensor -> padded tile; rows >= L are zero filled.  128 threads.
__device__ __forceinline__ void load_tile(unsigned char *tile, const __nv_bfloat16 *base, int64_t sl, int row0, int L, int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = tid + i * 128, r = idx >> 2, c = idx & 3;
    const bool ok = row0 + r < L;
    cp16(smem_u32(tile + r * ROWB + c * 16), base + (ok ? (int64_t)(row0 + r) * sl + c * 8 : 0), ok);
  }
}

struct Params {
  const __nv_bfloat16 *q, *k, *v, *dout;
  __nv_bfloat16 *out, *dk, *dv;
  float *lse, *part_o, *part_ml, *dq32;
  const float *delta;
  const uint32_t *mask;
  int B, H, Lq, Lk, nsplit, nw;
  int64_t q_sl, q_sb, k_sl, k_sb, v_sl, v_sb, o_sl, o_sb, dk_sl, dk_sb, dv_sl, dv_sb, mask_sb;
  float scale, scale_log2;
};

// =====================================================================================================================
// forward: CTA = 64 queries (4 warps x 16 rows) x one key split of one (image, head)
// =====================================================================================================================
__global__ void __launch_bounds__(128) attn_fwd_kernel(const Params p) {
  __shared__ __align__(16) unsigned char Ks[2][TILE], Vs[2][TILE];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H, split = blockIdx.z;
  const int r0 = blockIdx.x * 64 + warp * 16 + g, r1 = r0 + 8;
  const __nv_bfloat16 *qb = p.q + b * p.q_sb + h * D, *kb_ = p.k + b * p.k_sb + h * D, *vb = p.v + b * p.v_sb + h * D;

  uint32_t qa[2][4];
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    const int c = kk * 16 + 2 * t;
    qa[kk][0] = r0 < p.Lq ? *reinterpret_cast<const uint32_t *>(qb + (int64_t)r0 * p.q_sl + c) : 0u;
    qa[kk][1] = r1 < p.Lq ? *reinterpret_cast<const uint32_t *>(qb + (int64_t)r1 * p.q_sl + c) : 0u;
    qa[kk][2] = r0 < p.Lq ? *reinterpret_cast<const uint32_t *>(qb + (int64_t)r0 * p.q_sl + c + 8) : 0u;
    qa[kk][3] = r1 < p.Lq ? *reinterpret_cast<const uint32_t *>(qb + (int64_t)r1 * p.q_sl + c + 8) : 0u;
  }
  const int tiles = (p.Lk + 63) / 64, tps = (tiles + p.nsplit - 1) / p.nsplit;
  const int kt0 = split * tps, kt1 = min(kt0 + tps, tiles);
  const uint32_t *m0p = nullptr, *m1p = nullptr;
  if (p.mask) {
    const uint32_t *mb = p.mask + b * p.mask_sb;
    if (r0 < p.Lq) m0p = mb + (int64_t)r0 * p.nw;
    if (r1 < p.Lq) m1p = mb + (int64_t)r1 * p.nw;
  }
  float mx[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f}, o[4][4];
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[n][e] = 0.f;

  if (kt0 < kt1) {
    load_tile(Ks[0], kb_, p.k_sl, kt0 * 64, p.Lk, tid);
    load_tile(Vs[0], vb, p.v_sl, kt0 * 64, p.Lk, tid);
  }
  cp_commit();
  for (int kt = kt0; kt < kt1; ++kt) {
    const int buf = (kt - kt0) & 1;
    cp_wait_all();
    __syncthreads();
    if (kt + 1 < kt1) {
      load_tile(Ks[buf ^ 1], kb_, p.k_sl, (kt + 1) * 64, p.Lk, tid);
      load_tile(Vs[buf ^ 1], vb, p.v_sl, (kt + 1) * 64, p.Lk, tid);
    }
    cp_commit();
    const int kbase = kt * 64;
    uint32_t w[2][2] = {{0u, 0u}, {0u, 0u}};      // mask words [row][key half]
    if (m0p) {
      w[0][0] = __ldg(m0p + kbase / 32);
      if (kbase / 32 + 1 < p.nw) w[0][1] = __ldg(m0p + kbase / 32 + 1);
    }
    if (m1p) {
      w[1][0] = __ldg(m1p + kbase / 32);
      if (kbase / 32 + 1 < p.nw) w[1][1] = __ldg(m1p + kbase / 32 + 1);
    }
    // S = Q K^T
    float s[8][4];
    const uint32_t kaddr = smem_u32(Ks[buf]) + (lane & 7) * ROWB + (lane >> 3) * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t b0, b1, b2, b3;
      ldsm4(kaddr + j * 8 * ROWB, b0, b1, b2, b3);
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      mma(s[j], qa[0], b0, b1);
      mma(s[j], qa[1], b2, b3);
    }
    // scale (log2 domain) + mask + running max
    float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = j * 8 + 2 * t + (e & 1), r = e >> 1;
        const bool off = kbase + c >= p.Lk || ((w[r][c >> 5] >> (c & 31)) & 1u);
        s[j][e] = off ? -INFINITY : s[j][e] * p.scale_log2;
        tmax[r] = fmaxf(tmax[r], s[j][e]);
      }
    float base[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
      const float mnew = fmaxf(mx[r], tmax[r]);
      base[r] = mnew == -INFINITY ? 0.f : mnew;
      const float alpha = exp2f(mx[r] - base[r]);
      mx[r] = mnew;
      l[r] *= alpha;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        o[n][2 * r] *= alpha;
        o[n][2 * r + 1] *= alpha;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s[j][e] = exp2f(s[j][e] - base[e >> 1]);
        l[e >> 1] += s[j][e];
      }
    // O += P V
    const uint32_t vaddr = smem_u32(Vs[buf]) + ((lane & 7) + ((lane >> 3) & 1) * 8) * ROWB + ((lane >> 4) & 1) * 16;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4] = {pack2(s[2 * kk][0], s[2 * kk][1]), pack2(s[2 * kk][2], s[2 * kk][3]),
                       pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm4t(vaddr + kk * 16 * ROWB + np * 32, b0, b1, b2, b3);
        mma(o[2 * np], a, b0, b1);
        mma(o[2 * np + 1], a, b2, b3);
      }
    }
  }
  cp_wait_all();
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
  const int rows[2] = {r0, r1};
  if (p.nsplit == 1) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (rows[r] >= p.Lq) continue;
      const float inv = l[r] > 0.f ? 1.f / l[r] : 0.f;
      __nv_bfloat16 *op = p.out + (int64_t)rows[r] * p.o_sl + b * p.o_sb + h * D + 2 * t;
#pragma unroll
      for (int n = 0; n < 4; ++n) *reinterpret_cast<uint32_t *>(op + n * 8) = pack2(o[n][2 * r] * inv, o[n][2 * r + 1] * inv);
      if (t == 0) p.lse[(int64_t)bh * p.Lq + rows[r]] = l[r] > 0.f ? mx[r] + log2f(l[r]) : INFINITY;
    }
  } else {
    const int64_t slot = (int64_t)split * p.B * p.H + bh;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (rows[r] >= p.Lq) continue;
      float *op = p.part_o + (slot * p.Lq + rows[r]) * D + 2 * t;
#pragma unroll
      for (int n = 0; n < 4; ++n) *reinterpret_cast<float2 *>(op + n * 8) = make_float2(o[n][2 * r], o[n][2 * r + 1]);
      if (t == 0) *reinterpret_cast<float2 *>(p.part_ml + (slot * p.Lq + rows[r]) * 2) = make_float2(mx[r], l[r]);
    }
  }
}

// merge of the key splits: warp = one (image, head, query) row, lane = channel
__global__ void __launch_bounds__(256) attn_combine_kernel(const Params p) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t nrow = (int64_t)p.B * p.H * p.Lq;
  if (row >= nrow) return;
  const int bh = (int)(row / p.Lq), r = (int)(row % p.Lq), b = bh / p.H, h = bh % p.H;
  float m = -INFINITY;
  for (int s = 0; s < p.nsplit; ++s) m = fmaxf(m, p.part_ml[((int64_t)s * p.B * p.H * p.Lq + row) * 2]);
  const float base = m == -INFINITY ? 0.f : m;
  float l = 0.f, acc = 0.f;
  for (int s = 0; s < p.nsplit; ++s) {
    const int64_t i = (int64_t)s * p.B * p.H * p.Lq + row;
    const float2 ml = *reinterpret_cast<const float2 *>(p.part_ml + i * 2);
    const float wgt = exp2f(ml.x - base);
    l += wgt * ml.y;
    acc += wgt * p.part_o[i * D + lane];
  }
  p.out[(int64_t)r * p.o_sl + b * p.o_sb + h * D + lane] = __float2bfloat16_rn(l > 0.f ? acc / l : 0.f);
  if (lane == 0) p.lse[row] = l > 0.f ? m + log2f(l) : INFINITY;
}

// delta[bh, r] = sum_d dO * O   (the softmax-backward row term)
__global__ void __launch_bounds__(256) attn_delta_kernel(const Params p, float *delta) {
  const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (row >= (int64_t)p.B * p.H * p.Lq) return;
  const int bh = (int)(row / p.Lq), r = (int)(row % p.Lq), b = bh / p.H, h = bh % p.H;
  const int64_t off = (int64_t)r * p.o_sl + b * p.o_sb + h * D;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p.out + off) + c), d = __ldg(reinterpret_cast<const uint4 *>(p.dout + off) + c);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&av[e]));
      const float2 fd = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&dv[e]));
      acc += fa.x * fd.x + fa.y * fd.y;
    }
  }
  delta[row] = acc;
}

// =====================================================================================================================
// backward: CTA = 64 keys (4 warps x 16 keys) of one (image, head), loop over the query tiles
// =====================================================================================================================
__global__ void __launch_bounds__(128) attn_bwd_kernel(const Params p) {
  __shared__ __align__(16) unsigned char Ks[TILE], Vs[TILE], Qs[2][TILE], Os[2][TILE], dSs[64 * DSB];
  __shared__ float lse_s[2][64], delta_s[2][64];
  __shared__ uint32_t msk_s[2][128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int kbase = blockIdx.x * 64;
  const __nv_bfloat16 *qb = p.q + b * p.q_sb + h * D, *dob = p.dout + b * p.o_sb + h * D;
  const uint32_t *mb = p.mask ? p.mask + b * p.mask_sb : nullptr;
  const int qtiles = (p.Lq + 63) / 64;

  auto stage = [&](int qt, int buf) {       // queries of tile qt -> shared memory buffer `buf`
    load_tile(Qs[buf], qb, p.q_sl, qt * 64, p.Lq, tid);
    load_tile(Os[buf], dob, p.o_sl, qt * 64, p.Lq, tid);
    const int r = qt * 64 + (tid >> 1);
    uint32_t word = 0u;
    if (mb && r < p.Lq && kbase / 32 + (tid & 1) < p.nw) word = __ldg(mb + (int64_t)r * p.nw + kbase / 32 + (tid & 1));
    msk_s[buf][tid] = word;
    if (tid < 64) {
      const int rr = qt * 64 + tid;
      lse_s[buf][tid] = rr < p.Lq ? p.lse[(int64_t)bh * p.Lq + rr] : INFINITY;      // P = exp2(s - inf) = 0 on the tail rows
      delta_s[buf][tid] = rr < p.Lq ? p.delta[(int64_t)bh * p.Lq + rr] : 0.f;
    }
  };
  load_tile(Ks, p.k + b * p.k_sb + h * D, p.k_sl, kbase, p.Lk, tid);
  load_tile(Vs, p.v + b * p.v_sb + h * D, p.v_sl, kbase, p.Lk, tid);
  stage(0, 0);
  cp_commit();
  cp_wait_all();
  __syncthreads();
  // this warp's 16 keys as A operands (K for S^T, V for dP^T)
  uint32_t ka[2][4], va[2][4];
  {
    const uint32_t off = (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ROWB + ((lane >> 4) & 1) * 16;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      ldsm4(smem_u32(Ks) + off + kk * 32, ka[kk][0], ka[kk][1], ka[kk][2], ka[kk][3]);
      ldsm4(smem_u32(Vs) + off + kk * 32, va[kk][0], va[kk][1], va[kk][2], va[kk][3]);
    }
  }
  float dk[4][4], dv[4][4];
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) dk[n][e] = dv[n][e] = 0.f;
  const int key_in_tile[2] = {warp * 16 + g, warp * 16 + g + 8};
  const bool key_ok[2] = {kbase + key_in_tile[0] < p.Lk, kbase + key_in_tile[1] < p.Lk};

  for (int qt = 0; qt < qtiles; ++qt) {
    const int buf = qt & 1;
    if (qt > 0) {
      cp_wait_all();
      __syncthreads();
    }
    if (qt + 1 < qtiles) stage(qt + 1, buf ^ 1);
    cp_commit();
    // S^T = K Q^T and dP^T = V dO^T   (16 keys x 64 queries per warp)
    float st[8][4], dpt[8][4];
    const uint32_t qaddr = smem_u32(Qs[buf]) + (lane & 7) * ROWB + (lane >> 3) * 16;
    const uint32_t oaddr = smem_u32(Os[buf]) + (lane & 7) * ROWB + (lane >> 3) * 16;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      uint32_t b0, b1, b2, b3;
      ldsm4(qaddr + n * 8 * ROWB, b0, b1, b2, b3);
      st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f;
      mma(st[n], ka[0], b0, b1);
      mma(st[n], ka[1], b2, b3);
      ldsm4(oaddr + n * 8 * ROWB, b0, b1, b2, b3);
      dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f;
      mma(dpt[n], va[0], b0, b1);
      mma(dpt[n], va[1], b2, b3);
    }
    // P^T = exp2(S^T c - lse_q), dS^T = P^T (dP^T - delta_q) / sqrt(d)
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qc = n * 8 + 2 * t + (e & 1), r = e >> 1, kin = key_in_tile[r];
        const bool off = !key_ok[r] || ((msk_s[buf][qc * 2 + (kin >> 5)] >> (kin & 31)) & 1u);
        const float pv = off ? 0.f : exp2f(st[n][e] * p.scale_log2 - lse_s[buf][qc]);
        st[n][e] = pv;
        dpt[n][e] = pv * (dpt[n][e] - delta_s[buf][qc]) * p.scale;
      }
    // dV += P^T dO, dK += dS^T Q   (A operands straight from the accumulator registers)
    const uint32_t toff = ((lane & 7) + ((lane >> 3) & 1) * 8) * ROWB + ((lane >> 4) & 1) * 16;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4] = {pack2(st[2 * kk][0], st[2 * kk][1]), pack2(st[2 * kk][2], st[2 * kk][3]),
                        pack2(st[2 * kk + 1][0], st[2 * kk + 1][1]), pack2(st[2 * kk + 1][2], st[2 * kk + 1][3])};
      uint32_t sa[4] = {pack2(dpt[2 * kk][0], dpt[2 * kk][1]), pack2(dpt[2 * kk][2], dpt[2 * kk][3]),
                        pack2(dpt[2 * kk + 1][0], dpt[2 * kk + 1][1]), pack2(dpt[2 * kk + 1][2], dpt[2 * kk + 1][3])};
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm4t(smem_u32(Os[buf]) + toff + kk * 16 * ROWB + np * 32, b0, b1, b2, b3);
        mma(dv[2 * np], pa, b0, b1);
        mma(dv[2 * np + 1], pa, b2, b3);
        ldsm4t(smem_u32(Qs[buf]) + toff + kk * 16 * ROWB + np * 32, b0, b1, b2, b3);
        mma(dk[2 * np], sa, b0, b1);
        mma(dk[2 * np + 1], sa, b2, b3);
      }
    }
    // dS^T -> shared memory [key][query]
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      *reinterpret_cast<uint32_t *>(dSs + key_in_tile[0] * DSB + (n * 8 + 2 * t) * 2) = pack2(dpt[n][0], dpt[n][1]);
      *reinterpret_cast<uint32_t *>(dSs + key_in_tile[1] * DSB + (n * 8 + 2 * t) * 2) = pack2(dpt[n][2], dpt[n][3]);
    }
    __syncthreads();
    // dQ[16 queries of this warp] = dS K over the CTA's 64 keys
    float dq[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      ldsm4t(smem_u32(dSs) + (kk * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * DSB + (warp * 16 + ((lane >> 3) & 1) * 8) * 2, a[0], a[1],
             a[2], a[3]);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm4t(smem_u32(Ks) + toff + kk * 16 * ROWB + np * 32, b0, b1, b2, b3);
        mma(dq[2 * np], a, b0, b1);
        mma(dq[2 * np + 1], a, b2, b3);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = qt * 64 + warp * 16 + g + r * 8;
      if (row >= p.Lq) continue;
      float *dst = p.dq32 + ((int64_t)row * p.B + b) * p.H * D + h * D + 2 * t;
#pragma unroll
      for (int n = 0; n < 4; ++n)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst + n * 8), "f"(dq[n][2 * r]), "f"(dq[n][2 * r + 1]) : "memory");
    }
  }
  cp_wait_all();
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (!key_ok[r]) continue;
    const int key = kbase + key_in_tile[r];
    __nv_bfloat16 *kp = p.dk + (int64_t)key * p.dk_sl + b * p.dk_sb + h * D + 2 * t;
    __nv_bfloat16 *vp = p.dv + (int64_t)key * p.dv_sl + b * p.dv_sb + h * D + 2 * t;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      *reinterpret_cast<uint32_t *>(kp + n * 8) = pack2(dk[n][2 * r], dk[n][2 * r + 1]);
      *reinterpret_cast<uint32_t *>(vp + n * 8) = pack2(dv[n][2 * r], dv[n][2 * r + 1]);
    }
  }
}

// =====================================================================================================================
// masks as bit words: bit (k & 31) of word [row][k >> 5] set = key k is NOT attended by that query row
// =====================================================================================================================
__device__ __forceinline__ void bil_src(int o, float scale, int in, int &i0, int &i1, float &lam) {
  float s = scale * (o + 0.5f) - 0.5f;       // (same source-coordinate rule as rsc_bilinear_fwd / F.interpolate)
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 < in - 1 ? i0 + 1 : i0;
  lam = s - i0;
}

// CTA = one (image, query) row of mask_pred (Hi x Wi logits) -> nw words at the (Ho x Wo) key grid
template <typename T>
__global__ void __launch_bounds__(256) m2f_mask_bits_kernel(const T *__restrict__ mask_pred, uint32_t *__restrict__ bits, int Hi, int Wi,
                                                            int Ho, int Wo, int nw, float sh, float sw) {
  const int64_t row = blockIdx.x;
  const T *src = mask_pred + row * Hi * Wi;
  uint32_t *dst = bits + row * nw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, Lk = Ho * Wo;
  int open = 0;
  for (int w = warp; w < nw; w += 8) {
    const int key = w * 32 + lane;
    bool masked = false;
    if (key < Lk) {
      int y0, y1, x0, x1;
      float ly, lx;
      bil_src(key / Wo, sh, Hi, y0, y1, ly);
      bil_src(key % Wo, sw, Wi, x0, x1, lx);
      const float v00 = to_f<T>(src[y0 * Wi + x0]), v01 = to_f<T>(src[y0 * Wi + x1]);
      const float v10 = to_f<T>(src[y1 * Wi + x0]), v11 = to_f<T>(src[y1 * Wi + x1]);
      const float val = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
      masked = val < 0.f;                    // sigmoid(val) < 0.5
      open |= !masked;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, masked);
    if (lane == 0) dst[w] = word;
  }
  if (!__syncthreads_or(open))               // every key masked: the reference un-masks the whole row
    for (int w = threadIdx.x; w < nw; w += 256) dst[w] = 0u;
}

__global__ void __launch_bounds__(256) pack_mask_bits_kernel(const unsigned char *__restrict__ m, uint32_t *__restrict__ bits, int64_t rows,
                                                             int Lk, int nw) {
  const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= rows * nw) return;
  const int64_t row = wid / nw;
  const int key = (int)(wid % nw) * 32 + (threadIdx.x & 31);
  const uint32_t word = __ballot_sync(0xffffffffu, key < Lk && m[row * Lk + key] != 0);
  if ((threadIdx.x & 31) == 0) bits[wid] = word;
}

static int check_common(const char *fn, int B, int H, int Lq, int Lk, int head_dim, const int64_t *strides, int n) {
  RSC_CHECK_ARG(B > 0 && H > 0 && Lq > 0 && Lk > 0, "%s: bad shape (B=%d,H=%d,Lq=%d,Lk=%d)", fn, B, H, Lq, Lk);
  RSC_CHECK_ARG(head_dim == D, "%s: head_dim %d is not supported (32 only)", fn, head_dim);
  for (int i = 0; i < n; ++i) RSC_CHECK_ARG(strides[i] % 8 == 0, "%s: strides must be multiples of 8 elements (16-byte rows)", fn);
  return RSC_OK;
}

}  // namespace attn
}  // namespace rsc

using namespace rsc;
using namespace rsc::attn;

extern "C" int rsc_attn_nsplit(int B, int H, int Lq, int Lk) {
  const int ctas = ((Lq + 63) / 64) * B * H, tiles = (Lk + 63) / 64;
  int ns = (2 * kNumSMs + ctas - 1) / ctas;
  ns = ns < 1 ? 1 : ns;
  const int cap = (tiles + 3) / 4;            // at least 4 key tiles per split
  ns = ns > cap ? cap : ns;
  return ns < 1 ? 1 : (ns > 32 ? 32 : ns);
}

extern "C" int rsc_attn_fwd(const void *q, const void *k, const void *v, const void *mask_bits, void *out, float *lse, float *ws, int B,
                            int H, int Lq, int Lk, int head_dim, int64_t q_sl, int64_t q_sb, int64_t k_sl, int64_t k_sb,
                            int64_t v_sl, int64_t v_sb, int64_t o_sl, int64_t o_sb, int64_t mask_sb, int nsplit, float scale,
                            void *stream) {
  const int64_t st[8] = {q_sl, q_sb, k_sl, k_sb, v_sl, v_sb, o_sl, o_sb};
  if (int e = check_common("rsc_attn_fwd", B, H, Lq, Lk, head_dim, st, 8)) return e;
  RSC_CHECK_ARG(q && k && v && out && lse, "rsc_attn_fwd: null pointer");
  RSC_CHECK_ARG(nsplit >= 1 && nsplit <= 32 && (nsplit == 1 || ws), "rsc_attn_fwd: bad nsplit %d / missing workspace", nsplit);
  Params p = {};
  p.q = (const __nv_bfloat16 *)q, p.k = (const __nv_bfloat16 *)k, p.v = (const __nv_bfloat16 *)v, p.out = (__nv_bfloat16 *)out;
  p.lse = lse, p.mask = (const uint32_t *)mask_bits;
  p.B = B, p.H = H, p.Lq = Lq, p.Lk = Lk, p.nsplit = nsplit, p.nw = (Lk + 31) / 32;
  p.q_sl = q_sl, p.q_sb = q_sb, p.k_sl = k_sl, p.k_sb = k_sb, p.v_sl = v_sl, p.v_sb = v_sb, p.o_sl = o_sl, p.o_sb = o_sb;
  p.mask_sb = mask_sb, p.scale = scale, p.scale_log2 = scale * 1.4426950408889634f;
  if (nsplit > 1) {
    p.part_o = ws;
    p.part_ml = ws + (int64_t)nsplit * B * H * Lq * D;
  }
  attn_fwd_kernel<<<dim3((Lq + 63) / 64, B * H, nsplit), 128, 0, (cudaStream_t)stream>>>(p);
  RSC_CHECK_LAUNCH("rsc_attn_fwd");
  if (nsplit > 1) {
    attn_combine_kernel<<<(unsigned)(((int64_t)B * H * Lq + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
    RSC_CHECK_LAUNCH("rsc_attn_fwd(combine)");
  }
  return RSC_OK;
}

extern "C" int rsc_attn_bwd(const void *q, const void *k, const void *v, const void *mask_bits, const void *out, const void *dout,
                            const float *lse, float *delta_ws, float *dq32, void *dk, void *dv, int B, int H, int Lq, int Lk,
                            int head_dim, int64_t q_sl, int64_t q_sb, int64_t k_sl, int64_t k_sb, int64_t v_sl, int64_t v_sb,
                            int64_t o_sl, int64_t o_sb, int64_t dk_sl, int64_t dk_sb, int64_t dv_sl, int64_t dv_sb, int64_t mask_sb,
                            float scale, void *stream) {
  const int64_t st[12] = {q_sl, q_sb, k_sl, k_sb, v_sl, v_sb, o_sl, o_sb, dk_sl, dk_sb, dv_sl, dv_sb};
  if (int e = check_common("rsc_attn_bwd", B, H, Lq, Lk, head_dim, st, 12)) return e;
  RSC_CHECK_ARG(q && k && v && out && dout && lse && delta_ws && dq32 && dk && dv, "rsc_attn_bwd: null pointer");
  Params p = {};
  p.q = (const __nv_bfloat16 *)q, p.k = (const __nv_bfloat16 *)k, p.v = (const __nv_bfloat16 *)v;
  p.out = (__nv_bfloat16 *)const_cast<void *>(out), p.dout = (const __nv_bfloat16 *)dout;
  p.lse = const_cast<float *>(lse), p.delta = delta_ws, p.dq32 = dq32, p.dk = (__nv_bfloat16 *)dk, p.dv = (__nv_bfloat16 *)dv;
  p.mask = (const uint32_t *)mask_bits;
  p.B = B, p.H = H, p.Lq = Lq, p.Lk = Lk, p.nsplit = 1, p.nw = (Lk + 31) / 32;
  p.q_sl = q_sl, p.q_sb = q_sb, p.k_sl = k_sl, p.k_sb = k_sb, p.v_sl = v_sl, p.v_sb = v_sb, p.o_sl = o_sl, p.o_sb = o_sb;
  p.dk_sl = dk_sl, p.dk_sb = dk_sb, p.dv_sl = dv_sl, p.dv_sb = dv_sb;
  p.mask_sb = mask_sb, p.scale = scale, p.scale_log2 = scale * 1.4426950408889634f;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(dq32, 0, (size_t)Lq * B * H * D * sizeof(float), s) != cudaSuccess) {
    set_error("rsc_attn_bwd: memset failed");
    return RSC_ERR_CUDA;
  }
  attn_delta_kernel<<<(unsigned)(((int64_t)B * H * Lq + 255) / 256), 256, 0, s>>>(p, delta_ws);
  RSC_CHECK_LAUNCH("rsc_attn_bwd(delta)");
  attn_bwd_kernel<<<dim3((Lk + 63) / 64, B * H), 128, 0, s>>>(p);
  RSC_CHECK_LAUNCH("rsc_attn_bwd");
  return RSC_OK;
}

extern "C" int rsc_m2f_mask_bits(const void *mask_pred, void *bits, int rows, int Hi, int Wi, int Ho, int Wo, int dtype, void *stream) {
  RSC_CHECK_ARG(rows > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "rsc_m2f_mask_bits: bad shape");
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "rsc_m2f_mask_bits: bad dtype %d", dtype);
  RSC_CHECK_ARG(mask_pred && bits, "rsc_m2f_mask_bits: null pointer");
  const int nw = (Ho * Wo + 31) / 32;
  const float sh = (float)Hi / Ho, sw = (float)Wi / Wo;
  if (dtype == RSC_F32)
    m2f_mask_bits_kernel<float><<<rows, 256, 0, (cudaStream_t)stream>>>((const float *)mask_pred, (uint32_t *)bits, Hi, Wi, Ho, Wo, nw, sh, sw);
  else
    m2f_mask_bits_kernel<__nv_bfloat16>
        <<<rows, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)mask_pred, (uint32_t *)bits, Hi, Wi, Ho, Wo, nw, sh, sw);
  RSC_CHECK_LAUNCH("rsc_m2f_mask_bits");
  return RSC_OK;
}

extern "C" int rsc_pack_mask_bits(const void *mask_u8, void *bits, int64_t rows, int Lk, void *stream) {
  RSC_CHECK_ARG(rows > 0 && Lk > 0, "rsc_pack_mask_bits: bad shape");
  RSC_CHECK_ARG(mask_u8 && bits, "rsc_pack_mask_bits: null pointer");
  const int nw = (Lk + 31) / 32;
  pack_mask_bits_kernel<<<(unsigned)((rows * nw + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const unsigned char *)mask_u8, (uint32_t *)bits,
                                                                                        rows, Lk, nw);
  RSC_CHECK_LAUNCH("rsc_pack_mask_bits");
  return RSC_OK;
}
