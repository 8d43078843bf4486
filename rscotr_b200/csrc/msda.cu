// Multi-scale deformable attention, forward and backward.
// Drop-in for mmcv-full 1.6.1 ext_module.ms_deform_attn_{forward,backward}.
//
// Mapping: one warp = one query x 4 consecutive heads; an 8-lane group owns one
// head, each lane 4 of its 32 channels (16 B fp32 / 8 B bf16 vector loads, a
// corner read is one fully used 128 B / 64 B segment per group).  The warp's
// sampling locations (4*LP*2 floats) and attention weights (4*LP floats) are
// contiguous in HBM: they are fetched once with coalesced float4 loads into a
// per-warp smem slab and broadcast from there.  Backward reduces d(loc) and
// d(weight) over the 32 channels with 3 xor-shuffles inside the 8-lane group
// and scatters d(value) with vector red.global.add.f32.
#include "common.cuh"

namespace rsc {

constexpr int MSDA_WARPS = 8;
constexpr int MSDA_MAX_L = 8;
constexpr int MSDA_MAX_LP = 64;

struct Levels {
  int h[MSDA_MAX_L], w[MSDA_MAX_L], start[MSDA_MAX_L];
};

__device__ __forceinline__ void load_levels(Levels &lv, const int64_t *shapes, const int64_t *starts, int L) {
  if (threadIdx.x < L) {
    lv.h[threadIdx.x] = (int)shapes[2 * threadIdx.x];
    lv.w[threadIdx.x] = (int)shapes[2 * threadIdx.x + 1];
    lv.start[threadIdx.x] = (int)starts[threadIdx.x];
  }
}

// item -> (head group, image * Nq + query, image): 32-bit divisions where the item count allows it (the emulated 64-bit
// ones are ~10 % of a warp item's instructions)
__device__ __forceinline__ void split_item(int64_t item, int64_t total, int hgroups, int Nq, int &hg, int64_t &bq, int &b) {
  if (total < ((int64_t)1 << 31)) {
    const unsigned q = (unsigned)item / (unsigned)hgroups;
    hg = (int)((unsigned)item - q * (unsigned)hgroups);
    bq = q;
    b = (int)(q / (unsigned)Nq);
  } else {
    hg = (int)(item % hgroups);
    bq = item / hgroups;
    b = (int)(bq / Nq);
  }
}

__device__ __forceinline__ float group_sum8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ void fma4(float4 &acc, float s, float4 v) {
  acc.x = fmaf(s, v.x, acc.x), acc.y = fmaf(s, v.y, acc.y), acc.z = fmaf(s, v.z, acc.z), acc.w = fmaf(s, v.w, acc.w);
}

template <typename T>
__global__ void __launch_bounds__(MSDA_WARPS * 32)
    msda_fwd_kernel(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ starts,
                    const float *__restrict__ loc, const float *__restrict__ aw, T *__restrict__ out, int B, int Nv,
                    int Nq, int heads, int L, int P) {
  extern __shared__ __align__(16) float smem[];
  __shared__ Levels lv;
  load_levels(lv, shapes, starts, L);
  __syncthreads();
  const int LP = L * P;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *sloc = smem + warp * (12 * LP);  // [4][LP][2]
  float *sw = sloc + 8 * LP;              // [4][LP]
  const int g = lane >> 3, t = lane & 7;
  const int hgroups = heads >> 2;
  const int64_t total = (int64_t)B * Nq * hgroups;
  const int64_t nwarps = (int64_t)gridDim.x * MSDA_WARPS;
  const int vstride = heads * 32;
  for (int64_t item = (int64_t)blockIdx.x * MSDA_WARPS + warp; item < total; item += nwarps) {
    int hg, b;
    int64_t bq;
    split_item(item, total, hgroups, Nq, hg, bq, b);
    const int64_t slab = (bq * heads + hg * 4) * LP;
    const float4 *gl = reinterpret_cast<const float4 *>(loc + slab * 2);
    const float4 *gw = reinterpret_cast<const float4 *>(aw + slab);
    for (int v = lane; v < 2 * LP; v += 32) reinterpret_cast<float4 *>(sloc)[v] = __ldg(gl + v);
    for (int v = lane; v < LP; v += 32) reinterpret_cast<float4 *>(sw)[v] = __ldg(gw + v);
    __syncwarp();
    const int head = hg * 4 + g;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *ml = sloc + g * 2 * LP;
    const float *mw = sw + g * LP;
    for (int l = 0; l < L; ++l) {
      const int Hl = lv.h[l], Wl = lv.w[l];
      const T *vb = value + ((int64_t)b * Nv + lv.start[l]) * vstride + head * 32 + t * 4;
#pragma unroll 4
      for (int p = 0; p < P; ++p) {
        const float2 xy = *reinterpret_cast<const float2 *>(ml + (l * P + p) * 2);
        const float wgt = mw[l * P + p];
        const float h_im = xy.y * Hl - 0.5f, w_im = xy.x * Wl - 0.5f;
        if (h_im > -1.f && w_im > -1.f && h_im < Hl && w_im < Wl) {
          const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
          const float lh = h_im - h_low, lw = w_im - w_low, hh = 1.f - lh, hw = 1.f - lw;
          const bool h0 = h_low >= 0, h1 = h_low + 1 <= Hl - 1, w0 = w_low >= 0, w1 = w_low + 1 <= Wl - 1;
          const T *p00 = vb + ((int64_t)h_low * Wl + w_low) * vstride;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (h0 && w0) fma4(v, hh * hw, load4<T>(p00));
          if (h0 && w1) fma4(v, hh * lw, load4<T>(p00 + vstride));
          if (h1 && w0) fma4(v, lh * hw, load4<T>(p00 + (int64_t)Wl * vstride));
          if (h1 && w1) fma4(v, lh * lw, load4<T>(p00 + (int64_t)(Wl + 1) * vstride));
          fma4(acc, wgt, v);
        }
      }
    }
    store4<T>(out + bq * vstride + head * 32 + t * 4, acc);
    __syncwarp();
  }
}

template <typename T>
__global__ void __launch_bounds__(MSDA_WARPS * 32)
    msda_bwd_kernel(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ starts,
                    const float *__restrict__ loc, const float *__restrict__ aw, const T *__restrict__ gout,
                    float *__restrict__ gvalue, float *__restrict__ gloc, float *__restrict__ gaw, int B, int Nv,
                    int Nq, int heads, int L, int P) {
  extern __shared__ __align__(16) float smem[];
  __shared__ Levels lv;
  load_levels(lv, shapes, starts, L);
  __syncthreads();
  const int LP = L * P;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *sloc = smem + warp * (24 * LP);  // [4][LP][2]
  float *sw = sloc + 8 * LP;              // [4][LP]
  float *sgl = sw + 4 * LP;               // [4][LP][2] grad loc
  float *sgw = sgl + 8 * LP;              // [4][LP]    grad weight
  const int g = lane >> 3, t = lane & 7;
  const int hgroups = heads >> 2;
  const int64_t total = (int64_t)B * Nq * hgroups;
  const int64_t nwarps = (int64_t)gridDim.x * MSDA_WARPS;
  const int vstride = heads * 32;
  for (int64_t item = (int64_t)blockIdx.x * MSDA_WARPS + warp; item < total; item += nwarps) {
    int hg, b;
    int64_t bq;
    split_item(item, total, hgroups, Nq, hg, bq, b);
    const int64_t slab = (bq * heads + hg * 4) * LP;
    const float4 *gl = reinterpret_cast<const float4 *>(loc + slab * 2);
    const float4 *gw = reinterpret_cast<const float4 *>(aw + slab);
    for (int v = lane; v < 2 * LP; v += 32) reinterpret_cast<float4 *>(sloc)[v] = __ldg(gl + v);
    for (int v = lane; v < LP; v += 32) reinterpret_cast<float4 *>(sw)[v] = __ldg(gw + v);
    __syncwarp();
    const int head = hg * 4 + g;
    const float4 go = load4<T>(gout + bq * vstride + head * 32 + t * 4);
    const float *ml = sloc + g * 2 * LP;
    const float *mw = sw + g * LP;
    for (int l = 0; l < L; ++l) {
      const int Hl = lv.h[l], Wl = lv.w[l];
      const int64_t voff = ((int64_t)b * Nv + lv.start[l]) * vstride + head * 32 + t * 4;
      for (int p = 0; p < P; ++p) {
        const float2 xy = *reinterpret_cast<const float2 *>(ml + (l * P + p) * 2);
        const float wgt = mw[l * P + p];
        const float h_im = xy.y * Hl - 0.5f, w_im = xy.x * Wl - 0.5f;
        float gx = 0.f, gy = 0.f, ga = 0.f;
        if (h_im > -1.f && w_im > -1.f && h_im < Hl && w_im < Wl) {  // uniform within the 8-lane group
          const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
          const float lh = h_im - h_low, lw = w_im - w_low, hh = 1.f - lh, hw = 1.f - lw;
          const bool h0 = h_low >= 0, h1 = h_low + 1 <= Hl - 1, w0 = w_low >= 0, w1 = w_low + 1 <= Wl - 1;
          const int64_t o00 = voff + ((int64_t)h_low * Wl + w_low) * vstride;
          const float4 tw = make_float4(go.x * wgt, go.y * wgt, go.z * wgt, go.w * wgt);  // top_grad * attn_weight
          float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
          float gh_w = 0.f, gw_w = 0.f;  // sum_c tw_c * d val_c / d{h,w}
          auto corner = [&](bool ok, int64_t off, float wc, float dh, float dw) {
            if (!ok) return;
            const float4 v = load4<T>(value + off);
            const float tv = dot4(tw, v);
            gh_w = fmaf(dh, tv, gh_w);
            gw_w = fmaf(dw, tv, gw_w);
            fma4(val, wc, v);
            atomicAdd(reinterpret_cast<float4 *>(gvalue + off), make_float4(wc * tw.x, wc * tw.y, wc * tw.z, wc * tw.w));
          };
          corner(h0 && w0, o00, hh * hw, -hw, -hh);
          corner(h0 && w1, o00 + vstride, hh * lw, -lw, hh);
          corner(h1 && w0, o00 + (int64_t)Wl * vstride, lh * hw, hw, -lh);
          corner(h1 && w1, o00 + (int64_t)(Wl + 1) * vstride, lh * lw, lw, lh);
          gx = Wl * gw_w;
          gy = Hl * gh_w;
          ga = dot4(go, val);
        }
        gx = group_sum8(gx);
        gy = group_sum8(gy);
        ga = group_sum8(ga);
        if (t == 0) {
          sgl[g * 2 * LP + (l * P + p) * 2] = gx;
          sgl[g * 2 * LP + (l * P + p) * 2 + 1] = gy;
          sgw[g * LP + l * P + p] = ga;
        }
      }
    }
    __syncwarp();
    float4 *ogl = reinterpret_cast<float4 *>(gloc + slab * 2);
    float4 *ogw = reinterpret_cast<float4 *>(gaw + slab);
    for (int v = lane; v < 2 * LP; v += 32) ogl[v] = reinterpret_cast<const float4 *>(sgl)[v];
    for (int v = lane; v < LP; v += 32) ogw[v] = reinterpret_cast<const float4 *>(sgw)[v];
    __syncwarp();
  }
}

// =============================================================================================
// Fused variant: the whole tail of mmcv MultiScaleDeformableAttention.forward between the two small Linears and
// the output projection (ops/multi_scale_deform_attn.py: softmax over the L*P attention logits, sampling_locations
// = reference_points + offsets / (W_l, H_l)  [or + offsets / P * ref_wh * 0.5 for 4-d references], then the
// sampling op) in ONE kernel, forward and backward.  The eager chain materialises (B,Nq,heads,L,P,2) fp32
// tensors five times per layer (float(), div, add, softmax, their backward twins); here the raw Linear outputs are
// read once.  Same warp mapping as above (query x 4 heads), restructured to cut instructions:
//   phase 1  one lane per (head, point) pair: softmax weight (16-lane shuffles), location, the 4 corner token
//            offsets and bilinear weights -> a 32-byte entry in shared memory (computed once, not by all 8 lanes);
//   phase 2  8-lane group per head, 4 channels per lane: branch-free gather-accumulate over the 64 entries;
//   phase 3  (backward) one lane per pair again: softmax backward and d(offsets) from the reduced gradients.
// L*P must be 16 (L=4 levels x P=4 points of every reference config).
// =============================================================================================
namespace msf {

constexpr int WARPS = 8;
constexpr int LP = 16;
constexpr int GSTRIDE = LP + 1;   // entries per head group (+1: the four groups' LDS.128 hit different banks)

struct __align__(16) Entry {
  int off[4];      // token index (within the image's Nv tokens) of the corners 00, 01, 10, 11
  float w[4];      // fwd: bilinear weight x attention weight | bwd: {lh, lw, attention weight, corner-valid bits}
};

template <typename T>
__device__ __forceinline__ float ldf(const T *p, int64_t i) { return to_f<T>(p[i]); }

// location + corner geometry of one (head, point) pair
struct Geo {
  int off[4];
  float lh, lw;
  unsigned mask;   // bit k: corner k is inside the level (0 for a sample outside the level)
  float sx, sy;    // d(loc)/d(offset) per axis
  int Hl, Wl;
};

__device__ __forceinline__ Geo pair_geo(const Levels &lv, const float *ref, int64_t bq, int L, int P, int R, int lp,
                                        float ox, float oy) {
  Geo g;
  const int l = lp / P;
  g.Hl = lv.h[l], g.Wl = lv.w[l];
  const float *r = ref + (bq * L + l) * R;
  float lx, ly;
  if (R == 2) {
    g.sx = 1.0f / g.Wl, g.sy = 1.0f / g.Hl;
    lx = r[0] + ox / g.Wl, ly = r[1] + oy / g.Hl;
  } else {
    g.sx = r[2] * 0.5f / P, g.sy = r[3] * 0.5f / P;
    lx = r[0] + ox / P * r[2] * 0.5f, ly = r[1] + oy / P * r[3] * 0.5f;
  }
  const float h_im = ly * g.Hl - 0.5f, w_im = lx * g.Wl - 0.5f;
  const int start = lv.start[l];
  g.mask = 0;
  g.lh = g.lw = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) g.off[k] = start;
  if (h_im > -1.f && w_im > -1.f && h_im < g.Hl && w_im < g.Wl) {
    const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
    g.lh = h_im - h_low, g.lw = w_im - w_low;
    const bool h0 = h_low >= 0, h1 = h_low + 1 <= g.Hl - 1, w0 = w_low >= 0, w1 = w_low + 1 <= g.Wl - 1;
    const int base = start + h_low * g.Wl + w_low;
    if (h0 && w0) g.mask |= 1u, g.off[0] = base;
    if (h0 && w1) g.mask |= 2u, g.off[1] = base + 1;
    if (h1 && w0) g.mask |= 4u, g.off[2] = base + g.Wl;
    if (h1 && w1) g.mask |= 8u, g.off[3] = base + g.Wl + 1;
  }
  return g;
}

__device__ __forceinline__ float half_max(float v) {   // over the 16 lanes of a half warp
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, typename TO>
__global__ void __launch_bounds__(WARPS * 32)
    msda_fused_fwd_kernel(const T *__restrict__ value, const int64_t *__restrict__ shapes,
                          const int64_t *__restrict__ starts, const TO *__restrict__ offs, const TO *__restrict__ logits,
                          const float *__restrict__ ref, T *__restrict__ out, int B, int Nv, int Nq, int heads, int L,
                          int P, int R, int64_t ostride, int64_t lstride) {
  __shared__ Entry ent[WARPS][4 * GSTRIDE];
  __shared__ Levels lv;
  load_levels(lv, shapes, starts, L);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 3, t = lane & 7;
  const int hgroups = heads >> 2;
  const int64_t total = (int64_t)B * Nq * hgroups;
  const int64_t nwarps = (int64_t)gridDim.x * WARPS;
  const int vstride = heads * 32;
  Entry *my = ent[warp];
  for (int64_t item = (int64_t)blockIdx.x * WARPS + warp; item < total; item += nwarps) {
    int hg, b;
    int64_t bq;
    split_item(item, total, hgroups, Nq, hg, bq, b);
    // row strides: a query's offsets / logits may be column ranges of one wider matrix (the output of ONE GEMM)
    const int64_t lrow = bq * lstride + hg * 4 * LP, orow = bq * ostride + hg * 8 * LP;
    // ---- phase 1: two (head, point) pairs per lane ----
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int q = lane + 32 * it, hq = q >> 4, lp = q & 15;
      const float lg = ldf<TO>(logits, lrow + q);
      const float ox = ldf<TO>(offs, orow + q * 2), oy = ldf<TO>(offs, orow + q * 2 + 1);
      const float e = __expf(lg - half_max(lg));
      const float aw = e / half_sum(e);
      const Geo ge = pair_geo(lv, ref, bq, L, P, R, lp, ox, oy);
      const float hh = 1.f - ge.lh, hw = 1.f - ge.lw;
      Entry en;
#pragma unroll
      for (int k = 0; k < 4; ++k) en.off[k] = ge.off[k];
      en.w[0] = (ge.mask & 1u) ? hh * hw * aw : 0.f;
      en.w[1] = (ge.mask & 2u) ? hh * ge.lw * aw : 0.f;
      en.w[2] = (ge.mask & 4u) ? ge.lh * hw * aw : 0.f;
      en.w[3] = (ge.mask & 8u) ? ge.lh * ge.lw * aw : 0.f;
      my[hq * GSTRIDE + lp] = en;
    }
    __syncwarp();
    // ---- phase 2: gather-accumulate, branch free ----
    const int head = hg * 4 + g;
    const T *vb = value + (int64_t)b * Nv * vstride + head * 32 + t * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int lp = 0; lp < LP; ++lp) {
      const Entry en = my[g * GSTRIDE + lp];
#pragma unroll
      for (int k = 0; k < 4; ++k) fma4(acc, en.w[k], load4<T>(vb + (int64_t)en.off[k] * vstride));
    }
    store4<T>(out + bq * vstride + head * 32 + t * 4, acc);
    __syncwarp();
  }
}

// four elements of T as loaded (bf16: two packed words -- half the registers of four floats)
template <typename T>
struct RawV;
template <>
struct RawV<float> {
  float4 v;
  __device__ __forceinline__ void load(const float *p) { v = __ldg(reinterpret_cast<const float4 *>(p)); }
  __device__ __forceinline__ float4 get() const { return v; }
};
template <>
struct RawV<__nv_bfloat16> {
  uint2 v;
  __device__ __forceinline__ void load(const __nv_bfloat16 *p) { v = __ldg(reinterpret_cast<const uint2 *>(p)); }
  __device__ __forceinline__ float4 get() const {
    return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                       __uint_as_float(v.y & 0xffff0000u));
  }
};

// BF = branch-free point loop, two points per trip: every corner's value load is issued unconditionally (an invalid
// corner points at the level's first token and carries zero weights), so the eight loads of two points are in flight
// together instead of one point's four behind a divergent `if (mask)`; only the reductions stay predicated.
template <typename T, typename TO, bool BF>
__global__ void __launch_bounds__(WARPS * 32, 4)
    msda_fused_bwd_kernel(const T *__restrict__ value, const int64_t *__restrict__ shapes,
                          const int64_t *__restrict__ starts, const TO *__restrict__ offs, const TO *__restrict__ logits,
                          const float *__restrict__ ref, const T *__restrict__ gout, float *__restrict__ gvalue,
                          TO *__restrict__ goffs, TO *__restrict__ glogits, int B, int Nv, int Nq, int heads, int L, int P,
                          int R, int64_t ostride, int64_t lstride) {
  __shared__ Entry ent[WARPS][4 * GSTRIDE];
  __shared__ float red[WARPS][4 * LP][3];   // per pair: d/d(loc x), d/d(loc y), d/d(attention weight)
  __shared__ Levels lv;
  load_levels(lv, shapes, starts, L);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 3, t = lane & 7;
  const int hgroups = heads >> 2;
  const int64_t total = (int64_t)B * Nq * hgroups;
  const int64_t nwarps = (int64_t)gridDim.x * WARPS;
  const int vstride = heads * 32;
  Entry *my = ent[warp];
  for (int64_t item = (int64_t)blockIdx.x * WARPS + warp; item < total; item += nwarps) {
    int hg, b;
    int64_t bq;
    split_item(item, total, hgroups, Nq, hg, bq, b);
    const int64_t lrow = bq * lstride + hg * 4 * LP, orow = bq * ostride + hg * 8 * LP;
    float aw_[2], sx_[2], sy_[2];
    // ---- phase 1 ----
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int q = lane + 32 * it, hq = q >> 4, lp = q & 15;
      const float lg = ldf<TO>(logits, lrow + q);
      const float ox = ldf<TO>(offs, orow + q * 2), oy = ldf<TO>(offs, orow + q * 2 + 1);
      const float e = __expf(lg - half_max(lg));
      const float aw = e / half_sum(e);
      const Geo ge = pair_geo(lv, ref, bq, L, P, R, lp, ox, oy);
      aw_[it] = aw, sx_[it] = ge.sx * ge.Wl, sy_[it] = ge.sy * ge.Hl;   // d(w_im)/d(offset_x), d(h_im)/d(offset_y)
      Entry en;
#pragma unroll
      for (int k = 0; k < 4; ++k) en.off[k] = ge.off[k];
      en.w[0] = ge.lh, en.w[1] = ge.lw, en.w[2] = aw, en.w[3] = __uint_as_float(ge.mask);
      my[hq * GSTRIDE + lp] = en;
    }
    __syncwarp();
    // ---- phase 2: per point, the 8-lane group reduces over its 32 channels ----
    const int head = hg * 4 + g;
    const int64_t vbase = (int64_t)b * Nv * vstride + head * 32 + t * 4;
    const float4 go = load4<T>(gout + bq * vstride + head * 32 + t * 4);
    if (BF) {
      for (int lp0 = 0; lp0 < LP; lp0 += 2) {
        // the eight value loads of two points first (kept in storage type), then the arithmetic and the reductions:
        // written out by hand because the compiler does not move loads across the atomics
        Entry en[2];
        RawV<T> rv[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          en[u] = my[g * GSTRIDE + lp0 + u];
#pragma unroll
          for (int k = 0; k < 4; ++k) rv[u][k].load(value + vbase + (int64_t)en[u].off[k] * vstride);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float lh = en[u].w[0], lw = en[u].w[1], aw = en[u].w[2], hh = 1.f - lh, hw = 1.f - lw;
          const unsigned mask = __float_as_uint(en[u].w[3]);
          const float4 tw = make_float4(go.x * aw, go.y * aw, go.z * aw, go.w * aw);   // top_grad * attention weight
          const float wc[4] = {hh * hw, hh * lw, lh * hw, lh * lw};
          const float dh[4] = {-hw, -lw, hw, lw}, dw[4] = {-hh, hh, -lh, lh};
          float gh_w = 0.f, gw_w = 0.f;
          float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool on = (mask >> k) & 1u;
            const float4 v = rv[u][k].get();
            const float w = on ? wc[k] : 0.f;
            const float tv = on ? dot4(tw, v) : 0.f;
            gh_w = fmaf(dh[k], tv, gh_w);
            gw_w = fmaf(dw[k], tv, gw_w);
            fma4(val, w, v);
            if (on)
              atomicAdd(reinterpret_cast<float4 *>(gvalue + vbase + (int64_t)en[u].off[k] * vstride),
                        make_float4(w * tw.x, w * tw.y, w * tw.z, w * tw.w));
          }
          float ga = dot4(go, val);
          gw_w = group_sum8(gw_w);
          gh_w = group_sum8(gh_w);
          ga = group_sum8(ga);
          if (t == 0) {
            float *r = red[warp][g * LP + lp0 + u];
            r[0] = gw_w, r[1] = gh_w, r[2] = ga;
          }
        }
      }
    } else
    for (int lp = 0; lp < LP; ++lp) {
      const Entry en = my[g * GSTRIDE + lp];
      const float lh = en.w[0], lw = en.w[1], aw = en.w[2], hh = 1.f - lh, hw = 1.f - lw;
      const unsigned mask = __float_as_uint(en.w[3]);
      float gh_w = 0.f, gw_w = 0.f, ga = 0.f;
      if (mask) {   // uniform within the 8-lane group
        const float4 tw = make_float4(go.x * aw, go.y * aw, go.z * aw, go.w * aw);   // top_grad * attention weight
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        auto corner = [&](unsigned bit, int k, float wc, float dh, float dw) {
          if (!(mask & bit)) return;
          const int64_t off = vbase + (int64_t)en.off[k] * vstride;
          const float4 v = load4<T>(value + off);
          const float tv = dot4(tw, v);
          gh_w = fmaf(dh, tv, gh_w);
          gw_w = fmaf(dw, tv, gw_w);
          fma4(val, wc, v);
          atomicAdd(reinterpret_cast<float4 *>(gvalue + off), make_float4(wc * tw.x, wc * tw.y, wc * tw.z, wc * tw.w));
        };
        corner(1u, 0, hh * hw, -hw, -hh);
        corner(2u, 1, hh * lw, -lw, hh);
        corner(4u, 2, lh * hw, hw, -lh);
        corner(8u, 3, lh * lw, lw, lh);
        ga = dot4(go, val);
      }
      gw_w = group_sum8(gw_w);
      gh_w = group_sum8(gh_w);
      ga = group_sum8(ga);
      if (t == 0) {
        float *r = red[warp][g * LP + lp];
        r[0] = gw_w, r[1] = gh_w, r[2] = ga;
      }
    }
    __syncwarp();
    // ---- phase 3: softmax backward and d(offsets), one lane per pair ----
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int q = lane + 32 * it;
      const float *r = red[warp][q];
      const float dot = half_sum(aw_[it] * r[2]);
      glogits[lrow + q] = from_f<TO>(aw_[it] * (r[2] - dot));
      goffs[orow + q * 2] = from_f<TO>(r[0] * sx_[it]);
      goffs[orow + q * 2 + 1] = from_f<TO>(r[1] * sy_[it]);
    }
    __syncwarp();
  }
}

}  // namespace msf

static int msda_check(const char *fn, int B, int Nv, int Nq, int heads, int L, int P, int dtype) {
  RSC_CHECK_ARG(B > 0 && Nv > 0 && Nq > 0, "%s: empty tensor (B=%d,Nv=%d,Nq=%d)", fn, B, Nv, Nq);
  RSC_CHECK_ARG(heads > 0 && heads % 4 == 0, "%s: num_heads must be a multiple of 4 (got %d)", fn, heads);
  RSC_CHECK_ARG(L > 0 && L <= MSDA_MAX_L && P > 0 && L * P <= MSDA_MAX_LP, "%s: need L<=%d, L*P<=%d (L=%d,P=%d)", fn,
                MSDA_MAX_L, MSDA_MAX_LP, L, P);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "%s: bad dtype %d", fn, dtype);
  return RSC_OK;
}

static int msda_grid(int64_t items) {
  int64_t blocks = (items + MSDA_WARPS - 1) / MSDA_WARPS;
  int64_t cap = (int64_t)kNumSMs * 32;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace rsc

using namespace rsc;

extern "C" int rsc_msda_fwd(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                            const float *sampling_loc, const float *attn_weight, void *out, int B, int Nv, int Nq,
                            int heads, int L, int P, int im2col_step, int dtype, void *stream) {
  (void)im2col_step;
  if (int e = msda_check("rsc_msda_fwd", B, Nv, Nq, heads, L, P, dtype)) return e;
  RSC_CHECK_ARG(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && out,
                "rsc_msda_fwd: null pointer");
  int64_t items = (int64_t)B * Nq * (heads / 4);
  size_t smem = sizeof(float) * MSDA_WARPS * 12 * L * P;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RSC_F32)
    msda_fwd_kernel<float><<<msda_grid(items), MSDA_WARPS * 32, smem, st>>>(
        (const float *)value, spatial_shapes, level_start_index, sampling_loc, attn_weight, (float *)out, B, Nv, Nq,
        heads, L, P);
  else
    msda_fwd_kernel<__nv_bfloat16><<<msda_grid(items), MSDA_WARPS * 32, smem, st>>>(
        (const __nv_bfloat16 *)value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
        (__nv_bfloat16 *)out, B, Nv, Nq, heads, L, P);
  RSC_CHECK_LAUNCH("rsc_msda_fwd");
  return RSC_OK;
}

extern "C" int rsc_msda_bwd(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                            const float *sampling_loc, const float *attn_weight, const void *grad_out,
                            float *grad_value, float *grad_loc, float *grad_weight, int B, int Nv, int Nq, int heads,
                            int L, int P, int im2col_step, int dtype, void *stream) {
  (void)im2col_step;
  if (int e = msda_check("rsc_msda_bwd", B, Nv, Nq, heads, L, P, dtype)) return e;
  RSC_CHECK_ARG(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && grad_out &&
                    grad_value && grad_loc && grad_weight,
                "rsc_msda_bwd: null pointer");
  int64_t items = (int64_t)B * Nq * (heads / 4);
  size_t smem = sizeof(float) * MSDA_WARPS * 24 * L * P;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RSC_F32)
    msda_bwd_kernel<float><<<msda_grid(items), MSDA_WARPS * 32, smem, st>>>(
        (const float *)value, spatial_shapes, level_start_index, sampling_loc, attn_weight, (const float *)grad_out,
        grad_value, grad_loc, grad_weight, B, Nv, Nq, heads, L, P);
  else
    msda_bwd_kernel<__nv_bfloat16><<<msda_grid(items), MSDA_WARPS * 32, smem, st>>>(
        (const __nv_bfloat16 *)value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
        (const __nv_bfloat16 *)grad_out, grad_value, grad_loc, grad_weight, B, Nv, Nq, heads, L, P);
  RSC_CHECK_LAUNCH("rsc_msda_bwd");
  return RSC_OK;
}

static int msda_fused_check(const char *fn, int B, int Nv, int Nq, int heads, int L, int P, int R, int dtype, int odt) {
  if (int e = msda_check(fn, B, Nv, Nq, heads, L, P, dtype)) return e;
  RSC_CHECK_ARG(L * P == msf::LP, "%s: the fused kernel needs L*P == 16 (got L=%d, P=%d)", fn, L, P);
  RSC_CHECK_ARG(R == 2 || R == 4, "%s: reference points must be 2-d or 4-d (got %d)", fn, R);
  RSC_CHECK_ARG(odt == RSC_F32 || odt == RSC_BF16, "%s: bad offsets dtype %d", fn, odt);
  return RSC_OK;
}

// row strides (elements) of the offsets / logits matrices and of their gradients; 0 = dense rows.  In-place: the
// defaults are substituted.
static int msda_fused_strides(const char *fn, int heads, int &ostride, int &lstride) {
  const int od = heads * msf::LP * 2, ld = heads * msf::LP;
  if (ostride == 0) ostride = od;
  if (lstride == 0) lstride = ld;
  RSC_CHECK_ARG(ostride >= od && lstride >= ld, "%s: row strides (%d, %d) shorter than a row (%d, %d)", fn, ostride,
                lstride, od, ld);
  return RSC_OK;
}

extern "C" int rsc_msda_fused_fwd(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                                  const void *offsets, const void *logits, const float *ref, void *out, int B, int Nv,
                                  int Nq, int heads, int L, int P, int ref_dim, int dtype, int off_dtype,
                                  int off_row_stride, int logit_row_stride, void *stream) {
  if (int e = msda_fused_check("rsc_msda_fused_fwd", B, Nv, Nq, heads, L, P, ref_dim, dtype, off_dtype)) return e;
  if (int e = msda_fused_strides("rsc_msda_fused_fwd", heads, off_row_stride, logit_row_stride)) return e;
  RSC_CHECK_ARG(value && spatial_shapes && level_start_index && offsets && logits && ref && out,
                "rsc_msda_fused_fwd: null pointer");
  const int64_t items = (int64_t)B * Nq * (heads / 4);
  const int grid = msda_grid(items);
  cudaStream_t st = (cudaStream_t)stream;
#define MFF(T, TO)                                                                                                  \
  msf::msda_fused_fwd_kernel<T, TO><<<grid, msf::WARPS * 32, 0, st>>>((const T *)value, spatial_shapes, level_start_index, \
                                                                       (const TO *)offsets, (const TO *)logits, ref,  \
                                                                       (T *)out, B, Nv, Nq, heads, L, P, ref_dim,     \
                                                                       (int64_t)off_row_stride, (int64_t)logit_row_stride)
  if (dtype == RSC_F32) {
    if (off_dtype == RSC_F32) MFF(float, float); else MFF(float, __nv_bfloat16);
  } else {
    if (off_dtype == RSC_F32) MFF(__nv_bfloat16, float); else MFF(__nv_bfloat16, __nv_bfloat16);
  }
#undef MFF
  RSC_CHECK_LAUNCH("rsc_msda_fused_fwd");
  return RSC_OK;
}

extern "C" int rsc_msda_fused_bwd(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                                  const void *offsets, const void *logits, const float *ref, const void *grad_out,
                                  float *grad_value, void *grad_offsets, void *grad_logits, int B, int Nv, int Nq,
                                  int heads, int L, int P, int ref_dim, int dtype, int off_dtype, int off_row_stride,
                                  int logit_row_stride, void *stream) {
  if (int e = msda_fused_check("rsc_msda_fused_bwd", B, Nv, Nq, heads, L, P, ref_dim, dtype, off_dtype)) return e;
  if (int e = msda_fused_strides("rsc_msda_fused_bwd", heads, off_row_stride, logit_row_stride)) return e;
  RSC_CHECK_ARG(value && spatial_shapes && level_start_index && offsets && logits && ref && grad_out && grad_value &&
                    grad_offsets && grad_logits,
                "rsc_msda_fused_bwd: null pointer");
  const int64_t items = (int64_t)B * Nq * (heads / 4);
  const int grid = msda_grid(items);
  cudaStream_t st = (cudaStream_t)stream;
  static const bool bf = [] {       // RSC_MSDA_BWD_BF=0: the branching one-point loop (A/B switch)
    const char *e = getenv("RSC_MSDA_BWD_BF");
    return !(e && e[0] == '0');
  }();
#define MFB_(T, TO, BF)                                                                                               \
  msf::msda_fused_bwd_kernel<T, TO, BF><<<grid, msf::WARPS * 32, 0, st>>>(                                            \
      (const T *)value, spatial_shapes, level_start_index, (const TO *)offsets, (const TO *)logits, ref,              \
      (const T *)grad_out, grad_value, (TO *)grad_offsets, (TO *)grad_logits, B, Nv, Nq, heads, L, P, ref_dim,        \
      (int64_t)off_row_stride, (int64_t)logit_row_stride)
#define MFB(T, TO)                                                                                                    \
  do {                                                                                                                \
    if (bf) MFB_(T, TO, true);                                                                                        \
    else MFB_(T, TO, false);                                                                                          \
  } while (0)
  if (dtype == RSC_F32) {
    if (off_dtype == RSC_F32) MFB(float, float); else MFB(float, __nv_bfloat16);
  } else {
    if (off_dtype == RSC_F32) MFB(__nv_bfloat16, float); else MFB(__nv_bfloat16, __nv_bfloat16);
  }
#undef MFB
#undef MFB_
  RSC_CHECK_LAUNCH("rsc_msda_fused_bwd");
  return RSC_OK;
}
