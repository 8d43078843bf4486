// Bandwidth-bound helper kernels of the co-training path:
//   PatchMerging gather + LayerNorm (a6), global average pool (a12),
//   bilinear resize align_corners=False (a18/a19), sigmoid focal loss (a16).
#include <float.h>

#include "common.cuh"

namespace rsc {

// ===========================================================================
// PatchMerging: 2x2 gather in nn.Unfold channel order (c*4 + kh*2 + kw) fused
// with LayerNorm(4C).  One warp per output token; a lane owns 4 consecutive
// source channels of each of the 4 source tokens = 16 consecutive output
// channels, so both sides move as 16 B / 8 B vectors.
// ===========================================================================
struct PMGeom {
  int B, H, W, C, Ho, Wo;
};

template <typename T>
__device__ __forceinline__ float4 pm_load(const T *x, const PMGeom &g, int b, int h, int w, int c) {
  if (h >= g.H || w >= g.W) return make_float4(0.f, 0.f, 0.f, 0.f);  // 'corner' zero padding
  return load4<T>(x + (((int64_t)b * g.H + h) * g.W + w) * g.C + c);
}

template <typename T, int ITER>
__global__ void __launch_bounds__(256)
    patch_merge_ln_fwd_kernel(const T *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                              T *__restrict__ y, float *__restrict__ mean, float *__restrict__ rstd, PMGeom g,
                              float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t tokens = (int64_t)g.B * g.Ho * g.Wo;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float invn = 1.0f / (4 * g.C);
  for (int64_t tok = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tok < tokens; tok += nwarps) {
    const int j = (int)(tok % g.Wo), i = (int)((tok / g.Wo) % g.Ho), b = (int)(tok / ((int64_t)g.Wo * g.Ho));
    float4 v[ITER][4];
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < g.C) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[it][k] = pm_load<T>(x, g, b, 2 * i + (k >> 1), 2 * j + (k & 1), c);
          s += v[it][k].x + v[it][k].y + v[it][k].z + v[it][k].w;
        }
      }
    }
    const float mu = warp_sum(s) * invn;
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < g.C) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = v[it][k].x - mu, b2 = v[it][k].y - mu, c2 = v[it][k].z - mu, d = v[it][k].w - mu;
          q += a * a + b2 * b2 + c2 * c2 + d * d;
        }
      }
    }
    const float rs = rsqrtf(warp_sum(q) * invn + eps);
    if (lane == 0) {
      mean[tok] = mu;
      rstd[tok] = rs;
    }
    T *yo = y + tok * 4 * g.C;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < g.C) {
        const float *vf = reinterpret_cast<const float *>(&v[it][0]);  // vf[k*4 + cc]
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int o = (c + cc) * 4;
          const float4 ga = __ldg(reinterpret_cast<const float4 *>(gamma + o));
          const float4 be = __ldg(reinterpret_cast<const float4 *>(beta + o));
          float4 r;
          r.x = (vf[0 * 4 + cc] - mu) * rs * ga.x + be.x;
          r.y = (vf[1 * 4 + cc] - mu) * rs * ga.y + be.y;
          r.z = (vf[2 * 4 + cc] - mu) * rs * ga.z + be.z;
          r.w = (vf[3 * 4 + cc] - mu) * rs * ga.w + be.w;
          store4<T>(yo + o, r);
        }
      }
    }
  }
}

template <typename T>
__device__ __forceinline__ void pm_store(T *dx, const PMGeom &g, int b, int h, int w, int c, float4 v) {
  if (h >= g.H || w >= g.W) return;
  store4<T>(dx + (((int64_t)b * g.H + h) * g.W + w) * g.C + c, v);
}

template <typename T, int ITER>
__global__ void __launch_bounds__(256)
    patch_merge_ln_bwd_kernel(const T *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ mean,
                              const float *__restrict__ rstd, const T *__restrict__ dy, T *__restrict__ dx,
                              float *__restrict__ dgamma, float *__restrict__ dbeta, PMGeom g) {
  const int lane = threadIdx.x & 31;
  const int64_t tokens = (int64_t)g.B * g.Ho * g.Wo;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float invn = 1.0f / (4 * g.C);
  float dg[ITER][16], db[ITER][16];  // [cc*4 + k] = output channel (c+cc)*4 + k
#pragma unroll
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int e = 0; e < 16; ++e) dg[it][e] = 0.f, db[it][e] = 0.f;

  for (int64_t tok = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tok < tokens; tok += nwarps) {
    const int j = (int)(tok % g.Wo), i = (int)((tok / g.Wo) % g.Ho), b = (int)(tok / ((int64_t)g.Wo * g.Ho));
    const float mu = mean[tok], rs = rstd[tok];
    const T *dyo = dy + tok * 4 * g.C;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < g.C) {
        float4 xv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) xv[k] = pm_load<T>(x, g, b, 2 * i + (k >> 1), 2 * j + (k & 1), c);
        const float *xf = reinterpret_cast<const float *>(&xv[0]);  // xf[k*4 + cc]
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int o = (c + cc) * 4;
          const float4 d4 = load4<T>(dyo + o);
          const float4 ga = __ldg(reinterpret_cast<const float4 *>(gamma + o));
          const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
          const float gv[4] = {ga.x, ga.y, ga.z, ga.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float xh = (xf[k * 4 + cc] - mu) * rs;
            const float dxh = dv[k] * gv[k];
            s1 += dxh;
            s2 = fmaf(dxh, xh, s2);
            dg[it][cc * 4 + k] = fmaf(dv[k], xh, dg[it][cc * 4 + k]);
            db[it][cc * 4 + k] += dv[k];
          }
        }
      }
    }
    s1 = warp_sum(s1) * invn;
    s2 = warp_sum(s2) * invn;
    // second sweep: re-read x / dy (L1 hits) and emit dx
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < g.C) {
        float4 xv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) xv[k] = pm_load<T>(x, g, b, 2 * i + (k >> 1), 2 * j + (k & 1), c);
        float *xf = reinterpret_cast<float *>(&xv[0]);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int o = (c + cc) * 4;
          const float4 d4 = load4<T>(dyo + o);
          const float4 ga = __ldg(reinterpret_cast<const float4 *>(gamma + o));
          const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
          const float gv[4] = {ga.x, ga.y, ga.z, ga.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float xh = (xf[k * 4 + cc] - mu) * rs;
            xf[k * 4 + cc] = rs * (dv[k] * gv[k] - s1 - xh * s2);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) pm_store<T>(dx, g, b, 2 * i + (k >> 1), 2 * j + (k & 1), c, xv[k]);
      }
    }
  }
  // per-lane partial sums -> per-CTA shared table -> ONE global atomic per channel per CTA
  extern __shared__ float red[];   // [2][4C]
  for (int i2 = threadIdx.x; i2 < 8 * g.C; i2 += blockDim.x) red[i2] = 0.f;
  __syncthreads();
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int c = it * 128 + lane * 4;
    if (c < g.C) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        atomicAdd(red + c * 4 + e, dg[it][e]);
        atomicAdd(red + 4 * g.C + c * 4 + e, db[it][e]);
      }
    }
  }
  __syncthreads();
  for (int i2 = threadIdx.x; i2 < 4 * g.C; i2 += blockDim.x) {
    atomicAdd(dgamma + i2, red[i2]);
    atomicAdd(dbeta + i2, red[4 * g.C + i2]);
  }
}

// ---------------------------------------------------------------------------
// Round-2 rewrite of the two kernels above (same arithmetic, same lane <-> channel mapping).  The first version
// walked ONE token per warp step through load -> reduce -> reload -> store: three dependent memory round trips per
// 2.3 KB, i.e. latency-bound at 0.16-0.28 of the HBM roofline (profiles/r02_ncu_launches_timed_region.csv.gz:
// 288 us for the 369 MB stage-0 backward).  Here a warp step covers TPW tokens whose loads are all issued before the
// first reduction, the operands stay in registers in their STORAGE type (bf16 pairs: half the registers of fp32)
// for the second sweep instead of being re-read, and 128-thread CTAs let the register-heavy variants keep more
// warps resident.  The cross-warp fold of d(gamma) / d(beta) is a plain shared-memory read-modify-write per warp
// in turn (shared fp32 atomicAdd is a compare-and-swap loop on this architecture).
// ---------------------------------------------------------------------------
template <typename T>
struct Raw4;
template <>
struct Raw4<float> {
  float4 v;
  __device__ __forceinline__ void load(const float *p) { v = __ldg(reinterpret_cast<const float4 *>(p)); }
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void touch() {}
  __device__ __forceinline__ float4 get() const { return v; }
};
template <>
struct Raw4<__nv_bfloat16> {
  uint2 v;
  __device__ __forceinline__ void load(const __nv_bfloat16 *p) { v = __ldg(reinterpret_cast<const uint2 *>(p)); }
  __device__ __forceinline__ void zero() { v = make_uint2(0u, 0u); }
  // opaque to the optimiser: a later get() is re-evaluated from the packed words instead of keeping the four fp32
  // values of an earlier get() alive (registers decide how many warps stay resident)
  __device__ __forceinline__ void touch() { asm volatile("" : "+r"(v.x), "+r"(v.y)); }
  __device__ __forceinline__ float4 get() const {
    // bf16 -> fp32 is a 16-bit shift
    return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                       __uint_as_float(v.y & 0xffff0000u));
  }
};

// 16 consecutive elements (a lane's output channels of one merged token) <-> four Raw4: bf16 moves them as two 16-byte
// vectors (four 8-byte accesses per lane, 32 bytes apart between lanes, would touch every 32-byte sector four times)
template <typename T>
__device__ __forceinline__ void load_row16(const T *p, Raw4<T> (&r)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) r[k].load(p + 4 * k);
}
template <>
__device__ __forceinline__ void load_row16<__nv_bfloat16>(const __nv_bfloat16 *p, Raw4<__nv_bfloat16> (&r)[4]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p)), b = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
  r[0].v = make_uint2(a.x, a.y), r[1].v = make_uint2(a.z, a.w), r[2].v = make_uint2(b.x, b.y), r[3].v = make_uint2(b.z, b.w);
}
template <typename T>
__device__ __forceinline__ void store_row16(T *p, const float4 (&r)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) store4<T>(p + 4 * k, r[k]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}
template <>
__device__ __forceinline__ void store_row16<__nv_bfloat16>(__nv_bfloat16 *p, const float4 (&r)[4]) {
  reinterpret_cast<uint4 *>(p)[0] = make_uint4(pack_bf16x2(r[0].x, r[0].y), pack_bf16x2(r[0].z, r[0].w),
                                               pack_bf16x2(r[1].x, r[1].y), pack_bf16x2(r[1].z, r[1].w));
  reinterpret_cast<uint4 *>(p)[1] = make_uint4(pack_bf16x2(r[2].x, r[2].y), pack_bf16x2(r[2].z, r[2].w),
                                               pack_bf16x2(r[3].x, r[3].y), pack_bf16x2(r[3].z, r[3].w));
}

constexpr int PM_THREADS = 128;

// merged token index -> (image, row, column) with two 32-bit divisions (the 64-bit `%` / `/` of the first version were
// most of the kernels' instructions: three emulated divisions per token, per lane); tokens < 2^31 is checked at launch
__device__ __forceinline__ void pm_decompose(unsigned tok, const PMGeom &g, int &b, int &i, int &j) {
  const unsigned q = tok / (unsigned)g.Wo;
  j = (int)(tok - q * (unsigned)g.Wo);
  const unsigned bb = q / (unsigned)g.Ho;
  i = (int)(q - bb * (unsigned)g.Ho);
  b = (int)bb;
}

template <typename T, int ITER, int TPW>
__global__ void __launch_bounds__(PM_THREADS)
    patch_merge_ln_fwd2_kernel(const T *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                               T *__restrict__ y, float *__restrict__ mean, float *__restrict__ rstd, PMGeom g,
                               float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t tokens = (int64_t)g.B * g.Ho * g.Wo;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float invn = 1.0f / (4 * g.C);
  const int64_t rowC = (int64_t)g.W * g.C;
  for (int64_t tok0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * TPW; tok0 < tokens;
       tok0 += nwarps * TPW) {
    Raw4<T> rx[TPW][ITER][4];
    float s[TPW];
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      const int64_t tok = tok0 + t;
      const bool valid = tok < tokens;
      int b, i, j;
      pm_decompose((unsigned)tok, g, b, i, j);
      const T *xb = x + ((int64_t)b * g.H + 2 * i) * rowC + (int64_t)2 * j * g.C;
      const bool h1 = 2 * i + 1 < g.H, w1 = 2 * j + 1 < g.W;
#pragma unroll
      for (int it = 0; it < ITER; ++it) {
        const int c = it * 128 + lane * 4;
        const bool on = valid && c < g.C;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (on && ((k >> 1) == 0 || h1) && ((k & 1) == 0 || w1)) rx[t][it][k].load(xb + (k >> 1) * rowC + (k & 1) * g.C + c);
          else rx[t][it][k].zero();
        }
      }
    }
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      float a = 0.f;
#pragma unroll
      for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 v = rx[t][it][k].get();
          a += (v.x + v.y) + (v.z + v.w);
        }
      s[t] = a;
    }
    float mu[TPW], rs[TPW];
#pragma unroll
    for (int t = 0; t < TPW; ++t) mu[t] = warp_sum(s[t]) * invn;
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      float q = 0.f;
#pragma unroll
      for (int it = 0; it < ITER; ++it) {
        if (it * 128 + lane * 4 < g.C) {      // (inactive lanes hold zeros, which are NOT zero after centring)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 v = rx[t][it][k].get();
            const float a = v.x - mu[t], b2 = v.y - mu[t], c2 = v.z - mu[t], d = v.w - mu[t];
            q += a * a + b2 * b2 + c2 * c2 + d * d;
          }
        }
      }
      s[t] = q;
    }
#pragma unroll
    for (int t = 0; t < TPW; ++t) rs[t] = rsqrtf(warp_sum(s[t]) * invn + eps);
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      const int64_t tok = tok0 + t;
      if (tok >= tokens) break;
      if (lane == 0) {
        mean[tok] = mu[t];
        rstd[tok] = rs[t];
      }
      T *yo = y + tok * 4 * g.C;
#pragma unroll
      for (int it = 0; it < ITER; ++it) {
        const int c = it * 128 + lane * 4;
        if (c < g.C) {
          float4 xv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) xv[k] = rx[t][it][k].get();
          const float *vf = reinterpret_cast<const float *>(&xv[0]);  // vf[k*4 + cc]
          float4 r[4];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int o = (c + cc) * 4;
            const float4 ga = __ldg(reinterpret_cast<const float4 *>(gamma + o));
            const float4 be = __ldg(reinterpret_cast<const float4 *>(beta + o));
            r[cc].x = (vf[0 * 4 + cc] - mu[t]) * rs[t] * ga.x + be.x;
            r[cc].y = (vf[1 * 4 + cc] - mu[t]) * rs[t] * ga.y + be.y;
            r[cc].z = (vf[2 * 4 + cc] - mu[t]) * rs[t] * ga.z + be.z;
            r[cc].w = (vf[3 * 4 + cc] - mu[t]) * rs[t] * ga.w + be.w;
          }
          store_row16<T>(yo + c * 4, r);
        }
      }
    }
  }
}

template <typename T, int ITER, int TPW>
__global__ void __launch_bounds__(PM_THREADS, ITER <= 3 ? 3 : 2)   // (3 CTAs: 168 registers, <= 152 B of spills)
    patch_merge_ln_bwd2_kernel(const T *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ mean,
                               const float *__restrict__ rstd, const T *__restrict__ dy, T *__restrict__ dx,
                               float *__restrict__ dgamma, float *__restrict__ dbeta, PMGeom g) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tokens = (int64_t)g.B * g.Ho * g.Wo;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float invn = 1.0f / (4 * g.C);
  const int64_t rowC = (int64_t)g.W * g.C;
  float dg[ITER][16], db[ITER][16];  // [cc*4 + k] = output channel (c+cc)*4 + k
#pragma unroll
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int e = 0; e < 16; ++e) dg[it][e] = 0.f, db[it][e] = 0.f;

  for (int64_t tok0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * TPW; tok0 < tokens;
       tok0 += nwarps * TPW) {
    Raw4<T> rx[TPW][ITER][4], rd[TPW][ITER][4];   // rx[..][k]: source token k; rd[..][cc]: output channels (c+cc)*4 .. +3
    float mu[TPW], rs[TPW], s1[TPW], s2[TPW];
    int64_t xoff[TPW];      // offset of the 2x2 patch's first source token (x and dx share the layout)
    unsigned edge[TPW];     // bit 0: the patch has a second row, bit 1: a second column
    // every load of the TPW tokens is in flight before the first use
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      const int64_t tok = tok0 + t;
      const bool valid = tok < tokens;
      int b, i, j;
      pm_decompose((unsigned)tok, g, b, i, j);
      xoff[t] = ((int64_t)b * g.H + 2 * i) * rowC + (int64_t)2 * j * g.C;
      const T *xb = x + xoff[t];
      const T *dyo = dy + tok * 4 * g.C;
      const bool h1 = 2 * i + 1 < g.H, w1 = 2 * j + 1 < g.W;
      edge[t] = (h1 ? 1u : 0u) | (w1 ? 2u : 0u);
      mu[t] = valid ? __ldg(mean + tok) : 0.f;
      rs[t] = valid ? __ldg(rstd + tok) : 0.f;
#pragma unroll
      for (int it = 0; it < ITER; ++it) {
        const int c = it * 128 + lane * 4;
        const bool on = valid && c < g.C;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (on && ((k >> 1) == 0 || h1) && ((k & 1) == 0 || w1)) rx[t][it][k].load(xb + (k >> 1) * rowC + (k & 1) * g.C + c);
          else rx[t][it][k].zero();
        }
        if (on) load_row16<T>(dyo + c * 4, rd[t][it]);
        else {
#pragma unroll
          for (int k = 0; k < 4; ++k) rd[t][it][k].zero();
        }
      }
    }
    // sweep 1: row sums of the LayerNorm backward + d(gamma) / d(beta) partials
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int it = 0; it < ITER; ++it) {
        const int c = it * 128 + lane * 4;
        if (c < g.C) {
          float4 xv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) xv[k] = rx[t][it][k].get();
          const float *xf = reinterpret_cast<const float *>(&xv[0]);  // xf[k*4 + cc]
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const float4 d4 = rd[t][it][cc].get();
            const float4 ga = __ldg(reinterpret_cast<const float4 *>(gamma + (c + cc) * 4));
            const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
            const float gv[4] = {ga.x, ga.y, ga.z, ga.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float xh = (xf[k * 4 + cc] - mu[t]) * rs[t];
              const float dxh = dv[k] * gv[k];
              a1 += dxh;
              a2 = fmaf(dxh, xh, a2);
              dg[it][cc * 4 + k] = fmaf(dv[k], xh, dg[it][cc * 4 + k]);
              db[it][cc * 4 + k] += dv[k];
            }
          }
        }
      }
      s1[t] = a1, s2[t] = a2;
    }
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      s1[t] = warp_sum(s1[t]) * invn;
      s2[t] = warp_sum(s2[t]) * invn;
    }
    // sweep 2 from the same registers: dx
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      const int64_t tok = tok0 + t;
      if (tok >= tokens) break;
      T *dxb = dx + xoff[t];
      const bool h1 = edge[t] & 1u, w1 = edge[t] & 2u;
#pragma unroll
      for (int it = 0; it < ITER; ++it) {
        const int c = it * 128 + lane * 4;
        if (c < g.C) {
          float4 xv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            rx[t][it][k].touch();
            rd[t][it][k].touch();
            xv[k] = rx[t][it][k].get();
          }
          float *xf = reinterpret_cast<float *>(&xv[0]);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const float4 d4 = rd[t][it][cc].get();
            const float4 ga = __ldg(reinterpret_cast<const float4 *>(gamma + (c + cc) * 4));
            const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
            const float gv[4] = {ga.x, ga.y, ga.z, ga.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float xh = (xf[k * 4 + cc] - mu[t]) * rs[t];
              xf[k * 4 + cc] = rs[t] * (dv[k] * gv[k] - s1[t] - xh * s2[t]);
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (((k >> 1) == 0 || h1) && ((k & 1) == 0 || w1)) store4<T>(dxb + (k >> 1) * rowC + (k & 1) * g.C + c, xv[k]);
        }
      }
    }
  }
  // per-lane partial sums -> per-CTA table (each warp in turn, plain read-modify-write: a lane owns its channels)
  // -> ONE global atomic per channel per CTA
  extern __shared__ float red[];   // [2][4C]
  for (int i2 = threadIdx.x; i2 < 8 * g.C; i2 += blockDim.x) red[i2] = 0.f;
  __syncthreads();
  for (int w = 0; w < PM_THREADS / 32; ++w) {
    if (warp == w) {
#pragma unroll
      for (int it = 0; it < ITER; ++it) {
        const int c = it * 128 + lane * 4;
        if (c < g.C) {
          float4 *rg = reinterpret_cast<float4 *>(red + c * 4), *rb = reinterpret_cast<float4 *>(red + 4 * g.C + c * 4);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            float4 a = rg[cc], b2 = rb[cc];
            a.x += dg[it][cc * 4 + 0], a.y += dg[it][cc * 4 + 1], a.z += dg[it][cc * 4 + 2], a.w += dg[it][cc * 4 + 3];
            b2.x += db[it][cc * 4 + 0], b2.y += db[it][cc * 4 + 1], b2.z += db[it][cc * 4 + 2], b2.w += db[it][cc * 4 + 3];
            rg[cc] = a, rb[cc] = b2;
          }
        }
      }
    }
    __syncthreads();
  }
  for (int i2 = threadIdx.x; i2 < 4 * g.C; i2 += blockDim.x) {
    atomicAdd(dgamma + i2, red[i2]);
    atomicAdd(dbeta + i2, red[4 * g.C + i2]);
  }
}

// ===========================================================================
// Global average pool
// ===========================================================================
template <typename T>
__global__ void gap_nchw_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, int64_t planes, int HW) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < planes; p += nwarps) {
    const T *src = x + p * HW;
    float s = 0.f;
    for (int i = lane; i < HW; i += 32) s += to_f<T>(src[i]);
    s = warp_sum(s);
    if (lane == 0) y[p] = from_f<T>(s / HW);
  }
}

// x (B,HW,C): block = 32 channel-quads x 8 row slices; grid (C/128, B)
template <typename T>
__global__ void __launch_bounds__(256) gap_nhwc_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, int C, int HW) {
  __shared__ float4 red[8][32];
  const int c = blockIdx.x * 128 + threadIdx.x * 4;
  const int b = blockIdx.y;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    const T *src = x + (int64_t)b * HW * C + c;
    for (int r = threadIdx.y; r < HW; r += 8) {
      const float4 v = load4<T>(src + (int64_t)r * C);
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 v = red[k][threadIdx.x];
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
    const float inv = 1.0f / HW;
    store4<T>(y + (int64_t)b * C + c, make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv));
  }
}

template <typename T>
__global__ void gap_nchw_bwd_kernel(const T *__restrict__ dy, T *__restrict__ dx, int64_t total, int HW) {
  const float inv = 1.0f / HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = from_f<T>(to_f<T>(dy[i / HW]) * inv);
}

template <typename T>
__global__ void gap_nhwc_bwd_kernel(const T *__restrict__ dy, T *__restrict__ dx, int64_t total4, int C, int HW) {
  const float inv = 1.0f / HW;
  const int C4 = C / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const int64_t b = i / ((int64_t)C4 * HW);
    const float4 v = load4<T>(dy + b * C + c4 * 4);
    store4<T>(dx + i * 4, make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv));
  }
}

// ===========================================================================
// Bilinear resize (align_corners=False), planes (N,H,W)
// ===========================================================================
__device__ __forceinline__ void bil_src(int o, float scale, int in, int &i0, int &i1, float &lam) {
  float s = scale * (o + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 < in - 1 ? i0 + 1 : i0;
  lam = s - i0;
}

template <typename T>
__global__ void bilinear_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, int64_t total, int Hi, int Wi, int Ho,
                                    int Wo, float sh, float sw) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
    const int64_t n = i / ((int64_t)Wo * Ho);
    int y0, y1, x0, x1;
    float ly, lx;
    bil_src(oy, sh, Hi, y0, y1, ly);
    bil_src(ox, sw, Wi, x0, x1, lx);
    const T *p = x + n * Hi * Wi;
    const float v00 = to_f<T>(p[y0 * Wi + x0]), v01 = to_f<T>(p[y0 * Wi + x1]);
    const float v10 = to_f<T>(p[y1 * Wi + x0]), v11 = to_f<T>(p[y1 * Wi + x1]);
    y[i] = from_f<T>((1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11));
  }
}

// weight with which output index o (of an axis) reads input index i
__device__ __forceinline__ float bil_w(int o, int i, float scale, int in) {
  int i0, i1;
  float lam;
  bil_src(o, scale, in, i0, i1, lam);
  float w = 0.f;
  if (i0 == i) w += 1.f - lam;
  if (i1 == i) w += lam;
  return w;
}

// gather formulation: thread per INPUT pixel, loops over the outputs that read it
template <typename T>
__global__ void bilinear_bwd_kernel(const T *__restrict__ dy, T *__restrict__ dx, int64_t total, int Hi, int Wi,
                                    int Ho, int Wo, float sh, float sw) {
  const float rh = 1.f / sh, rw = 1.f / sw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ix = (int)(i % Wi), iy = (int)((i / Wi) % Hi);
    const int64_t n = i / ((int64_t)Wi * Hi);
    // outputs o with source coordinate in (i-1, i+1) (plus border clamping): conservative range
    int oy0 = (int)floorf((iy - 1 + 0.5f) * rh - 0.5f) - 1, oy1 = (int)ceilf((iy + 1 + 0.5f) * rh - 0.5f) + 1;
    int ox0 = (int)floorf((ix - 1 + 0.5f) * rw - 0.5f) - 1, ox1 = (int)ceilf((ix + 1 + 0.5f) * rw - 0.5f) + 1;
    if (iy == 0) oy0 = 0;
    if (ix == 0) ox0 = 0;
    if (iy == Hi - 1) oy1 = Ho - 1;
    if (ix == Wi - 1) ox1 = Wo - 1;
    oy0 = max(oy0, 0), ox0 = max(ox0, 0), oy1 = min(oy1, Ho - 1), ox1 = min(ox1, Wo - 1);
    const T *p = dy + n * Ho * Wo;
    float acc = 0.f;
    for (int oy = oy0; oy <= oy1; ++oy) {
      const float wy = bil_w(oy, iy, sh, Hi);
      if (wy == 0.f) continue;
      float row = 0.f;
      for (int ox = ox0; ox <= ox1; ++ox) {
        const float wx = bil_w(ox, ix, sw, Wi);
        if (wx != 0.f) row = fmaf(wx, to_f<T>(p[(int64_t)oy * Wo + ox]), row);
      }
      acc = fmaf(wy, row, acc);
    }
    dx[i] = from_f<T>(acc);
  }
}

// channels-last variants: x (B,Hi,Wi,C) -> y (B,Ho,Wo,C); a thread moves one 16-byte channel vector (VEC elements), a
// warp covers consecutive channels of one pixel -> coalesced both ways, and the maps the convolutions / norms produce
// (channels-last) need no transpose in front of the resize
template <typename T, int VEC>
__device__ __forceinline__ void ldvec(const T *p, float (&v)[VEC]);
template <>
__device__ __forceinline__ void ldvec<float, 4>(const float *p, float (&v)[4]) {
  const float4 r = __ldg(reinterpret_cast<const float4 *>(p));
  v[0] = r.x, v[1] = r.y, v[2] = r.z, v[3] = r.w;
}
template <>
__device__ __forceinline__ void ldvec<__nv_bfloat16, 8>(const __nv_bfloat16 *p, float (&v)[8]) {
  const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w[i]));
    v[2 * i] = f.x, v[2 * i + 1] = f.y;
  }
}
template <typename T, int VEC>
__device__ __forceinline__ void stvec(T *p, const float (&v)[VEC]);
template <>
__device__ __forceinline__ void stvec<float, 4>(float *p, const float (&v)[4]) {
  *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void stvec<__nv_bfloat16, 8>(__nv_bfloat16 *p, const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t *>(&h);
  }
  *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

template <typename T, int VEC>
__global__ void bilinear_cl_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, int64_t total, int C, int Hi, int Wi, int Ho,
                                       int Wo, float sh, float sw) {
  const int CV = C / VEC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    int64_t t = i / CV;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho);
    const int64_t b = t / Ho;
    int y0, y1, x0, x1;
    float ly, lx;
    bil_src(oy, sh, Hi, y0, y1, ly);
    bil_src(ox, sw, Wi, x0, x1, lx);
    const T *p = x + b * Hi * Wi * C + cv * VEC;
    float v00[VEC], v01[VEC], v10[VEC], v11[VEC], o[VEC];
    ldvec<T, VEC>(p + ((int64_t)y0 * Wi + x0) * C, v00);
    ldvec<T, VEC>(p + ((int64_t)y0 * Wi + x1) * C, v01);
    ldvec<T, VEC>(p + ((int64_t)y1 * Wi + x0) * C, v10);
    ldvec<T, VEC>(p + ((int64_t)y1 * Wi + x1) * C, v11);
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      o[e] = (1.f - ly) * ((1.f - lx) * v00[e] + lx * v01[e]) + ly * ((1.f - lx) * v10[e] + lx * v11[e]);
    stvec<T, VEC>(y + i * VEC, o);
  }
}

template <typename T, int VEC>
__global__ void bilinear_cl_bwd_kernel(const T *__restrict__ dy, T *__restrict__ dx, int64_t total, int C, int Hi, int Wi, int Ho,
                                       int Wo, float sh, float sw) {
  const int CV = C / VEC;
  const float rh = 1.f / sh, rw = 1.f / sw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    int64_t t = i / CV;
    const int ix = (int)(t % Wi);
    t /= Wi;
    const int iy = (int)(t % Hi);
    const int64_t b = t / Hi;
    int oy0 = (int)floorf((iy - 1 + 0.5f) * rh - 0.5f) - 1, oy1 = (int)ceilf((iy + 1 + 0.5f) * rh - 0.5f) + 1;
    int ox0 = (int)floorf((ix - 1 + 0.5f) * rw - 0.5f) - 1, ox1 = (int)ceilf((ix + 1 + 0.5f) * rw - 0.5f) + 1;
    if (iy == 0) oy0 = 0;
    if (ix == 0) ox0 = 0;
    if (iy == Hi - 1) oy1 = Ho - 1;
    if (ix == Wi - 1) ox1 = Wo - 1;
    oy0 = max(oy0, 0), ox0 = max(ox0, 0), oy1 = min(oy1, Ho - 1), ox1 = min(ox1, Wo - 1);
    const T *p = dy + b * Ho * Wo * C + cv * VEC;
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    for (int oy = oy0; oy <= oy1; ++oy) {
      const float wy = bil_w(oy, iy, sh, Hi);
      if (wy == 0.f) continue;
      for (int ox = ox0; ox <= ox1; ++ox) {
        const float wgt = wy * bil_w(ox, ix, sw, Wi);
        if (wgt == 0.f) continue;
        float v[VEC];
        ldvec<T, VEC>(p + ((int64_t)oy * Wo + ox) * C, v);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = fmaf(wgt, v[e], acc[e]);
      }
    }
    stvec<T, VEC>(dx + i * VEC, acc);
  }
}

// ===========================================================================
// Iterative box refinement of the DINO decoder / head: out = sigmoid(tmp + inverse_sigmoid(ref, eps)), fp32 out
// (models/multi/bbox_head/dino_head.py forward + DinoTransformerDecoder.forward; mmdet inverse_sigmoid:
//  x = clamp(ref, 0, 1); log(max(x, eps) / max(1 - x, eps))).  One launch instead of the 8-9 element-wise kernels of the
// eager chain; the backward follows torch's clamp conventions (gradient passes where min <= x <= max).
// ===========================================================================
template <typename T>
__global__ void box_refine_fwd_kernel(const T *__restrict__ tmp, const float *__restrict__ ref, float *__restrict__ out, int64_t n,
                                      float eps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = fminf(fmaxf(ref[i], 0.f), 1.f);
    const float inv = logf(fmaxf(x, eps) / fmaxf(1.f - x, eps));
    out[i] = 1.f / (1.f + expf(-(to_f<T>(tmp[i]) + inv)));
  }
}
template <typename T>
__global__ void box_refine_bwd_kernel(const float *__restrict__ out, const float *__restrict__ ref, const float *__restrict__ dout,
                                      T *__restrict__ dtmp, float *__restrict__ dref, int64_t n, float eps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = out[i], g = dout[i] * s * (1.f - s);
    dtmp[i] = from_f<T>(g);
    if (dref) {
      const float r = ref[i];
      float d = 0.f;
      if (r >= 0.f && r <= 1.f) {
        const float x = r, y = 1.f - r;
        if (x >= eps) d += 1.f / x;
        if (y >= eps) d += 1.f / y;
      }
      dref[i] = g * d;
    }
  }
}

// ===========================================================================
// Sigmoid focal loss (element-wise, mmcv semantics)
// ===========================================================================
template <typename T, bool BWD>
__global__ void focal_kernel(const T *__restrict__ in, const int64_t *__restrict__ target, float *__restrict__ out,
                             int64_t total, int C, float gamma, float alpha) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t t = target[i / C];
    const float x = to_f<T>(in[i]);
    const float p = 1.f / (1.f + expf(-x));
    const float lp = logf(fmaxf(p, FLT_MIN)), ln = logf(fmaxf(1.f - p, FLT_MIN));
    const float pw_p = powf(1.f - p, gamma), pw_n = powf(p, gamma);
    float r;
    if (!BWD) {
      r = (t == c) ? -alpha * pw_p * lp : -(1.f - alpha) * pw_n * ln;
    } else {
      r = (t == c) ? -alpha * pw_p * (1.f - p - gamma * p * lp) : -(1.f - alpha) * pw_n * (gamma * (1.f - p) * ln - p);
    }
    out[i] = r;
  }
}

// ===========================================================================
// Patch embedding gather: non-overlapping 4x4 patches of an NCHW image -> token-major (B*Ho*Wo, Cin*16) rows in
// Conv2d weight order (c, kh, kw), cast to the GEMM dtype on the way.  mmdet PatchEmbed = Conv2d(3, 96, k4, s4):
// with this gather it is ONE plain GEMM with K = 48 (the reference's cuDNN path converts the 123 MB fp32 batch
// to channels-last first and runs an implicit-GEMM convolution forward and backward).
// ===========================================================================
template <typename TI, typename TO>
__global__ void patchify4_kernel(const TI *__restrict__ x, TO *__restrict__ y, int64_t total, int Cin, int H, int W) {
  const int Ho = H >> 2, Wo = W >> 2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % Wo);
    const int c = (int)((idx / Wo) % Cin);
    const int i = (int)((idx / ((int64_t)Wo * Cin)) % Ho);
    const int64_t b = idx / ((int64_t)Wo * Cin * Ho);
    const TI *src = x + ((b * Cin + c) * H + 4 * i) * W + 4 * j;
    TO *dst = y + ((b * Ho + i) * Wo + j) * (Cin * 16) + c * 16;
    float4 r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = load4<TI>(src + (int64_t)k * W);
#pragma unroll
    for (int k = 0; k < 4; ++k) store4<TO>(dst + 4 * k, r[k]);
  }
}

static inline int ew_grid(int64_t n, int per_block) {
  int64_t blocks = (n + per_block - 1) / per_block;
  int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace rsc

using namespace rsc;

// ===========================================================================
// uint8 image batch -> normalised float batch (device half of a deferred Normalize): channel flip (BGR -> RGB),
// (x - mean) / std per channel, zero outside each image's valid (h, w) -- the reference pads AFTER normalising
// ===========================================================================
template <typename TO>
__global__ void normalize_u8_kernel(const unsigned char *__restrict__ img, TO *__restrict__ out, const float *__restrict__ mean,
                                    const float *__restrict__ inv_std, const int *__restrict__ valid_hw, int64_t total4, int C,
                                    int H, int W, int flip) {
  const int W4 = W / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int x4 = (int)(i % W4);
    int64_t t = i / W4;
    const int y = (int)(t % H);
    t /= H;
    const int c = (int)(t % C), b = (int)(t / C);
    const int cs = flip ? C - 1 - c : c;
    const uchar4 v = __ldg(reinterpret_cast<const uchar4 *>(img + (((int64_t)b * C + cs) * H + y) * W) + x4);
    const float m = mean[c], is = inv_std[c];
    const int vh = valid_hw ? valid_hw[2 * b] : H, vw = valid_hw ? valid_hw[2 * b + 1] : W;
    const unsigned char px[4] = {v.x, v.y, v.z, v.w};
    TO *o = out + i * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = from_f<TO>((y < vh && x4 * 4 + e < vw) ? ((float)px[e] - m) * is : 0.f);
  }
}

#define DISPATCH_T(dtype, ...)                    \
  if (dtype == RSC_F32) {                         \
    using T = float;                              \
    __VA_ARGS__;                                  \
  } else {                                        \
    using T = __nv_bfloat16;                      \
    __VA_ARGS__;                                  \
  }

template <typename T>
static void pm_fwd_launch(int iters, int grid, cudaStream_t st, const void *x, const float *gamma, const float *beta,
                          void *y, float *mean, float *rstd, PMGeom g, float eps) {
#define PM_CASE(N)                                                                                                  \
  case N:                                                                                                           \
    patch_merge_ln_fwd_kernel<T, N><<<grid, 256, 0, st>>>((const T *)x, gamma, beta, (T *)y, mean, rstd, g, eps);   \
    break;
  switch (iters) { PM_CASE(1) PM_CASE(2) PM_CASE(3) PM_CASE(4) }
#undef PM_CASE
}
template <typename T>
static void pm_bwd_launch(int iters, int grid, cudaStream_t st, const void *x, const float *gamma, const float *mean,
                          const float *rstd, const void *dy, void *dx, float *dgamma, float *dbeta, PMGeom g) {
#define PM_CASE(N)                                                                                                 \
  case N:                                                                                                          \
    patch_merge_ln_bwd_kernel<T, N>                                                                                \
        <<<grid, 256, 8 * g.C * sizeof(float), st>>>((const T *)x, gamma, mean, rstd, (const T *)dy, (T *)dx,      \
                                                     dgamma, dbeta, g);                                            \
    break;
  switch (iters) { PM_CASE(1) PM_CASE(2) PM_CASE(3) PM_CASE(4) }
#undef PM_CASE
}

// grid of the round-2 kernels: every resident CTA slot of the device, once (persistent warps), or fewer for few tokens
template <typename K>
static int pm_grid(K kernel, size_t smem, int64_t tokens, int tpw) {
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, PM_THREADS, smem) != cudaSuccess || occ < 1) occ = 1;
  const int64_t need = (tokens + (int64_t)tpw * (PM_THREADS / 32) - 1) / ((int64_t)tpw * (PM_THREADS / 32));
  const int64_t cap = (int64_t)kNumSMs * occ;
  return (int)(need < cap ? need : cap);
}

template <typename T>
static void pm_fwd2_launch(int iters, cudaStream_t st, const void *x, const float *gamma, const float *beta, void *y,
                           float *mean, float *rstd, PMGeom g, float eps) {
  const int64_t tokens = (int64_t)g.B * g.Ho * g.Wo;
#define PM_CASE(N, TPW)                                                                                             \
  case N: {                                                                                                         \
    auto k = patch_merge_ln_fwd2_kernel<T, N, TPW>;                                                                 \
    k<<<pm_grid(k, 0, tokens, TPW), PM_THREADS, 0, st>>>((const T *)x, gamma, beta, (T *)y, mean, rstd, g, eps);    \
  } break;
  switch (iters) { PM_CASE(1, 4) PM_CASE(2, 4) PM_CASE(3, 2) PM_CASE(4, 2) }
#undef PM_CASE
}
template <typename T>
static void pm_bwd2_launch(int iters, cudaStream_t st, const void *x, const float *gamma, const float *mean,
                           const float *rstd, const void *dy, void *dx, float *dgamma, float *dbeta, PMGeom g) {
  const int64_t tokens = (int64_t)g.B * g.Ho * g.Wo;
  const size_t smem = 8 * g.C * sizeof(float);
#define PM_CASE(N, TPW)                                                                                             \
  case N: {                                                                                                         \
    auto k = patch_merge_ln_bwd2_kernel<T, N, TPW>;                                                                 \
    k<<<pm_grid(k, smem, tokens, TPW), PM_THREADS, smem, st>>>((const T *)x, gamma, mean, rstd, (const T *)dy,      \
                                                               (T *)dx, dgamma, dbeta, g);                          \
  } break;
  switch (iters) { PM_CASE(1, 4) PM_CASE(2, 2) PM_CASE(3, 1) PM_CASE(4, 1) }
#undef PM_CASE
}

// RSC_PATCH_MERGE_V2=1 selects the round-2 kernels.  They are 2x faster on the backward (+1.5 % on the whole step,
// profiles/r02_ab_round2b.log), pass every parity test and compute-sanitizer (initcheck, racecheck, memcheck), return
// bit-identical y / dx on repeated calls and equal the round-1 kernels to 1e-8 (tools/pm_determinism.py,
// profiles/r02_pm_determinism.log).  Still, repeated fp32 trainings of the small test models, which agree to 2e-7 with
// the round-1 kernels, land on one of TWO outcomes ~5e-4 apart in single iterations with these
// (tools/replay_diag*.py, profiles/r02_replay_diag_*.log): their different last bits put those trainings next to
// something that amplifies last-bit differences, and what that is was not found before the round's GPU budget ended.
// The default stays with the kernels under which the replay tests reproduce.
static int g_pm_variant = 0;   // 0 = environment, 1 / 2 = set by rsc_set_patch_merge_variant
static bool pm_v1() {
  static const bool env_v1 = [] {
    const char *e = getenv("RSC_PATCH_MERGE_V2");
    return !(e && e[0] == '1');
  }();
  return g_pm_variant ? g_pm_variant == 1 : env_v1;
}

// the round-2 kernels move the merged-token side as 16-byte vectors
static bool pm_aligned16(const void *a, const void *b, const void *c) {
  return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}
// ... and decompose token indices in 32 bits
static bool pm_v2_ok(int64_t tokens, const void *a, const void *b, const void *c) {
  return !pm_v1() && tokens < ((int64_t)1 << 31) && pm_aligned16(a, b, c);
}

static int pm_check(const char *fn, int B, int H, int W, int C, int dtype) {
  RSC_CHECK_ARG(B > 0 && H > 0 && W > 0, "%s: empty tensor (B=%d,H=%d,W=%d)", fn, B, H, W);
  RSC_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 512, "%s: C must be a multiple of 4, <= 512 (got %d)", fn, C);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "%s: bad dtype %d", fn, dtype);
  return RSC_OK;
}

extern "C" int rsc_set_patch_merge_variant(int variant) {
  RSC_CHECK_ARG(variant >= 0 && variant <= 2, "rsc_set_patch_merge_variant: 0 (environment), 1 or 2 (got %d)", variant);
  g_pm_variant = variant;
  return RSC_OK;
}

extern "C" int rsc_patch_merge_ln_fwd(const void *x, const float *gamma, const float *beta, void *y, float *mean,
                                      float *rstd, int B, int H, int W, int C, float eps, int dtype, void *stream) {
  if (int e = pm_check("rsc_patch_merge_ln_fwd", B, H, W, C, dtype)) return e;
  RSC_CHECK_ARG(x && gamma && beta && y && mean && rstd, "rsc_patch_merge_ln_fwd: null pointer");
  PMGeom g{B, H, W, C, (H + 1) / 2, (W + 1) / 2};
  int64_t tokens = (int64_t)B * g.Ho * g.Wo;
  int grid = ew_grid(tokens, 8);
  if (!pm_v2_ok(tokens, x, y, nullptr)) {
    DISPATCH_T(dtype, pm_fwd_launch<T>((C + 127) / 128, grid, (cudaStream_t)stream, x, gamma, beta, y, mean, rstd, g, eps));
  } else {
    DISPATCH_T(dtype, pm_fwd2_launch<T>((C + 127) / 128, (cudaStream_t)stream, x, gamma, beta, y, mean, rstd, g, eps));
  }
  RSC_CHECK_LAUNCH("rsc_patch_merge_ln_fwd");
  return RSC_OK;
}

extern "C" int rsc_patch_merge_ln_bwd(const void *x, const float *gamma, const float *mean, const float *rstd,
                                      const void *dy, void *dx, float *dgamma, float *dbeta, int B, int H, int W,
                                      int C, int dtype, void *stream) {
  if (int e = pm_check("rsc_patch_merge_ln_bwd", B, H, W, C, dtype)) return e;
  RSC_CHECK_ARG(x && gamma && mean && rstd && dy && dx && dgamma && dbeta, "rsc_patch_merge_ln_bwd: null pointer");
  PMGeom g{B, H, W, C, (H + 1) / 2, (W + 1) / 2};
  int64_t tokens = (int64_t)B * g.Ho * g.Wo;
  int64_t blocks = (tokens + 7) / 8;
  int grid = (int)(blocks < kNumSMs * 4 ? blocks : kNumSMs * 4);  // few warps -> few dgamma atomics
  if (!pm_v2_ok(tokens, x, dy, dx)) {
    DISPATCH_T(dtype, pm_bwd_launch<T>((C + 127) / 128, grid, (cudaStream_t)stream, x, gamma, mean, rstd, dy, dx, dgamma,
                                       dbeta, g));
  } else {
    DISPATCH_T(dtype, pm_bwd2_launch<T>((C + 127) / 128, (cudaStream_t)stream, x, gamma, mean, rstd, dy, dx, dgamma,
                                        dbeta, g));
  }
  RSC_CHECK_LAUNCH("rsc_patch_merge_ln_bwd");
  return RSC_OK;
}

static int gap_check(const char *fn, int B, int C, int HW, int channels_last, int dtype) {
  RSC_CHECK_ARG(B > 0 && C > 0 && HW > 0, "%s: empty tensor (B=%d,C=%d,HW=%d)", fn, B, C, HW);
  RSC_CHECK_ARG(!channels_last || C % 4 == 0, "%s: channels_last needs C %% 4 == 0 (C=%d)", fn, C);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "%s: bad dtype %d", fn, dtype);
  return RSC_OK;
}

extern "C" int rsc_gap_fwd(const void *x, void *y, int B, int C, int HW, int channels_last, int dtype, void *stream) {
  if (int e = gap_check("rsc_gap_fwd", B, C, HW, channels_last, dtype)) return e;
  RSC_CHECK_ARG(x && y, "rsc_gap_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (channels_last) {
    dim3 grid((C + 127) / 128, B), block(32, 8);
    DISPATCH_T(dtype, gap_nhwc_fwd_kernel<T><<<grid, block, 0, st>>>((const T *)x, (T *)y, C, HW));
  } else {
    int64_t planes = (int64_t)B * C;
    DISPATCH_T(dtype, gap_nchw_fwd_kernel<T><<<ew_grid(planes, 8), 256, 0, st>>>((const T *)x, (T *)y, planes, HW));
  }
  RSC_CHECK_LAUNCH("rsc_gap_fwd");
  return RSC_OK;
}

extern "C" int rsc_gap_bwd(const void *dy, void *dx, int B, int C, int HW, int channels_last, int dtype,
                           void *stream) {
  if (int e = gap_check("rsc_gap_bwd", B, C, HW, channels_last, dtype)) return e;
  RSC_CHECK_ARG(dy && dx, "rsc_gap_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)B * C * HW;
  if (channels_last) {
    DISPATCH_T(dtype,
               gap_nhwc_bwd_kernel<T><<<ew_grid(total / 4, 256), 256, 0, st>>>((const T *)dy, (T *)dx, total / 4, C, HW));
  } else {
    DISPATCH_T(dtype, gap_nchw_bwd_kernel<T><<<ew_grid(total, 256), 256, 0, st>>>((const T *)dy, (T *)dx, total, HW));
  }
  RSC_CHECK_LAUNCH("rsc_gap_bwd");
  return RSC_OK;
}

static int bil_check(const char *fn, int N, int Hi, int Wi, int Ho, int Wo, int dtype) {
  RSC_CHECK_ARG(N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "%s: empty tensor (N=%d, %dx%d -> %dx%d)", fn, N, Hi, Wi,
                Ho, Wo);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "%s: bad dtype %d", fn, dtype);
  return RSC_OK;
}

extern "C" int rsc_bilinear_fwd(const void *x, void *y, int N, int Hi, int Wi, int Ho, int Wo, int dtype,
                                void *stream) {
  if (int e = bil_check("rsc_bilinear_fwd", N, Hi, Wi, Ho, Wo, dtype)) return e;
  RSC_CHECK_ARG(x && y, "rsc_bilinear_fwd: null pointer");
  int64_t total = (int64_t)N * Ho * Wo;
  float sh = (float)Hi / Ho, sw = (float)Wi / Wo;
  DISPATCH_T(dtype, bilinear_fwd_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
                        (const T *)x, (T *)y, total, Hi, Wi, Ho, Wo, sh, sw));
  RSC_CHECK_LAUNCH("rsc_bilinear_fwd");
  return RSC_OK;
}

extern "C" int rsc_bilinear_bwd(const void *dy, void *dx, int N, int Hi, int Wi, int Ho, int Wo, int dtype,
                                void *stream) {
  if (int e = bil_check("rsc_bilinear_bwd", N, Hi, Wi, Ho, Wo, dtype)) return e;
  RSC_CHECK_ARG(dy && dx, "rsc_bilinear_bwd: null pointer");
  int64_t total = (int64_t)N * Hi * Wi;
  float sh = (float)Hi / Ho, sw = (float)Wi / Wo;
  DISPATCH_T(dtype, bilinear_bwd_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
                        (const T *)dy, (T *)dx, total, Hi, Wi, Ho, Wo, sh, sw));
  RSC_CHECK_LAUNCH("rsc_bilinear_bwd");
  return RSC_OK;
}

static int focal_check(const char *fn, int N, int C, int dtype) {
  RSC_CHECK_ARG(N >= 0 && C > 0, "%s: bad shape (N=%d,C=%d)", fn, N, C);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "%s: bad dtype %d", fn, dtype);
  return RSC_OK;
}

extern "C" int rsc_sigmoid_focal_loss_fwd(const void *input, const int64_t *target, float *output, int N, int C,
                                          float gamma, float alpha, int dtype, void *stream) {
  if (int e = focal_check("rsc_sigmoid_focal_loss_fwd", N, C, dtype)) return e;
  if (N == 0) return RSC_OK;
  RSC_CHECK_ARG(input && target && output, "rsc_sigmoid_focal_loss_fwd: null pointer");
  int64_t total = (int64_t)N * C;
  DISPATCH_T(dtype, focal_kernel<T, false><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
                        (const T *)input, target, output, total, C, gamma, alpha));
  RSC_CHECK_LAUNCH("rsc_sigmoid_focal_loss_fwd");
  return RSC_OK;
}

extern "C" int rsc_sigmoid_focal_loss_bwd(const void *input, const int64_t *target, float *grad_input, int N, int C,
                                          float gamma, float alpha, int dtype, void *stream) {
  if (int e = focal_check("rsc_sigmoid_focal_loss_bwd", N, C, dtype)) return e;
  if (N == 0) return RSC_OK;
  RSC_CHECK_ARG(input && target && grad_input, "rsc_sigmoid_focal_loss_bwd: null pointer");
  int64_t total = (int64_t)N * C;
  DISPATCH_T(dtype, focal_kernel<T, true><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
                        (const T *)input, target, grad_input, total, C, gamma, alpha));
  RSC_CHECK_LAUNCH("rsc_sigmoid_focal_loss_bwd");
  return RSC_OK;
}

extern "C" int rsc_patchify4(const void *x, void *y, int B, int Cin, int H, int W, int in_dtype, int out_dtype,
                             void *stream) {
  RSC_CHECK_ARG(B > 0 && Cin > 0 && H > 0 && W > 0, "rsc_patchify4: empty tensor (B=%d,C=%d,H=%d,W=%d)", B, Cin, H, W);
  RSC_CHECK_ARG(H % 4 == 0 && W % 4 == 0, "rsc_patchify4: H and W must be multiples of 4 (pad first; got %dx%d)", H, W);
  RSC_CHECK_ARG((in_dtype == RSC_F32 || in_dtype == RSC_BF16) && (out_dtype == RSC_F32 || out_dtype == RSC_BF16),
                "rsc_patchify4: bad dtype");
  RSC_CHECK_ARG(x && y, "rsc_patchify4: null pointer");
  const int64_t total = (int64_t)B * Cin * (H / 4) * (W / 4);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ew_grid(total, 256);
  if (in_dtype == RSC_F32 && out_dtype == RSC_F32)
    patchify4_kernel<float, float><<<grid, 256, 0, st>>>((const float *)x, (float *)y, total, Cin, H, W);
  else if (in_dtype == RSC_F32)
    patchify4_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float *)x, (__nv_bfloat16 *)y, total, Cin, H, W);
  else if (out_dtype == RSC_F32)
    patchify4_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, (float *)y, total, Cin, H, W);
  else
    patchify4_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y,
                                                                         total, Cin, H, W);
  RSC_CHECK_LAUNCH("rsc_patchify4");
  return RSC_OK;
}

extern "C" int rsc_normalize_u8(const void *img, void *out, const float *mean, const float *inv_std, const int *valid_hw, int B, int C,
                                int H, int W, int flip, int out_dtype, void *stream) {
  RSC_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && W % 4 == 0, "rsc_normalize_u8: bad shape (B=%d,C=%d,H=%d,W=%d; W %% 4 == 0)", B, C, H, W);
  RSC_CHECK_ARG(out_dtype == RSC_F32 || out_dtype == RSC_BF16, "rsc_normalize_u8: bad dtype %d", out_dtype);
  RSC_CHECK_ARG(img && out && mean && inv_std, "rsc_normalize_u8: null pointer");
  const int64_t total4 = (int64_t)B * C * H * (W / 4);
  if (out_dtype == RSC_F32)
    normalize_u8_kernel<float><<<ew_grid(total4, 256), 256, 0, (cudaStream_t)stream>>>((const unsigned char *)img, (float *)out, mean, inv_std,
                                                                                      valid_hw, total4, C, H, W, flip);
  else
    normalize_u8_kernel<__nv_bfloat16><<<ew_grid(total4, 256), 256, 0, (cudaStream_t)stream>>>(
        (const unsigned char *)img, (__nv_bfloat16 *)out, mean, inv_std, valid_hw, total4, C, H, W, flip);
  RSC_CHECK_LAUNCH("rsc_normalize_u8");
  return RSC_OK;
}

static int bil_cl_check(const char *fn, int B, int C, int Hi, int Wi, int Ho, int Wo, int dtype) {
  RSC_CHECK_ARG(B > 0 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "%s: empty tensor (B=%d, C=%d, %dx%d -> %dx%d)", fn, B, C,
                Hi, Wi, Ho, Wo);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "%s: bad dtype %d", fn, dtype);
  RSC_CHECK_ARG(C % (dtype == RSC_F32 ? 4 : 8) == 0, "%s: C=%d must be a multiple of the 16-byte channel vector", fn, C);
  return RSC_OK;
}

extern "C" int rsc_bilinear_cl_fwd(const void *x, void *y, int B, int C, int Hi, int Wi, int Ho, int Wo, int dtype, void *stream) {
  if (int e = bil_cl_check("rsc_bilinear_cl_fwd", B, C, Hi, Wi, Ho, Wo, dtype)) return e;
  RSC_CHECK_ARG(x && y, "rsc_bilinear_cl_fwd: null pointer");
  const float sh = (float)Hi / Ho, sw = (float)Wi / Wo;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RSC_F32) {
    const int64_t total = (int64_t)B * Ho * Wo * (C / 4);
    bilinear_cl_fwd_kernel<float, 4><<<ew_grid(total, 256), 256, 0, st>>>((const float *)x, (float *)y, total, C, Hi, Wi, Ho, Wo, sh, sw);
  } else {
    const int64_t total = (int64_t)B * Ho * Wo * (C / 8);
    bilinear_cl_fwd_kernel<__nv_bfloat16, 8>
        <<<ew_grid(total, 256), 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y, total, C, Hi, Wi, Ho, Wo, sh, sw);
  }
  RSC_CHECK_LAUNCH("rsc_bilinear_cl_fwd");
  return RSC_OK;
}

extern "C" int rsc_bilinear_cl_bwd(const void *dy, void *dx, int B, int C, int Hi, int Wi, int Ho, int Wo, int dtype, void *stream) {
  if (int e = bil_cl_check("rsc_bilinear_cl_bwd", B, C, Hi, Wi, Ho, Wo, dtype)) return e;
  RSC_CHECK_ARG(dy && dx, "rsc_bilinear_cl_bwd: null pointer");
  const float sh = (float)Hi / Ho, sw = (float)Wi / Wo;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RSC_F32) {
    const int64_t total = (int64_t)B * Hi * Wi * (C / 4);
    bilinear_cl_bwd_kernel<float, 4><<<ew_grid(total, 256), 256, 0, st>>>((const float *)dy, (float *)dx, total, C, Hi, Wi, Ho, Wo, sh, sw);
  } else {
    const int64_t total = (int64_t)B * Hi * Wi * (C / 8);
    bilinear_cl_bwd_kernel<__nv_bfloat16, 8>
        <<<ew_grid(total, 256), 256, 0, st>>>((const __nv_bfloat16 *)dy, (__nv_bfloat16 *)dx, total, C, Hi, Wi, Ho, Wo, sh, sw);
  }
  RSC_CHECK_LAUNCH("rsc_bilinear_cl_bwd");
  return RSC_OK;
}

extern "C" int rsc_box_refine_fwd(const void *tmp, const float *ref, float *out, int64_t n, float eps, int dtype, void *stream) {
  RSC_CHECK_ARG(n > 0 && tmp && ref && out, "rsc_box_refine_fwd: null pointer / empty tensor");
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "rsc_box_refine_fwd: bad dtype %d", dtype);
  DISPATCH_T(dtype, box_refine_fwd_kernel<T><<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>((const T *)tmp, ref, out, n, eps));
  RSC_CHECK_LAUNCH("rsc_box_refine_fwd");
  return RSC_OK;
}

extern "C" int rsc_box_refine_bwd(const float *out, const float *ref, const float *dout, void *dtmp, float *dref, int64_t n, float eps,
                                  int dtype, void *stream) {
  RSC_CHECK_ARG(n > 0 && out && ref && dout && dtmp, "rsc_box_refine_bwd: null pointer / empty tensor");
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "rsc_box_refine_bwd: bad dtype %d", dtype);
  DISPATCH_T(dtype, box_refine_bwd_kernel<T><<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(out, ref, dout, (T *)dtmp, dref, n, eps));
  RSC_CHECK_LAUNCH("rsc_box_refine_bwd");
  return RSC_OK;
}
