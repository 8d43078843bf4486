// LayerNorm forward / backward over the last dimension (norm1/norm2 of every
// Swin block, the per-stage output norms, the encoder / decoder post-norms).  Replaces ATen's
// native_layer_norm(+backward): one warp per row, 16-byte / 8-byte vector I/O, statistics in fp32,
// input and output element types independent (fp32 residual stream in -> bf16 GEMM operand out),
// d(gamma)/d(beta) accumulated per lane in registers over the persistent row loop, reduced through
// shared memory, one atomic per channel per CTA.
#include "common.cuh"

namespace rsc {

// Sub-warp mapping: L lanes share one row (L = C / (8*EV), a power of two), a warp handles
// 32/L rows at a time, every lane owns EV vectors of 8 consecutive channels, interleaved by L so
// that one load instruction of the L lanes covers 16*L (bf16) contiguous bytes of the row.
//   C = 96 -> L=4, EV=3 | 192 -> 8,3 | 384 -> 16,3 | 768 -> 32,3 | 256 -> 8,4 | 128 -> 4,4 | 512 -> 16,4 | 1024 -> 32,4
constexpr int LN_THREADS = 128;

template <typename T>
__device__ __forceinline__ void load8(const T *p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float *p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16 *p, float (&v)[8]) {
  const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w[i]));
    v[2 * i] = f.x, v[2 * i + 1] = f.y;
  }
}
template <typename T>
__device__ __forceinline__ void store8(T *p, const float (&v)[8]);
template <>
__device__ __forceinline__ void store8<float>(float *p, const float (&v)[8]) {
  reinterpret_cast<float4 *>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4 *>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16 *p, const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t *>(&h);
  }
  *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

// VW-element vector access (VW = 8: 16 B bf16 / 2 x 16 B fp32; VW = 4: 8 B bf16 / 16 B fp32)
template <typename T, int VW>
__device__ __forceinline__ void loadv(const T *p, float (&v)[VW]) {
  if constexpr (VW == 8) {
    load8<T>(p, v);
  } else {
    const float4 f = load4<T>(p);
    v[0] = f.x, v[1] = f.y, v[2] = f.z, v[3] = f.w;
  }
}
template <typename T, int VW>
__device__ __forceinline__ void storev(T *p, const float (&v)[VW]) {
  if constexpr (VW == 8) {
    store8<T>(p, v);
  } else {
    store4<T>(p, make_float4(v[0], v[1], v[2], v[3]));
  }
}

template <int L>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename TI, typename TO, int VW, int EV, int L>
__global__ void __launch_bounds__(LN_THREADS)
    ln_fwd_kernel(const TI *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                  TO *__restrict__ y, float *__restrict__ mean, float *__restrict__ rstd, int64_t rows, float eps) {
  constexpr int C = VW * EV * L, RPW = 32 / L;
  const int lane = threadIdx.x & 31, sub = lane % L, rw = lane / L;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float ga[EV][VW], be[EV][VW];
#pragma unroll
  for (int k = 0; k < EV; ++k) {
    loadv<float, VW>(gamma + (sub + k * L) * VW, ga[k]);
    loadv<float, VW>(beta + (sub + k * L) * VW, be[k]);
  }
  for (int64_t r0 = warp * RPW; r0 < rows; r0 += nwarps * RPW) {
    const int64_t r = r0 + rw;
    const bool ok = r < rows;
    float v[EV][VW];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < EV; ++k) {
      if (ok) loadv<TI, VW>(x + r * C + (sub + k * L) * VW, v[k]);
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        if (!ok) v[k][e] = 0.f;
        s += v[k][e];
      }
    }
    const float mu = group_sum<L>(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < EV; ++k)
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        const float d = v[k][e] - mu;
        q = fmaf(d, d, q);
      }
    const float rs = rsqrtf(group_sum<L>(q) * (1.0f / C) + eps);
    if (ok) {
      if (sub == 0) {
        mean[r] = mu;
        rstd[r] = rs;
      }
#pragma unroll
      for (int k = 0; k < EV; ++k) {
        float o[VW];
#pragma unroll
        for (int e = 0; e < VW; ++e) o[e] = fmaf((v[k][e] - mu) * rs, ga[k][e], be[k][e]);
        storev<TO, VW>(y + r * C + (sub + k * L) * VW, o);
      }
    }
  }
}

template <typename TI, typename TO, int VW, int EV, int L>
__global__ void __launch_bounds__(LN_THREADS)
    ln_bwd_kernel(const TI *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ mean,
                  const float *__restrict__ rstd, const TO *__restrict__ dy, TI *__restrict__ dx,
                  float *__restrict__ dgamma, float *__restrict__ dbeta, int64_t rows) {
  constexpr int C = VW * EV * L, RPW = 32 / L;
  __shared__ float red[2 * C];   // per-CTA partial d(gamma) | d(beta)
  const int lane = threadIdx.x & 31, sub = lane % L, rw = lane / L;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float ga[EV][VW], dg[EV][VW], db[EV][VW];
#pragma unroll
  for (int k = 0; k < EV; ++k) {
    loadv<float, VW>(gamma + (sub + k * L) * VW, ga[k]);
#pragma unroll
    for (int e = 0; e < VW; ++e) dg[k][e] = 0.f, db[k][e] = 0.f;
  }
  for (int64_t r0 = warp * RPW; r0 < rows; r0 += nwarps * RPW) {
    const int64_t r = r0 + rw;
    const bool ok = r < rows;
    const float mu = ok ? mean[r] : 0.f, rs = ok ? rstd[r] : 0.f;
    float xh[EV][VW], g[EV][VW];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < EV; ++k) {
      float d[VW];
      if (ok) {
        loadv<TI, VW>(x + r * C + (sub + k * L) * VW, xh[k]);
        loadv<TO, VW>(dy + r * C + (sub + k * L) * VW, d);
      }
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        if (!ok) xh[k][e] = 0.f, d[e] = 0.f;
        xh[k][e] = (xh[k][e] - mu) * rs;
        dg[k][e] = fmaf(d[e], xh[k][e], dg[k][e]);
        db[k][e] += d[e];
        g[k][e] = d[e] * ga[k][e];
        s1 += g[k][e];
        s2 = fmaf(g[k][e], xh[k][e], s2);
      }
    }
    s1 = group_sum<L>(s1) * (1.0f / C);
    s2 = group_sum<L>(s2) * (1.0f / C);
    if (ok) {
#pragma unroll
      for (int k = 0; k < EV; ++k) {
        float o[VW];
#pragma unroll
        for (int e = 0; e < VW; ++e) o[e] = rs * (g[k][e] - s1 - xh[k][e] * s2);
        storev<TI, VW>(dx + r * C + (sub + k * L) * VW, o);
      }
    }
  }
  // lanes with equal `sub` (different rows of the warp) hold partials of the same channels
#pragma unroll
  for (int k = 0; k < EV; ++k)
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      float a = dg[k][e], b = db[k][e];
#pragma unroll
      for (int o = 16; o >= L; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (rw == 0) {
        atomicAdd(red + (sub + k * L) * VW + e, a);
        atomicAdd(red + C + (sub + k * L) * VW + e, b);
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[C + i]);
  }
}

// (VW, EV, L) for a supported C (fewest elements per lane), or false -> generic warp-per-row kernels below
static bool ln_shape(int C, int &vw, int &ev, int &l) {
  static const int table[8][4] = {{96, 4, 3, 8},   {192, 4, 3, 16}, {384, 4, 3, 32}, {768, 8, 3, 32},
                                  {128, 8, 1, 16}, {256, 8, 1, 32}, {512, 8, 2, 32}, {1024, 8, 4, 32}};
  for (int i = 0; i < 8; ++i)
    if (table[i][0] == C) {
      vw = table[i][1], ev = table[i][2], l = table[i][3];
      return true;
    }
  return false;
}

#define LN_SHAPES(M) M(4, 3, 8) M(4, 3, 16) M(4, 3, 32) M(8, 3, 32) M(8, 1, 16) M(8, 1, 32) M(8, 2, 32) M(8, 4, 32)

template <typename TI, typename TO>
static bool ln_fast_fwd(int C, int grid, cudaStream_t st, const void *x, const float *gamma, const float *beta, void *y,
                        float *mean, float *rstd, int64_t rows, float eps) {
  int vw, ev, l;
  if (!ln_shape(C, vw, ev, l)) return false;
#define LNF(V, E, LL)                                                                                      \
  if (vw == V && ev == E && l == LL) {                                                                     \
    ln_fwd_kernel<TI, TO, V, E, LL><<<grid, LN_THREADS, 0, st>>>((const TI *)x, gamma, beta, (TO *)y, mean, rstd, rows, eps); \
    return true;                                                                                           \
  }
  LN_SHAPES(LNF)
#undef LNF
  return false;
}

template <typename TI, typename TO>
static bool ln_fast_bwd(int C, int grid, cudaStream_t st, const void *x, const float *gamma, const float *mean,
                        const float *rstd, const void *dy, void *dx, float *dgamma, float *dbeta, int64_t rows) {
  int vw, ev, l;
  if (!ln_shape(C, vw, ev, l)) return false;
#define LNB(V, E, LL)                                                                                       \
  if (vw == V && ev == E && l == LL) {                                                                      \
    ln_bwd_kernel<TI, TO, V, E, LL><<<grid, LN_THREADS, 0, st>>>((const TI *)x, gamma, mean, rstd, (const TO *)dy, \
                                                                  (TI *)dx, dgamma, dbeta, rows);           \
    return true;                                                                                            \
  }
  LN_SHAPES(LNB)
#undef LNB
  return false;
}

// ---------------------------------------------------------------------------
// generic path (any C % 4 == 0, C <= 1024): one warp per row
// ---------------------------------------------------------------------------
constexpr int LN_MAX_ITER = 8;   // C <= 1024
constexpr int LN_WARPS = 8;

template <typename TI, typename TO, int ITER>
__global__ void __launch_bounds__(LN_WARPS * 32)
    layernorm_fwd_kernel(const TI *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                         TO *__restrict__ y, float *__restrict__ mean, float *__restrict__ rstd, int64_t rows, int C,
                         float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float invn = 1.0f / C;
  float4 ga[ITER], be[ITER];
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int c = it * 128 + lane * 4;
    if (c < C) {
      ga[it] = __ldg(reinterpret_cast<const float4 *>(gamma + c));
      be[it] = __ldg(reinterpret_cast<const float4 *>(beta + c));
    }
  }
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += nwarps) {
    const TI *xr = x + r * C;
    float4 v[ITER];
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < C) {
        v[it] = load4<TI>(xr + c);
        s += v[it].x + v[it].y + v[it].z + v[it].w;
      }
    }
    const float mu = warp_sum(s) * invn;
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < C) {
        const float a = v[it].x - mu, b = v[it].y - mu, cc = v[it].z - mu, d = v[it].w - mu;
        q += a * a + b * b + cc * cc + d * d;
      }
    }
    const float rs = rsqrtf(warp_sum(q) * invn + eps);
    if (lane == 0) {
      mean[r] = mu;
      rstd[r] = rs;
    }
    TO *yr = y + r * C;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < C) {
        float4 o;
        o.x = (v[it].x - mu) * rs * ga[it].x + be[it].x;
        o.y = (v[it].y - mu) * rs * ga[it].y + be[it].y;
        o.z = (v[it].z - mu) * rs * ga[it].z + be[it].z;
        o.w = (v[it].w - mu) * rs * ga[it].w + be[it].w;
        store4<TO>(yr + c, o);
      }
    }
  }
}

template <typename TI, typename TO, int ITER>
__global__ void __launch_bounds__(LN_WARPS * 32)
    layernorm_bwd_kernel(const TI *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ mean,
                         const float *__restrict__ rstd, const TO *__restrict__ dy, TI *__restrict__ dx,
                         float *__restrict__ dgamma, float *__restrict__ dbeta, int64_t rows, int C) {
  extern __shared__ float red[];   // [2][C] per-CTA partial sums
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float invn = 1.0f / C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float4 ga[ITER], dg[ITER], db[ITER];
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int c = it * 128 + lane * 4;
    dg[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) ga[it] = __ldg(reinterpret_cast<const float4 *>(gamma + c));
  }
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += nwarps) {
    const float mu = mean[r], rs = rstd[r];
    const TI *xr = x + r * C;
    const TO *dyr = dy + r * C;
    float4 xh[ITER], g[ITER];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < C) {
        const float4 xv = load4<TI>(xr + c);
        const float4 d = load4<TO>(dyr + c);
        xh[it] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        dg[it].x = fmaf(d.x, xh[it].x, dg[it].x), dg[it].y = fmaf(d.y, xh[it].y, dg[it].y);
        dg[it].z = fmaf(d.z, xh[it].z, dg[it].z), dg[it].w = fmaf(d.w, xh[it].w, dg[it].w);
        db[it].x += d.x, db[it].y += d.y, db[it].z += d.z, db[it].w += d.w;
        g[it] = make_float4(d.x * ga[it].x, d.y * ga[it].y, d.z * ga[it].z, d.w * ga[it].w);
        s1 += g[it].x + g[it].y + g[it].z + g[it].w;
        s2 += g[it].x * xh[it].x + g[it].y * xh[it].y + g[it].z * xh[it].z + g[it].w * xh[it].w;
      }
    }
    s1 = warp_sum(s1) * invn;
    s2 = warp_sum(s2) * invn;
    TI *dxr = dx + r * C;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < C) {
        float4 o;
        o.x = rs * (g[it].x - s1 - xh[it].x * s2);
        o.y = rs * (g[it].y - s1 - xh[it].y * s2);
        o.z = rs * (g[it].z - s1 - xh[it].z * s2);
        o.w = rs * (g[it].w - s1 - xh[it].w * s2);
        store4<TI>(dxr + c, o);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int c = it * 128 + lane * 4;
    if (c < C) {
      atomicAdd(red + c, dg[it].x), atomicAdd(red + c + 1, dg[it].y), atomicAdd(red + c + 2, dg[it].z),
          atomicAdd(red + c + 3, dg[it].w);
      atomicAdd(red + C + c, db[it].x), atomicAdd(red + C + c + 1, db[it].y), atomicAdd(red + C + c + 2, db[it].z),
          atomicAdd(red + C + c + 3, db[it].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[C + i]);
  }
}

static int ln_check(const char *fn, int64_t rows, int C, int a, int b) {
  RSC_CHECK_ARG(rows > 0, "%s: empty tensor (rows=%lld)", fn, (long long)rows);
  RSC_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 128 * LN_MAX_ITER, "%s: C must be a multiple of 4, <= %d (got %d)", fn,
                128 * LN_MAX_ITER, C);
  RSC_CHECK_ARG((a == RSC_F32 || a == RSC_BF16) && (b == RSC_F32 || b == RSC_BF16), "%s: bad dtype", fn);
  return RSC_OK;
}

template <typename TI, typename TO>
static void ln_fwd_launch(int iters, int grid, cudaStream_t st, const void *x, const float *gamma, const float *beta,
                          void *y, float *mean, float *rstd, int64_t rows, int C, float eps) {
#define LN_CASE(N)                                                                                                  \
  case N:                                                                                                           \
    layernorm_fwd_kernel<TI, TO, N><<<grid, LN_WARPS * 32, 0, st>>>((const TI *)x, gamma, beta, (TO *)y, mean, rstd, \
                                                                     rows, C, eps);                                 \
    break;
  switch (iters) { LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8) }
#undef LN_CASE
}

template <typename TI, typename TO>
static void ln_bwd_launch(int iters, int grid, cudaStream_t st, const void *x, const float *gamma, const float *mean,
                          const float *rstd, const void *dy, void *dx, float *dgamma, float *dbeta, int64_t rows,
                          int C) {
  size_t smem = sizeof(float) * 2 * C;
#define LN_CASE(N)                                                                                          \
  case N:                                                                                                   \
    layernorm_bwd_kernel<TI, TO, N><<<grid, LN_WARPS * 32, smem, st>>>(                                     \
        (const TI *)x, gamma, mean, rstd, (const TO *)dy, (TI *)dx, dgamma, dbeta, rows, C);                \
    break;
  switch (iters) { LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8) }
#undef LN_CASE
}

#define LN_DISPATCH(in_dt, out_dt, FN, ...)                                   \
  if (in_dt == RSC_F32 && out_dt == RSC_F32) FN<float, float>(__VA_ARGS__);   \
  else if (in_dt == RSC_F32) FN<float, __nv_bfloat16>(__VA_ARGS__);           \
  else if (out_dt == RSC_F32) FN<__nv_bfloat16, float>(__VA_ARGS__);          \
  else FN<__nv_bfloat16, __nv_bfloat16>(__VA_ARGS__);

}  // namespace rsc

using namespace rsc;

extern "C" int rsc_layernorm_fwd(const void *x, const float *gamma, const float *beta, void *y, float *mean,
                                 float *rstd, int64_t rows, int C, float eps, int in_dtype, int out_dtype,
                                 void *stream) {
  if (int e = ln_check("rsc_layernorm_fwd", rows, C, in_dtype, out_dtype)) return e;
  RSC_CHECK_ARG(x && gamma && beta && y && mean && rstd, "rsc_layernorm_fwd: null pointer");
  {
    int vw = 0, ev = 0, l = 32;
    if (ln_shape(C, vw, ev, l)) {
      int64_t fb = (rows + (LN_THREADS / 32) * (32 / l) - 1) / ((LN_THREADS / 32) * (32 / l));
      int fgrid = (int)(fb < kNumSMs * 12 ? fb : kNumSMs * 12);
      bool done = false;
      cudaStream_t st = (cudaStream_t)stream;
      if (in_dtype == RSC_F32 && out_dtype == RSC_F32) done = ln_fast_fwd<float, float>(C, fgrid, st, x, gamma, beta, y, mean, rstd, rows, eps);
      else if (in_dtype == RSC_F32) done = ln_fast_fwd<float, __nv_bfloat16>(C, fgrid, st, x, gamma, beta, y, mean, rstd, rows, eps);
      else if (out_dtype == RSC_F32) done = ln_fast_fwd<__nv_bfloat16, float>(C, fgrid, st, x, gamma, beta, y, mean, rstd, rows, eps);
      else done = ln_fast_fwd<__nv_bfloat16, __nv_bfloat16>(C, fgrid, st, x, gamma, beta, y, mean, rstd, rows, eps);
      if (done) {
        RSC_CHECK_LAUNCH("rsc_layernorm_fwd");
        return RSC_OK;
      }
    }
  }
  int64_t blocks = (rows + LN_WARPS - 1) / LN_WARPS;
  int grid = (int)(blocks < kNumSMs * 8 ? blocks : kNumSMs * 8);
  LN_DISPATCH(in_dtype, out_dtype, ln_fwd_launch, (C + 127) / 128, grid, (cudaStream_t)stream, x, gamma, beta, y, mean,
              rstd, rows, C, eps);
  RSC_CHECK_LAUNCH("rsc_layernorm_fwd");
  return RSC_OK;
}

extern "C" int rsc_layernorm_bwd(const void *x, const float *gamma, const float *mean, const float *rstd,
                                 const void *dy, void *dx, float *dgamma, float *dbeta, int64_t rows, int C,
                                 int in_dtype, int out_dtype, void *stream) {
  if (int e = ln_check("rsc_layernorm_bwd", rows, C, in_dtype, out_dtype)) return e;
  RSC_CHECK_ARG(x && gamma && mean && rstd && dy && dx && dgamma && dbeta, "rsc_layernorm_bwd: null pointer");
  {
    int vw = 0, ev = 0, l = 32;
    if (ln_shape(C, vw, ev, l)) {
      int rpb = (LN_THREADS / 32) * (32 / l);
      if (rpb < 32) rpb = 32;        // >= 32 rows per CTA: each CTA ends with 2*C global atomics (d gamma | d beta)
      int64_t fb = (rows + rpb - 1) / rpb;
      int fgrid = (int)(fb < kNumSMs * 8 ? fb : kNumSMs * 8);
      bool done = false;
      cudaStream_t st = (cudaStream_t)stream;
      if (in_dtype == RSC_F32 && out_dtype == RSC_F32) done = ln_fast_bwd<float, float>(C, fgrid, st, x, gamma, mean, rstd, dy, dx, dgamma, dbeta, rows);
      else if (in_dtype == RSC_F32) done = ln_fast_bwd<float, __nv_bfloat16>(C, fgrid, st, x, gamma, mean, rstd, dy, dx, dgamma, dbeta, rows);
      else if (out_dtype == RSC_F32) done = ln_fast_bwd<__nv_bfloat16, float>(C, fgrid, st, x, gamma, mean, rstd, dy, dx, dgamma, dbeta, rows);
      else done = ln_fast_bwd<__nv_bfloat16, __nv_bfloat16>(C, fgrid, st, x, gamma, mean, rstd, dy, dx, dgamma, dbeta, rows);
      if (done) {
        RSC_CHECK_LAUNCH("rsc_layernorm_bwd");
        return RSC_OK;
      }
    }
  }
  int64_t blocks = (rows + 127) / 128;
  int grid = (int)(blocks < kNumSMs * 4 ? blocks : kNumSMs * 4);
  LN_DISPATCH(in_dtype, out_dtype, ln_bwd_launch, (C + 127) / 128, grid, (cudaStream_t)stream, x, gamma, mean, rstd, dy,
              dx, dgamma, dbeta, rows, C);
  RSC_CHECK_LAUNCH("rsc_layernorm_bwd");
  return RSC_OK;
}
