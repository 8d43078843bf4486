// LayerNorm forward / backward over the last dimension (norm1/norm2 of every
// Swin block, the per-stage output norms, the encoder / decoder post-norms).  Replaces ATen's
// native_layer_norm(+backward): one warp per row, 16-byte / 8-byte vector I/O, statistics in fp32,
// input and output element types independent (fp32 residual stream in -> bf16 GEMM operand out),
// d(gamma)/d(beta) accumulated per lane in registers over the persistent row loop, reduced through
// shared memory, one atomic per channel per CTA.
#include "common.cuh"

namespace rsc {

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
  int64_t blocks = (rows + LN_WARPS - 1) / LN_WARPS;
  int grid = (int)(blocks < kNumSMs * 4 ? blocks : kNumSMs * 4);
  LN_DISPATCH(in_dtype, out_dtype, ln_bwd_launch, (C + 127) / 128, grid, (cudaStream_t)stream, x, gamma, mean, rstd, dy,
              dx, dgamma, dbeta, rows, C);
  RSC_CHECK_LAUNCH("rsc_layernorm_bwd");
  return RSC_OK;
}
