// Fused element-wise passes of the Swin block's residual stream (mmdet SwinBlock =
// x + DropPath(attn(LN(x))), x + DropPath(FFN(LN(x))) with FFN = Linear -> GELU -> Linear).  Every tensor here
// is streamed at HBM speed, so the only lever is the NUMBER of passes:
//
//   rsc_add_ln_{fwd,bwd}   r = identity + (x + bias) * scale[sample] ;  n = LayerNorm(r)
//       = bias add of the producing Linear + DropPath + residual add + the NEXT LayerNorm in one pass
//       (eager: addmm epilogue, mul, add, layer_norm = 5 tensor passes forward -> 4; backward LN-bwd, add, mul,
//       column sum = 9 passes -> 5, and the Linear's bias gradient comes out of the same pass).
//   rsc_bias_gelu_{fwd,bwd} y = gelu(h + bias) (erf form) ; backward also yields d(bias) = column sums of dh,
//       i.e. the bias gradient of the first FFN Linear without another pass over the 4C-wide tensor.
#include "common.cuh"

namespace rsc {
namespace few {

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

template <int L>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Same sub-warp mapping as layernorm.cu: L lanes share a row (C = 8*EV*L), a warp covers 32/L rows per step.
template <typename T, int EV, int L>
__global__ void __launch_bounds__(LN_THREADS)
    add_ln_fwd_kernel(const T *__restrict__ identity, const T *__restrict__ x, const float *__restrict__ bias,
                      const float *__restrict__ scale, const float *__restrict__ gamma, const float *__restrict__ beta,
                      T *__restrict__ r_out, T *__restrict__ n_out, float *__restrict__ mean, float *__restrict__ rstd,
                      int64_t rows, int64_t rows_per_sample, float eps) {
  constexpr int C = 8 * EV * L, RPW = 32 / L;
  const int lane = threadIdx.x & 31, sub = lane % L, rw = lane / L;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float ga[EV][8], be[EV][8], bi[EV][8];
#pragma unroll
  for (int k = 0; k < EV; ++k) {
    load8<float>(gamma + (sub + k * L) * 8, ga[k]);
    load8<float>(beta + (sub + k * L) * 8, be[k]);
    if (bias) load8<float>(bias + (sub + k * L) * 8, bi[k]);
    else {
#pragma unroll
      for (int e = 0; e < 8; ++e) bi[k][e] = 0.f;
    }
  }
  for (int64_t r0 = warp * RPW; r0 < rows; r0 += nwarps * RPW) {
    const int64_t r = r0 + rw;
    const bool ok = r < rows;
    const float sc = (ok && scale) ? __ldg(scale + r / rows_per_sample) : 1.0f;
    float v[EV][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < EV; ++k) {
      float a[8], b[8];
      if (ok) {
        load8<T>(identity + r * C + (sub + k * L) * 8, a);
        load8<T>(x + r * C + (sub + k * L) * 8, b);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[k][e] = ok ? fmaf(b[e] + bi[k][e], sc, a[e]) : 0.f;
        // the stored residual is the (possibly bf16-rounded) value the rest of the network sees:
        // normalise exactly that value so that forward and backward agree
        v[k][e] = to_f<T>(from_f<T>(v[k][e]));
        s += v[k][e];
      }
    }
    const float mu = group_sum<L>(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < EV; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
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
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf((v[k][e] - mu) * rs, ga[k][e], be[k][e]);
        store8<T>(r_out + r * C + (sub + k * L) * 8, v[k]);
        store8<T>(n_out + r * C + (sub + k * L) * 8, o);
      }
    }
  }
}

// dr = dr_ext + LNbwd(dn) ; d_identity = dr ; dx = dr * scale ; dbias += colsum(dx) ; dgamma, dbeta
template <typename T, int EV, int L>
__global__ void __launch_bounds__(LN_THREADS)
    add_ln_bwd_kernel(const T *__restrict__ r_in, const float *__restrict__ gamma, const float *__restrict__ mean,
                      const float *__restrict__ rstd, const T *__restrict__ dn, const T *__restrict__ dr_ext,
                      const float *__restrict__ scale, T *__restrict__ d_identity, T *__restrict__ dx,
                      float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ dbias, int64_t rows,
                      int64_t rows_per_sample) {
  constexpr int C = 8 * EV * L, RPW = 32 / L;
  __shared__ float red[3 * C];   // per-CTA partial d(gamma) | d(beta) | d(bias)
  const int lane = threadIdx.x & 31, sub = lane % L, rw = lane / L;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float ga[EV][8], dg[EV][8], db[EV][8], dbi[EV][8];
#pragma unroll
  for (int k = 0; k < EV; ++k) {
    load8<float>(gamma + (sub + k * L) * 8, ga[k]);
#pragma unroll
    for (int e = 0; e < 8; ++e) dg[k][e] = 0.f, db[k][e] = 0.f, dbi[k][e] = 0.f;
  }
  for (int64_t r0 = warp * RPW; r0 < rows; r0 += nwarps * RPW) {
    const int64_t r = r0 + rw;
    const bool ok = r < rows;
    const float mu = ok ? mean[r] : 0.f, rs = ok ? rstd[r] : 0.f;
    const float sc = (ok && scale) ? __ldg(scale + r / rows_per_sample) : 1.0f;
    float xh[EV][8], g[EV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < EV; ++k) {
      float d[8];
      if (ok) {
        load8<T>(r_in + r * C + (sub + k * L) * 8, xh[k]);
        load8<T>(dn + r * C + (sub + k * L) * 8, d);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
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
        float o[8], e8[8];
        if (dr_ext) load8<T>(dr_ext + r * C + (sub + k * L) * 8, e8);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o[e] = rs * (g[k][e] - s1 - xh[k][e] * s2) + (dr_ext ? e8[e] : 0.f);
        }
        store8<T>(d_identity + r * C + (sub + k * L) * 8, o);
        if (dx || dbias) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            o[e] = to_f<T>(from_f<T>(o[e])) * sc;   // dx = (stored dr) * scale
            dbi[k][e] += o[e];
          }
          if (dx) store8<T>(dx + r * C + (sub + k * L) * 8, o);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < EV; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a = dg[k][e], b = db[k][e], c = dbi[k][e];
#pragma unroll
      for (int o = 16; o >= L; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
      }
      if (rw == 0) {
        atomicAdd(red + (sub + k * L) * 8 + e, a);
        atomicAdd(red + C + (sub + k * L) * 8 + e, b);
        atomicAdd(red + 2 * C + (sub + k * L) * 8 + e, c);
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[C + i]);
    if (dbias) atomicAdd(dbias + i, red[2 * C + i]);
  }
}

static bool ln_shape(int C, int &ev, int &l) {
  for (int e = 3; e <= 4; ++e)
    for (int ll = 4; ll <= 32; ll *= 2)
      if (C == 8 * e * ll) {
        ev = e, l = ll;
        return true;
      }
  return false;
}

// ---------------------------------------------------------------------------------------------
// bias + GELU (erf).  Block = 32 column octets x 8 row lanes over a slab of rows.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

template <typename T>
__global__ void __launch_bounds__(256)
    bias_gelu_fwd_kernel(const T *__restrict__ h, const float *__restrict__ bias, T *__restrict__ y, int64_t rows, int C) {
  const int c = (blockIdx.x * 32 + threadIdx.x) * 8;
  if (c >= C) return;
  float b[8];
  load8<float>(bias + c, b);
  for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < rows; r += (int64_t)gridDim.y * 8) {
    float v[8];
    load8<T>(h + r * C + c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = gelu_f(v[e] + b[e]);
    store8<T>(y + r * C + c, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    bias_gelu_bwd_kernel(const T *__restrict__ h, const float *__restrict__ bias, const T *__restrict__ dy,
                         T *__restrict__ dh, float *__restrict__ dbias, int64_t rows, int C) {
  __shared__ float red[8][32][8 + 1];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (c < C) {
    float b[8];
    load8<float>(bias + c, b);
    for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < rows; r += (int64_t)gridDim.y * 8) {
      float v[8], d[8];
      load8<T>(h + r * C + c, v);
      load8<T>(dy + r * C + c, d);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        d[e] = to_f<T>(from_f<T>(d[e] * gelu_grad(v[e] + b[e])));
        acc[e] += d[e];
      }
      store8<T>(dh + r * C + c, d);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.y][threadIdx.x][e] = acc[e];
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x][e];
      atomicAdd(dbias + c + e, s);
    }
  }
}

}  // namespace few
}  // namespace rsc

using namespace rsc;

extern "C" int rsc_add_ln_supported(int C) {
  int ev, l;
  return few::ln_shape(C, ev, l) ? 1 : 0;
}

extern "C" int rsc_add_ln_fwd(const void *identity, const void *x, const float *bias, const float *scale,
                              const float *gamma, const float *beta, void *r_out, void *n_out, float *mean, float *rstd,
                              int64_t rows, int64_t rows_per_sample, int C, float eps, int dtype, void *stream) {
  RSC_CHECK_ARG(rows > 0 && rows_per_sample > 0, "rsc_add_ln_fwd: empty tensor");
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "rsc_add_ln_fwd: bad dtype %d", dtype);
  int ev = 0, l = 0;
  RSC_CHECK_ARG(few::ln_shape(C, ev, l), "rsc_add_ln_fwd: unsupported channel count %d (need 8*{3,4}*{4,8,16,32})", C);
  RSC_CHECK_ARG(identity && x && gamma && beta && r_out && n_out && mean && rstd, "rsc_add_ln_fwd: null pointer");
  const int rpb = (few::LN_THREADS / 32) * (32 / l);
  int64_t fb = (rows + rpb - 1) / rpb;
  const int grid = (int)(fb < kNumSMs * 12 ? fb : kNumSMs * 12);
  cudaStream_t st = (cudaStream_t)stream;
#define ALF(T, E, LL)                                                                                                   \
  if (ev == E && l == LL) {                                                                                             \
    few::add_ln_fwd_kernel<T, E, LL><<<grid, few::LN_THREADS, 0, st>>>((const T *)identity, (const T *)x, bias, scale,   \
                                                                        gamma, beta, (T *)r_out, (T *)n_out, mean, rstd, \
                                                                        rows, rows_per_sample, eps);                    \
  }
#define ALF_ALL(T) ALF(T, 3, 4) ALF(T, 3, 8) ALF(T, 3, 16) ALF(T, 3, 32) ALF(T, 4, 4) ALF(T, 4, 8) ALF(T, 4, 16) ALF(T, 4, 32)
  if (dtype == RSC_F32) { ALF_ALL(float) } else { ALF_ALL(__nv_bfloat16) }
#undef ALF
#undef ALF_ALL
  RSC_CHECK_LAUNCH("rsc_add_ln_fwd");
  return RSC_OK;
}

extern "C" int rsc_add_ln_bwd(const void *r, const float *gamma, const float *mean, const float *rstd, const void *dn,
                              const void *dr_ext, const float *scale, void *d_identity, void *dx, float *dgamma,
                              float *dbeta, float *dbias, int64_t rows, int64_t rows_per_sample, int C, int dtype,
                              void *stream) {
  RSC_CHECK_ARG(rows > 0 && rows_per_sample > 0, "rsc_add_ln_bwd: empty tensor");
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "rsc_add_ln_bwd: bad dtype %d", dtype);
  int ev = 0, l = 0;
  RSC_CHECK_ARG(few::ln_shape(C, ev, l), "rsc_add_ln_bwd: unsupported channel count %d", C);
  RSC_CHECK_ARG(r && gamma && mean && rstd && dn && d_identity && dgamma && dbeta, "rsc_add_ln_bwd: null pointer");
  RSC_CHECK_ARG(!(scale && !dx), "rsc_add_ln_bwd: a per-sample scale needs a separate dx buffer");
  const int rpb = (few::LN_THREADS / 32) * (32 / l);
  int64_t fb = (rows + rpb - 1) / rpb;
  const int grid = (int)(fb < kNumSMs * 6 ? fb : kNumSMs * 6);
  cudaStream_t st = (cudaStream_t)stream;
#define ALB(T, E, LL)                                                                                                  \
  if (ev == E && l == LL) {                                                                                            \
    few::add_ln_bwd_kernel<T, E, LL><<<grid, few::LN_THREADS, 0, st>>>(                                                \
        (const T *)r, gamma, mean, rstd, (const T *)dn, (const T *)dr_ext, scale, (T *)d_identity, (T *)dx, dgamma,    \
        dbeta, dbias, rows, rows_per_sample);                                                                          \
  }
#define ALB_ALL(T) ALB(T, 3, 4) ALB(T, 3, 8) ALB(T, 3, 16) ALB(T, 3, 32) ALB(T, 4, 4) ALB(T, 4, 8) ALB(T, 4, 16) ALB(T, 4, 32)
  if (dtype == RSC_F32) { ALB_ALL(float) } else { ALB_ALL(__nv_bfloat16) }
#undef ALB
#undef ALB_ALL
  RSC_CHECK_LAUNCH("rsc_add_ln_bwd");
  return RSC_OK;
}

static int bg_grid_y(int64_t rows, int cblocks) {
  int64_t want = (int64_t)kNumSMs * 8 / cblocks;
  if (want < 1) want = 1;
  int64_t slabs = (rows + 7) / 8;
  return (int)(slabs < want ? slabs : want);
}

extern "C" int rsc_bias_gelu_fwd(const void *h, const float *bias, void *y, int64_t rows, int C, int dtype, void *stream) {
  RSC_CHECK_ARG(rows > 0 && C > 0 && C % 8 == 0, "rsc_bias_gelu_fwd: need rows > 0, C %% 8 == 0 (rows=%lld, C=%d)",
                (long long)rows, C);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "rsc_bias_gelu_fwd: bad dtype %d", dtype);
  RSC_CHECK_ARG(h && bias && y, "rsc_bias_gelu_fwd: null pointer");
  const int cblocks = (C + 255) / 256;
  dim3 grid(cblocks, bg_grid_y(rows, cblocks)), block(32, 8);
  if (dtype == RSC_F32)
    few::bias_gelu_fwd_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float *)h, bias, (float *)y, rows, C);
  else
    few::bias_gelu_fwd_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)h, bias, (__nv_bfloat16 *)y, rows, C);
  RSC_CHECK_LAUNCH("rsc_bias_gelu_fwd");
  return RSC_OK;
}

extern "C" int rsc_bias_gelu_bwd(const void *h, const float *bias, const void *dy, void *dh, float *dbias, int64_t rows,
                                 int C, int dtype, void *stream) {
  RSC_CHECK_ARG(rows > 0 && C > 0 && C % 8 == 0, "rsc_bias_gelu_bwd: need rows > 0, C %% 8 == 0 (rows=%lld, C=%d)",
                (long long)rows, C);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "rsc_bias_gelu_bwd: bad dtype %d", dtype);
  RSC_CHECK_ARG(h && bias && dy && dh && dbias, "rsc_bias_gelu_bwd: null pointer");
  const int cblocks = (C + 255) / 256;
  dim3 grid(cblocks, bg_grid_y(rows, cblocks)), block(32, 8);
  if (dtype == RSC_F32)
    few::bias_gelu_bwd_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float *)h, bias, (const float *)dy,
                                                                                 (float *)dh, dbias, rows, C);
  else
    few::bias_gelu_bwd_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)h, bias, (const __nv_bfloat16 *)dy, (__nv_bfloat16 *)dh, dbias, rows, C);
  RSC_CHECK_LAUNCH("rsc_bias_gelu_bwd");
  return RSC_OK;
}
