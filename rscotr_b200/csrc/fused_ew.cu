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
#include <cstdlib>
#include <stdlib.h>

#include "common.cuh"

namespace rsc {
namespace few {

constexpr int LN_THREADS = 128;

// CTAs per SM of the grid cap (tuning knob, read once from the environment)
static int grid_mult(const char *env, int dflt) {
  const char *v = getenv(env);
  const int m = v ? atoi(v) : 0;
  return m > 0 ? m : dflt;
}

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

// sample index of a row: one 32-bit division where the row count allows it (the emulated 64-bit division is ~60
// instructions per row group, a third of the forward's loop body)
__device__ __forceinline__ int64_t sample_of(int64_t r, int64_t rows, int64_t rows_per_sample) {
  if (rows < ((int64_t)1 << 31)) return (int64_t)((unsigned)r / (unsigned)rows_per_sample);
  return r / rows_per_sample;
}

template <int L>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Same sub-warp mapping as layernorm.cu: L lanes share a row (C = 8*EV*L), a warp covers 32/L rows per step.
template <typename T, int VW, int EV, int L>
__global__ void __launch_bounds__(LN_THREADS)
    add_ln_fwd_kernel(const T *__restrict__ identity, const T *__restrict__ x, const float *__restrict__ bias,
                      const float *__restrict__ scale, const float *__restrict__ gamma, const float *__restrict__ beta,
                      T *__restrict__ r_out, T *__restrict__ n_out, float *__restrict__ mean, float *__restrict__ rstd,
                      int64_t rows, int64_t rows_per_sample, float eps) {
  constexpr int C = VW * EV * L, RPW = 32 / L;
  const int lane = threadIdx.x & 31, sub = lane % L, rw = lane / L;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float ga[EV][VW], be[EV][VW], bi[EV][VW];
#pragma unroll
  for (int k = 0; k < EV; ++k) {
    loadv<float, VW>(gamma + (sub + k * L) * VW, ga[k]);
    loadv<float, VW>(beta + (sub + k * L) * VW, be[k]);
    if (bias) loadv<float, VW>(bias + (sub + k * L) * VW, bi[k]);
    else {
#pragma unroll
      for (int e = 0; e < VW; ++e) bi[k][e] = 0.f;
    }
  }
  for (int64_t r0 = warp * RPW; r0 < rows; r0 += nwarps * RPW) {
    const int64_t r = r0 + rw;
    const bool ok = r < rows;
    const float sc = (ok && scale) ? __ldg(scale + sample_of(r, rows, rows_per_sample)) : 1.0f;
    float v[EV][VW];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < EV; ++k) {
      float a[VW], b[VW];
      if (ok) {
        loadv<T, VW>(identity + r * C + (sub + k * L) * VW, a);
        loadv<T, VW>(x + r * C + (sub + k * L) * VW, b);
      }
#pragma unroll
      for (int e = 0; e < VW; ++e) {
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
        storev<T, VW>(r_out + r * C + (sub + k * L) * VW, v[k]);
        storev<T, VW>(n_out + r * C + (sub + k * L) * VW, o);
      }
    }
  }
}

// dr = dr_ext + LNbwd(dn) ; d_identity = dr ; dx = dr * scale ; dbias += colsum(dx) ; dgamma, dbeta
template <typename T, int VW, int EV, int L>
__global__ void __launch_bounds__(LN_THREADS)
    add_ln_bwd_kernel(const T *__restrict__ r_in, const float *__restrict__ gamma, const float *__restrict__ mean,
                      const float *__restrict__ rstd, const T *__restrict__ dn, const T *__restrict__ dr_ext,
                      const float *__restrict__ scale, T *__restrict__ d_identity, T *__restrict__ dx,
                      float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ dbias, int64_t rows,
                      int64_t rows_per_sample) {
  constexpr int C = VW * EV * L, RPW = 32 / L;
  __shared__ float red[3 * C];   // per-CTA partial d(gamma) | d(beta) | d(bias)
  const int lane = threadIdx.x & 31, sub = lane % L, rw = lane / L;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float ga[EV][VW], dg[EV][VW], db[EV][VW], dbi[EV][VW];
#pragma unroll
  for (int k = 0; k < EV; ++k) {
    loadv<float, VW>(gamma + (sub + k * L) * VW, ga[k]);
#pragma unroll
    for (int e = 0; e < VW; ++e) dg[k][e] = 0.f, db[k][e] = 0.f, dbi[k][e] = 0.f;
  }
  for (int64_t r0 = warp * RPW; r0 < rows; r0 += nwarps * RPW) {
    const int64_t r = r0 + rw;
    const bool ok = r < rows;
    const float mu = ok ? mean[r] : 0.f, rs = ok ? rstd[r] : 0.f;
    const float sc = (ok && scale) ? __ldg(scale + sample_of(r, rows, rows_per_sample)) : 1.0f;
    float xh[EV][VW], g[EV][VW], ex[EV][VW];
    float s1 = 0.f, s2 = 0.f;
    // all loads of the row are issued before the first reduction (memory-level parallelism)
#pragma unroll
    for (int k = 0; k < EV; ++k) {
      if (ok) {
        loadv<T, VW>(r_in + r * C + (sub + k * L) * VW, xh[k]);
        loadv<T, VW>(dn + r * C + (sub + k * L) * VW, g[k]);
        if (dr_ext) loadv<T, VW>(dr_ext + r * C + (sub + k * L) * VW, ex[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < EV; ++k) {
      float d[VW];
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        d[e] = g[k][e];
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
        for (int e = 0; e < VW; ++e) {
          o[e] = rs * (g[k][e] - s1 - xh[k][e] * s2) + (dr_ext ? ex[k][e] : 0.f);
        }
        storev<T, VW>(d_identity + r * C + (sub + k * L) * VW, o);
        if (dx || dbias) {
#pragma unroll
          for (int e = 0; e < VW; ++e) {
            o[e] = to_f<T>(from_f<T>(o[e])) * sc;   // dx = (stored dr) * scale
            dbi[k][e] += o[e];
          }
          if (dx) storev<T, VW>(dx + r * C + (sub + k * L) * VW, o);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < EV; ++k)
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      float a = dg[k][e], b = db[k][e], c = dbi[k][e];
#pragma unroll
      for (int o = 16; o >= L; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
      }
      if (rw == 0) {
        atomicAdd(red + (sub + k * L) * VW + e, a);
        atomicAdd(red + C + (sub + k * L) * VW + e, b);
        atomicAdd(red + 2 * C + (sub + k * L) * VW + e, c);
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[C + i]);
    if (dbias) atomicAdd(dbias + i, red[2 * C + i]);
  }
}

// (VW, EV, L) with VW*EV*L == C that keeps the fewest elements per lane (registers!) with 16-byte vectors
// preferred: 96 -> (4,3,8) | 192 -> (4,3,16) | 384 -> (4,3,32) | 768 -> (8,3,32) | 128 -> (8,1,16) | 256 -> (8,1,32)
// | 512 -> (8,2,32) | 1024 -> (8,4,32)
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

// ---------------------------------------------------------------------------------------------
// bias + GELU (erf).  Block = 32 column octets x 8 row lanes over a slab of rows.
// ---------------------------------------------------------------------------------------------
// Activations.  GELU is the exact erf form (torch.nn.GELU default).  These passes are ALU-bound with libm's erff
// (~45 instructions per element for value + derivative), so the bf16 path evaluates erf with Abramowitz & Stegun
// 7.1.26 (|error| < 1.5e-7, four orders below bf16 resolution) sharing ONE exponential between erf and the
// Gaussian density: u = exp(-x^2/2) = exp(-z^2) with z = x/sqrt(2).  fp32 keeps erff (the 1e-3 parity path).
// ACT_GELU_SIG (bf16 only, opt-in): Phi(x) ~ 1 / (1 + 2^(-x (c0 + c1 x^2 + c2 x^4))) -- a tanh-form fit of the normal CDF
// written as a logistic, |GELU error| < 2.6e-5 (two orders below bf16 resolution), a few percent relative accuracy
// in the left tail down to x = -5 (the exponential carries it), 9 instructions instead of 17 for the value.  x^2 is clamped at 49 (the
// quartic fit turns over beyond |x| ~ 8; at the clamp the argument is already +-35, i.e. Phi = 0 or 1 to 1e-11).
enum { ACT_GELU = 0, ACT_RELU = 1, ACT_GELU_SIG = 2 };

__device__ __forceinline__ float gelu_sig_cdf(float x) {
  const float x2 = fminf(x * x, 49.0f);
  float p = fmaf(x2, 0.0010142630552444944f, -0.10677572400272597f);
  p = fmaf(x2, p, -2.3011213394566354f);                            // -(c0 + c1 x^2 + c2 x^4)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + exp2f(x * p)));
  return r;
}

template <bool FAST>
__device__ __forceinline__ void gelu_parts(float x, float &cdf, float &u) {
  if (FAST) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));   // MUFU.RCP (an IEEE reciprocal is ~8 instructions; these passes are issue-bound)
    u = exp2f(-0.72134752044448170f * x * x);                      // exp(-x^2/2)
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float half_erfc = 0.5f * poly * t * u;                    // 0.5 * erfc(|z|)
    cdf = x >= 0.f ? 1.0f - half_erfc : half_erfc;
  } else {
    cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
    u = __expf(-0.5f * x * x);
  }
}

template <int ACT, bool FAST>
__device__ __forceinline__ float act_f(float x) {
  if (ACT == ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == ACT_GELU_SIG && FAST) return x * gelu_sig_cdf(x);
  float cdf, u;
  gelu_parts<FAST>(x, cdf, u);
  return x * cdf;
}
template <int ACT, bool FAST>
__device__ __forceinline__ float act_grad(float x) {
  if (ACT == ACT_RELU) return x > 0.f ? 1.f : 0.f;
  if (ACT == ACT_GELU_SIG && FAST) return fmaf(x * 0.3989422804014327f, exp2f(-0.72134752044448170f * x * x), gelu_sig_cdf(x));
  float cdf, u;
  gelu_parts<FAST>(x, cdf, u);
  return fmaf(x * 0.3989422804014327f, u, cdf);
}

template <typename T>
struct FastAct {
  static constexpr bool value = false;
};
template <>
struct FastAct<__nv_bfloat16> {
  static constexpr bool value = true;
};

// Flat mapping: the (rows, C) tensor is a stream of 8-element vectors; thread t owns vectors t, t + T, t + 2T ...
// with T = total threads a multiple of C/8, so that a thread always sees the SAME 8 columns (its bias values and
// its partial d(bias) sums live in registers) and every warp access is a contiguous 512-byte run.
template <typename T, int ACT>
__global__ void __launch_bounds__(256)
    bias_act_fwd_kernel(const T *__restrict__ h, const float *__restrict__ bias, T *__restrict__ y, int64_t nvec, int C8) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x, T_ = (int64_t)gridDim.x * 256;
  float b[8];
  load8<float>(bias + (t % C8) * 8, b);
  int64_t v = t;
  for (; v + T_ < nvec; v += 2 * T_) {   // two independent vectors in flight
    float a0[8], a1[8];
    load8<T>(h + v * 8, a0);
    load8<T>(h + (v + T_) * 8, a1);
#pragma unroll
    for (int e = 0; e < 8; ++e) a0[e] = act_f<ACT, FastAct<T>::value>(a0[e] + b[e]), a1[e] = act_f<ACT, FastAct<T>::value>(a1[e] + b[e]);
    store8<T>(y + v * 8, a0);
    store8<T>(y + (v + T_) * 8, a1);
  }
  if (v < nvec) {
    float a0[8];
    load8<T>(h + v * 8, a0);
#pragma unroll
    for (int e = 0; e < 8; ++e) a0[e] = act_f<ACT, FastAct<T>::value>(a0[e] + b[e]);
    store8<T>(y + v * 8, a0);
  }
}

template <typename T, int ACT>
__global__ void __launch_bounds__(256)
    bias_act_bwd_kernel(const T *__restrict__ h, const float *__restrict__ bias, const T *__restrict__ dy,
                         T *__restrict__ dh, float *__restrict__ dbias, int64_t nvec, int C8) {
  extern __shared__ float red[];   // [C] per-CTA partial d(bias)
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x, T_ = (int64_t)gridDim.x * 256;
  const int c8 = (int)(t % C8);
  for (int i = threadIdx.x; i < C8 * 8; i += 256) red[i] = 0.f;
  __syncthreads();
  float b[8], acc[8];
  load8<float>(bias + c8 * 8, b);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  int64_t v = t;
  for (; v + T_ < nvec; v += 2 * T_) {
    float a0[8], d0[8], a1[8], d1[8];
    load8<T>(h + v * 8, a0);
    load8<T>(dy + v * 8, d0);
    load8<T>(h + (v + T_) * 8, a1);
    load8<T>(dy + (v + T_) * 8, d1);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      d0[e] = to_f<T>(from_f<T>(d0[e] * act_grad<ACT, FastAct<T>::value>(a0[e] + b[e])));
      d1[e] = to_f<T>(from_f<T>(d1[e] * act_grad<ACT, FastAct<T>::value>(a1[e] + b[e])));
      acc[e] += d0[e] + d1[e];
    }
    store8<T>(dh + v * 8, d0);
    store8<T>(dh + (v + T_) * 8, d1);
  }
  if (v < nvec) {
    float a0[8], d0[8];
    load8<T>(h + v * 8, a0);
    load8<T>(dy + v * 8, d0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      d0[e] = to_f<T>(from_f<T>(d0[e] * act_grad<ACT, FastAct<T>::value>(a0[e] + b[e])));
      acc[e] += d0[e];
    }
    store8<T>(dh + v * 8, d0);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) atomicAdd(red + c8 * 8 + e, acc[e]);
  __syncthreads();
  for (int i = threadIdx.x; i < C8 * 8; i += 256) {
    const float s = red[i];
    if (s != 0.f) atomicAdd(dbias + i, s);
  }
}

}  // namespace few
}  // namespace rsc

using namespace rsc;

extern "C" int rsc_add_ln_supported(int C) {
  int vw, ev, l;
  return few::ln_shape(C, vw, ev, l) ? 1 : 0;
}

extern "C" int rsc_add_ln_fwd(const void *identity, const void *x, const float *bias, const float *scale,
                              const float *gamma, const float *beta, void *r_out, void *n_out, float *mean, float *rstd,
                              int64_t rows, int64_t rows_per_sample, int C, float eps, int dtype, void *stream) {
  RSC_CHECK_ARG(rows > 0 && rows_per_sample > 0, "rsc_add_ln_fwd: empty tensor");
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "rsc_add_ln_fwd: bad dtype %d", dtype);
  int vw = 0, ev = 0, l = 0;
  RSC_CHECK_ARG(few::ln_shape(C, vw, ev, l), "rsc_add_ln_fwd: unsupported channel count %d (96..1024 Swin / transformer widths)", C);
  RSC_CHECK_ARG(identity && x && gamma && beta && r_out && n_out && mean && rstd, "rsc_add_ln_fwd: null pointer");
  const int rpb = (few::LN_THREADS / 32) * (32 / l);
  int64_t fb = (rows + rpb - 1) / rpb;
  // 95 registers -> 5 resident CTAs / SM: for long tensors exactly one resident wave (measured +10 % at the stage-0 / stage-1
  // cls shapes, tools/lnbench.py); short ones finish sooner with more, smaller CTAs
  const int cap = kNumSMs * few::grid_mult("RSC_ADDLN_FWD_MULT", rows >= (1 << 17) ? 5 : 12);
  const int grid = (int)(fb < cap ? fb : cap);
  cudaStream_t st = (cudaStream_t)stream;
#define ALF(T, V, E, LL)                                                                                                \
  if (vw == V && ev == E && l == LL) {                                                                                  \
      few::add_ln_fwd_kernel<T, V, E, LL><<<grid, few::LN_THREADS, 0, st>>>((const T *)identity, (const T *)x, bias, scale, \
                                                                          gamma, beta, (T *)r_out, (T *)n_out, mean, rstd, \
                                                                          rows, rows_per_sample, eps);                  \
  }
#define ALF_ALL(T) \
  ALF(T, 4, 3, 8) ALF(T, 4, 3, 16) ALF(T, 4, 3, 32) ALF(T, 8, 3, 32) ALF(T, 8, 1, 16) ALF(T, 8, 1, 32) ALF(T, 8, 2, 32) ALF(T, 8, 4, 32)
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
  int vw = 0, ev = 0, l = 0;
  RSC_CHECK_ARG(few::ln_shape(C, vw, ev, l), "rsc_add_ln_bwd: unsupported channel count %d", C);
  RSC_CHECK_ARG(r && gamma && mean && rstd && dn && d_identity && dgamma && dbeta, "rsc_add_ln_bwd: null pointer");
  RSC_CHECK_ARG(!(scale && !dx), "rsc_add_ln_bwd: a per-sample scale needs a separate dx buffer");
  // every CTA ends with 3*C global atomics (d gamma | d beta | d bias): keep >= 32 rows per CTA so that small
  // tensors (13 294-token encoder maps) are not dominated by thousand-way same-address atomic contention
  int rpb = (few::LN_THREADS / 32) * (32 / l);
  if (rpb < 32) rpb = 32;
  int64_t fb = (rows + rpb - 1) / rpb;
  const int cap = kNumSMs * few::grid_mult("RSC_ADDLN_BWD_MULT", 8);
  const int grid = (int)(fb < cap ? fb : cap);
  cudaStream_t st = (cudaStream_t)stream;
#define ALB(T, V, E, LL)                                                                                               \
  if (vw == V && ev == E && l == LL) {                                                                                 \
    few::add_ln_bwd_kernel<T, V, E, LL><<<grid, few::LN_THREADS, 0, st>>>(                                                \
        (const T *)r, gamma, mean, rstd, (const T *)dn, (const T *)dr_ext, scale, (T *)d_identity, (T *)dx, dgamma,    \
        dbeta, dbias, rows, rows_per_sample);                                                                          \
  }
#define ALB_ALL(T) \
  ALB(T, 4, 3, 8) ALB(T, 4, 3, 16) ALB(T, 4, 3, 32) ALB(T, 8, 3, 32) ALB(T, 8, 1, 16) ALB(T, 8, 1, 32) ALB(T, 8, 2, 32) ALB(T, 8, 4, 32)
  if (dtype == RSC_F32) { ALB_ALL(float) } else { ALB_ALL(__nv_bfloat16) }
#undef ALB
#undef ALB_ALL
  RSC_CHECK_LAUNCH("rsc_add_ln_bwd");
  return RSC_OK;
}

// grid of 256-thread CTAs whose total thread count is a multiple of C/8 (see the kernels), ~8 CTAs per SM
static int bg_grid(int64_t nvec, int C8) {
  int a = C8, b = 256;
  while (b) {
    const int t = a % b;
    a = b, b = t;
  }
  const int m = C8 / a;                                   // grid must be a multiple of m
  int64_t want = (nvec + 16 * 256 - 1) / (16 * 256);      // >= 8 trips of two vectors per thread: the backward ends
                                                          // with C global atomics per CTA
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (want > cap) want = cap;
  want = (want + m - 1) / m * m;
  return (int)(want < m ? m : want);
}

static int bias_act_check(const char *fn, int64_t rows, int C, int act, int dtype) {
  RSC_CHECK_ARG(rows > 0 && C > 0 && C % 8 == 0 && C <= 8192, "%s: need rows > 0, C %% 8 == 0, C <= 8192 (rows=%lld, C=%d)", fn,
                (long long)rows, C);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "%s: bad dtype %d", fn, dtype);
  RSC_CHECK_ARG(act == few::ACT_GELU || act == few::ACT_RELU || (act == few::ACT_GELU_SIG && dtype == RSC_BF16),
                "%s: act must be 0 (gelu), 1 (relu) or 2 (gelu, logistic fit, bf16 only), got %d", fn, act);
  return RSC_OK;
}

extern "C" int rsc_bias_act_fwd(const void *h, const float *bias, void *y, int64_t rows, int C, int act, int dtype,
                                void *stream) {
  if (int e = bias_act_check("rsc_bias_act_fwd", rows, C, act, dtype)) return e;
  RSC_CHECK_ARG(h && bias && y, "rsc_bias_act_fwd: null pointer");
  const int64_t nvec = rows * (C / 8);
  const int grid = bg_grid(nvec, C / 8);
  cudaStream_t st = (cudaStream_t)stream;
#define BAF(T, A) few::bias_act_fwd_kernel<T, A><<<grid, 256, 0, st>>>((const T *)h, bias, (T *)y, nvec, C / 8)
  if (dtype == RSC_F32) {
    if (act == few::ACT_GELU) BAF(float, few::ACT_GELU); else BAF(float, few::ACT_RELU);
  } else {
    if (act == few::ACT_GELU) BAF(__nv_bfloat16, few::ACT_GELU);
    else if (act == few::ACT_GELU_SIG) BAF(__nv_bfloat16, few::ACT_GELU_SIG);
    else BAF(__nv_bfloat16, few::ACT_RELU);
  }
#undef BAF
  RSC_CHECK_LAUNCH("rsc_bias_act_fwd");
  return RSC_OK;
}

extern "C" int rsc_bias_act_bwd(const void *h, const float *bias, const void *dy, void *dh, float *dbias, int64_t rows,
                                int C, int act, int dtype, void *stream) {
  if (int e = bias_act_check("rsc_bias_act_bwd", rows, C, act, dtype)) return e;
  RSC_CHECK_ARG(h && bias && dy && dh && dbias, "rsc_bias_act_bwd: null pointer");
  const int64_t nvec = rows * (C / 8);
  const int grid = bg_grid(nvec, C / 8);
  const size_t smem = (size_t)C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define BAB(T, A) \
  few::bias_act_bwd_kernel<T, A><<<grid, 256, smem, st>>>((const T *)h, bias, (const T *)dy, (T *)dh, dbias, nvec, C / 8)
  if (dtype == RSC_F32) {
    if (act == few::ACT_GELU) BAB(float, few::ACT_GELU); else BAB(float, few::ACT_RELU);
  } else {
    if (act == few::ACT_GELU) BAB(__nv_bfloat16, few::ACT_GELU);
    else if (act == few::ACT_GELU_SIG) BAB(__nv_bfloat16, few::ACT_GELU_SIG);
    else BAB(__nv_bfloat16, few::ACT_RELU);
  }
#undef BAB
  RSC_CHECK_LAUNCH("rsc_bias_act_bwd");
  return RSC_OK;
}
