// GroupNorm / BatchNorm (training statistics) on channels-last maps, with the ReLU that follows them in mmcv's
// ConvModule fused in (SURVEY 8a rows a8 ChannelMapper GN-32, a17 pixel-decoder GN-32 (+ReLU), a20 UPerHead / FCNHead
// conv-BN-ReLU).  Replaces nn.GroupNorm (ATen RowwiseMoments + elementwise chains on NCHW-strided tensors) and
// nn.BatchNorm2d (cuDNN) + nn.ReLU behind mmdet ChannelMapper (cfg MTL_slvlcls_...py:26-33),
// seg_head/pixel_decoder.py:39-64 and mmseg UPerHead.
//
// One formulation for both: x is (R, P, C) with C innermost; a statistic group is (row r, Cg consecutive channels) over
// the P pixels of the row.  GroupNorm(G): R = batch, P = H*W, Cg = C/G.  BatchNorm2d in training mode: R = 1,
// P = batch*H*W, Cg = 1.  HBM/L2-bound element-wise work: every thread moves 16-byte channel vectors; statistics are
// shifted sums (shift = the group's first element, so E[d^2] - E[d]^2 does not cancel) reduced through shared memory
// and a few global atomics per CTA.
//   fwd: stats (per-channel shifted sums) -> finalize (mean, rstd per group; BN running statistics) -> apply
//   bwd: reduce (per-channel sum dy, sum dy*xhat) -> finalize (group sums; d(gamma) / d(beta) accumulated) -> apply
#include "common.cuh"

namespace rsc {
namespace nrm {

template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float *p, float (&v)[4]) {
    const float4 r = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = r.x, v[1] = r.y, v[2] = r.z, v[3] = r.w;
  }
  __device__ static void store(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16 *p, float (&v)[8]) {
    const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w[i]));
      v[2 * i] = f.x, v[2 * i + 1] = f.y;
    }
  }
  __device__ static void store(__nv_bfloat16 *p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t *>(&h);
    }
    *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

constexpr int MAXC = 2048;      // channels (shared-memory reduction arrays)

// per-channel shifted sums of one pixel chunk of row r:  ws[r][c] += (sum (x - K), sum (x - K)^2),  K = x[r][0][group start]
template <typename T>
__global__ void __launch_bounds__(256) stats_kernel(const T *__restrict__ x, float *__restrict__ ws, int P, int C, int Cg, int chunk) {
  constexpr int VEC = Vec<T>::N;
  extern __shared__ float red[];      // [2][C]
  const int V = C / VEC, r = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += 256) red[i] = 0.f;
  __syncthreads();
  const T *xr = x + (int64_t)r * P * C;
  const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
  {      // one vector column per thread (V divides 256), 256 / V pixels per pass
    const int vi = threadIdx.x % V, po = threadIdx.x / V, pstep = 256 / V;
    float K[VEC], s1[VEC], s2[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      K[e] = to_f<T>(xr[(vi * VEC + e) / Cg * Cg]);
      s1[e] = s2[e] = 0.f;
    }
    for (int p = p0 + po; p < p1; p += pstep) {
      float v[VEC];
      Vec<T>::load(xr + (int64_t)p * C + vi * VEC, v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float d = v[e] - K[e];
        s1[e] += d;
        s2[e] += d * d;
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      atomicAdd(&red[vi * VEC + e], s1[e]);
      atomicAdd(&red[C + vi * VEC + e], s2[e]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    atomicAdd(&ws[((int64_t)r * C + c) * 2], red[c]);
    atomicAdd(&ws[((int64_t)r * C + c) * 2 + 1], red[C + c]);
  }
}

// thread per (row, group): mean / rstd; BatchNorm's running statistics (momentum update, unbiased variance)
template <typename T>
__global__ void finalize_fwd_kernel(const T *__restrict__ x, const float *__restrict__ ws, float *__restrict__ stats, int R, int P, int C,
                                    int Cg, float eps, float *run_mean, float *run_var, float momentum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, G = C / Cg;
  if (i >= R * G) return;
  const int r = i / G, g = i % G;
  float s1 = 0.f, s2 = 0.f;
  for (int c = g * Cg; c < (g + 1) * Cg; ++c) {
    s1 += ws[((int64_t)r * C + c) * 2];
    s2 += ws[((int64_t)r * C + c) * 2 + 1];
  }
  const float n = (float)P * Cg, K = to_f<T>(x[(int64_t)r * P * C + g * Cg]);
  const float m1 = s1 / n, var = fmaxf(s2 / n - m1 * m1, 0.f);
  stats[i * 2] = K + m1;
  stats[i * 2 + 1] = rsqrtf(var + eps);
  if (run_mean) {      // (BatchNorm: R == 1, Cg == 1, i == channel)
    run_mean[i] = (1.f - momentum) * run_mean[i] + momentum * (K + m1);
    run_var[i] = (1.f - momentum) * run_var[i] + momentum * var * (n / fmaxf(n - 1.f, 1.f));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) apply_fwd_kernel(const T *__restrict__ x, const float *__restrict__ stats,
                                                        const float *__restrict__ gamma, const float *__restrict__ beta, T *__restrict__ y,
                                                        int64_t total, int P, int C, int Cg, int relu) {
  constexpr int VEC = Vec<T>::N;
  const int V = C / VEC, G = C / Cg;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int vi = (int)(i % V);
    const int r = (int)(i / ((int64_t)V * P));
    float v[VEC];
    Vec<T>::load(x + i * VEC, v);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int c = vi * VEC + e;
      const float2 st = __ldg(reinterpret_cast<const float2 *>(stats) + r * G + c / Cg);
      float o = (v[e] - st.x) * st.y * __ldg(gamma + c) + __ldg(beta + c);
      v[e] = relu ? fmaxf(o, 0.f) : o;
    }
    Vec<T>::store(y + i * VEC, v);
  }
}

// wsb[r][c] += (sum dy', sum dy' * xhat) over one pixel chunk;  dy' = dy masked by the fused ReLU
template <typename T>
__global__ void __launch_bounds__(256) reduce_bwd_kernel(const T *__restrict__ x, const T *__restrict__ dy, const float *__restrict__ stats,
                                                         const float *__restrict__ gamma, const float *__restrict__ beta,
                                                         float *__restrict__ wsb, int P, int C, int Cg, int chunk, int relu) {
  constexpr int VEC = Vec<T>::N;
  extern __shared__ float red[];      // [2][C]
  const int V = C / VEC, G = C / Cg, r = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += 256) red[i] = 0.f;
  __syncthreads();
  const int vi = threadIdx.x % V, po = threadIdx.x / V, pstep = 256 / V;
  const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
  float mean[VEC], rstd[VEC], ga[VEC], be[VEC], a[VEC], b[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    const int c = vi * VEC + e;
    const float2 st = __ldg(reinterpret_cast<const float2 *>(stats) + r * G + c / Cg);
    mean[e] = st.x, rstd[e] = st.y, ga[e] = gamma[c], be[e] = beta[c];
    a[e] = b[e] = 0.f;
  }
  for (int p = p0 + po; p < p1; p += pstep) {
    const int64_t off = ((int64_t)r * P + p) * C + vi * VEC;
    float v[VEC], d[VEC];
    Vec<T>::load(x + off, v);
    Vec<T>::load(dy + off, d);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float xh = (v[e] - mean[e]) * rstd[e];
      const float g = (relu && xh * ga[e] + be[e] <= 0.f) ? 0.f : d[e];
      a[e] += g;
      b[e] += g * xh;
    }
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    atomicAdd(&red[vi * VEC + e], a[e]);
    atomicAdd(&red[C + vi * VEC + e], b[e]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    atomicAdd(&wsb[((int64_t)r * C + c) * 2], red[c]);
    atomicAdd(&wsb[((int64_t)r * C + c) * 2 + 1], red[C + c]);
  }
}

// thread per (row, group): gsum[r][g] = (sum_c gamma_c * sum dy', sum_c gamma_c * sum dy' xhat); threads of row 0 also fold
// the rows into d(beta) / d(gamma) (ACCUMULATED)
__global__ void finalize_bwd_kernel(const float *__restrict__ wsb, const float *__restrict__ gamma, float *__restrict__ gsum,
                                    float *__restrict__ dgamma, float *__restrict__ dbeta, int R, int C, int Cg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, G = C / Cg;
  if (i >= R * G) return;
  const int r = i / G, g = i % G;
  float sa = 0.f, sb = 0.f;
  for (int c = g * Cg; c < (g + 1) * Cg; ++c) {
    sa += gamma[c] * wsb[((int64_t)r * C + c) * 2];
    sb += gamma[c] * wsb[((int64_t)r * C + c) * 2 + 1];
  }
  gsum[i * 2] = sa;
  gsum[i * 2 + 1] = sb;
  if (r == 0)
    for (int c = g * Cg; c < (g + 1) * Cg; ++c) {
      float da = 0.f, db = 0.f;
      for (int rr = 0; rr < R; ++rr) {
        da += wsb[((int64_t)rr * C + c) * 2];
        db += wsb[((int64_t)rr * C + c) * 2 + 1];
      }
      dbeta[c] += da;
      dgamma[c] += db;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) apply_bwd_kernel(const T *__restrict__ x, const T *__restrict__ dy, const float *__restrict__ stats,
                                                        const float *__restrict__ gsum, const float *__restrict__ gamma,
                                                        const float *__restrict__ beta, T *__restrict__ dx, int64_t total, int P, int C,
                                                        int Cg, int relu) {
  constexpr int VEC = Vec<T>::N;
  const int V = C / VEC, G = C / Cg;
  const float inv_n = 1.f / ((float)P * Cg);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int vi = (int)(i % V);
    const int r = (int)(i / ((int64_t)V * P));
    float v[VEC], d[VEC];
    Vec<T>::load(x + i * VEC, v);
    Vec<T>::load(dy + i * VEC, d);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int c = vi * VEC + e;
      const float2 st = __ldg(reinterpret_cast<const float2 *>(stats) + r * G + c / Cg);
      const float2 gs = __ldg(reinterpret_cast<const float2 *>(gsum) + r * G + c / Cg);
      const float ga = __ldg(gamma + c), xh = (v[e] - st.x) * st.y;
      const float g = (relu && xh * ga + __ldg(beta + c) <= 0.f) ? 0.f : d[e];
      v[e] = st.y * (ga * g - (gs.x + xh * gs.y) * inv_n);
    }
    Vec<T>::store(dx + i * VEC, v);
  }
}

static int check(const char *fn, int R, int P, int C, int Cg, int dtype) {
  RSC_CHECK_ARG(R > 0 && P > 0 && C > 0 && Cg > 0 && C % Cg == 0, "%s: bad shape (R=%d,P=%d,C=%d,Cg=%d)", fn, R, P, C, Cg);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "%s: bad dtype %d", fn, dtype);
  const int vec = dtype == RSC_F32 ? 4 : 8;
  RSC_CHECK_ARG(C % vec == 0 && C <= MAXC && 256 % (C / vec) == 0, "%s: C=%d unsupported (C/%d must divide 256, C <= %d)", fn, C, vec, MAXC);
  return RSC_OK;
}

static inline int chunk_of(int P, int R) {
  // enough CTAs to fill the machine, at least 64 pixels each
  int chunks = (4 * kNumSMs + R - 1) / R;
  int chunk = (P + chunks - 1) / chunks;
  return chunk < 64 ? 64 : chunk;
}

static inline unsigned ew_grid(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace nrm
}  // namespace rsc

using namespace rsc;
using namespace rsc::nrm;

extern "C" int rsc_norm_supported(int C, int dtype) {
  const int vec = dtype == RSC_F32 ? 4 : 8;
  return (dtype == RSC_F32 || dtype == RSC_BF16) && C % vec == 0 && C <= MAXC && 256 % (C / vec) == 0;
}

extern "C" int rsc_groupnorm_fwd(const void *x, const float *gamma, const float *beta, void *y, float *stats, float *ws, int R, int P,
                                 int C, int Cg, float eps, int relu, float *run_mean, float *run_var, float momentum, int dtype,
                                 void *stream) {
  if (int e = check("rsc_groupnorm_fwd", R, P, C, Cg, dtype)) return e;
  RSC_CHECK_ARG(x && gamma && beta && y && stats && ws, "rsc_groupnorm_fwd: null pointer");
  RSC_CHECK_ARG(!run_mean || (R == 1 && Cg == 1 && run_var), "rsc_groupnorm_fwd: running statistics need R == 1, Cg == 1");
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(ws, 0, (size_t)R * C * 2 * sizeof(float), s) != cudaSuccess) {
    set_error("rsc_groupnorm_fwd: memset failed");
    return RSC_ERR_CUDA;
  }
  const int chunk = chunk_of(P, R), G = C / Cg;
  const dim3 grid((P + chunk - 1) / chunk, R);
  const int64_t total = (int64_t)R * P * C / (dtype == RSC_F32 ? 4 : 8);
  if (dtype == RSC_F32) {
    stats_kernel<float><<<grid, 256, 2 * C * sizeof(float), s>>>((const float *)x, ws, P, C, Cg, chunk);
    finalize_fwd_kernel<float><<<(R * G + 127) / 128, 128, 0, s>>>((const float *)x, ws, stats, R, P, C, Cg, eps, run_mean, run_var, momentum);
    apply_fwd_kernel<float><<<ew_grid(total), 256, 0, s>>>((const float *)x, stats, gamma, beta, (float *)y, total, P, C, Cg, relu);
  } else {
    using B16 = __nv_bfloat16;
    stats_kernel<B16><<<grid, 256, 2 * C * sizeof(float), s>>>((const B16 *)x, ws, P, C, Cg, chunk);
    finalize_fwd_kernel<B16><<<(R * G + 127) / 128, 128, 0, s>>>((const B16 *)x, ws, stats, R, P, C, Cg, eps, run_mean, run_var, momentum);
    apply_fwd_kernel<B16><<<ew_grid(total), 256, 0, s>>>((const B16 *)x, stats, gamma, beta, (B16 *)y, total, P, C, Cg, relu);
  }
  RSC_CHECK_LAUNCH("rsc_groupnorm_fwd");
  count_launch(2);
  return RSC_OK;
}

extern "C" int rsc_groupnorm_bwd(const void *x, const void *dy, const float *gamma, const float *beta, const float *stats, void *dx,
                                 float *dgamma, float *dbeta, float *ws, int R, int P, int C, int Cg, int relu, int dtype,
                                 void *stream) {
  if (int e = check("rsc_groupnorm_bwd", R, P, C, Cg, dtype)) return e;
  RSC_CHECK_ARG(x && dy && gamma && beta && stats && dx && dgamma && dbeta && ws, "rsc_groupnorm_bwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int G = C / Cg;
  // ws: [R][C][2] channel sums, then [R][G][2] group sums
  if (cudaMemsetAsync(ws, 0, (size_t)R * C * 2 * sizeof(float), s) != cudaSuccess) {
    set_error("rsc_groupnorm_bwd: memset failed");
    return RSC_ERR_CUDA;
  }
  float *gsum = ws + (size_t)R * C * 2;
  const int chunk = chunk_of(P, R);
  const dim3 grid((P + chunk - 1) / chunk, R);
  const int64_t total = (int64_t)R * P * C / (dtype == RSC_F32 ? 4 : 8);
  if (dtype == RSC_F32) {
    reduce_bwd_kernel<float><<<grid, 256, 2 * C * sizeof(float), s>>>((const float *)x, (const float *)dy, stats, gamma, beta, ws, P, C, Cg,
                                                                    chunk, relu);
    finalize_bwd_kernel<<<(R * G + 127) / 128, 128, 0, s>>>(ws, gamma, gsum, dgamma, dbeta, R, C, Cg);
    apply_bwd_kernel<float><<<ew_grid(total), 256, 0, s>>>((const float *)x, (const float *)dy, stats, gsum, gamma, beta, (float *)dx, total,
                                                         P, C, Cg, relu);
  } else {
    using B16 = __nv_bfloat16;
    reduce_bwd_kernel<B16><<<grid, 256, 2 * C * sizeof(float), s>>>((const B16 *)x, (const B16 *)dy, stats, gamma, beta, ws, P, C, Cg, chunk,
                                                                  relu);
    finalize_bwd_kernel<<<(R * G + 127) / 128, 128, 0, s>>>(ws, gamma, gsum, dgamma, dbeta, R, C, Cg);
    apply_bwd_kernel<B16><<<ew_grid(total), 256, 0, s>>>((const B16 *)x, (const B16 *)dy, stats, gsum, gamma, beta, (B16 *)dx, total, P, C,
                                                       Cg, relu);
  }
  RSC_CHECK_LAUNCH("rsc_groupnorm_bwd");
  count_launch(2);
  return RSC_OK;
}
