// Convolutions as im2col GEMMs + the PPM pooling (SURVEY 8a rows a8 / a20, 8b minimum export list):
//   rsc_im2col_{fwd,bwd}          channels-last patch gather (k x k, stride, zero padding) and its adjoint; the GEMM
//                                 itself is rsc_linear_{fwd,dx,dw} (csrc/gemm_tc.cu) on the gathered matrix.  A 1x1
//                                 convolution needs no gather at all: it IS the GEMM on the (B*H*W, C) token matrix.
//   rsc_adaptive_avgpool_{fwd,bwd} nn.AdaptiveAvgPool2d on channels-last maps (PPM scales 1, 2, 3, 6)
// Replaces nn.Conv2d (cuDNN) / nn.AdaptiveAvgPool2d (ATen) behind mmdet ChannelMapper (cfg :26-33), the FPN / mask
// feature convs of seg_head/pixel_decoder.py:39-64,158-170 and mmseg UPerHead / PPM.  HBM-bound gathers: every thread
// moves one 16-byte channel octet, a warp covers 256 consecutive channels of one tap -> fully coalesced both ways.
#include "common.cuh"

namespace rsc {
namespace conv {

// col (B*Ho*Wo, kh*kw*C) <- x (B,H,W,C): column index = (tap, c), tap = i*kw + j; elements of 16 bytes (VEC channels)
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
    im2col_kernel(const T *__restrict__ x, T *__restrict__ col, int B, int H, int W, int C, int kh, int kw, int stride, int pad,
                  int Ho, int Wo, int64_t total) {
  const int CV = C / VEC;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV);
    int64_t t = idx / CV;
    const int tap = (int)(t % (kh * kw));
    t /= kh * kw;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho), b = (int)(t / Ho);
    const int h = ho * stride - pad + tap / kw, w = wo * stride - pad + tap % kw;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (h >= 0 && h < H && w >= 0 && w < W)
      v = __ldg(reinterpret_cast<const uint4 *>(x + (((int64_t)b * H + h) * W + w) * C + cv * VEC));
    *reinterpret_cast<uint4 *>(col + idx * VEC) = v;
  }
}

// dx (B,H,W,C) = adjoint: every input pixel gathers the taps that read it (deterministic, no atomics)
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
    col2im_kernel(const T *__restrict__ dcol, T *__restrict__ dx, int B, int H, int W, int C, int kh, int kw, int stride, int pad,
                  int Ho, int Wo, int64_t total) {
  const int CV = C / VEC;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV);
    int64_t t = idx / CV;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H), b = (int)(t / H);
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    for (int i = 0; i < kh; ++i) {
      const int hn = h + pad - i;
      if (hn < 0 || hn % stride) continue;
      const int ho = hn / stride;
      if (ho >= Ho) continue;
      for (int j = 0; j < kw; ++j) {
        const int wn = w + pad - j;
        if (wn < 0 || wn % stride) continue;
        const int wo = wn / stride;
        if (wo >= Wo) continue;
        const T *src = dcol + ((((int64_t)b * Ho + ho) * Wo + wo) * (kh * kw) + i * kw + j) * C + cv * VEC;
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(src));
        const T *e4 = reinterpret_cast<const T *>(&q);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] += to_f<T>(e4[e]);
      }
    }
    uint4 o;
    T *oe = reinterpret_cast<T *>(&o);
#pragma unroll
    for (int e = 0; e < VEC; ++e) oe[e] = from_f<T>(acc[e]);
    *reinterpret_cast<uint4 *>(dx + idx * VEC) = o;
  }
}

// nn.AdaptiveAvgPool2d bins: [floor(i*H/S), ceil((i+1)*H/S))
__device__ __forceinline__ int bin_lo(int i, int n, int s) { return (i * n) / s; }
__device__ __forceinline__ int bin_hi(int i, int n, int s) { return ((i + 1) * n + s - 1) / s; }

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
    avgpool_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, int B, int H, int W, int C, int S, int64_t total) {
  const int CV = C / VEC;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV);
    int64_t t = idx / CV;
    const int j = (int)(t % S);
    t /= S;
    const int i = (int)(t % S), b = (int)(t / S);
    const int h0 = bin_lo(i, H, S), h1 = bin_hi(i, H, S), w0 = bin_lo(j, W, S), w1 = bin_hi(j, W, S);
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    for (int h = h0; h < h1; ++h)
      for (int w = w0; w < w1; ++w) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(x + (((int64_t)b * H + h) * W + w) * C + cv * VEC));
        const T *e4 = reinterpret_cast<const T *>(&q);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] += to_f<T>(e4[e]);
      }
    const float inv = 1.0f / (float)((h1 - h0) * (w1 - w0));
    uint4 o;
    T *oe = reinterpret_cast<T *>(&o);
#pragma unroll
    for (int e = 0; e < VEC; ++e) oe[e] = from_f<T>(acc[e] * inv);
    *reinterpret_cast<uint4 *>(y + idx * VEC) = o;
  }
}

// dx(b,h,w,:) = sum over the bins that contain (h,w) of dy(bin) / area(bin)   (bins overlap when H % S != 0)
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
    avgpool_bwd_kernel(const T *__restrict__ dy, T *__restrict__ dx, int B, int H, int W, int C, int S, int64_t total) {
  const int CV = C / VEC;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV);
    int64_t t = idx / CV;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H), b = (int)(t / H);
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    // bins that contain h: floor(i H / S) <= h < ceil((i+1) H / S)  <=>  floor(h S / H) <= i <= ceil((h+1) S / H) - 1
    const int i0 = (h * S) / H, i1 = ((h + 1) * S + H - 1) / H - 1, j0 = (w * S) / W, j1 = ((w + 1) * S + W - 1) / W - 1;
    for (int i = i0; i <= min(S - 1, i1); ++i) {
      const int h0 = bin_lo(i, H, S), h1 = bin_hi(i, H, S);
      for (int j = j0; j <= min(S - 1, j1); ++j) {
        const int w0 = bin_lo(j, W, S), w1 = bin_hi(j, W, S);
        const float inv = 1.0f / (float)((h1 - h0) * (w1 - w0));
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(dy + (((int64_t)b * S + i) * S + j) * C + cv * VEC));
        const T *e4 = reinterpret_cast<const T *>(&q);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = fmaf(to_f<T>(e4[e]), inv, acc[e]);
      }
    }
    uint4 o;
    T *oe = reinterpret_cast<T *>(&o);
#pragma unroll
    for (int e = 0; e < VEC; ++e) oe[e] = from_f<T>(acc[e]);
    *reinterpret_cast<uint4 *>(dx + idx * VEC) = o;
  }
}

static int grid_for(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(g < cap ? g : cap);
}

}  // namespace conv
}  // namespace rsc

using namespace rsc;

static int check_conv(const char *fn, const void *a, const void *b, int B, int H, int W, int C, int kh, int kw, int stride, int pad,
                      int dtype) {
  RSC_CHECK_ARG(a && b, "%s: null pointer", fn);
  RSC_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, "%s: bad geometry", fn);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "%s: bad dtype %d", fn, dtype);
  RSC_CHECK_ARG(C % (dtype == RSC_BF16 ? 8 : 4) == 0, "%s: C = %d must be a multiple of %d (16-byte channel groups)", fn, C,
                dtype == RSC_BF16 ? 8 : 4);
  RSC_CHECK_ARG((((uintptr_t)a | (uintptr_t)b) & 15) == 0, "%s: 16-byte alignment", fn);
  RSC_CHECK_ARG((H + 2 * pad - kh) / stride + 1 > 0 && (W + 2 * pad - kw) / stride + 1 > 0, "%s: empty output", fn);
  return 0;
}

// col (B*Ho*Wo, kh*kw*C) <- x (B,H,W,C) channels-last; Ho = (H + 2 pad - kh) / stride + 1 (nn.Conv2d, dilation 1)
extern "C" int rsc_im2col_fwd(const void *x, void *col, int B, int H, int W, int C, int kh, int kw, int stride, int pad, int dtype,
                              void *stream) {
  if (int e = check_conv("rsc_im2col_fwd", x, col, B, H, W, C, kh, kw, stride, pad, dtype)) return e;
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RSC_BF16) {
    const int64_t total = (int64_t)B * Ho * Wo * kh * kw * (C / 8);
    conv::im2col_kernel<__nv_bfloat16, 8><<<conv::grid_for(total), 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)col, B, H,
                                                                                 W, C, kh, kw, stride, pad, Ho, Wo, total);
  } else {
    const int64_t total = (int64_t)B * Ho * Wo * kh * kw * (C / 4);
    conv::im2col_kernel<float, 4><<<conv::grid_for(total), 256, 0, st>>>((const float *)x, (float *)col, B, H, W, C, kh, kw, stride, pad,
                                                                         Ho, Wo, total);
  }
  RSC_CHECK_LAUNCH("rsc_im2col_fwd");
  return RSC_OK;
}

// dx (B,H,W,C) <- dcol (B*Ho*Wo, kh*kw*C): the adjoint of rsc_im2col_fwd, fully written
extern "C" int rsc_im2col_bwd(const void *dcol, void *dx, int B, int H, int W, int C, int kh, int kw, int stride, int pad, int dtype,
                              void *stream) {
  if (int e = check_conv("rsc_im2col_bwd", dcol, dx, B, H, W, C, kh, kw, stride, pad, dtype)) return e;
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RSC_BF16) {
    const int64_t total = (int64_t)B * H * W * (C / 8);
    conv::col2im_kernel<__nv_bfloat16, 8><<<conv::grid_for(total), 256, 0, st>>>((const __nv_bfloat16 *)dcol, (__nv_bfloat16 *)dx, B, H,
                                                                                 W, C, kh, kw, stride, pad, Ho, Wo, total);
  } else {
    const int64_t total = (int64_t)B * H * W * (C / 4);
    conv::col2im_kernel<float, 4><<<conv::grid_for(total), 256, 0, st>>>((const float *)dcol, (float *)dx, B, H, W, C, kh, kw, stride,
                                                                         pad, Ho, Wo, total);
  }
  RSC_CHECK_LAUNCH("rsc_im2col_bwd");
  return RSC_OK;
}

// y (B,S,S,C) = AdaptiveAvgPool2d(S)(x (B,H,W,C)), channels-last
extern "C" int rsc_adaptive_avgpool_fwd(const void *x, void *y, int B, int H, int W, int C, int S, int dtype, void *stream) {
  if (int e = check_conv("rsc_adaptive_avgpool_fwd", x, y, B, H, W, C, 1, 1, 1, 0, dtype)) return e;
  RSC_CHECK_ARG(S > 0, "rsc_adaptive_avgpool_fwd: output size must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RSC_BF16) {
    const int64_t total = (int64_t)B * S * S * (C / 8);
    conv::avgpool_fwd_kernel<__nv_bfloat16, 8><<<conv::grid_for(total), 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y, B, H,
                                                                                      W, C, S, total);
  } else {
    const int64_t total = (int64_t)B * S * S * (C / 4);
    conv::avgpool_fwd_kernel<float, 4><<<conv::grid_for(total), 256, 0, st>>>((const float *)x, (float *)y, B, H, W, C, S, total);
  }
  RSC_CHECK_LAUNCH("rsc_adaptive_avgpool_fwd");
  return RSC_OK;
}

extern "C" int rsc_adaptive_avgpool_bwd(const void *dy, void *dx, int B, int H, int W, int C, int S, int dtype, void *stream) {
  if (int e = check_conv("rsc_adaptive_avgpool_bwd", dy, dx, B, H, W, C, 1, 1, 1, 0, dtype)) return e;
  RSC_CHECK_ARG(S > 0, "rsc_adaptive_avgpool_bwd: output size must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RSC_BF16) {
    const int64_t total = (int64_t)B * H * W * (C / 8);
    conv::avgpool_bwd_kernel<__nv_bfloat16, 8><<<conv::grid_for(total), 256, 0, st>>>((const __nv_bfloat16 *)dy, (__nv_bfloat16 *)dx, B,
                                                                                      H, W, C, S, total);
  } else {
    const int64_t total = (int64_t)B * H * W * (C / 4);
    conv::avgpool_bwd_kernel<float, 4><<<conv::grid_for(total), 256, 0, st>>>((const float *)dy, (float *)dx, B, H, W, C, S, total);
  }
  RSC_CHECK_LAUNCH("rsc_adaptive_avgpool_bwd");
  return RSC_OK;
}
