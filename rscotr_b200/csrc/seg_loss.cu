// Fused bilinear upsample (align_corners=False) + softmax cross-entropy of the segmentation head
// (SURVEY 8a row a19, 8f rank 2).  Reference: mmseg 0.28 BaseDecodeHead.losses called from
// models/multi/seg_head/mask2former_head.py:204 -- resize(seg_logit, label size) materialises a
// (B,100,800,800) tensor (256 MB / image in fp32) and CrossEntropyLoss streams it three more times.
// Here the upsampled logits never exist: every output pixel interpolates its 4 taps per channel from a
// shared-memory patch of the low-resolution logits, and the backward is a GATHER per low-resolution
// pixel (deterministic, no atomics on the gradient).
//
//   rsc_upsample_ce_fwd : logits (B,C,h,w) -> stats[3] += {sum of per-pixel CE, #correct, #valid},
//                         lse (B,H,W) saved for the backward (+inf marks ignored pixels)
//   rsc_upsample_ce_bwd : dlogits (B,C,h,w) = scale[0] * sum_{output pixels} w * (softmax - onehot)
#include <float.h>

#include "common.cuh"

namespace rsc {
namespace segl {

__device__ __forceinline__ void bil_src(int o, float scale, int in, int &i0, int &i1, float &lam) {
  float s = scale * (o + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 < in - 1 ? i0 + 1 : i0;
  lam = s - i0;
}

constexpr int TX = 32, TY = 8;

template <typename T>
__global__ void __launch_bounds__(TX *TY)
    upsample_ce_fwd_kernel(const T *__restrict__ logits, const int64_t *__restrict__ label, float *__restrict__ lse_out,
                           float *__restrict__ stats, int C, int h, int w, int H, int W, float sh, float sw,
                           int ignore_index, int patch_cap) {
  extern __shared__ float patch[];   // [C][ph][pw]
  __shared__ float red[3][TX * TY / 32];
  const int b = blockIdx.z;
  const int ox0 = blockIdx.x * TX, oy0 = blockIdx.y * TY;
  const int ox_last = min(ox0 + TX, W) - 1, oy_last = min(oy0 + TY, H) - 1;
  int py0, py1, px0, px1, t0, t1;
  float tl;
  bil_src(oy0, sh, h, py0, t1, tl);
  bil_src(oy_last, sh, h, t0, py1, tl);
  bil_src(ox0, sw, w, px0, t1, tl);
  bil_src(ox_last, sw, w, t0, px1, tl);
  const int ph = py1 - py0 + 1, pw = px1 - px0 + 1;
  const int tid = threadIdx.y * TX + threadIdx.x;
  const T *src = logits + (int64_t)b * C * h * w;
  const int pp = ph * pw;
  if (pp * C > patch_cap) __trap();   // host sized the patch for this scale; cannot happen
  for (int idx = tid; idx < pp * C; idx += TX * TY) {
    const int c = idx / pp, r = idx - c * pp;
    const int yy = r / pw, xx = r - yy * pw;
    patch[idx] = to_f<T>(src[((int64_t)c * h + py0 + yy) * w + px0 + xx]);
  }
  __syncthreads();
  const int ox = ox0 + threadIdx.x, oy = oy0 + threadIdx.y;
  float ce = 0.f, correct = 0.f, valid = 0.f;
  if (ox < W && oy < H) {
    int y0, y1, x0, x1;
    float ly, lx;
    bil_src(oy, sh, h, y0, y1, ly);
    bil_src(ox, sw, w, x0, x1, lx);
    const int o00 = (y0 - py0) * pw + (x0 - px0), o01 = (y0 - py0) * pw + (x1 - px0);
    const int o10 = (y1 - py0) * pw + (x0 - px0), o11 = (y1 - py0) * pw + (x1 - px0);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const int64_t pix = ((int64_t)b * H + oy) * W + ox;
    const int64_t lab64 = label[pix];
    const bool ok = lab64 != ignore_index && lab64 >= 0 && lab64 < C;
    const int lab = ok ? (int)lab64 : -1;
    float m = -INFINITY, vlab = 0.f;
    int amax = 0;
    for (int c = 0; c < C; ++c) {
      const float *p = patch + c * pp;
      const float v = w00 * p[o00] + w01 * p[o01] + w10 * p[o10] + w11 * p[o11];
      if (v > m) m = v, amax = c;
      if (c == lab) vlab = v;
    }
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float *p = patch + c * pp;
      const float v = w00 * p[o00] + w01 * p[o01] + w10 * p[o10] + w11 * p[o11];
      s += expf(v - m);
    }
    const float lse = m + logf(s);
    lse_out[pix] = ok ? lse : INFINITY;
    if (ok) ce = lse - vlab, valid = 1.f, correct = amax == lab ? 1.f : 0.f;
  }
  ce = warp_sum(ce), correct = warp_sum(correct), valid = warp_sum(valid);
  if ((tid & 31) == 0) red[0][tid >> 5] = ce, red[1][tid >> 5] = correct, red[2][tid >> 5] = valid;
  __syncthreads();
  if (tid < 3) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < TX * TY / 32; ++k) s += red[tid][k];
    if (s != 0.f) atomicAdd(stats + tid, s);
  }
}

// one warp per low-resolution pixel, lanes over channels; the (<= FP x FP) footprint of output pixels that
// read this pixel is staged per warp in shared memory (lse, label, per-axis tap geometry)
constexpr int BW = 8;     // warps per block
constexpr int FP = 24;    // max footprint per axis (2*scale + 5 -> scale up to 9.5)

template <typename T>
__global__ void __launch_bounds__(BW * 32)
    upsample_ce_bwd_kernel(const T *__restrict__ logits, const int64_t *__restrict__ label, const float *__restrict__ lse,
                           const float *__restrict__ gscale, T *__restrict__ dlogits, int B, int C, int h, int w, int H,
                           int W, float sh, float sw, int64_t npix) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per-warp staging: lse[FP*FP], lab[FP*FP], then per-axis {r0 (0/1), r1 (1/2), lam, wgt} for y and x
  float *s_lse = sm + (size_t)warp * (2 * FP * FP + 8 * FP);
  int *s_lab = reinterpret_cast<int *>(s_lse + FP * FP);
  float *s_y = s_lse + 2 * FP * FP;   // [4][FP]
  float *s_x = s_y + 4 * FP;          // [4][FP]
  const float g = gscale[0];
  const float rh = 1.f / sh, rw = 1.f / sw;
  for (int64_t sp = (int64_t)blockIdx.x * BW + warp; sp < npix; sp += (int64_t)gridDim.x * BW) {
    const int sx = (int)(sp % w), sy = (int)((sp / w) % h), b = (int)(sp / ((int64_t)w * h));
    int oy0 = (int)floorf((sy - 1 + 0.5f) * rh - 0.5f) - 1, oy1 = (int)ceilf((sy + 1 + 0.5f) * rh - 0.5f) + 1;
    int ox0 = (int)floorf((sx - 1 + 0.5f) * rw - 0.5f) - 1, ox1 = (int)ceilf((sx + 1 + 0.5f) * rw - 0.5f) + 1;
    if (sy == 0) oy0 = 0;
    if (sx == 0) ox0 = 0;
    if (sy == h - 1) oy1 = H - 1;
    if (sx == w - 1) ox1 = W - 1;
    oy0 = max(oy0, 0), ox0 = max(ox0, 0), oy1 = min(oy1, H - 1), ox1 = min(ox1, W - 1);
    // trim the conservative range to the rows / columns that really read (sy, sx)
    const int ny = oy1 - oy0 + 1, nx = ox1 - ox0 + 1;   // <= FP (checked on the host)
    __syncwarp();
    for (int k = lane; k < ny + nx; k += 32) {
      const bool isy = k < ny;
      const int o = isy ? oy0 + k : ox0 + (k - ny);
      int i0, i1;
      float lam;
      bil_src(o, isy ? sh : sw, isy ? h : w, i0, i1, lam);
      const int s = isy ? sy : sx;
      float wgt = 0.f;
      if (i0 == s) wgt += 1.f - lam;
      if (i1 == s) wgt += lam;
      float *d = isy ? s_y + k : s_x + (k - ny);
      d[0] = (float)(i0 - (s - 1));        // 0 or 1 when wgt != 0
      d[FP] = (float)(i1 - (s - 1));       // 1 or 2
      d[2 * FP] = lam;
      d[3 * FP] = wgt;
    }
    for (int k = lane; k < ny * nx; k += 32) {
      const int yy = k / nx, xx = k - yy * nx;
      const int64_t pix = ((int64_t)b * H + oy0 + yy) * W + ox0 + xx;
      const float l = lse[pix];
      s_lse[yy * FP + xx] = l;
      s_lab[yy * FP + xx] = l == INFINITY ? -1 : (int)label[pix];
    }
    __syncwarp();
    const int ym = max(sy - 1, 0), yp = min(sy + 1, h - 1), xm = max(sx - 1, 0), xp = min(sx + 1, w - 1);
    for (int c = lane; c < C; c += 32) {
      const T *p = logits + ((int64_t)b * C + c) * h * w;
      float L[3][3];
      const int ys[3] = {ym, sy, yp}, xs[3] = {xm, sx, xp};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int d = 0; d < 3; ++d) L[a][d] = to_f<T>(p[ys[a] * w + xs[d]]);
      float acc = 0.f;
      for (int yy = 0; yy < ny; ++yy) {
        const float wy = s_y[3 * FP + yy];
        if (wy == 0.f) continue;
        const bool a0 = s_y[yy] != 0.f, a1 = s_y[FP + yy] == 2.f;
        const float ly = s_y[2 * FP + yy];
        float r[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) r[d] = (1.f - ly) * (a0 ? L[1][d] : L[0][d]) + ly * (a1 ? L[2][d] : L[1][d]);
        float row = 0.f;
        for (int xx = 0; xx < nx; ++xx) {
          const float wx = s_x[3 * FP + xx];
          if (wx == 0.f) continue;
          const bool b0 = s_x[xx] != 0.f, b1 = s_x[FP + xx] == 2.f;
          const float lx = s_x[2 * FP + xx];
          const float v = (1.f - lx) * (b0 ? r[1] : r[0]) + lx * (b1 ? r[2] : r[1]);
          const float pr = __expf(v - s_lse[yy * FP + xx]);            // 0 for ignored pixels (lse = +inf)
          row = fmaf(wx, pr - (s_lab[yy * FP + xx] == c ? 1.f : 0.f), row);
        }
        acc = fmaf(wy, row, acc);
      }
      dlogits[((int64_t)b * C + c) * h * w + (int64_t)sy * w + sx] = from_f<T>(acc * g);
    }
  }
}

}  // namespace segl
}  // namespace rsc

using namespace rsc;

static int uce_check(const char *fn, int B, int C, int h, int w, int H, int W, int dtype) {
  RSC_CHECK_ARG(B > 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0, "%s: bad shape (B=%d,C=%d,h=%d,w=%d,H=%d,W=%d)", fn, B, C, h,
                w, H, W);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "%s: bad dtype %d", fn, dtype);
  RSC_CHECK_ARG(H >= h && W >= w, "%s: only up-sampling is fused (got %dx%d -> %dx%d)", fn, h, w, H, W);
  const float rh = (float)H / h, rw = (float)W / w;
  RSC_CHECK_ARG(2.f * rh + 5.f <= segl::FP && 2.f * rw + 5.f <= segl::FP, "%s: scale factor too large (max 9.5x)", fn);
  return RSC_OK;
}

extern "C" int rsc_upsample_ce_fwd(const void *logits, const int64_t *label, float *lse, float *stats, int B, int C, int h,
                                   int w, int H, int W, int ignore_index, int dtype, void *stream) {
  if (int e = uce_check("rsc_upsample_ce_fwd", B, C, h, w, H, W, dtype)) return e;
  RSC_CHECK_ARG(logits && label && lse && stats, "rsc_upsample_ce_fwd: null pointer");
  const float sh = (float)h / H, sw = (float)w / W;
  const int ph = (int)(segl::TY * sh) + 3, pw = (int)(segl::TX * sw) + 3;
  const int cap = ph * pw * C;
  const size_t smem = (size_t)cap * sizeof(float);
  RSC_CHECK_ARG(smem <= 160 * 1024, "rsc_upsample_ce_fwd: logits patch (%d x %d x %d) does not fit shared memory", ph, pw, C);
  dim3 grid((W + segl::TX - 1) / segl::TX, (H + segl::TY - 1) / segl::TY, B), block(segl::TX, segl::TY);
  if (dtype == RSC_F32) {
    auto k = segl::upsample_ce_fwd_kernel<float>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, block, smem, (cudaStream_t)stream>>>((const float *)logits, label, lse, stats, C, h, w, H, W, sh, sw,
                                                    ignore_index, cap);
  } else {
    auto k = segl::upsample_ce_fwd_kernel<__nv_bfloat16>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, block, smem, (cudaStream_t)stream>>>((const __nv_bfloat16 *)logits, label, lse, stats, C, h, w, H, W, sh,
                                                    sw, ignore_index, cap);
  }
  RSC_CHECK_LAUNCH("rsc_upsample_ce_fwd");
  return RSC_OK;
}

extern "C" int rsc_upsample_ce_bwd(const void *logits, const int64_t *label, const float *lse, const float *gscale,
                                   void *dlogits, int B, int C, int h, int w, int H, int W, int dtype, void *stream) {
  if (int e = uce_check("rsc_upsample_ce_bwd", B, C, h, w, H, W, dtype)) return e;
  RSC_CHECK_ARG(logits && label && lse && gscale && dlogits, "rsc_upsample_ce_bwd: null pointer");
  const float sh = (float)h / H, sw = (float)w / W;
  const int64_t npix = (int64_t)B * h * w;
  const size_t smem = (size_t)segl::BW * (2 * segl::FP * segl::FP + 8 * segl::FP) * sizeof(float);
  int64_t blocks = (npix + segl::BW - 1) / segl::BW;
  const int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  if (dtype == RSC_F32) {
    auto k = segl::upsample_ce_bwd_kernel<float>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, segl::BW * 32, smem, (cudaStream_t)stream>>>((const float *)logits, label, lse, gscale, (float *)dlogits, B,
                                                            C, h, w, H, W, sh, sw, npix);
  } else {
    auto k = segl::upsample_ce_bwd_kernel<__nv_bfloat16>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, segl::BW * 32, smem, (cudaStream_t)stream>>>((const __nv_bfloat16 *)logits, label, lse, gscale,
                                                            (__nv_bfloat16 *)dlogits, B, C, h, w, H, W, sh, sw, npix);
  }
  RSC_CHECK_LAUNCH("rsc_upsample_ce_bwd");
  return RSC_OK;
}
