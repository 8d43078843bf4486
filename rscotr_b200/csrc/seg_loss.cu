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

// Backward.  Every output pixel interpolates between the SAME four low-resolution logits as its neighbours inside one
// "cell" (cy, cx) = {output pixels whose upper-left tap is (cy, cx)}; the cells partition the output.  One warp per
// cell, lanes over channels: softmax - onehot is evaluated ONCE per (output pixel, channel) (a gather per low-resolution
// pixel visits every output pixel from its four neighbours: 4x the exponentials, and this pass is MUFU-bound), and the
// four tap-weighted sums of the cell go to a workspace [B][h][w][4][C]; a second pass adds the <= 4 cells around each
// low-resolution pixel (deterministic, no atomics).
constexpr int BW = 8;     // warps per block
constexpr int FP = 24;    // max output pixels per cell and axis (first / last cells hold the clamped border: 1.5 * scale + 1)

__device__ __forceinline__ int tap0(int o, float scale, int in) {
  int i0, i1;
  float lam;
  bil_src(o, scale, in, i0, i1, lam);
  return i0;
}
// output index range [lo, hi] of the cell whose upper tap is k (tap0 is monotone in o)
__device__ __forceinline__ void cell_range(int k, float scale, int in, int out, int &lo, int &hi) {
  auto first_ge = [&](int kk) {      // first o with tap0(o) >= kk
    int o = (int)ceilf((kk + 0.5f) / scale - 0.5f);
    o = min(max(o, 0), out);
    while (o > 0 && tap0(o - 1, scale, in) >= kk) --o;
    while (o < out && tap0(o, scale, in) < kk) ++o;
    return o;
  };
  lo = k == 0 ? 0 : first_ge(k);
  hi = k == in - 1 ? out - 1 : first_ge(k + 1) - 1;
}

template <typename T>
__global__ void __launch_bounds__(BW * 32)
    upsample_ce_bwd_cells_kernel(const T *__restrict__ logits, const int64_t *__restrict__ label, const float *__restrict__ lse,
                                 float *__restrict__ ws, int B, int C, int h, int w, int H, int W, float sh, float sw,
                                 int64_t ncell) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *s_lse = sm + (size_t)warp * (2 * FP * FP + 2 * FP);
  int *s_lab = reinterpret_cast<int *>(s_lse + FP * FP);
  float *s_ly = s_lse + 2 * FP * FP, *s_lx = s_ly + FP;
  for (int64_t cell = (int64_t)blockIdx.x * BW + warp; cell < ncell; cell += (int64_t)gridDim.x * BW) {
    const int cx = (int)(cell % w), cy = (int)((cell / w) % h), b = (int)(cell / ((int64_t)w * h));
    int oy0, oy1, ox0, ox1;
    cell_range(cy, sh, h, H, oy0, oy1);
    cell_range(cx, sw, w, W, ox0, ox1);
    const int ny = oy1 - oy0 + 1, nx = ox1 - ox0 + 1;      // <= FP (checked on the host); 0 when the cell is empty
    __syncwarp();
    for (int k = lane; k < ny + nx; k += 32) {
      const bool isy = k < ny;
      int i0, i1;
      float lam;
      bil_src(isy ? oy0 + k : ox0 + (k - ny), isy ? sh : sw, isy ? h : w, i0, i1, lam);
      (isy ? s_ly : s_lx - ny)[k] = lam;
    }
    for (int k = lane; k < ny * nx; k += 32) {
      const int yy = k / nx, xx = k - yy * nx;
      const int64_t pix = ((int64_t)b * H + oy0 + yy) * W + ox0 + xx;
      const float l = lse[pix];
      s_lse[yy * FP + xx] = l;
      s_lab[yy * FP + xx] = l == INFINITY ? -1 : (int)label[pix];
    }
    __syncwarp();
    const int cy1 = min(cy + 1, h - 1), cx1 = min(cx + 1, w - 1);
    float *out = ws + cell * 4 * C;
    for (int c = lane; c < C; c += 32) {
      const T *p = logits + ((int64_t)b * C + c) * h * w;
      const float L00 = to_f<T>(p[cy * w + cx]), L01 = to_f<T>(p[cy * w + cx1]);
      const float L10 = to_f<T>(p[cy1 * w + cx]), L11 = to_f<T>(p[cy1 * w + cx1]);
      float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
      for (int yy = 0; yy < ny; ++yy) {
        const float ly = s_ly[yy];
        const float r0 = L00 + ly * (L10 - L00), r1 = L01 + ly * (L11 - L01);
        float all = 0.f, right = 0.f;
        for (int xx = 0; xx < nx; ++xx) {
          const float lx = s_lx[xx];
          const float v = fmaf(lx, r1 - r0, r0);
          const float g = __expf(v - s_lse[yy * FP + xx]) - (s_lab[yy * FP + xx] == c ? 1.f : 0.f);   // 0 for ignored pixels
          all += g;
          right = fmaf(lx, g, right);
        }
        const float left = all - right;
        a00 = fmaf(1.f - ly, left, a00);
        a01 = fmaf(1.f - ly, right, a01);
        a10 = fmaf(ly, left, a10);
        a11 = fmaf(ly, right, a11);
      }
      if (cy1 == cy) a00 += a10, a01 += a11, a10 = a11 = 0.f;      // clamped last row / column: both taps are the same pixel
      if (cx1 == cx) a00 += a01, a10 += a11, a01 = a11 = 0.f;
      out[c] = a00, out[C + c] = a01, out[2 * C + c] = a10, out[3 * C + c] = a11;
    }
  }
}

// dlogits[b][c][y][x] = g * (cell(y,x).a00 + cell(y,x-1).a01 + cell(y-1,x).a10 + cell(y-1,x-1).a11); thread order (b,y,x,c)
template <typename T>
__global__ void upsample_ce_bwd_gather_kernel(const float *__restrict__ ws, const float *__restrict__ gscale, T *__restrict__ dlogits,
                                              int C, int h, int w, int64_t total) {
  const float g = gscale[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t pix = i / C;
    const int x = (int)(pix % w), y = (int)((pix / w) % h);
    const int64_t b = pix / ((int64_t)w * h);
    float acc = ws[pix * 4 * C + c];
    if (x > 0) acc += ws[(pix - 1) * 4 * C + C + c];
    if (y > 0) acc += ws[(pix - w) * 4 * C + 2 * C + c];
    if (x > 0 && y > 0) acc += ws[(pix - w - 1) * 4 * C + 3 * C + c];
    dlogits[((b * C + c) * h + y) * w + x] = from_f<T>(acc * g);
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
                                   void *dlogits, float *ws, int B, int C, int h, int w, int H, int W, int dtype, void *stream) {
  if (int e = uce_check("rsc_upsample_ce_bwd", B, C, h, w, H, W, dtype)) return e;
  RSC_CHECK_ARG(logits && label && lse && gscale && dlogits && ws, "rsc_upsample_ce_bwd: null pointer");
  const float sh = (float)h / H, sw = (float)w / W;
  const int64_t ncell = (int64_t)B * h * w, total = ncell * C;
  const size_t smem = (size_t)segl::BW * (2 * segl::FP * segl::FP + 2 * segl::FP) * sizeof(float);
  int64_t blocks = (ncell + segl::BW - 1) / segl::BW;
  const int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
  int64_t gb = (total + 255) / 256;
  const int ggrid = (int)(gb < (int64_t)kNumSMs * 16 ? gb : (int64_t)kNumSMs * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RSC_F32) {
    auto k = segl::upsample_ce_bwd_cells_kernel<float>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, segl::BW * 32, smem, st>>>((const float *)logits, label, lse, ws, B, C, h, w, H, W, sh, sw, ncell);
    segl::upsample_ce_bwd_gather_kernel<float><<<ggrid, 256, 0, st>>>(ws, gscale, (float *)dlogits, C, h, w, total);
  } else {
    auto k = segl::upsample_ce_bwd_cells_kernel<__nv_bfloat16>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, segl::BW * 32, smem, st>>>((const __nv_bfloat16 *)logits, label, lse, ws, B, C, h, w, H, W, sh, sw, ncell);
    segl::upsample_ce_bwd_gather_kernel<__nv_bfloat16><<<ggrid, 256, 0, st>>>(ws, gscale, (__nv_bfloat16 *)dlogits, C, h, w, total);
  }
  RSC_CHECK_LAUNCH("rsc_upsample_ce_bwd");
  count_launch(1);
  return RSC_OK;
}
