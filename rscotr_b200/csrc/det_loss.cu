// Detection-loss hot path of the DINO head on the GPU:
//
//   rsc_det_match      Hungarian matching of every (decoder layer, image) pair in ONE launch: the
//                      FocalLossCost + BBoxL1Cost(xywh) + IoUCost(giou) cost matrix (mmdet 2.25.1
//                      core/bbox/match_costs, reference cfg :170-174) and the rectangular linear sum
//                      assignment (the shortest-augmenting-path algorithm scipy.optimize.linear_sum_assignment
//                      implements; reference call site mmdet_detr_head/detr_head.py:513 via HungarianAssigner)
//                      -- one CTA per problem, duals in fp64 shared memory, column scans in parallel.
//                      The reference does 7 x B scipy calls, each behind a device->host sync.
//   rsc_det_loss_fwd   sigmoid focal loss + L1 + GIoU of all layers of one query segment, already divided
//   rsc_det_loss_bwd   by the averaging factors and multiplied by the loss weights -> (rows, 3) sums; the
//                      backward writes d(cls logits) and d(boxes) element-wise (detr_head.py:333-416,
//                      dino_head.py:236-309 loss_single / loss_dn_single, ~60 tiny ATen kernels per call).
#include <float.h>

#include "common.cuh"

namespace rsc {
namespace det {

constexpr int THREADS = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float pow_gamma(float x, float gamma) { return gamma == 2.f ? x * x : powf(x, gamma); }

template <typename T>
__device__ __forceinline__ float ldf(const void *p, int64_t i) {
  return to_f<T>(reinterpret_cast<const T *>(p)[i]);
}

// mmdet bbox_overlaps(mode='giou', eps) on two xyxy boxes
__device__ __forceinline__ float giou_xyxy(const float a[4], const float b[4], float eps) {
  const float a1 = (a[2] - a[0]) * (a[3] - a[1]), a2 = (b[2] - b[0]) * (b[3] - b[1]);
  const float iw = fmaxf(fminf(a[2], b[2]) - fmaxf(a[0], b[0]), 0.f), ih = fmaxf(fminf(a[3], b[3]) - fmaxf(a[1], b[1]), 0.f);
  const float ov = iw * ih;
  const float un = fmaxf(a1 + a2 - ov, eps);
  const float ew = fmaxf(fmaxf(a[2], b[2]) - fminf(a[0], b[0]), 0.f), eh = fmaxf(fmaxf(a[3], b[3]) - fminf(a[1], b[1]), 0.f);
  const float ea = fmaxf(ew * eh, eps);
  return ov / un - (ea - un) / ea;
}

// --------------------------------------------------------------------------------------------
// cost matrix + linear sum assignment, one CTA per (layer, image) problem
// --------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(THREADS)
    det_match_kernel(const void *__restrict__ cls, const float *__restrict__ box, const int64_t *__restrict__ gt_labels,
                     const float *__restrict__ gt_boxes, const int *__restrict__ gt_start,
                     const float *__restrict__ img_wh, int B, int NqTot, int q0, int Nq, int C, int max_gt, float w_cls,
                     float w_reg, float w_iou, float alpha, float gamma, float eps, float *__restrict__ cost,
                     int *__restrict__ assign, float *__restrict__ gt_norm) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double *v = reinterpret_cast<double *>(smraw);   // [Nq]  column duals
  double *sp = v + Nq;                             // [Nq]  shortest path costs
  double *u = sp + Nq;                             // [max_gt] row duals
  int *path = reinterpret_cast<int *>(u + max_gt); // [Nq]
  int *row4col = path + Nq;                        // [Nq]
  int *col4row = row4col + Nq;                     // [max_gt]
  unsigned char *SC = reinterpret_cast<unsigned char *>(col4row + max_gt);   // [Nq]
  unsigned char *SR = SC + Nq;                                               // [max_gt]
  __shared__ double red_v[THREADS / 32];
  __shared__ int red_k[THREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = blockIdx.x, b = p % B;
  const int g0 = gt_start[b], n = gt_start[b + 1] - g0;
  const float Wi = img_wh[2 * b], Hi = img_wh[2 * b + 1];
  float *cst = cost + (size_t)p * max_gt * Nq;
  int *asg = assign + (size_t)p * Nq;

  if (gt_norm && p == 0) {   // normalised cxcywh targets of every gt box (what the loss kernels regress to)
    const int G = gt_start[B];
    for (int g = tid; g < G; g += THREADS) {
      int bb = 0;
      while (bb + 1 < B && g >= gt_start[bb + 1]) ++bb;
      const float w = img_wh[2 * bb], h = img_wh[2 * bb + 1];
      const float x1 = gt_boxes[4 * g] / w, y1 = gt_boxes[4 * g + 1] / h, x2 = gt_boxes[4 * g + 2] / w,
                  y2 = gt_boxes[4 * g + 3] / h;
      gt_norm[4 * g] = (x1 + x2) / 2, gt_norm[4 * g + 1] = (y1 + y2) / 2, gt_norm[4 * g + 2] = x2 - x1,
                  gt_norm[4 * g + 3] = y2 - y1;
    }
  }
  for (int j = tid; j < Nq; j += THREADS) asg[j] = -1;
  if (n <= 0) return;

  // ---- cost[i][q] ----
  for (int idx = tid; idx < n * Nq; idx += THREADS) {
    const int i = idx / Nq, q = idx - i * Nq;
    const int64_t row = (int64_t)p * NqTot + q0 + q;
    const int lab = (int)gt_labels[g0 + i];
    const float pr = sigmoidf_(ldf<T>(cls, row * C + lab));
    const float neg = -logf(1.f - pr + eps) * (1.f - alpha) * pow_gamma(pr, gamma);
    const float pos = -logf(pr + eps) * alpha * pow_gamma(1.f - pr, gamma);
    const float c_cls = (pos - neg) * w_cls;
    const float *bp = box + row * 4;
    const float *gb = gt_boxes + (size_t)(g0 + i) * 4;
    const float gx1 = gb[0] / Wi, gy1 = gb[1] / Hi, gx2 = gb[2] / Wi, gy2 = gb[3] / Hi;
    const float c_reg = (fabsf(bp[0] - (gx1 + gx2) / 2) + fabsf(bp[1] - (gy1 + gy2) / 2) + fabsf(bp[2] - (gx2 - gx1)) +
                         fabsf(bp[3] - (gy2 - gy1))) * w_reg;
    const float pb[4] = {(bp[0] - 0.5f * bp[2]) * Wi, (bp[1] - 0.5f * bp[3]) * Hi, (bp[0] + 0.5f * bp[2]) * Wi,
                         (bp[1] + 0.5f * bp[3]) * Hi};
    const float gbx[4] = {gb[0], gb[1], gb[2], gb[3]};
    const float c_iou = -giou_xyxy(pb, gbx, 1e-6f) * w_iou;
    cst[(size_t)i * Nq + q] = c_cls + c_reg + c_iou;
  }
  for (int j = tid; j < Nq; j += THREADS) v[j] = 0.0, row4col[j] = -1;
  for (int i = tid; i < n; i += THREADS) u[i] = 0.0, col4row[i] = -1;
  __syncthreads();

  // ---- shortest augmenting path per row (gt), columns (queries) scanned in parallel ----
  for (int cur = 0; cur < n; ++cur) {
    for (int j = tid; j < Nq; j += THREADS) sp[j] = INFINITY, SC[j] = 0;
    for (int i = tid; i < n; i += THREADS) SR[i] = 0;
    __syncthreads();
    double minVal = 0.0;
    int i = cur, sink = -1;
    while (sink < 0) {
      if (tid == 0) SR[i] = 1;
      const double ui = u[i];
      const float *crow = cst + (size_t)i * Nq;
      double bestv = INFINITY;
      int bestk = 0x7fffffff;
      for (int j = tid; j < Nq; j += THREADS) {
        if (SC[j]) continue;
        const double r = minVal + (double)crow[j] - ui - v[j];
        double s = sp[j];
        if (r < s) s = r, sp[j] = r, path[j] = i;
        const int key = (row4col[j] >= 0 ? (1 << 24) : 0) | j;   // among equal minima prefer a free column
        if (s < bestv || (s == bestv && key < bestk)) bestv = s, bestk = key;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bestv, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bestk, o);
        if (ov < bestv || (ov == bestv && ok < bestk)) bestv = ov, bestk = ok;
      }
      if (lane == 0) red_v[warp] = bestv, red_k[warp] = bestk;
      __syncthreads();
      bestv = red_v[0], bestk = red_k[0];
#pragma unroll
      for (int w = 1; w < THREADS / 32; ++w)
        if (red_v[w] < bestv || (red_v[w] == bestv && red_k[w] < bestk)) bestv = red_v[w], bestk = red_k[w];
      if (bestk == 0x7fffffff || !(bestv < INFINITY)) {   // infeasible (NaN / inf costs): leave the rest unmatched
        sink = -2;
        break;
      }
      minVal = bestv;
      const int j = bestk & 0xffffff;
      const int r4c = row4col[j];
      if (r4c < 0) sink = j; else i = r4c;
      __syncthreads();            // every thread has read red_* / row4col before the next writes
      if (tid == 0) SC[j] = 1;
      __syncthreads();
    }
    if (sink < 0) break;
    if (tid == 0) u[cur] += minVal;
    for (int r = tid; r < n; r += THREADS)
      if (SR[r] && r != cur) u[r] += minVal - sp[col4row[r]];
    for (int j = tid; j < Nq; j += THREADS)
      if (SC[j]) v[j] -= minVal - sp[j];
    __syncthreads();
    if (tid == 0) {
      int j = sink;
      while (true) {
        const int r = path[j];
        row4col[j] = r;
        const int t = col4row[r];
        col4row[r] = j;
        j = t;
        if (r == cur) break;
      }
    }
    __syncthreads();
  }
  for (int r = tid; r < n; r += THREADS)
    if (col4row[r] >= 0) asg[col4row[r]] = g0 + r;
}

// --------------------------------------------------------------------------------------------
// fused focal + L1 + GIoU of one query segment of all layers
//   element (l, b, q): row = (l*B + b)*NqTot + q0 + q ;  target gt = assign[l*a_ls + b*Nq + q] (-1: background)
//   out row of layer l: out_row0 + (last_first ? (l == L-1 ? 0 : l + 1) : l)
// --------------------------------------------------------------------------------------------
struct LossArgs {
  int L, B, NqTot, q0, Nq, C, a_ls, out_row0, last_first;
  float gamma, alpha, w_cls, w_l1, w_iou, eps;
};

__device__ __forceinline__ int out_row(const LossArgs &a, int l) {
  return a.out_row0 + (a.last_first ? (l == a.L - 1 ? 0 : l + 1) : l);
}

template <typename T>
__global__ void __launch_bounds__(THREADS)
    det_loss_fwd_kernel(const void *__restrict__ cls, const float *__restrict__ box, const int *__restrict__ assign,
                        const int64_t *__restrict__ gt_labels, const float *__restrict__ gt_norm,
                        const float *__restrict__ img_wh, const float *__restrict__ cls_factor,
                        const float *__restrict__ pos_factor, float *__restrict__ out, LossArgs a) {
  __shared__ float red[3][THREADS / 32];
  const int lb = blockIdx.y, l = lb / a.B, b = lb - l * a.B;
  const int q = blockIdx.x * THREADS + threadIdx.x;
  float s_cls = 0.f, s_l1 = 0.f, s_iou = 0.f;
  if (q < a.Nq) {
    const int64_t row = (int64_t)lb * a.NqTot + a.q0 + q;
    const int g = assign[(int64_t)l * a.a_ls + (int64_t)b * a.Nq + q];
    const int t = g >= 0 ? (int)gt_labels[g] : a.C;
    for (int c = 0; c < a.C; ++c) {
      const float p = sigmoidf_(ldf<T>(cls, row * a.C + c));
      s_cls += (t == c) ? -a.alpha * pow_gamma(1.f - p, a.gamma) * logf(fmaxf(p, FLT_MIN))
                        : -(1.f - a.alpha) * pow_gamma(p, a.gamma) * logf(fmaxf(1.f - p, FLT_MIN));
    }
    if (g >= 0) {
      const float *bp = box + row * 4, *gt = gt_norm + (size_t)g * 4;
      const float Wi = img_wh[2 * b], Hi = img_wh[2 * b + 1];
      s_l1 = fabsf(bp[0] - gt[0]) + fabsf(bp[1] - gt[1]) + fabsf(bp[2] - gt[2]) + fabsf(bp[3] - gt[3]);
      const float pb[4] = {(bp[0] - 0.5f * bp[2]) * Wi, (bp[1] - 0.5f * bp[3]) * Hi, (bp[0] + 0.5f * bp[2]) * Wi,
                           (bp[1] + 0.5f * bp[3]) * Hi};
      const float tb[4] = {(gt[0] - 0.5f * gt[2]) * Wi, (gt[1] - 0.5f * gt[3]) * Hi, (gt[0] + 0.5f * gt[2]) * Wi,
                           (gt[1] + 0.5f * gt[3]) * Hi};
      s_iou = 1.f - giou_xyxy(pb, tb, a.eps);
    }
  }
  s_cls = warp_sum(s_cls), s_l1 = warp_sum(s_l1), s_iou = warp_sum(s_iou);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[0][warp] = s_cls, red[1][warp] = s_l1, red[2][warp] = s_iou;
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) s += red[threadIdx.x][w];
    const float scale = threadIdx.x == 0 ? a.w_cls / cls_factor[l] : (threadIdx.x == 1 ? a.w_l1 : a.w_iou) / pos_factor[l];
    if (s != 0.f) atomicAdd(out + out_row(a, l) * 3 + threadIdx.x, s * scale);
  }
}

// torch semantics of the piecewise ops at ties: maximum/minimum split the gradient, clamp(min) passes it at equality
__device__ __forceinline__ float gsel_gt(float x, float y) { return x > y ? 1.f : (x == y ? 0.5f : 0.f); }

template <typename T>
__global__ void __launch_bounds__(THREADS)
    det_loss_bwd_kernel(const void *__restrict__ cls, const float *__restrict__ box, const int *__restrict__ assign,
                        const int64_t *__restrict__ gt_labels, const float *__restrict__ gt_norm,
                        const float *__restrict__ img_wh, const float *__restrict__ cls_factor,
                        const float *__restrict__ pos_factor, const float *__restrict__ dout, void *__restrict__ dcls,
                        float *__restrict__ dbox, LossArgs a) {
  const int lb = blockIdx.y, l = lb / a.B, b = lb - l * a.B;
  const int q = blockIdx.x * THREADS + threadIdx.x;
  if (q >= a.Nq) return;
  const int64_t row = (int64_t)lb * a.NqTot + a.q0 + q;
  const int g = assign[(int64_t)l * a.a_ls + (int64_t)b * a.Nq + q];
  const int t = g >= 0 ? (int)gt_labels[g] : a.C;
  const float *go = dout + out_row(a, l) * 3;
  const float k_cls = go[0] * a.w_cls / cls_factor[l];
  for (int c = 0; c < a.C; ++c) {
    const float p = sigmoidf_(ldf<T>(cls, row * a.C + c));
    const float lp = logf(fmaxf(p, FLT_MIN)), ln = logf(fmaxf(1.f - p, FLT_MIN));
    const float d = (t == c) ? -a.alpha * pow_gamma(1.f - p, a.gamma) * (1.f - p - a.gamma * p * lp)
                             : -(1.f - a.alpha) * pow_gamma(p, a.gamma) * (a.gamma * (1.f - p) * ln - p);
    reinterpret_cast<T *>(dcls)[row * a.C + c] = from_f<T>(d * k_cls);
  }
  float d4[4] = {0.f, 0.f, 0.f, 0.f};
  if (g >= 0) {
    const float *bp = box + row * 4, *gt = gt_norm + (size_t)g * 4;
    const float Wi = img_wh[2 * b], Hi = img_wh[2 * b + 1];
    const float k_l1 = go[1] * a.w_l1 / pos_factor[l], k_iou = go[2] * a.w_iou / pos_factor[l];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float df = bp[k] - gt[k];
      d4[k] = k_l1 * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
    }
    const float px1 = (bp[0] - 0.5f * bp[2]) * Wi, py1 = (bp[1] - 0.5f * bp[3]) * Hi, px2 = (bp[0] + 0.5f * bp[2]) * Wi,
                py2 = (bp[1] + 0.5f * bp[3]) * Hi;
    const float tx1 = (gt[0] - 0.5f * gt[2]) * Wi, ty1 = (gt[1] - 0.5f * gt[3]) * Hi, tx2 = (gt[0] + 0.5f * gt[2]) * Wi,
                ty2 = (gt[1] + 0.5f * gt[3]) * Hi;
    const float pw = px2 - px1, ph = py2 - py1;
    const float a1 = pw * ph, a2 = (tx2 - tx1) * (ty2 - ty1);
    const float iw0 = fminf(px2, tx2) - fmaxf(px1, tx1), ih0 = fminf(py2, ty2) - fmaxf(py1, ty1);
    const float iw = fmaxf(iw0, 0.f), ih = fmaxf(ih0, 0.f), ov = iw * ih;
    const float un0 = a1 + a2 - ov, un = fmaxf(un0, a.eps);
    const float ew0 = fmaxf(px2, tx2) - fminf(px1, tx1), eh0 = fmaxf(py2, ty2) - fminf(py1, ty1);
    const float ew = fmaxf(ew0, 0.f), eh = fmaxf(eh0, 0.f), ea0 = ew * eh, ea = fmaxf(ea0, a.eps);
    // giou = ov/un + un/ea - 1 ;  loss = 1 - giou
    float d_ov = 1.f / un, d_un = -ov / (un * un) + 1.f / ea, d_ea = -un / (ea * ea);
    const float d_un0 = un0 >= a.eps ? d_un : 0.f;
    const float d_a1 = d_un0;
    d_ov -= d_un0;
    const float d_ea0 = ea0 >= a.eps ? d_ea : 0.f;
    const float d_ew = ew0 >= 0.f ? d_ea0 * eh : 0.f, d_eh = eh0 >= 0.f ? d_ea0 * ew : 0.f;
    const float d_iw = iw0 >= 0.f ? d_ov * ih : 0.f, d_ih = ih0 >= 0.f ? d_ov * iw : 0.f;
    // x: ex2 = max(px2,tx2), ex1 = min(px1,tx1), rbx = min(px2,tx2), ltx = max(px1,tx1)
    float gx1 = -d_ew * gsel_gt(tx1, px1) - d_iw * gsel_gt(px1, tx1) - d_a1 * ph;
    float gx2 = d_ew * gsel_gt(px2, tx2) + d_iw * gsel_gt(tx2, px2) + d_a1 * ph;
    float gy1 = -d_eh * gsel_gt(ty1, py1) - d_ih * gsel_gt(py1, ty1) - d_a1 * pw;
    float gy2 = d_eh * gsel_gt(py2, ty2) + d_ih * gsel_gt(ty2, py2) + d_a1 * pw;
    // loss = 1 - giou
    gx1 = -gx1 * k_iou, gx2 = -gx2 * k_iou, gy1 = -gy1 * k_iou, gy2 = -gy2 * k_iou;
    d4[0] += (gx1 + gx2) * Wi;
    d4[1] += (gy1 + gy2) * Hi;
    d4[2] += (gx2 - gx1) * 0.5f * Wi;
    d4[3] += (gy2 - gy1) * 0.5f * Hi;
  }
  *reinterpret_cast<float4 *>(dbox + row * 4) = make_float4(d4[0], d4[1], d4[2], d4[3]);
}

}  // namespace det
}  // namespace rsc

using namespace rsc;

extern "C" int rsc_det_match(const void *cls, const float *box, const int64_t *gt_labels, const float *gt_boxes,
                             const int *gt_start, const float *img_wh, int P, int B, int NqTot, int q0, int Nq, int C,
                             int max_gt, float w_cls, float w_reg, float w_iou, float alpha, float gamma, float eps,
                             float *cost, int *assign, float *gt_norm, int dtype, void *stream) {
  RSC_CHECK_ARG(P >= 0 && B > 0 && Nq > 0 && C > 0 && q0 >= 0 && q0 + Nq <= NqTot && max_gt >= 0,
                "rsc_det_match: bad shape (P=%d,B=%d,NqTot=%d,q0=%d,Nq=%d,C=%d,max_gt=%d)", P, B, NqTot, q0, Nq, C, max_gt);
  RSC_CHECK_ARG(P % B == 0, "rsc_det_match: P (%d) must be a multiple of B (%d)", P, B);
  RSC_CHECK_ARG(max_gt <= Nq, "rsc_det_match: more ground-truth boxes (%d) than queries (%d)", max_gt, Nq);
  RSC_CHECK_ARG(Nq < (1 << 24), "rsc_det_match: Nq too large");
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "rsc_det_match: bad dtype %d", dtype);
  if (P == 0) return RSC_OK;
  RSC_CHECK_ARG(cls && box && gt_start && img_wh && assign && (max_gt == 0 || (gt_labels && gt_boxes && cost)),
                "rsc_det_match: null pointer");
  const int mg = max_gt > 0 ? max_gt : 1;
  size_t smem = (size_t)Nq * (8 + 8 + 4 + 4 + 1) + (size_t)mg * (8 + 4 + 1) + 16;
  RSC_CHECK_ARG(smem <= 200 * 1024, "rsc_det_match: problem too large for shared memory (Nq=%d, max_gt=%d)", Nq, max_gt);
  // keep the int arrays 4-byte aligned: doubles first (8-aligned), then ints; u has max_gt doubles
  if (dtype == RSC_F32) {
    auto k = det::det_match_kernel<float>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<P, det::THREADS, smem, (cudaStream_t)stream>>>(cls, box, gt_labels, gt_boxes, gt_start, img_wh, B, NqTot, q0, Nq,
                                                        C, mg, w_cls, w_reg, w_iou, alpha, gamma, eps, cost, assign,
                                                        gt_norm);
  } else {
    auto k = det::det_match_kernel<__nv_bfloat16>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<P, det::THREADS, smem, (cudaStream_t)stream>>>(cls, box, gt_labels, gt_boxes, gt_start, img_wh, B, NqTot, q0, Nq,
                                                        C, mg, w_cls, w_reg, w_iou, alpha, gamma, eps, cost, assign,
                                                        gt_norm);
  }
  RSC_CHECK_LAUNCH("rsc_det_match");
  return RSC_OK;
}

static int det_loss_args(const char *fn, det::LossArgs &a, int L, int B, int NqTot, int q0, int Nq, int C,
                         int assign_layer_stride, int out_row0, int last_first, float gamma, float alpha, float w_cls,
                         float w_l1, float w_iou, float eps, int dtype) {
  RSC_CHECK_ARG(L >= 0 && B > 0 && Nq >= 0 && C > 0 && q0 >= 0 && q0 + Nq <= NqTot,
                "%s: bad shape (L=%d,B=%d,NqTot=%d,q0=%d,Nq=%d,C=%d)", fn, L, B, NqTot, q0, Nq, C);
  RSC_CHECK_ARG(assign_layer_stride == 0 || assign_layer_stride == B * Nq, "%s: assign_layer_stride must be 0 or B*Nq", fn);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "%s: bad dtype %d", fn, dtype);
  RSC_CHECK_ARG(out_row0 >= 0, "%s: bad out_row0", fn);
  a = det::LossArgs{L, B, NqTot, q0, Nq, C, assign_layer_stride, out_row0, last_first, gamma, alpha, w_cls, w_l1, w_iou, eps};
  return RSC_OK;
}

extern "C" int rsc_det_loss_fwd(const void *cls, const float *box, const int *assign, const int64_t *gt_labels,
                                const float *gt_norm, const float *img_wh, const float *cls_factor,
                                const float *pos_factor, float *out, int L, int B, int NqTot, int q0, int Nq, int C,
                                int assign_layer_stride, int out_row0, int last_first, float gamma, float alpha,
                                float w_cls, float w_l1, float w_iou, float eps, int dtype, void *stream) {
  det::LossArgs a;
  if (int e = det_loss_args("rsc_det_loss_fwd", a, L, B, NqTot, q0, Nq, C, assign_layer_stride, out_row0, last_first,
                            gamma, alpha, w_cls, w_l1, w_iou, eps, dtype))
    return e;
  if (L == 0 || Nq == 0) return RSC_OK;
  RSC_CHECK_ARG(cls && box && assign && img_wh && cls_factor && pos_factor && out, "rsc_det_loss_fwd: null pointer");
  dim3 grid((Nq + det::THREADS - 1) / det::THREADS, L * B);
  if (dtype == RSC_F32)
    det::det_loss_fwd_kernel<float><<<grid, det::THREADS, 0, (cudaStream_t)stream>>>(
        cls, box, assign, gt_labels, gt_norm, img_wh, cls_factor, pos_factor, out, a);
  else
    det::det_loss_fwd_kernel<__nv_bfloat16><<<grid, det::THREADS, 0, (cudaStream_t)stream>>>(
        cls, box, assign, gt_labels, gt_norm, img_wh, cls_factor, pos_factor, out, a);
  RSC_CHECK_LAUNCH("rsc_det_loss_fwd");
  return RSC_OK;
}

extern "C" int rsc_det_loss_bwd(const void *cls, const float *box, const int *assign, const int64_t *gt_labels,
                                const float *gt_norm, const float *img_wh, const float *cls_factor,
                                const float *pos_factor, const float *dout, void *dcls, float *dbox, int L, int B,
                                int NqTot, int q0, int Nq, int C, int assign_layer_stride, int out_row0, int last_first,
                                float gamma, float alpha, float w_cls, float w_l1, float w_iou, float eps, int dtype,
                                void *stream) {
  det::LossArgs a;
  if (int e = det_loss_args("rsc_det_loss_bwd", a, L, B, NqTot, q0, Nq, C, assign_layer_stride, out_row0, last_first,
                            gamma, alpha, w_cls, w_l1, w_iou, eps, dtype))
    return e;
  if (L == 0 || Nq == 0) return RSC_OK;
  RSC_CHECK_ARG(cls && box && assign && img_wh && cls_factor && pos_factor && dout && dcls && dbox,
                "rsc_det_loss_bwd: null pointer");
  RSC_CHECK_ARG((uintptr_t)dbox % 16 == 0, "rsc_det_loss_bwd: dbox must be 16-byte aligned");
  dim3 grid((Nq + det::THREADS - 1) / det::THREADS, L * B);
  if (dtype == RSC_F32)
    det::det_loss_bwd_kernel<float><<<grid, det::THREADS, 0, (cudaStream_t)stream>>>(
        cls, box, assign, gt_labels, gt_norm, img_wh, cls_factor, pos_factor, dout, dcls, dbox, a);
  else
    det::det_loss_bwd_kernel<__nv_bfloat16><<<grid, det::THREADS, 0, (cudaStream_t)stream>>>(
        cls, box, assign, gt_labels, gt_norm, img_wh, cls_factor, pos_factor, dout, dcls, dbox, a);
  RSC_CHECK_LAUNCH("rsc_det_loss_bwd");
  return RSC_OK;
}
