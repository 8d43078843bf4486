// Flat fused optimiser step and bias-gradient column sums (mmcv OptimizerHook =
// clip_grad_norm_ + torch AdamW over 63 M parameters; the reference runs one kernel chain per tensor).
//
// rsc_adamw_step: ONE pass over a contiguous fp32 range of the flat parameter buffer:
//   g = grad * clip_coef ; p *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)          (torch.optim.AdamW)
// and (optionally) refreshes the bf16 shadow of the parameters that the next step's GEMMs read, so no
// per-weight fp32->bf16 cast kernel ever runs (autocast launches one per weight per step).
// lr, t and clip_coef are DEVICE scalars so that the launch is CUDA-graph replayable.
#include "common.cuh"

namespace rsc {

__global__ void __launch_bounds__(256)
    adamw_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                 int64_t n4, const float *__restrict__ lr_ptr, float lr_mult, float beta1, float beta2, float eps,
                 float wd, const float *__restrict__ step_ptr, const float *__restrict__ clip_ptr,
                 __nv_bfloat16 *__restrict__ p_lp) {
  const float lr = __ldg(lr_ptr) * lr_mult;
  const float t = __ldg(step_ptr);
  const float clip = clip_ptr ? __ldg(clip_ptr) : 1.0f;
  const float bc1 = 1.0f - powf(beta1, t), bc2 = 1.0f - powf(beta2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.0f - lr * wd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4 *>(p)[i];
    const float4 gv4 = reinterpret_cast<const float4 *>(g)[i];
    float4 mv = reinterpret_cast<float4 *>(m)[i];
    float4 vv = reinterpret_cast<float4 *>(v)[i];
    float *pp = &pv.x, *mm = &mv.x, *vq = &vv.x;
    const float *gg = &gv4.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gg[k] * clip;
      mm[k] = beta1 * mm[k] + (1.0f - beta1) * gk;
      vq[k] = beta2 * vq[k] + (1.0f - beta2) * gk * gk;
      const float denom = sqrtf(vq[k]) * inv_sqrt_bc2 + eps;
      pp[k] = pp[k] * decay - step_size * (mm[k] / denom);
    }
    reinterpret_cast<float4 *>(p)[i] = pv;
    if (p_lp) {   // compute-dtype shadow of the master weights (what the GEMMs of the next step read)
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      uint2 pk = make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
      reinterpret_cast<uint2 *>(p_lp)[i] = pk;
    }
    reinterpret_cast<float4 *>(m)[i] = mv;
    reinterpret_cast<float4 *>(v)[i] = vv;
  }
}

// y[c] (+)= sum_r x[r][c]: bias gradient of a Linear layer.  Block = 32 column-quads x 8 row slices over a
// slab of rows; partial sums through shared memory, one atomic per column per block.
template <typename T>
__global__ void __launch_bounds__(256)
    colsum_kernel(const T *__restrict__ x, float *__restrict__ y, int64_t rows, int C, int rows_per_block) {
  __shared__ float4 red[8][32];
  const int c = blockIdx.x * 128 + threadIdx.x * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    const T *src = x + c;
#pragma unroll 4
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const float4 v = load4<T>(src + r * C);
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
    atomicAdd(y + c, s.x), atomicAdd(y + c + 1, s.y), atomicAdd(y + c + 2, s.z), atomicAdd(y + c + 3, s.w);
  }
}

}  // namespace rsc

using namespace rsc;

extern "C" int rsc_adamw_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                              const float *lr, float lr_mult, float beta1, float beta2, float eps,
                              float weight_decay, const float *step, const float *clip_coef, void *param_bf16,
                              void *stream) {
  RSC_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && lr && step, "rsc_adamw_step: null pointer");
  RSC_CHECK_ARG(n > 0 && n % 4 == 0, "rsc_adamw_step: n must be a positive multiple of 4 (got %lld)", (long long)n);
  RSC_CHECK_ARG(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
                "rsc_adamw_step: buffers must be 16-byte aligned");
  RSC_CHECK_ARG((uintptr_t)param_bf16 % 8 == 0, "rsc_adamw_step: param_bf16 must be 8-byte aligned");
  int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  int grid = (int)(blocks < kNumSMs * 8 ? blocks : kNumSMs * 8);
  adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n4, lr, lr_mult, beta1, beta2,
                                                        eps, weight_decay, step, clip_coef,
                                                        (__nv_bfloat16 *)param_bf16);
  RSC_CHECK_LAUNCH("rsc_adamw_step");
  return RSC_OK;
}

extern "C" int rsc_colsum(const void *x, float *y, int64_t rows, int C, int dtype, void *stream) {
  RSC_CHECK_ARG(x && y, "rsc_colsum: null pointer");
  RSC_CHECK_ARG(rows > 0 && C > 0 && C % 4 == 0, "rsc_colsum: need rows > 0, C %% 4 == 0 (rows=%lld, C=%d)",
                (long long)rows, C);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "rsc_colsum: bad dtype %d", dtype);
  const int cblocks = (C + 127) / 128;
  int64_t want = (int64_t)kNumSMs * 4 / cblocks;          // ~4 blocks per SM in total
  if (want < 1) want = 1;
  int64_t rpb = (rows + want - 1) / want;
  if (rpb < 64) rpb = 64;
  const int rblocks = (int)((rows + rpb - 1) / rpb);
  dim3 grid(cblocks, rblocks), block(32, 8);
  if (dtype == RSC_F32)
    colsum_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float *)x, y, rows, C, (int)rpb);
  else
    colsum_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x, y, rows, C, (int)rpb);
  RSC_CHECK_LAUNCH("rsc_colsum");
  return RSC_OK;
}
