// Shared helpers for the rscotr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rscotr.h"

namespace rsc {

constexpr int kNumSMs = 148;  // B200

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define RSC_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      rsc::set_error(__VA_ARGS__);        \
      return RSC_ERR_INVALID;             \
    }                                     \
  } while (0)

#define RSC_CHECK_LAUNCH(name)                                               \
  do {                                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      rsc::set_error("%s: CUDA error: %s", name, cudaGetErrorString(e__));   \
      return RSC_ERR_CUDA;                                                   \
    }                                                                        \
    rsc::count_launch();                                                     \
  } while (0)

// ---- element <-> float conversion -----------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 4-element vector load/store (16 B for float, 8 B for bf16); pointer must be
// aligned to the vector size.
template <typename T>
__device__ __forceinline__ float4 load4(const T *p);
template <>
__device__ __forceinline__ float4 load4<float>(const float *p) {
  return __ldg(reinterpret_cast<const float4 *>(p));
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16 *p) {
  uint2 r = __ldg(reinterpret_cast<const uint2 *>(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162 *>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162 *>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T>
__device__ __forceinline__ void store4(T *p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float *p, float4 v) {
  *reinterpret_cast<float4 *>(p) = v;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16 *p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t *>(&a);
  r.y = *reinterpret_cast<uint32_t *>(&b);
  *reinterpret_cast<uint2 *>(p) = r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- window geometry (shared by the index kernels and the fused attention) --
struct WinGeom {
  int B, H, W, ws, shift;
  int Hp, Wp, nWh, nWw;
  __host__ __device__ WinGeom(int B_, int H_, int W_, int ws_, int shift_)
      : B(B_), H(H_), W(W_), ws(ws_), shift(shift_) {
    Hp = (H + ws - 1) / ws * ws;
    Wp = (W + ws - 1) / ws * ws;
    nWh = Hp / ws;
    nWw = Wp / ws;
  }
  // source (h,w) in the un-padded map for slot (i,j) of window (wh,ww) of the
  // padded tensor after roll(-shift); returns false for a padded slot.
  __host__ __device__ __forceinline__ bool source(int wh, int ww, int i, int j, int &h, int &w) const {
    h = wh * ws + i + shift;
    if (h >= Hp) h -= Hp;
    w = ww * ws + j + shift;
    if (w >= Wp) w -= Wp;
    return h < H && w < W;
  }
  // region id (0..8) of the shift mask for slot (i,j) of window (wh,ww):
  // slices (0,-ws), (-ws,-shift), (-shift,None) on the PADDED rolled tensor.
  __host__ __device__ __forceinline__ int region(int wh, int ww, int i, int j) const {
    int hp = wh * ws + i, wp = ww * ws + j;
    int rh = hp < Hp - ws ? 0 : (hp < Hp - shift ? 1 : 2);
    int rw = wp < Wp - ws ? 0 : (wp < Wp - shift ? 1 : 2);
    return rh * 3 + rw;
  }
};

}  // namespace rsc
