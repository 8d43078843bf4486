// Window index maps and stand-alone partition / reverse copies.
// pad -> roll(-shift) -> window_partition and its inverse, expressed on
// indices (mmdet ShiftWindowMSA.forward).  The same
// WinGeom::source() is used by the fused attention kernels, so the bit-exact
// index test covers them too.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace rsc {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

__global__ void window_index_partition_kernel(int64_t *__restrict__ out, WinGeom g, int64_t total) {
  const int n = g.ws * g.ws;
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(s % n);
    int64_t win = s / n;
    int ww = (int)(win % g.nWw);
    int wh = (int)((win / g.nWw) % g.nWh);
    int b = (int)(win / ((int64_t)g.nWw * g.nWh));
    int h, w;
    bool ok = g.source(wh, ww, t / g.ws, t % g.ws, h, w);
    out[s] = ok ? ((int64_t)b * g.H + h) * g.W + w : -1;
  }
}

// inverse map: token (b,h,w) -> slot.  rolled position hp = (h - shift) mod Hp.
__device__ __forceinline__ int64_t token_to_slot(const WinGeom &g, int b, int h, int w) {
  int hp = h - g.shift;
  if (hp < 0) hp += g.Hp;
  int wp = w - g.shift;
  if (wp < 0) wp += g.Wp;
  int64_t win = ((int64_t)b * g.nWh + hp / g.ws) * g.nWw + wp / g.ws;
  return win * (g.ws * g.ws) + (hp % g.ws) * g.ws + (wp % g.ws);
}

__global__ void window_index_reverse_kernel(int64_t *__restrict__ out, WinGeom g, int64_t total) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x) {
    int w = (int)(s % g.W);
    int h = (int)((s / g.W) % g.H);
    int b = (int)(s / ((int64_t)g.W * g.H));
    out[s] = token_to_slot(g, b, h, w);
  }
}

// one warp per window slot / token, lanes over channels (4 elements per lane)
template <typename T>
__global__ void window_partition_kernel(const T *__restrict__ x, T *__restrict__ win, WinGeom g, int C, int64_t slots) {
  const int lane = threadIdx.x & 31;
  const int n = g.ws * g.ws;
  int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = warp; s < slots; s += nwarps) {
    int t = (int)(s % n);
    int64_t wi = s / n;
    int ww = (int)(wi % g.nWw);
    int wh = (int)((wi / g.nWw) % g.nWh);
    int b = (int)(wi / ((int64_t)g.nWw * g.nWh));
    int h, w;
    bool ok = g.source(wh, ww, t / g.ws, t % g.ws, h, w);
    const T *src = x + (((int64_t)b * g.H + h) * g.W + w) * C;
    T *dst = win + s * C;
    for (int c = lane * 4; c < C; c += 128) {
      float4 v = ok ? load4<T>(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      store4<T>(dst + c, v);
    }
  }
}

template <typename T>
__global__ void window_reverse_kernel(const T *__restrict__ win, T *__restrict__ x, WinGeom g, int C, int64_t tokens) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = warp; s < tokens; s += nwarps) {
    int w = (int)(s % g.W);
    int h = (int)((s / g.W) % g.H);
    int b = (int)(s / ((int64_t)g.W * g.H));
    const T *src = win + token_to_slot(g, b, h, w) * C;
    T *dst = x + s * C;
    for (int c = lane * 4; c < C; c += 128) store4<T>(dst + c, load4<T>(src + c));
  }
}

static int check_geom(const char *fn, int B, int H, int W, int ws, int shift) {
  RSC_CHECK_ARG(B > 0 && H > 0 && W > 0, "%s: empty tensor (B=%d,H=%d,W=%d)", fn, B, H, W);
  RSC_CHECK_ARG(ws > 0 && shift >= 0 && shift < ws, "%s: need 0 <= shift < ws (ws=%d, shift=%d)", fn, ws, shift);
  return RSC_OK;
}

static inline int grid_for(int64_t work_items, int per_block) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace rsc

using namespace rsc;

extern "C" {

const char *rsc_last_error(void) { return g_err; }
int rsc_version(void) { return 100; }
int64_t rsc_launch_count(void) { return g_launches.load(); }
void rsc_reset_launch_count(void) { g_launches.store(0); }

int rsc_window_index_partition(int64_t *idx_out, int B, int H, int W, int ws, int shift, void *stream) {
  if (int e = check_geom("rsc_window_index_partition", B, H, W, ws, shift)) return e;
  RSC_CHECK_ARG(idx_out, "rsc_window_index_partition: null output");
  WinGeom g(B, H, W, ws, shift);
  int64_t total = (int64_t)B * g.nWh * g.nWw * ws * ws;
  window_index_partition_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(idx_out, g, total);
  RSC_CHECK_LAUNCH("rsc_window_index_partition");
  return RSC_OK;
}

int rsc_window_index_reverse(int64_t *idx_out, int B, int H, int W, int ws, int shift, void *stream) {
  if (int e = check_geom("rsc_window_index_reverse", B, H, W, ws, shift)) return e;
  RSC_CHECK_ARG(idx_out, "rsc_window_index_reverse: null output");
  WinGeom g(B, H, W, ws, shift);
  int64_t total = (int64_t)B * H * W;
  window_index_reverse_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(idx_out, g, total);
  RSC_CHECK_LAUNCH("rsc_window_index_reverse");
  return RSC_OK;
}

int rsc_window_partition(const void *x, void *windows, int B, int H, int W, int C, int ws, int shift, int dtype,
                         void *stream) {
  if (int e = check_geom("rsc_window_partition", B, H, W, ws, shift)) return e;
  RSC_CHECK_ARG(x && windows, "rsc_window_partition: null pointer");
  RSC_CHECK_ARG(C > 0 && C % 4 == 0, "rsc_window_partition: C must be a positive multiple of 4 (C=%d)", C);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "rsc_window_partition: bad dtype %d", dtype);
  WinGeom g(B, H, W, ws, shift);
  int64_t slots = (int64_t)B * g.nWh * g.nWw * ws * ws;
  int grid = grid_for(slots, 8);
  if (dtype == RSC_F32)
    window_partition_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float *)x, (float *)windows, g, C, slots);
  else
    window_partition_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)x, (__nv_bfloat16 *)windows, g, C, slots);
  RSC_CHECK_LAUNCH("rsc_window_partition");
  return RSC_OK;
}

int rsc_window_reverse(const void *windows, void *x, int B, int H, int W, int C, int ws, int shift, int dtype,
                       void *stream) {
  if (int e = check_geom("rsc_window_reverse", B, H, W, ws, shift)) return e;
  RSC_CHECK_ARG(x && windows, "rsc_window_reverse: null pointer");
  RSC_CHECK_ARG(C > 0 && C % 4 == 0, "rsc_window_reverse: C must be a positive multiple of 4 (C=%d)", C);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "rsc_window_reverse: bad dtype %d", dtype);
  WinGeom g(B, H, W, ws, shift);
  int64_t tokens = (int64_t)B * H * W;
  int grid = grid_for(tokens, 8);
  if (dtype == RSC_F32)
    window_reverse_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float *)windows, (float *)x, g, C, tokens);
  else
    window_reverse_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)windows, (__nv_bfloat16 *)x, g, C, tokens);
  RSC_CHECK_LAUNCH("rsc_window_reverse");
  return RSC_OK;
}

}  // extern "C"
