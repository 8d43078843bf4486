// Fused (shifted-)window attention core, SIMT fp32-accumulate version.
//
// One kernel = pad + cyclic shift + window_partition + per-head
// softmax(q*scale @ k^T + rel_pos_bias + shift_mask) @ v + window_reverse +
// un-shift + crop (mmdet ShiftWindowMSA/WindowMSA).
// qkv is read ONCE from its natural (B,H,W,3C) layout through the shifted /
// padded coordinates; the attention matrix never leaves the SM.
//
// This file is the exact-arithmetic path (fp32 I/O for the 1e-3 parity runs,
// bf16 I/O supported).  Work unit = (window, head): 49 tokens x 32 dims.
#include "common.cuh"

namespace rsc {

constexpr int WS = 7;
constexpr int NT = WS * WS;       // 49 tokens per window
constexpr int HD = 32;            // head dim (all Swin sizes)
constexpr int PITCH = HD + 4;     // smem row pitch (floats): LDS.128 conflict free
constexpr int GROUP = 64;         // threads per (window, head) unit
constexpr int TBL = (2 * WS - 1) * (2 * WS - 1);  // 169

// cooperative load of the q|k|v rows of `nh` consecutive heads of one window
// into smem[unit][part][token][PITCH]
template <typename T, int PARTS>
__device__ __forceinline__ void load_window_rows(const T *__restrict__ base, const float *__restrict__ pad_row,
                                                 float *smem, int unit_stride, const WinGeom &g, int b, int wh, int ww,
                                                 int C, int row_stride, int h0, int nh, int tid, int nthreads) {
  // base: tensor with `row_stride` elements per token, PARTS sections of C
  const int chunks_per_run = 8 * nh;
  const int total = NT * PARTS * chunks_per_run;
  for (int idx = tid; idx < total; idx += nthreads) {
    int c8 = idx % chunks_per_run;
    int part = (idx / chunks_per_run) % PARTS;
    int t = idx / (chunks_per_run * PARTS);
    int h, w;
    bool ok = g.source(wh, ww, t / WS, t % WS, h, w);
    int col = part * C + h0 * HD + c8 * 4;
    float4 v;
    if (ok) {
      v = load4<T>(base + (((int64_t)b * g.H + h) * g.W + w) * row_stride + col);
    } else if (pad_row) {
      v = __ldg(reinterpret_cast<const float4 *>(pad_row + col));
    } else {
      v = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float *dst = smem + (c8 >> 3) * unit_stride + (part * NT + t) * PITCH + (c8 & 7) * 4;
    *reinterpret_cast<float4 *>(dst) = v;
  }
}

__device__ __forceinline__ void decode_window(const WinGeom &g, int win, int &b, int &wh, int &ww) {
  ww = win % g.nWw;
  wh = (win / g.nWw) % g.nWh;
  b = win / (g.nWw * g.nWh);
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
constexpr int FWD_UNIT_FLOATS = 3 * NT * PITCH + TBL + 3;  // q|k|v rows + bias table (+pad to 4)

template <typename T, int HPB>
__global__ void __launch_bounds__(GROUP *HPB)
    wmsa_fwd_kernel(const T *__restrict__ qkv, const float *__restrict__ qkv_bias, const float *__restrict__ table,
                    T *__restrict__ out, WinGeom g, int C, int heads, float scale) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int unit = tid / GROUP;
  const int r = tid % GROUP;
  const int h0 = blockIdx.y * HPB;
  int b, wh, ww;
  decode_window(g, blockIdx.x, b, wh, ww);

  load_window_rows<T, 3>(qkv, qkv_bias, smem, FWD_UNIT_FLOATS, g, b, wh, ww, C, 3 * C, h0, HPB, tid, GROUP * HPB);
  float *my = smem + unit * FWD_UNIT_FLOATS;
  float *tbl = my + 3 * NT * PITCH;
  for (int i = r; i < TBL; i += GROUP) tbl[i] = __ldg(table + i * heads + h0 + unit);
  __syncthreads();

  const float *Q = my, *K = my + NT * PITCH, *V = my + 2 * NT * PITCH;
  float o[HD];
  float l = 0.f;
  if (r < NT) {
    const int ri = r / WS, ci = r % WS;
    const int reg_i = g.shift > 0 ? g.region(wh, ww, ri, ci) : 0;
    float q[HD];
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      float4 t4 = *reinterpret_cast<const float4 *>(Q + r * PITCH + d);
      q[d] = t4.x * scale, q[d + 1] = t4.y * scale, q[d + 2] = t4.z * scale, q[d + 3] = t4.w * scale;
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    float m = -INFINITY;
#pragma unroll 1
    for (int jr = 0; jr < WS; ++jr) {
      float s[WS];
      float mc = m;
#pragma unroll
      for (int jc = 0; jc < WS; ++jc) {
        const float *kj = K + (jr * WS + jc) * PITCH;
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          float4 k4 = *reinterpret_cast<const float4 *>(kj + d);
          acc = fmaf(q[d], k4.x, acc);
          acc = fmaf(q[d + 1], k4.y, acc);
          acc = fmaf(q[d + 2], k4.z, acc);
          acc = fmaf(q[d + 3], k4.w, acc);
        }
        acc += tbl[(ri - jr + WS - 1) * (2 * WS - 1) + (ci - jc + WS - 1)];
        if (g.shift > 0 && g.region(wh, ww, jr, jc) != reg_i) acc += -100.0f;
        s[jc] = acc;
        mc = fmaxf(mc, acc);
      }
      const float corr = __expf(m - mc);  // m = -inf on the first chunk -> 0
      l *= corr;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] *= corr;
      m = mc;
#pragma unroll
      for (int jc = 0; jc < WS; ++jc) {
        const float p = __expf(s[jc] - m);
        l += p;
        const float *vj = V + (jr * WS + jc) * PITCH;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          float4 v4 = *reinterpret_cast<const float4 *>(vj + d);
          o[d] = fmaf(p, v4.x, o[d]);
          o[d + 1] = fmaf(p, v4.y, o[d + 1]);
          o[d + 2] = fmaf(p, v4.z, o[d + 2]);
          o[d + 3] = fmaf(p, v4.w, o[d + 3]);
        }
      }
    }
    const float inv = 1.0f / l;
    // stage the output row in this thread's own q row (read by nobody else)
    float *orow = my + r * PITCH;
#pragma unroll
    for (int d = 0; d < HD; d += 4)
      *reinterpret_cast<float4 *>(orow + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
  }
  __syncthreads();
  // coalesced store: per token a run of 32*HPB contiguous channels
  const int chunks_per_run = 8 * HPB;
  for (int idx = tid; idx < NT * chunks_per_run; idx += GROUP * HPB) {
    int c8 = idx % chunks_per_run;
    int t = idx / chunks_per_run;
    int h, w;
    if (!g.source(wh, ww, t / WS, t % WS, h, w)) continue;
    float4 v = *reinterpret_cast<const float4 *>(smem + (c8 >> 3) * FWD_UNIT_FLOATS + t * PITCH + (c8 & 7) * 4);
    store4<T>(out + (((int64_t)b * g.H + h) * g.W + w) * C + h0 * HD + c8 * 4, v);
  }
}

// ---------------------------------------------------------------------------
// backward: one (window, head) unit per 64-thread CTA, persistent over units
// with a fixed head per CTA so the bias-table gradient is reduced on chip.
// ---------------------------------------------------------------------------
constexpr int SP = NT;  // pitch of the 49x49 matrices (odd -> conflict free both ways)
constexpr int BWD_FLOATS = 4 * NT * PITCH + 3 * NT * SP + TBL + 3 * HD + 4;

template <typename T>
__global__ void __launch_bounds__(GROUP)
    wmsa_bwd_kernel(const T *__restrict__ qkv, const float *__restrict__ qkv_bias, const float *__restrict__ table,
                    const T *__restrict__ dout, T *__restrict__ dqkv, float *__restrict__ dtable,
                    float *__restrict__ dqkv_bias, WinGeom g, int C, int heads, float scale, int num_units) {
  extern __shared__ __align__(16) float smem[];
  float *Q = smem;                 // [3][NT][PITCH]: q|k|v
  float *K = Q + NT * PITCH;
  float *V = K + NT * PITCH;
  float *DO = V + NT * PITCH;      // [NT][PITCH]
  float *P = DO + NT * PITCH;      // [NT][SP]
  float *DS = P + NT * SP;         // [NT][SP]
  float *DB = DS + NT * SP;        // [NT][SP] running sum of dS over this CTA's units
  float *tbl = DB + NT * SP;       // [TBL] (+3 pad)
  float *padacc = tbl + TBL + 3;   // [3][HD] gradient reaching the qkv bias via padded rows
  const int r = threadIdx.x;
  const int head = blockIdx.x % heads;  // gridDim.x is a multiple of heads

  for (int i = r; i < NT * SP; i += GROUP) DB[i] = 0.f;
  for (int i = r; i < 3 * HD; i += GROUP) padacc[i] = 0.f;
  for (int i = r; i < TBL; i += GROUP) tbl[i] = __ldg(table + i * heads + head);

  for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
    int b, wh, ww;
    decode_window(g, u / heads, b, wh, ww);
    __syncthreads();  // previous unit fully consumed (and init visible)
    load_window_rows<T, 3>(qkv, qkv_bias, Q, 0, g, b, wh, ww, C, 3 * C, head, 1, r, GROUP);
    load_window_rows<T, 1>(dout, nullptr, DO, 0, g, b, wh, ww, C, C, head, 1, r, GROUP);
    __syncthreads();

    int sh = 0, sw = 0;
    bool valid = false;
    if (r < NT) valid = g.source(wh, ww, r / WS, r % WS, sh, sw);
    const int64_t tok = ((int64_t)b * g.H + sh) * g.W + sw;

    // ---- pass 1: thread = query row i ----
    if (r < NT) {
      const int ri = r / WS, ci = r % WS;
      const int reg_i = g.shift > 0 ? g.region(wh, ww, ri, ci) : 0;
      float q[HD], go[HD];
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        float4 t4 = *reinterpret_cast<const float4 *>(Q + r * PITCH + d);
        q[d] = t4.x * scale, q[d + 1] = t4.y * scale, q[d + 2] = t4.z * scale, q[d + 3] = t4.w * scale;
        float4 g4 = *reinterpret_cast<const float4 *>(DO + r * PITCH + d);
        go[d] = g4.x, go[d + 1] = g4.y, go[d + 2] = g4.z, go[d + 3] = g4.w;
      }
      float m = -INFINITY;
#pragma unroll 1
      for (int j = 0; j < NT; ++j) {
        const float *kj = K + j * PITCH;
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          float4 k4 = *reinterpret_cast<const float4 *>(kj + d);
          acc = fmaf(q[d], k4.x, acc);
          acc = fmaf(q[d + 1], k4.y, acc);
          acc = fmaf(q[d + 2], k4.z, acc);
          acc = fmaf(q[d + 3], k4.w, acc);
        }
        const int jr = j / WS, jc = j % WS;
        acc += tbl[(ri - jr + WS - 1) * (2 * WS - 1) + (ci - jc + WS - 1)];
        if (g.shift > 0 && g.region(wh, ww, jr, jc) != reg_i) acc += -100.0f;
        P[r * SP + j] = acc;
        m = fmaxf(m, acc);
      }
      float l = 0.f;
#pragma unroll 1
      for (int j = 0; j < NT; ++j) {
        float p = __expf(P[r * SP + j] - m);
        P[r * SP + j] = p;
        l += p;
      }
      const float inv = 1.0f / l;
      float Dsum = 0.f;
#pragma unroll 1
      for (int j = 0; j < NT; ++j) {
        const float *vj = V + j * PITCH;
        float dp = 0.f;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          float4 v4 = *reinterpret_cast<const float4 *>(vj + d);
          dp = fmaf(go[d], v4.x, dp);
          dp = fmaf(go[d + 1], v4.y, dp);
          dp = fmaf(go[d + 2], v4.z, dp);
          dp = fmaf(go[d + 3], v4.w, dp);
        }
        const float p = P[r * SP + j] * inv;
        P[r * SP + j] = p;
        DS[r * SP + j] = dp;
        Dsum = fmaf(p, dp, Dsum);
      }
      float dq[HD];
#pragma unroll
      for (int d = 0; d < HD; ++d) dq[d] = 0.f;
#pragma unroll 1
      for (int j = 0; j < NT; ++j) {
        const float ds = P[r * SP + j] * (DS[r * SP + j] - Dsum);
        DS[r * SP + j] = ds;
        DB[r * SP + j] += ds;
        const float *kj = K + j * PITCH;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          float4 k4 = *reinterpret_cast<const float4 *>(kj + d);
          dq[d] = fmaf(ds, k4.x, dq[d]);
          dq[d + 1] = fmaf(ds, k4.y, dq[d + 1]);
          dq[d + 2] = fmaf(ds, k4.z, dq[d + 2]);
          dq[d + 3] = fmaf(ds, k4.w, dq[d + 3]);
        }
      }
      if (valid) {
        T *dst = dqkv + tok * 3 * C + head * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 4)
          store4<T>(dst + d, make_float4(dq[d] * scale, dq[d + 1] * scale, dq[d + 2] * scale, dq[d + 3] * scale));
      } else if (dqkv_bias) {
#pragma unroll
        for (int d = 0; d < HD; ++d) atomicAdd(padacc + d, dq[d] * scale);
      }
    }
    __syncthreads();
    // ---- pass 2: thread = key row j ----
    if (r < NT) {
      float dk[HD], dv[HD];
#pragma unroll
      for (int d = 0; d < HD; ++d) dk[d] = 0.f, dv[d] = 0.f;
#pragma unroll 1
      for (int i = 0; i < NT; ++i) {
        const float ds = DS[i * SP + r];
        const float p = P[i * SP + r];
        const float *qi = Q + i * PITCH;
        const float *gi = DO + i * PITCH;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          float4 q4 = *reinterpret_cast<const float4 *>(qi + d);
          float4 g4 = *reinterpret_cast<const float4 *>(gi + d);
          dk[d] = fmaf(ds, q4.x, dk[d]);
          dk[d + 1] = fmaf(ds, q4.y, dk[d + 1]);
          dk[d + 2] = fmaf(ds, q4.z, dk[d + 2]);
          dk[d + 3] = fmaf(ds, q4.w, dk[d + 3]);
          dv[d] = fmaf(p, g4.x, dv[d]);
          dv[d + 1] = fmaf(p, g4.y, dv[d + 1]);
          dv[d + 2] = fmaf(p, g4.z, dv[d + 2]);
          dv[d + 3] = fmaf(p, g4.w, dv[d + 3]);
        }
      }
      if (valid) {
        T *dst = dqkv + tok * 3 * C + C + head * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          store4<T>(dst + d, make_float4(dk[d] * scale, dk[d + 1] * scale, dk[d + 2] * scale, dk[d + 3] * scale));
          store4<T>(dst + C + d, make_float4(dv[d], dv[d + 1], dv[d + 2], dv[d + 3]));
        }
      } else if (dqkv_bias) {
#pragma unroll
        for (int d = 0; d < HD; ++d) {
          atomicAdd(padacc + HD + d, dk[d] * scale);
          atomicAdd(padacc + 2 * HD + d, dv[d]);
        }
      }
    }
  }
  __syncthreads();
  // ---- flush: fold DB[i][j] onto the 169-entry table (reuse P as scratch) ----
  for (int i = r; i < TBL; i += GROUP) P[i] = 0.f;
  __syncthreads();
  for (int idx = r; idx < NT * NT; idx += GROUP) {
    int i = idx / NT, j = idx % NT;
    int t = (i / WS - j / WS + WS - 1) * (2 * WS - 1) + (i % WS - j % WS + WS - 1);
    atomicAdd(P + t, DB[i * SP + j]);
  }
  __syncthreads();
  for (int i = r; i < TBL; i += GROUP) atomicAdd(dtable + i * heads + head, P[i]);
  if (dqkv_bias)
    for (int i = r; i < 3 * HD; i += GROUP) {
      float v = padacc[i];
      if (v != 0.f) atomicAdd(dqkv_bias + (i / HD) * C + head * HD + (i % HD), v);
    }
}

static int check_args(const char *fn, int B, int H, int W, int C, int heads, int ws, int shift, int dtype) {
  RSC_CHECK_ARG(B > 0 && H > 0 && W > 0, "%s: empty tensor (B=%d,H=%d,W=%d)", fn, B, H, W);
  RSC_CHECK_ARG(ws == WS, "%s: window_size must be 7 (got %d)", fn, ws);
  RSC_CHECK_ARG(shift == 0 || shift == ws / 2, "%s: shift must be 0 or %d (got %d)", fn, ws / 2, shift);
  RSC_CHECK_ARG(heads > 0 && C == heads * HD, "%s: head_dim must be 32 (C=%d, heads=%d)", fn, C, heads);
  RSC_CHECK_ARG(dtype == RSC_F32 || dtype == RSC_BF16, "%s: bad dtype %d", fn, dtype);
  return RSC_OK;
}

template <typename T, int HPB>
static int launch_fwd(const void *qkv, const float *qkv_bias, const float *table, void *out, const WinGeom &g, int C,
                      int heads, float scale, cudaStream_t st) {
  size_t smem = sizeof(float) * FWD_UNIT_FLOATS * HPB;
  auto kern = wmsa_fwd_kernel<T, HPB>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(g.B * g.nWh * g.nWw, heads / HPB);
  kern<<<grid, GROUP * HPB, smem, st>>>((const T *)qkv, qkv_bias, table, (T *)out, g, C, heads, scale);
  RSC_CHECK_LAUNCH("rsc_wmsa_fwd");
  return RSC_OK;
}

template <typename T>
static int dispatch_fwd(const void *qkv, const float *qkv_bias, const float *table, void *out, const WinGeom &g,
                        int C, int heads, float scale, cudaStream_t st) {
  if (heads % 4 == 0) return launch_fwd<T, 4>(qkv, qkv_bias, table, out, g, C, heads, scale, st);
  if (heads % 3 == 0) return launch_fwd<T, 3>(qkv, qkv_bias, table, out, g, C, heads, scale, st);
  if (heads % 2 == 0) return launch_fwd<T, 2>(qkv, qkv_bias, table, out, g, C, heads, scale, st);
  return launch_fwd<T, 1>(qkv, qkv_bias, table, out, g, C, heads, scale, st);
}

}  // namespace rsc

using namespace rsc;

// The tensor-core (tcgen05) bf16 kernels and the rsc_wmsa_{fwd,bwd} dispatch live in wmsa_tma.cu; these entries take
// fp32 (the exact-arithmetic parity path), RSC_WMSA_SIMT=1 and every argument error.
extern "C" int rsc_wmsa_fwd_simt(const void *qkv, const float *qkv_bias, const float *bias_table, void *out, int B,
                                 int H, int W, int C, int heads, int ws, int shift, float scale, int dtype,
                                 void *stream) {
  if (int e = check_args("rsc_wmsa_fwd", B, H, W, C, heads, ws, shift, dtype)) return e;
  RSC_CHECK_ARG(qkv && bias_table && out, "rsc_wmsa_fwd: null pointer");
  WinGeom g(B, H, W, ws, shift);
  if (dtype == RSC_F32) return dispatch_fwd<float>(qkv, qkv_bias, bias_table, out, g, C, heads, scale, (cudaStream_t)stream);
  return dispatch_fwd<__nv_bfloat16>(qkv, qkv_bias, bias_table, out, g, C, heads, scale, (cudaStream_t)stream);
}

extern "C" int rsc_wmsa_bwd_simt(const void *qkv, const float *qkv_bias, const float *bias_table, const void *dout,
                            void *dqkv, float *dbias_table, float *dqkv_bias, int B, int H, int W, int C, int heads,
                            int ws, int shift, float scale, int dtype, void *stream) {
  if (int e = check_args("rsc_wmsa_bwd", B, H, W, C, heads, ws, shift, dtype)) return e;
  RSC_CHECK_ARG(qkv && bias_table && dout && dqkv && dbias_table, "rsc_wmsa_bwd: null pointer");
  RSC_CHECK_ARG(!(dqkv_bias && !qkv_bias), "rsc_wmsa_bwd: dqkv_bias given without qkv_bias");
  WinGeom g(B, H, W, ws, shift);
  int units = B * g.nWh * g.nWw * heads;
  size_t smem = sizeof(float) * BWD_FLOATS;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  int grid = kNumSMs * per_sm;
  grid = grid / heads * heads;
  if (grid < heads) grid = heads;
  int cap = (units + heads - 1) / heads * heads;
  if (grid > cap) grid = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RSC_F32) {
    auto kern = wmsa_bwd_kernel<float>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, GROUP, smem, st>>>((const float *)qkv, qkv_bias, bias_table, (const float *)dout, (float *)dqkv,
                                    dbias_table, dqkv_bias, g, C, heads, scale, units);
  } else {
    auto kern = wmsa_bwd_kernel<__nv_bfloat16>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, GROUP, smem, st>>>((const __nv_bfloat16 *)qkv, qkv_bias, bias_table, (const __nv_bfloat16 *)dout,
                                    (__nv_bfloat16 *)dqkv, dbias_table, dqkv_bias, g, C, heads, scale, units);
  }
  RSC_CHECK_LAUNCH("rsc_wmsa_bwd");
  return RSC_OK;
}

