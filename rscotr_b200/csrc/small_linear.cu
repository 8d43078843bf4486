// Backward of the SMALL Linear layers (fewer than ~4096 token rows: the decoders' projections, the reg / cls branches, the
// mask-embedding MLPs) in ONE launch:
//     dX (M,K) = dY (M,N) W (N,K)          bf16 out
//     dW (N,K) += dY^T X                    fp32, accumulated into the step engine's flat gradient buffer
//     db (N)   += column sums of dY         fp32, accumulated
// These GEMMs are 0.1 - 1 GFLOP: the time is launch latency, not math -- the library path costs three launches per layer
// (dX GEMM, dW GEMM, bias column sum; ~450 launches per det + seg step pair).  A persistent tcgen05 kernel does not pay at this
// size either (TMEM allocation, 200 KB of shared memory, pipeline fill), so this is a plain tiled warp-MMA kernel
// (mma.sync.m16n8k16, cp.async double buffering): CTAs [0, tiles_dx) compute 64x64 tiles of dX, the rest 64x64 tiles of dW
// over a slice of the token range (red.global.add.v2.f32), the dW CTAs of the first column tile also sum dY's columns.
#include "common.cuh"

namespace rsc {
namespace slin {

constexpr int BT = 64, BKK = 32;          // output tile 64 x 64, contraction step 32
constexpr int A_ROW = (BKK + 8) * 2;      // bytes per row of the dX role's A tile  [64 m][32 n]
constexpr int T_ROW = (BT + 8) * 2;       // bytes per row of the [32][64] tiles (B tiles of both roles, dY^T tile of the dW role)
constexpr int A_BYTES = BT * A_ROW > BKK * T_ROW ? BT * A_ROW : BKK * T_ROW;
constexpr int B_BYTES = BKK * T_ROW;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(uint32_t dst, const void *src, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct Params {
  const __nv_bfloat16 *dy, *x, *w;
  __nv_bfloat16 *dx;
  float *dw, *db;
  int M, N, K;
  int64_t lddy, ldx, ldw, lddx, lddw;
  int tiles_dx, kt_dx;            // dX role: tiles = ceil(M/64) * kt_dx
  int kt_dw, splits, rows_per_split;
};

// [rows][cols] bf16 sub-matrix of a row-major matrix -> padded tile; 16-byte chunks, zero fill outside (R, C)
template <int ROWS, int COLS, int ROWB>
__device__ __forceinline__ void load_tile(unsigned char *tile, const __nv_bfloat16 *base, int64_t ld, int r0, int c0, int R, int C,
                                          int tid) {
  constexpr int CH = COLS / 8, TOTAL = ROWS * CH;
#pragma unroll
  for (int i = 0; i < TOTAL / 128; ++i) {
    const int idx = tid + i * 128, r = idx / CH, c = idx % CH;
    const bool ok = r0 + r < R && c0 + c * 8 < C;
    cp16(smem_u32(tile + r * ROWB + c * 16), base + (ok ? (int64_t)(r0 + r) * ld + c0 + c * 8 : 0), ok);
  }
}

__global__ void __launch_bounds__(128) small_linear_bwd_kernel(const Params p) {
  __shared__ __align__(16) unsigned char As[2][A_BYTES], Bs[2][B_BYTES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

  const bool role_dx = (int)blockIdx.x < p.tiles_dx;
  int r0, c0, s0, s1;       // output tile origin (rows, cols) and the contraction range [s0, s1)
  bool do_db = false;
  if (role_dx) {
    r0 = ((int)blockIdx.x / p.kt_dx) * BT, c0 = ((int)blockIdx.x % p.kt_dx) * BT;   // (m0, k0), contraction over n
    s0 = 0, s1 = p.N;
  } else {
    int u = (int)blockIdx.x - p.tiles_dx;
    const int split = u % p.splits;
    u /= p.splits;
    r0 = (u / p.kt_dw) * BT, c0 = (u % p.kt_dw) * BT;                               // (n0, k0), contraction over m
    s0 = split * p.rows_per_split, s1 = min(p.M, s0 + p.rows_per_split);
    do_db = p.db != nullptr && c0 == 0;
  }
  auto stage = [&](int s, int buf) {
    if (role_dx) {
      load_tile<BT, BKK, A_ROW>(As[buf], p.dy, p.lddy, r0, s, p.M, p.N, tid);      // dY[m0.., n-step]
      load_tile<BKK, BT, T_ROW>(Bs[buf], p.w, p.ldw, s, c0, p.N, p.K, tid);        // W[n-step, k0..]
    } else {
      load_tile<BKK, BT, T_ROW>(As[buf], p.dy, p.lddy, s, r0, s1, p.N, tid);       // dY[m-step, n0..]   (A^T)
      load_tile<BKK, BT, T_ROW>(Bs[buf], p.x, p.ldx, s, c0, s1, p.K, tid);         // X[m-step, k0..]
    }
  };
  float cs = 0.f;
  if (s0 < s1) stage(s0, 0);
  cp_commit();
  int buf = 0;
  for (int s = s0; s < s1; s += BKK, buf ^= 1) {
    cp_wait_all();
    __syncthreads();
    if (s + BKK < s1) stage(s + BKK, buf ^ 1);
    cp_commit();
    const uint32_t a_base = smem_u32(As[buf]), b_base = smem_u32(Bs[buf]);
    if (do_db && tid < BT) {                       // column sums of the dY tile (rows = tokens, column tid = output feature)
#pragma unroll 8
      for (int r = 0; r < BKK; ++r) cs += __bfloat162float(*reinterpret_cast<const __nv_bfloat16 *>(As[buf] + r * T_ROW + tid * 2));
    }
#pragma unroll
    for (int kk = 0; kk < BKK / 16; ++kk) {
      uint32_t a[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (role_dx)
          ldsm4(a_base + (wm * 32 + i * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * A_ROW + (kk * 16 + ((lane >> 4) & 1) * 8) * 2, a[i]);
        else
          ldsm4t(a_base + (kk * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * T_ROW + (wm * 32 + i * 16 + ((lane >> 3) & 1) * 8) * 2, a[i]);
      }
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        uint32_t b[4];
        ldsm4t(b_base + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * T_ROW + (wn * 32 + jp * 16 + ((lane >> 4) & 1) * 8) * 2, b);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          mma(acc[i][2 * jp], a[i], b[0], b[1]);
          mma(acc[i][2 * jp + 1], a[i], b[2], b[3]);
        }
      }
    }
  }
  cp_wait_all();
  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = r0 + wm * 32 + i * 16 + g + h * 8;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = c0 + wn * 32 + j * 8 + 2 * t;
        const float v0 = acc[i][j][2 * h], v1 = acc[i][j][2 * h + 1];
        if (role_dx) {
          if (row < p.M && col < p.K) {
            __nv_bfloat162 o = __floats2bfloat162_rn(v0, v1);
            *reinterpret_cast<__nv_bfloat162 *>(p.dx + (int64_t)row * p.lddx + col) = o;
          }
        } else if (row < p.N && col < p.K) {
          asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p.dw + (int64_t)row * p.lddw + col), "f"(v0), "f"(v1) : "memory");
        }
      }
    }
  if (do_db && tid < BT && r0 + tid < p.N) atomicAdd(p.db + r0 + tid, cs);
}

}  // namespace slin
}  // namespace rsc

using namespace rsc;

// dx may be NULL (no input gradient wanted), dw may be NULL (frozen weight), db may be NULL
extern "C" int rsc_small_linear_bwd(const void *dy, const void *x, const void *w, void *dx, float *dw, float *db, int M, int N, int K,
                                    int64_t lddy, int64_t ldx, int64_t ldw, int64_t lddx, int64_t lddw, void *stream) {
  RSC_CHECK_ARG(dy && M > 0 && N > 0 && K > 0, "rsc_small_linear_bwd: null pointer / empty shape");
  RSC_CHECK_ARG((dx == nullptr || w != nullptr) && (dw == nullptr || x != nullptr), "rsc_small_linear_bwd: dx needs w, dw needs x");
  RSC_CHECK_ARG(N % 8 == 0 && K % 8 == 0 && lddy % 8 == 0 && (x == nullptr || ldx % 8 == 0) && (w == nullptr || ldw % 8 == 0) &&
                    (dx == nullptr || lddx % 2 == 0) && (dw == nullptr || lddw % 2 == 0),
                "rsc_small_linear_bwd: N, K and the leading dimensions must be multiples of 8 (16-byte rows)");
  RSC_CHECK_ARG((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)w) & 15) == 0 && (((uintptr_t)dx) & 3) == 0 && (((uintptr_t)dw) & 7) == 0,
                "rsc_small_linear_bwd: alignment");
  slin::Params p = {};
  p.dy = (const __nv_bfloat16 *)dy, p.x = (const __nv_bfloat16 *)x, p.w = (const __nv_bfloat16 *)w, p.dx = (__nv_bfloat16 *)dx;
  p.dw = dw, p.db = db, p.M = M, p.N = N, p.K = K;
  p.lddy = lddy, p.ldx = ldx, p.ldw = ldw, p.lddx = lddx, p.lddw = lddw;
  const int mt = (M + slin::BT - 1) / slin::BT, kt = (K + slin::BT - 1) / slin::BT, nt = (N + slin::BT - 1) / slin::BT;
  p.kt_dx = kt;
  p.tiles_dx = dx ? mt * kt : 0;
  int tiles_dw = 0;
  p.kt_dw = kt, p.splits = 1, p.rows_per_split = M;
  if (dw || db) {
    RSC_CHECK_ARG(dw != nullptr, "rsc_small_linear_bwd: db without dw is not supported");
    int splits = (M + 255) / 256;                    // >= 8 contraction steps per CTA
    splits = splits < 1 ? 1 : (splits > 32 ? 32 : splits);
    int rps = ((M + splits - 1) / splits + slin::BKK - 1) / slin::BKK * slin::BKK;
    splits = (M + rps - 1) / rps;
    p.splits = splits, p.rows_per_split = rps;
    tiles_dw = nt * kt * splits;
  }
  const int grid = p.tiles_dx + tiles_dw;
  if (grid == 0) return RSC_OK;
  slin::small_linear_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  RSC_CHECK_LAUNCH("rsc_small_linear_bwd");
  return RSC_OK;
}
