// Probe of tcgen05 facts the next window-attention kernel design depends on (run on a B200; one test per
// process because a faulting variant poisons the context):
//
//   tc_probe_m64 m64          M=64  MMA, D at lane 0:  which TMEM lanes receive which accumulator rows?
//   tc_probe_m64 m64_lane16   M=64  MMA, D at lane 16: accepted? where does it land?
//   tc_probe_m64 m64_lane64   M=64  MMA, D at lane 64
//   tc_probe_m64 m128_lane64  M=128 MMA over a 64-row tile with D at lane 64 (wraps? faults?)
//   tc_probe_m64 ld16x256     thread <-> (row, col) map of tcgen05.ld.16x256b.x1 after an M=128 MMA
//   tc_probe_m64 ld16x128     same for .16x128b.x1
//   tc_probe_m64 ld16x64      same for .16x64b.x1
//
// Accumulator value D[r][n] = (r + 1) + (n + 1) / 128 (exact in bf16 x bf16 -> fp32), TMEM is pre-filled
// with the sentinel -7 so untouched cells are visible.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe_m64 tools/tc_probe_m64.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../rscotr_b200/csrc/tc_common.cuh"

using namespace rsc::tc;

#define CHECK(x)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) {                                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);              \
      return 2;                                                                                    \
    }                                                                                              \
  } while (0)

// K-major operand tile with K = 16 (two 16-byte chunks per row): chunk c of row r at (r/8)*256 + c*128 + (r%8)*16
__device__ __host__ inline int tile_off16(int r, int c) { return (r / 8) * 256 + c * 128 + (r % 8) * 16; }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, float v) {
  const uint32_t u = __float_as_uint(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(u)
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

enum Mode { M64 = 0, M64_L16, M64_L64, M128_L64, LD16x256, LD16x128, LD16x64 };

__global__ void __launch_bounds__(128) probe(int mode, float *out /* [128 lanes][128 cols] or [128 thr][8] */) {
  __shared__ __align__(1024) uint8_t sA[128 * 32];   // up to 128 rows x 16 bf16
  __shared__ __align__(1024) uint8_t sB[128 * 32];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  // A[r][0] = r + 1, A[r][1] = 1; B[n][0] = 1, B[n][1] = (n + 1) / 128; everything else 0
  for (int r = tid; r < 128; r += 128) {
    __nv_bfloat16 a[16], b[16];
    for (int k = 0; k < 16; ++k) a[k] = b[k] = __float2bfloat16(0.f);
    a[0] = __float2bfloat16((float)(r + 1));
    a[1] = __float2bfloat16(1.f);
    b[0] = __float2bfloat16(1.f);
    b[1] = __float2bfloat16((float)(r + 1) / 128.f);
    for (int c = 0; c < 2; ++c) {
      *reinterpret_cast<uint4 *>(sA + tile_off16(r, c)) = *reinterpret_cast<uint4 *>(a + 8 * c);
      *reinterpret_cast<uint4 *>(sB + tile_off16(r, c)) = *reinterpret_cast<uint4 *>(b + 8 * c);
    }
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base;
  const uint32_t quad = tm + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 128; c += 32) tmem_st32(quad + c, -7.f);   // sentinel in all 128 lanes x 128 columns
  fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    fence_after_sync();
    const uint64_t ad = make_smem_desc(smem_u32(sA), 128, 256), bd = make_smem_desc(smem_u32(sB), 128, 256);
    uint32_t d = tm, idesc = make_idesc_bf16(64, 64, false, false);
    if (mode == M64_L16) d = tm + (16u << 16);
    if (mode == M64_L64) d = tm + (64u << 16);
    if (mode == M128_L64) d = tm + (64u << 16), idesc = make_idesc_bf16(128, 64, false, false);
    if (mode >= LD16x256) idesc = make_idesc_bf16(128, 64, false, false);
    mma_bf16_ss(d, ad, bd, idesc, 0);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  if (mode < LD16x256) {
    uint32_t r[32];
    for (int c = 0; c < 128; c += 32) {
      tmem_ld32(quad + c, r);
      tmem_ld_wait();
      for (int k = 0; k < 32; ++k) out[(warp * 32 + lane) * 128 + c + k] = __uint_as_float(r[k]);
    }
  } else {
    // two half-quadrant reads (lanes +0 and +16 of this warp's quadrant), columns 0..
    for (int half = 0; half < 2; ++half) {
      const uint32_t a = quad + ((uint32_t)(half * 16) << 16);
      uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
      if (mode == LD16x256)
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                     : "r"(a)
                     : "memory");
      else if (mode == LD16x128)
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a) : "memory");
      else
        asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(a) : "memory");
      tmem_ld_wait();
      float *o = out + (tid * 2 + half) * 4;
      o[0] = __uint_as_float(r0), o[1] = __uint_as_float(r1), o[2] = __uint_as_float(r2), o[3] = __uint_as_float(r3);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}

static void decode(float v, int &row, int &col) {
  row = (int)floorf(v);
  col = (int)lroundf((v - row) * 128.f) - 1;
  row -= 1;
}

int main(int argc, char **argv) {
  const char *names[] = {"m64", "m64_lane16", "m64_lane64", "m128_lane64", "ld16x256", "ld16x128", "ld16x64"};
  int mode = -1;
  for (int k = 0; k < 7; ++k)
    if (argc > 1 && !strcmp(argv[1], names[k])) mode = k;
  if (mode < 0) {
    printf("usage: %s m64|m64_lane16|m64_lane64|m128_lane64|ld16x256|ld16x128|ld16x64\n", argv[0]);
    return 1;
  }
  float *d_out;
  std::vector<float> h(128 * 128, 0.f);
  CHECK(cudaMalloc(&d_out, h.size() * sizeof(float)));
  CHECK(cudaMemset(d_out, 0, h.size() * sizeof(float)));
  probe<<<1, 128>>>(mode, d_out);
  CHECK(cudaGetLastError());
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
  printf("== %s ==\n", names[mode]);
  if (mode < LD16x256) {
    // per TMEM lane: which accumulator row sits there (from column 0) and over which columns
    for (int l = 0; l < 128; ++l) {
      int first = -1, last = -1, row = -1, bad = 0;
      for (int c = 0; c < 128; ++c) {
        const float v = h[l * 128 + c];
        if (v == -7.f) continue;
        int r, cc;
        decode(v, r, cc);
        if (first < 0) first = c, row = r;
        last = c;
        if (r != row || cc != c - first) ++bad;
      }
      if (first < 0)
        printf("lane %3d: untouched\n", l);
      else
        printf("lane %3d: row %3d  tmem cols %3d..%3d  (accumulator cols 0..%d)%s\n", l, row, first, last, last - first,
               bad ? "  [irregular]" : "");
    }
  } else {
    const int nreg = mode == LD16x256 ? 4 : mode == LD16x128 ? 2 : 1;
    for (int t = 0; t < 128; ++t)
      for (int half = 0; half < 2; ++half) {
        printf("thread %3d (warp %d lane %2d) half %d:", t, t / 32, t % 32, half);
        for (int k = 0; k < nreg; ++k) {
          const float v = h[(t * 2 + half) * 4 + k];
          int r, c;
          decode(v, r, c);
          if (v == -7.f)
            printf("  r%d=sentinel", k);
          else
            printf("  r%d=(row %d, col %d)", k, r, c);
        }
        printf("\n");
      }
  }
  return 0;
}
