// Weight-streaming GEMM for the AR decode step:  out[b, n] = sum_k X[b, k] * W[n, k],  b < 32 rows in flight.
//
// HBM-bound by construction: every fp16 weight is read exactly once per step with 128-bit
// ld.global.nc.L1::no_allocate loads straight into mma.sync A fragments (no smem staging for W).  The K index inside
// one m16n8k16 MMA is a free permutation as long as A and B agree, so lane (g,t) takes the 8 contiguous halves
// W[n0+g, k0+8t .. k0+8t+7] (one 16-byte load) and feeds two MMAs with them; the activation tile (<= 32 x K fp16)
// lives in shared memory with a 64-byte row skew (conflict-free 16-byte reads).  A CTA owns 16 weight rows, its
// 8 warps split K in interleaved 32-column chunks and reduce through smem in a fixed order (deterministic).
//
// Replaces, at M = batch rows: wq/wk/wv, wo, w1/w3 (SwiGLU fused via interleaved rows), w2 and the vocabulary
// projection of nn_future.py:241,274,297-298,398 during KV-cached decoding.
#include "m5_internal.h"
#include "ptx.cuh"

namespace m5 {

static constexpr int SK_THREADS = 256;
static constexpr int SK_WARPS = 8;
static constexpr int SK_ROWS = 16;
static constexpr int SK_UNROLL = 4;

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <int NT>  // number of 8-row batch tiles (B <= 8*NT)
__global__ void __launch_bounds__(SK_THREADS)
gemm_skinny_kernel(SkinnyCall p, int kblk, int n_kblk) {
  extern __shared__ __align__(16) uint8_t sk_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * SK_ROWS;
  const int xstride = kblk * 2 + 64;  // bytes; == 64 (mod 128) because kblk % 64 == 0 ... see host check
  uint8_t* sx = sk_smem;
  float* sred = reinterpret_cast<float*>(sk_smem + (size_t)(8 * NT) * xstride);

  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  const int r0 = min(n0 + g, p.N - 1), r1 = min(n0 + g + 8, p.N - 1);
  const __half* w0 = p.W + (size_t)r0 * p.K + 8 * t;
  const __half* w1 = p.W + (size_t)r1 * p.K + 8 * t;
  const int chunks = kblk / 32;

  for (int kb = 0; kb < n_kblk; ++kb) {
    const int kbase = kb * kblk;
    if (kb > 0) __syncthreads();
    // stage X[:, kbase : kbase+kblk] (rows >= B are zero-filled)
    const int vec_per_row = kblk / 8;
    for (int i = tid; i < 8 * NT * vec_per_row; i += SK_THREADS) {
      const int row = i / vec_per_row, v = i - row * vec_per_row;
      const bool ok = row < p.B;
      cp_async16(sx + (size_t)row * xstride + v * 16, ok ? (p.X + (size_t)row * p.K + kbase + v * 8) : p.X, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    for (int c0 = warp; c0 < chunks; c0 += SK_WARPS * SK_UNROLL) {
      uint4 wa[SK_UNROLL], wb[SK_UNROLL];
#pragma unroll
      for (int u = 0; u < SK_UNROLL; ++u) {
        const int c = c0 + u * SK_WARPS;
        if (c < chunks) {
          wa[u] = ldg_stream(w0 + kbase + c * 32);
          wb[u] = ldg_stream(w1 + kbase + c * 32);
        }
      }
#pragma unroll
      for (int u = 0; u < SK_UNROLL; ++u) {
        const int c = c0 + u * SK_WARPS;
        if (c < chunks) {
          const uint32_t a1[4] = {wa[u].x, wb[u].x, wa[u].y, wb[u].y};
          const uint32_t a2[4] = {wa[u].z, wb[u].z, wa[u].w, wb[u].w};
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const uint4 x = *reinterpret_cast<const uint4*>(sx + (size_t)(nt * 8 + g) * xstride + (c * 32 + 8 * t) * 2);
            mma_16816(acc[nt], a1, x.x, x.y);
            mma_16816(acc[nt], a2, x.z, x.w);
          }
        }
      }
    }
  }
  // ---- cross-warp reduction: sred[warp][row 0..15][batch 0..8NT)
  constexpr int BT = 8 * NT;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    float* s = sred + (size_t)warp * SK_ROWS * BT;
    s[g * BT + nt * 8 + 2 * t] = acc[nt][0];
    s[g * BT + nt * 8 + 2 * t + 1] = acc[nt][1];
    s[(g + 8) * BT + nt * 8 + 2 * t] = acc[nt][2];
    s[(g + 8) * BT + nt * 8 + 2 * t + 1] = acc[nt][3];
  }
  __syncthreads();
  if (!p.swiglu) {
    for (int i = tid; i < SK_ROWS * BT; i += SK_THREADS) {
      const int b = i / SK_ROWS, row = i - b * SK_ROWS;  // consecutive threads -> consecutive output features
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < SK_WARPS; ++w) v += sred[(size_t)w * SK_ROWS * BT + row * BT + b];
      const int n = n0 + row;
      if (b < p.B && n < p.N) {
        float* o = p.out_f32 + (size_t)b * p.ldc + n;
        *o = p.accumulate ? (*o + v) : v;
      }
    }
  } else {
    for (int i = tid; i < (SK_ROWS / 2) * BT; i += SK_THREADS) {
      const int b = i / (SK_ROWS / 2), pr = i - b * (SK_ROWS / 2);
      float a = 0.f, c = 0.f;
#pragma unroll
      for (int w = 0; w < SK_WARPS; ++w) {
        a += sred[(size_t)w * SK_ROWS * BT + (2 * pr) * BT + b];
        c += sred[(size_t)w * SK_ROWS * BT + (2 * pr + 1) * BT + b];
      }
      const int n = n0 + 2 * pr;
      if (b < p.B && n + 1 < p.N) p.out_f16[(size_t)b * p.ldc + (n >> 1)] = __float2half_rn((a / (1.f + __expf(-a))) * c);
    }
  }
}

int gemm_skinny(const SkinnyCall& c, cudaStream_t stream, int num_sms) {
  if (c.B <= 0 || c.N <= 0) return M5_OK;
  if (c.B > 32 || c.K % 32 != 0) return M5_ERR_ARG;
  // split K into blocks that fit shared memory; every block must be a multiple of 64 columns so that the X row
  // stride (2*kblk + 64 bytes) is 64 mod 128.
  int n_kblk = 1;
  while ((c.K / n_kblk) > 2048 || c.K % n_kblk != 0 || (c.K / n_kblk) % 64 != 0) {
    ++n_kblk;
    if (n_kblk > 64) return M5_ERR_ARG;
  }
  const int kblk = c.K / n_kblk;
  const int NT = c.B <= 8 ? 1 : (c.B <= 16 ? 2 : 4);
  const size_t smem = (size_t)(8 * NT) * (kblk * 2 + 64) + (size_t)SK_WARPS * SK_ROWS * 8 * NT * sizeof(float);
  const int grid = (c.N + SK_ROWS - 1) / SK_ROWS;
  static bool attr_set = false;
  if (!attr_set) {  // opt in to > 48 KB dynamic shared memory once for every instantiation
    cudaFuncSetAttribute(gemm_skinny_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gemm_skinny_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gemm_skinny_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  auto launch = [&](auto kern) { kern<<<grid, SK_THREADS, smem, stream>>>(c, kblk, n_kblk); };
  if (NT == 1) launch(gemm_skinny_kernel<1>);
  else if (NT == 2) launch(gemm_skinny_kernel<2>);
  else launch(gemm_skinny_kernel<4>);
  (void)num_sms;
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

}  // namespace m5
