// Weight-streaming GEMM for the AR decode step:  out[b, n] = sum_k X[b, k] * W[n, k],  b <= 32 rows in flight.
//
// HBM-bound by construction: every fp16 weight is read exactly once per step with 128-bit
// ld.global.nc.L1::no_allocate loads straight into mma.sync A fragments (no smem staging for W).  The K index inside
// one m16n8k16 MMA is a free permutation as long as A and B agree, so lane (g,t) takes the 8 contiguous halves
// W[n0+g, k0+8t .. k0+8t+7] (one 16-byte load) and feeds two MMAs with them.
//
// Decomposition: a CTA owns 128 weight rows x one K slice (split-K); each of its 8 warps owns 16 rows and walks the
// slice in 32-column chunks with 6 chunks (12 x 16 B per lane, 48 KB per CTA) in flight.  Only the K slice of the
// activations (<= 32 x kslice fp16, 64-byte row skew -> conflict-free 16-byte reads) is staged in shared memory, and
// the first weight loads are issued before that staging is waited for.  Partial sums go to an L2-resident scratch tile;
// the last CTA to finish a row tile (atomic ticket) adds the slices in a fixed order (deterministic) and applies the
// epilogue: plain store, residual accumulate, or SwiGLU over interleaved (W_j, V_j) rows.
//
// Replaces, at M = batch rows: wq/wk/wv, wo, w1/w3, w2 and the vocabulary projection of nn_future.py:241,274,297-298,398
// during KV-cached decoding.
#include "m5_internal.h"
#include "ptx.cuh"

namespace m5 {

static constexpr int SK_THREADS = 256;
static constexpr int SK_WARPS = 8;
static constexpr int SK_ROWS = 128;   // weight rows per CTA (16 per warp)
static constexpr int SK_UNROLL = 6;   // chunks of 32 columns in flight per warp

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <int NT>  // number of 8-row batch tiles (B <= 8*NT)
__global__ void __launch_bounds__(SK_THREADS)
gemm_skinny_kernel(SkinnyCall p, int kslice, int ksplit) {
  extern __shared__ __align__(16) uint8_t sk_smem[];
  __shared__ int s_ticket;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int tile = blockIdx.x / ksplit, ks = blockIdx.x - tile * ksplit;
  const int n0 = tile * SK_ROWS + warp * 16;
  const int kbase = ks * kslice;
  const int xstride = kslice * 2 + 64;  // bytes; kslice % 64 == 0 -> stride == 64 (mod 128)
  constexpr int BT = 8 * NT;

  pdl_launch_dependents();
  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const int r0 = min(n0 + g, p.N - 1), r1 = min(n0 + g + 8, p.N - 1);
  const __half* w0 = p.W + (size_t)r0 * p.K + kbase + 8 * t;
  const __half* w1 = p.W + (size_t)r1 * p.K + kbase + 8 * t;
  const int chunks = kslice / 32;

  // The weights do not depend on the previous kernel: the first batch is requested BEFORE the grid dependency is
  // waited for, i.e. while the producer of X may still be running (programmatic dependent launch).
  uint4 wa[SK_UNROLL], wb[SK_UNROLL];
#pragma unroll
  for (int u = 0; u < SK_UNROLL; ++u) {
    if (u < chunks) {
      wa[u] = ldg_stream(w0 + u * 32);
      wb[u] = ldg_stream(w1 + u * 32);
    }
  }
  pdl_wait();
  // stage X[:, kbase : kbase + kslice] (rows >= B are zero-filled)
  const int vec_per_row = kslice / 8;
  for (int i = tid; i < BT * vec_per_row; i += SK_THREADS) {
    const int row = i / vec_per_row, v = i - row * vec_per_row;
    const bool ok = row < p.B;
    cp_async16(sk_smem + (size_t)row * xstride + v * 16, ok ? (p.X + (size_t)row * p.K + kbase + v * 8) : p.X, ok);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  for (int c0 = 0; c0 < chunks; c0 += SK_UNROLL) {
    if (c0 > 0) {
#pragma unroll
      for (int u = 0; u < SK_UNROLL; ++u) {
        if (c0 + u < chunks) {
          wa[u] = ldg_stream(w0 + (c0 + u) * 32);
          wb[u] = ldg_stream(w1 + (c0 + u) * 32);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < SK_UNROLL; ++u) {
      const int c = c0 + u;
      if (c < chunks) {
        const uint32_t a1[4] = {wa[u].x, wb[u].x, wa[u].y, wb[u].y};
        const uint32_t a2[4] = {wa[u].z, wb[u].z, wa[u].w, wb[u].w};
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const uint4 x = *reinterpret_cast<const uint4*>(sk_smem + (size_t)(nt * 8 + g) * xstride + (c * 32 + 8 * t) * 2);
          mma_16816(acc[nt], a1, x.x, x.y);
          mma_16816(acc[nt], a2, x.z, x.w);
        }
      }
    }
  }
  // ---- partial tile -> scratch[tile][ks][b][row]
  float* part = p.scratch + ((size_t)(tile * ksplit + ks)) * (BT * SK_ROWS);
  const int rl = warp * 16 + g;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int b0 = nt * 8 + 2 * t;
    part[(size_t)b0 * SK_ROWS + rl] = acc[nt][0];
    part[(size_t)(b0 + 1) * SK_ROWS + rl] = acc[nt][1];
    part[(size_t)b0 * SK_ROWS + rl + 8] = acc[nt][2];
    part[(size_t)(b0 + 1) * SK_ROWS + rl + 8] = acc[nt][3];
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(p.counters + tile, 1);
  __syncthreads();
  if (s_ticket != ksplit - 1) return;
  // ---- last CTA of this row tile: ordered reduction over the K slices + epilogue
  __threadfence();
  if (tid == 0) p.counters[tile] = 0;  // ready for the next launch / graph replay
  const float* base = p.scratch + (size_t)tile * ksplit * (BT * SK_ROWS);
  const int nrow0 = tile * SK_ROWS;
  if (!p.swiglu) {
    for (int i = tid; i < BT * SK_ROWS; i += SK_THREADS) {
      const int b = i / SK_ROWS, row = i - b * SK_ROWS;
      const int n = nrow0 + row;
      if (b >= p.B || n >= p.N) continue;
      float v = 0.f;
      for (int s = 0; s < ksplit; ++s) v += __ldcg(base + (size_t)s * (BT * SK_ROWS) + i);
      float* o = p.out_f32 + (size_t)b * p.ldc + n;
      *o = p.accumulate ? (*o + v) : v;
    }
  } else {
    for (int i = tid; i < BT * (SK_ROWS / 2); i += SK_THREADS) {
      const int b = i / (SK_ROWS / 2), pr = i - b * (SK_ROWS / 2);
      const int n = nrow0 + 2 * pr;
      if (b >= p.B || n + 1 >= p.N) continue;
      float a = 0.f, c = 0.f;
      for (int s = 0; s < ksplit; ++s) {
        const float2 v2 = __ldcg(reinterpret_cast<const float2*>(base + (size_t)s * (BT * SK_ROWS) + (size_t)b * SK_ROWS + 2 * pr));
        a += v2.x;
        c += v2.y;
      }
      p.out_f16[(size_t)b * p.ldc + (n >> 1)] = __float2half_rn((a / (1.f + __expf(-a))) * c);
    }
  }
}

// K slice per CTA: a divisor of K/64 so that (row tiles x slices) fills about two CTAs per SM.
static void pick_split(int N, int K, int num_sms, int& kslice, int& ksplit) {
  const int tiles = (N + SK_ROWS - 1) / SK_ROWS;
  const int kb = K / 64;
  int best = 1;
  for (int s = 1; s <= kb; ++s) {
    if (kb % s) continue;
    // fill about two CTAs per SM, but keep slices >= 256 columns and at most 8 of them (the last CTA of a row tile
    // adds the slices serially)
    if ((long)tiles * s <= 2L * num_sms && s <= 8 && (kb / s) * 64 >= 256) best = s;
  }
  while ((K / best) > 2048 && best < kb) {  // shared-memory bound on the activation slice
    ++best;
    while (kb % best) ++best;
  }
  ksplit = best;
  kslice = K / best;
}

size_t gemm_skinny_scratch_bytes(int num_sms) { return (size_t)(2 * num_sms + 64) * 32 * SK_ROWS * sizeof(float); }

int gemm_skinny(const SkinnyCall& c, cudaStream_t stream, int num_sms) {
  if (c.B <= 0 || c.N <= 0) return M5_OK;
  if (c.B > 32 || c.K % 64 != 0 || !c.scratch || !c.counters) return M5_ERR_ARG;
  int kslice, ksplit;
  pick_split(c.N, c.K, num_sms, kslice, ksplit);
  const int tiles = (c.N + SK_ROWS - 1) / SK_ROWS;
  if (tiles > 1024 || (size_t)tiles * ksplit > (size_t)(2 * num_sms + 64)) return M5_ERR_ARG;
  const int NT = c.B <= 8 ? 1 : (c.B <= 16 ? 2 : 4);
  const size_t smem = (size_t)(8 * NT) * (kslice * 2 + 64);
  static DeviceOnce once;
  unsigned long long bit;
  if (once.needed(bit)) {  // opt in to > 48 KB dynamic shared memory once per device for every instantiation
    cudaFuncSetAttribute(gemm_skinny_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(gemm_skinny_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(gemm_skinny_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    once.done(bit);
  }
  const int grid = tiles * ksplit;
  cudaError_t e;
  if (NT == 1) e = launch_k(gemm_skinny_kernel<1>, dim3(grid), dim3(SK_THREADS), smem, stream, c, kslice, ksplit);
  else if (NT == 2) e = launch_k(gemm_skinny_kernel<2>, dim3(grid), dim3(SK_THREADS), smem, stream, c, kslice, ksplit);
  else e = launch_k(gemm_skinny_kernel<4>, dim3(grid), dim3(SK_THREADS), smem, stream, c, kslice, ksplit);
  return e == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

}  // namespace m5
