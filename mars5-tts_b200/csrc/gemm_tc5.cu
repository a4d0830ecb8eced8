// Persistent warp-specialised tcgen05 GEMM for sm_100a:   C[M,N] = epilogue( A[M,K] * W[N,K]^T )
//
//   * A (activations) and W (weights, nn.Linear layout [out,in]) are fp16, K-major; accumulation is fp32 in TMEM.
//   * warp 0      : TMA producer  (cp.async.bulk.tensor 2-D, 128B swizzle, mbarrier complete_tx)
//     warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16)
//     warps 2..5  : epilogue (tcgen05.ld 32x32b.x32 -> registers -> fused bias/activation/gate/residual -> global)
//   * 2 accumulator stages in TMEM (2*BLOCK_N columns) so the epilogue of tile i overlaps the main loop of tile i+1.
//   * Tile order keeps a band of 148 M-tiles (<= 39 MB of A at K=1024) L2-resident while sweeping N.
//
// This one kernel serves every dense contraction of the hot path whose M is large: the NAR encoder/decoder/speaker
// projections (reference: nn.MultiheadAttention in/out-proj, FNNSwiGLU nn_future.py:13-29, linear2, the 8 output heads
// model.py:234-240), the AR prefill projections (nn_future.py:241,274,297-298,398) and the Vocos pointwise convs.
// "split" mode (kwrap > 0): A holds [hi | lo] fp16 halves of an fp32 activation (K = 2*kwrap) and the W tile is
// re-read for the second half, giving fp32-class accuracy at 2x the tensor work.
#include <cuda.h>
#include <stdio.h>

#include "m5_internal.h"
#include "ptx.cuh"

namespace m5 {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;  // 64 fp16 = one 128-byte swizzle row
static constexpr int UMMA_K = 16;
static constexpr int GEMM_THREADS = 192;
static constexpr int M_BAND = 148;  // M-tiles kept L2-resident per sweep over N

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : ((BLOCK_N == 128) ? 6 : 8);
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;  // barriers + alignment slack
};

struct GemmEpi {
  const float* bias;      // [N] fp32 or null
  const float* colscale;  // [N] fp32 or null (Vocos layer-scale gamma)
  void* out;              // fp32 or fp16, row stride ldc (elements)
  void* out_lo;           // M5_OUT_F16_SPLIT: low halves
  int ldc;
  int mode;        // M5_OUT_*
  int act;         // M5_ACT_*
  int accumulate;  // fp32 out: out += value (residual stream update)
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == M5_ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  if (act == M5_ACT_SILU) return v / (1.0f + __expf(-v));
  return v;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M, int N,
                int K, int kwrap, int awrap, GemmEpi epi) {
  using S = GemmSmem<BLOCK_N>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = K / BLOCK_K;
  constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : 2 * BLOCK_N;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tc5_alloc(tmem_slot, TMEM_COLS);
  tc5_fence_before();
  __syncthreads();
  tc5_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_coords = [&](int t, int& m_blk, int& n_blk) {
    const int band_tiles = M_BAND * n_tiles;
    const int band = t / band_tiles;
    const int r = t - band * band_tiles;
    const int band_m = min(M_BAND, m_tiles - band * M_BAND);
    n_blk = r / band_m;
    m_blk = band * M_BAND + (r - n_blk * band_m);
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(t, m_blk, n_blk);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          const int k0 = kb * BLOCK_K;
          tma_load_2d(sa, &tmap_a, &full_bar[stage], awrap > 0 ? (k0 % awrap) : k0, m_blk * BLOCK_M);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kwrap > 0 ? (k0 % kwrap) : k0, n_blk * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc5_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc5_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + S::A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // +32 bytes along K inside the 128B swizzle row == +2 in the 16-byte start-address field
            tc5_mma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          tc5_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc5_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..5
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      int m_blk, n_blk;
      tile_coords(t, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc5_fence_after();
      const int row = m_blk * BLOCK_M + quad * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t taddr0 = tmem_base + acc * BLOCK_N + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        const int col0 = n_blk * BLOCK_N + c * 32;
        if (col0 >= N) break;  // warp-uniform
        __syncwarp();          // reconverge before the warp-collective TMEM load
        uint32_t r[32];
        tc5_ld_32x32(taddr0 + c * 32, r);
        tc5_wait_ld();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(r[j]);
          const int col = col0 + j;
          if (col < N) {
            if (epi.bias) x += __ldg(epi.bias + col);
            x = apply_act(x, epi.act);
            if (epi.colscale) x *= __ldg(epi.colscale + col);
          }
          v[j] = x;
        }
        if (!row_ok) continue;
        const bool full = (col0 + 32 <= N);
        if (epi.mode == M5_OUT_F32) {
          float* o = reinterpret_cast<float*>(epi.out) + (size_t)row * epi.ldc + col0;
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 w = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              if (epi.accumulate) {
                const float4 p = *reinterpret_cast<const float4*>(o + j);
                w.x += p.x; w.y += p.y; w.z += p.z; w.w += p.w;
              }
              *reinterpret_cast<float4*>(o + j) = w;
            }
          } else {
            for (int j = 0; j < 32 && col0 + j < N; ++j) o[j] = epi.accumulate ? o[j] + v[j] : v[j];
          }
        } else if (epi.mode == M5_OUT_F16) {
          __half* o = reinterpret_cast<__half*>(epi.out) + (size_t)row * epi.ldc + col0;
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 w;
              w.x = pack_half2(v[j], v[j + 1]);
              w.y = pack_half2(v[j + 2], v[j + 3]);
              w.z = pack_half2(v[j + 4], v[j + 5]);
              w.w = pack_half2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = w;
            }
          } else {
            for (int j = 0; j < 32 && col0 + j < N; ++j) o[j] = __float2half_rn(v[j]);
          }
        } else if (epi.mode == M5_OUT_F16_SPLIT) {
          __half* o = reinterpret_cast<__half*>(epi.out) + (size_t)row * epi.ldc + col0;
          __half* ol = reinterpret_cast<__half*>(epi.out_lo) + (size_t)row * epi.ldc + col0;
          for (int j = 0; j < 32 && col0 + j < N; ++j) {
            const __half h = __float2half_rn(v[j]);
            o[j] = h;
            ol[j] = __float2half_rn(v[j] - __half2float(h));
          }
        } else {  // M5_OUT_SWIGLU_F16 / _SPLIT: columns (2j, 2j+1) = (W_j x, V_j x) -> silu(Wx) * Vx, N/2 outputs
          const int ocol0 = col0 >> 1;
          __half* o = reinterpret_cast<__half*>(epi.out) + (size_t)row * epi.ldc + ocol0;
          float g[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = v[2 * j];
            g[j] = (a / (1.0f + __expf(-a))) * v[2 * j + 1];
          }
          if (epi.mode == M5_OUT_SWIGLU_F16) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 16; j += 8) {
                uint4 w;
                w.x = pack_half2(g[j], g[j + 1]);
                w.y = pack_half2(g[j + 2], g[j + 3]);
                w.z = pack_half2(g[j + 4], g[j + 5]);
                w.w = pack_half2(g[j + 6], g[j + 7]);
                *reinterpret_cast<uint4*>(o + j) = w;
              }
            } else {
              for (int j = 0; j < 16 && col0 + 2 * j + 1 < N; ++j) o[j] = __float2half_rn(g[j]);
            }
          } else {
            __half* ol = reinterpret_cast<__half*>(epi.out_lo) + (size_t)row * epi.ldc + ocol0;
            for (int j = 0; j < 16 && col0 + 2 * j + 1 < N; ++j) {
              const __half h = __float2half_rn(g[j]);
              o[j] = h;
              ol[j] = __float2half_rn(g[j] - __half2float(h));
            }
          }
        }
      }
      tc5_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc5_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc5_fence_after();
    tc5_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp16 row-major [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128B swizzle.
static int make_tmap(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return M5_ERR_CUDA;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? M5_OK : M5_ERR_CUDA;
}

template <int BLOCK_N>
static int launch_bn(const GemmCall& g, cudaStream_t stream, int num_sms) {
  using S = GemmSmem<BLOCK_N>;
  CUtensorMap ta, tb;
  const int Ka = g.awrap > 0 ? g.awrap : g.K;   // A's stored K extent
  const int Kb = g.kwrap > 0 ? g.kwrap : g.K;  // W's K extent
  if (make_tmap(&ta, g.A, g.M, Ka, g.lda, BLOCK_M) != M5_OK) return M5_ERR_CUDA;
  if (make_tmap(&tb, g.W, g.N, Kb, g.ldw, BLOCK_N) != M5_OK) return M5_ERR_CUDA;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_tc5_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) !=
        cudaSuccess)
      return M5_ERR_CUDA;
    attr_set = true;
  }
  const int m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (g.N + BLOCK_N - 1) / BLOCK_N;
  const int grid = min(num_sms, m_tiles * n_tiles);
  GemmEpi e;
  e.bias = g.bias; e.colscale = g.colscale; e.out = g.out; e.out_lo = g.out_lo; e.ldc = g.ldc;
  e.mode = g.mode; e.act = g.act; e.accumulate = g.accumulate;
  gemm_tc5_kernel<BLOCK_N><<<grid, GEMM_THREADS, S::TOTAL, stream>>>(ta, tb, g.M, g.N, g.K, g.kwrap, g.awrap, e);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

int gemm_tc5(const GemmCall& g, cudaStream_t stream, int num_sms) {
  if (g.M <= 0 || g.N <= 0) return M5_OK;
  if (g.K % BLOCK_K != 0 || g.lda % 8 != 0 || g.ldw % 8 != 0) return M5_ERR_ARG;
  if ((g.mode == M5_OUT_F32 && g.ldc % 4 != 0) || (g.mode != M5_OUT_F32 && g.ldc % 8 != 0)) return M5_ERR_ARG;
  // Pick the N tile: 256 when it does not waste much, else 128 / 64.
  const int m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
  auto waste = [&](int bn) { return (double)(((g.N + bn - 1) / bn) * bn) / g.N; };
  int bn = 256;
  if (g.N <= 64) bn = 64;
  else if (g.N <= 128) bn = 128;
  else if (waste(256) > 1.12 * waste(128) || (long)m_tiles * ((g.N + 255) / 256) < num_sms) bn = 128;
  if (g.force_bn) bn = g.force_bn;
  if (bn == 256) return launch_bn<256>(g, stream, num_sms);
  if (bn == 128) return launch_bn<128>(g, stream, num_sms);
  return launch_bn<64>(g, stream, num_sms);
}

}  // namespace m5
