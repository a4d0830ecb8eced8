// Persistent warp-specialised tcgen05 GEMM for sm_100a:   C[M,N] = epilogue( A[M,K] * W[N,K]^T )
//
//   * A (activations) and W (weights, nn.Linear layout [out,in]) are fp16, K-major; accumulation is fp32 in TMEM.
//   * warp 0      : TMA producer  (cp.async.bulk.tensor 2-D, 128B swizzle, mbarrier complete_tx)
//     warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16)
//     warps 2..5  : epilogue: tcgen05.ld 32x32b.x32 (lane = accumulator row) -> padded smem tile -> registers with
//                   lane = 4 consecutive columns, so every global access is a whole 128-byte row segment; bias / column
//                   scale of the tile are staged in smem while the MMAs of the tile are still running; the TMEM load of
//                   chunk c+1 is in flight while chunk c is stored; residual rows are prefetched before the adds.
//   * 2 accumulator stages in TMEM (2*BLOCK_N columns) so the epilogue of tile i overlaps the main loop of tile i+1.
//   * Tile order keeps a band of 148 M-tiles (<= 39 MB of A at K=1024) L2-resident while sweeping N.
//   * The epilogue is compiled per output kind (template) to keep the instruction footprint inside the I-cache.
//
// This one kernel serves every dense contraction of the hot path whose M is large: the NAR encoder/decoder/speaker
// projections (reference: nn.MultiheadAttention in/out-proj, FNNSwiGLU nn_future.py:13-29, linear2, the 8 output heads
// model.py:234-240), the AR prefill projections (nn_future.py:241,274,297-298,398) and the Vocos pointwise convs.
// "split" operands: kwrap > 0 -> A = [hi | lo] fp16 halves of an fp32 activation (K = 2*kwrap), W re-read modulo kwrap;
// awrap > 0 -> A re-read modulo awrap against W = [W_hi | W_hi | W_lo]: fp32-class accuracy at 2-3x the tensor work.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "gemm_common.cuh"

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
  static constexpr int EPI_OFF = BAR_OFF + 256;             // 4 epilogue warps x [32 rows][36 floats] transpose tiles
  static constexpr int EPI_BYTES = 4 * 32 * 36 * 4;
  static constexpr int BIAS_OFF = EPI_OFF + EPI_BYTES;      // per epilogue warp: bias[BLOCK_N] | colscale[BLOCK_N]
  static constexpr int BIAS_BYTES = 4 * 2 * BLOCK_N * 4;
  static constexpr int TOTAL = BIAS_OFF + BIAS_BYTES + 1024;  // + alignment slack
};

template <int BLOCK_N, int KIND>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M, int N,
                int K, int kwrap, int awrap, GemmEpi epi) {
  using S = GemmSmem<BLOCK_N>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzle tiles; pointer arithmetic on the __shared__ array keeps the address space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
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
    float* stg = reinterpret_cast<float*>(smem + S::EPI_OFF) + (warp - 2) * (32 * 36);
    float* sbias = reinterpret_cast<float*>(smem + S::BIAS_OFF) + (warp - 2) * (2 * BLOCK_N);
    float* sscale = sbias + BLOCK_N;
    const int c4 = (lane & 7) * 4, rsub = lane >> 3;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      int m_blk, n_blk;
      tile_coords(t, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      // stage this tile's bias / column scale while the MMAs of the tile are still running
      __syncwarp();
      for (int j = lane; j < BLOCK_N; j += 32) {
        const int col = n_blk * BLOCK_N + j;
        sbias[j] = (epi.bias && col < N) ? __ldg(epi.bias + col) : 0.f;
        if constexpr (KIND == E_F32_ACC || KIND == E_GENERIC)
          sscale[j] = (epi.colscale && col < N) ? __ldg(epi.colscale + col) : 1.f;
      }
      __syncwarp();
      mbar_wait(&tmem_full[acc], acc_phase);
      tc5_fence_after();
      const int row_base = m_blk * BLOCK_M + quad * 32;
      const uint32_t taddr0 = tmem_base + acc * BLOCK_N + ((uint32_t)(quad * 32) << 16);
      const int n_chunks = min(BLOCK_N / 32, (N - n_blk * BLOCK_N + 31) / 32);
      uint32_t r[32];
      tc5_ld_32x32(taddr0, r);
      float4 prev_next[8];
      if constexpr (KIND == E_F32_ACC) {
        const int col_first = n_blk * BLOCK_N + c4;
        if (col_first + 3 < N) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = row_base + rsub + 4 * i;
            prev_next[i] = row < M ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(epi.out) + (size_t)row * epi.ldc + col_first)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        const int col = n_blk * BLOCK_N + c * 32 + c4;  // first of this lane's 4 columns
        const bool full4 = col + 3 < N;
        tc5_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stg + lane * 36 + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        __syncwarp();
        if (c + 1 < n_chunks) {
          tc5_ld_32x32(taddr0 + (c + 1) * 32, r);  // in flight while this chunk is stored
        } else {
          // accumulator fully read out (tcgen05.wait::ld above): release it before the last chunk's stores
          tc5_fence_before();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        const float4 bb = *reinterpret_cast<const float4*>(sbias + c * 32 + c4);
        float s4[4] = {1.f, 1.f, 1.f, 1.f};
        if constexpr (KIND == E_F32_ACC || KIND == E_GENERIC) {
          const float4 sc = *reinterpret_cast<const float4*>(sscale + c * 32 + c4);
          s4[0] = sc.x; s4[1] = sc.y; s4[2] = sc.z; s4[3] = sc.w;
        }
        float4 prev[8];
        if constexpr (KIND == E_F32_ACC) {
          // residual rows of THIS chunk were requested one iteration ago (prev_next); request the next chunk's now so
          // that the HBM/L2 latency overlaps this chunk's arithmetic and stores
#pragma unroll
          for (int i = 0; i < 8; ++i) prev[i] = prev_next[i];
          const int coln = col + 32;
          if (c + 1 < n_chunks && coln + 3 < N) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = row_base + rsub + 4 * i;
              prev_next[i] = row < M ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(epi.out) + (size_t)row * epi.ldc + coln)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = rsub + 4 * i;
          const int row = row_base + rl;
          const float4 a4 = *reinterpret_cast<const float4*>(stg + rl * 36 + c4);
          float v[4] = {a4.x + bb.x, a4.y + bb.y, a4.z + bb.z, a4.w + bb.w};
          if (row < M && col < N) epi_store4<KIND>(epi, v, s4, prev[i], row, col, N, full4);
        }
        __syncwarp();  // every lane is done with the staging tile before the next chunk overwrites it
      }
    }
  }

  tc5_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc5_fence_after();
    tc5_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BLOCK_N, int KIND>
static int launch_kind(const GemmCall& g, cudaStream_t stream, int num_sms) {
  using S = GemmSmem<BLOCK_N>;
  CUtensorMap ta, tb;
  const int Ka = g.awrap > 0 ? g.awrap : g.K;  // A's stored K extent
  const int Kb = g.kwrap > 0 ? g.kwrap : g.K;  // W's stored K extent
  if (make_tmap_k64(&ta, g.A, g.M, Ka, g.lda, BLOCK_M) != M5_OK) return M5_ERR_CUDA;
  if (make_tmap_k64(&tb, g.W, g.N, Kb, g.ldw, BLOCK_N) != M5_OK) return M5_ERR_CUDA;
  static DeviceOnce once;   // one per template instantiation
  unsigned long long bit;
  if (once.needed(bit)) {
    if (cudaFuncSetAttribute(gemm_tc5_kernel<BLOCK_N, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
      return M5_ERR_CUDA;
    once.done(bit);
  }
  const int m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (g.N + BLOCK_N - 1) / BLOCK_N;
  const int grid = min(num_sms, m_tiles * n_tiles);
  GemmEpi e;
  e.bias = g.bias; e.colscale = g.colscale; e.out = g.out; e.out_lo = g.out_lo; e.lo_from_col = g.lo_from_col; e.ldc = g.ldc;
  e.out_lo8 = g.out_lo8; e.ldc8 = g.ldc8;
  e.mode = g.mode; e.act = g.act; e.accumulate = g.accumulate;
  gemm_tc5_kernel<BLOCK_N, KIND><<<grid, GEMM_THREADS, S::TOTAL, stream>>>(ta, tb, g.M, g.N, g.K, g.kwrap, g.awrap, e);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

template <int BLOCK_N>
static int launch_bn(const GemmCall& g, cudaStream_t stream, int num_sms) {
  switch (gemm_epi_kind(g)) {
    case E_F32: return launch_kind<BLOCK_N, E_F32>(g, stream, num_sms);
    case E_F32_ACC: return launch_kind<BLOCK_N, E_F32_ACC>(g, stream, num_sms);
    case E_F16: return launch_kind<BLOCK_N, E_F16>(g, stream, num_sms);
    case E_SWIGLU: return launch_kind<BLOCK_N, E_SWIGLU>(g, stream, num_sms);
    case E_F16_SPLIT: return launch_kind<BLOCK_N, E_F16_SPLIT>(g, stream, num_sms);
    case E_SWIGLU_SPLIT: return launch_kind<BLOCK_N, E_SWIGLU_SPLIT>(g, stream, num_sms);
    case E_SWIGLU_SPLIT8: return launch_kind<BLOCK_N, E_SWIGLU_SPLIT8>(g, stream, num_sms);
    default: return launch_kind<BLOCK_N, E_GENERIC>(g, stream, num_sms);
  }
}

static bool pair_shape_ok(int M, int N, int num_sms) {
  auto waste = [&](int bn) { return (double)(((N + bn - 1) / bn) * bn) / N; };
  static const bool no_pairs = getenv("M5_DISABLE_2CTA") != nullptr;
  return !no_pairs && N >= 256 && waste(256) <= 1.12 && (long)((M + 255) / 256) * ((N + 255) / 256) >= num_sms / 2;
}
bool gemm_f8lo_eligible(int M, int N, int num_sms) { return pair_shape_ok(M, N, num_sms); }

int gemm_tc5(const GemmCall& g, cudaStream_t stream, int num_sms) {
  if (g.M <= 0 || g.N <= 0) return M5_OK;
  if (g.K % BLOCK_K != 0 || g.lda % 8 != 0 || g.ldw % 8 != 0) return M5_ERR_ARG;
  if ((g.kwrap > 0 && g.kwrap % BLOCK_K != 0) || (g.awrap > 0 && g.awrap % BLOCK_K != 0)) return M5_ERR_ARG;
  if (g.mode == M5_OUT_F32 && g.ldc % 4 != 0) return M5_ERR_ARG;
  if ((g.mode == M5_OUT_F16 || g.mode == M5_OUT_F16_SPLIT) && g.ldc % 4 != 0) return M5_ERR_ARG;
  if ((g.mode == M5_OUT_SWIGLU_F16 || g.mode == M5_OUT_SWIGLU_F16_SPLIT) && (g.ldc % 2 != 0 || g.N % 2 != 0)) return M5_ERR_ARG;
  const int m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
  auto waste = [&](int bn) { return (double)(((g.N + bn - 1) / bn) * bn) / g.N; };
  // Large problems run on CTA pairs (cta_group::2, 256 x 256 tiles): halves the operand bytes each SM has to ingest.
  const bool pair_ok = pair_shape_ok(g.M, g.N, num_sms);
  if (g.A8 && !(pair_ok || g.force_bn == 512)) return M5_ERR_ARG;   // the fp8 lo pass exists in the CTA-pair kernel only
  if (g.force_bn == 512 || (g.force_bn == 0 && pair_ok)) return gemm_tc5_2cta(g, stream, num_sms);
  // Pick the N tile: 256 when it does not waste much, else 128 / 64.
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
