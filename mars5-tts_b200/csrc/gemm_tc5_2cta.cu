// CTA-pair (cta_group::2) variant of the tcgen05 GEMM for the large NAR / prefill contractions.
//
// ncu on the 1-CTA kernel (profiles/r1_gemm_tc5_ncu.txt) shows the tensor pipe 56 % busy with L2 and DRAM far from
// saturated: a 128x256 tile needs 48 KB of operands per 64-wide K step per SM, more than one SM can ingest while the
// MMAs of that step run.  Here two CTAs of a cluster (same TPC) own one 256 x 256 tile: each loads its 128 rows of A and
// its 128-row HALF of the W tile (32 KB per K step per SM), rank 0 issues UMMA 256x256x16 with cta_group::2 (each SM
// multiplies its A half against both W halves), both CTAs drain their own 128 x 256 TMEM accumulator.  Barrier traffic:
//   full[s]      lives in the leader; both CTAs' TMA loads complete_tx on it (peer-bit-masked address)
//   empty[s]     one per CTA, released by a multicast tcgen05.commit from the leader
//   tmem_full[a] one per CTA, multicast commit;  tmem_empty[a] in the leader, 8 arrivals (4 epilogue warps x 2 CTAs)
#include "gemm_common.cuh"

namespace m5 {

static constexpr int C2_BM = 128;      // rows of A per CTA (tile M = 256 per pair)
static constexpr int C2_BN = 256;      // tile N (each CTA stores half of the W tile: 128 rows)
static constexpr int C2_BK = 64;
static constexpr int C2_STAGES = 6;
static constexpr int C2_THREADS = 192;
static constexpr int C2_A_BYTES = C2_BM * C2_BK * 2;        // 16 KB
static constexpr int C2_B_BYTES = (C2_BN / 2) * C2_BK * 2;  // 16 KB
static constexpr int C2_STAGE_BYTES = C2_A_BYTES + C2_B_BYTES;
static constexpr int C2_BAR_OFF = C2_STAGES * C2_STAGE_BYTES;            // 192 KB
static constexpr int C2_EPI_OFF = C2_BAR_OFF + 256;
static constexpr int C2_EPI_BYTES = 4 * 32 * 36 * 4;
static constexpr int C2_BIAS_OFF = C2_EPI_OFF + C2_EPI_BYTES;
static constexpr int C2_BIAS_BYTES = 4 * 2 * C2_BN * 4;
static constexpr int C2_TOTAL = C2_BIAS_OFF + C2_BIAS_BYTES + 1024;
static constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader's copy

M5_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
M5_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
M5_DEVINL void tc5_alloc2(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
M5_DEVINL void tc5_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load whose completion is signalled on the LEADER CTA's mbarrier (same smem offset, rank bit cleared)
M5_DEVINL void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// fp8 operands (A e5m2, B e4m3 per the instruction descriptor): 32 K-elements = 32 bytes per instruction
M5_DEVINL void tc5_mma_f8_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
M5_DEVINL void tc5_mma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this smem offset in BOTH CTAs of the pair once all prior MMAs have completed
M5_DEVINL void tc5_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
// arrive on the leader's copy of a barrier from either CTA
M5_DEVINL void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(smem_u32(bar))
      : "memory");
}

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C2_THREADS, 1)
gemm_tc5_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_a8, const __grid_constant__ CUtensorMap tmap_b8, int M, int N,
                     int K, int kwrap, int awrap, int K8, GemmEpi epi) {
  extern __shared__ uint8_t smem_raw2[];
  uint8_t* smem = smem_raw2 + ((1024u - (smem_u32(smem_raw2) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C2_BAR_OFF);
  uint64_t* empty_bar = full_bar + C2_STAGES;
  uint64_t* tmem_full = empty_bar + C2_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m_tiles = (M + 2 * C2_BM - 1) / (2 * C2_BM);
  const int n_tiles = (N + C2_BN - 1) / C2_BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = K / C2_BK;
  const int k_blocks8 = K8 / 128;   // fp8 "lo" blocks: 128 K-elements per 128-byte row, same 16 KB tiles
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr int M_BAND2 = 74;  // pairs: 74 x 256 rows of A stay L2-resident while N is swept

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (k_blocks8 > 0) { tma_prefetch_desc(&tmap_a8); tma_prefetch_desc(&tmap_b8); }
    for (int s = 0; s < C2_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);
    }
    fence_barrier_init();
  }
  cluster_sync_all();  // barrier inits visible cluster-wide before any remote arrive / multicast
  if (warp == 1) tc5_alloc2(tmem_slot, TMEM_COLS);
  tc5_fence_before();
  cluster_sync_all();
  tc5_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_coords = [&](int t, int& m_blk, int& n_blk) {
    const int band_tiles = M_BAND2 * n_tiles;
    const int band = t / band_tiles;
    const int r = t - band * band_tiles;
    const int band_m = min(M_BAND2, m_tiles - band * M_BAND2);
    n_blk = r / band_m;
    m_blk = band * M_BAND2 + (r - n_blk * band_m);
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += n_clusters) {
        int m_blk, n_blk;
        tile_coords(t, m_blk, n_blk);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C2_STAGE_BYTES;
          uint8_t* sb = sa + C2_A_BYTES;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * C2_STAGE_BYTES);  // both CTAs' bytes land on this barrier
          const int k0 = kb * C2_BK;
          tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], awrap > 0 ? (k0 % awrap) : k0, m_blk * 2 * C2_BM + rank * C2_BM);
          tma_load_2d_2sm(sb, &tmap_b, &full_bar[stage], kwrap > 0 ? (k0 % kwrap) : k0, n_blk * C2_BN + rank * (C2_BN / 2));
          if (++stage == C2_STAGES) { stage = 0; phase ^= 1; }
        }
        for (int kb = 0; kb < k_blocks8; ++kb) {   // lo halves: e5m2 activations x e4m3 weights, byte-addressed maps
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C2_STAGE_BYTES;
          uint8_t* sb = sa + C2_A_BYTES;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * C2_STAGE_BYTES);
          tma_load_2d_2sm(sa, &tmap_a8, &full_bar[stage], kb * 128, m_blk * 2 * C2_BM + rank * C2_BM);
          tma_load_2d_2sm(sb, &tmap_b8, &full_bar[stage], kb * 128, n_blk * C2_BN + rank * (C2_BN / 2));
          if (++stage == C2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * C2_BM, C2_BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = cluster_id; t < num_tiles; t += n_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc5_fence_after();
        const uint32_t tmem_d = tmem_base + acc * C2_BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc5_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C2_STAGE_BYTES);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + C2_A_BYTES);
#pragma unroll
          for (int k = 0; k < C2_BK / 16; ++k) tc5_mma_f16_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          tc5_commit_mc(&empty_bar[stage]);
          if (++stage == C2_STAGES) { stage = 0; phase ^= 1; }
        }
        constexpr uint32_t idesc8 = umma_idesc_e5m2_e4m3(2 * C2_BM, C2_BN);
        for (int kb = 0; kb < k_blocks8; ++kb) {   // same accumulator: the fp8 products are at true scale (2^-2 x 2^+2)
          mbar_wait(&full_bar[stage], phase);
          tc5_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C2_STAGE_BYTES);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + C2_A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc5_mma_f8_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc8, 1u);   // 32 B = 32 fp8 per UMMA
          tc5_commit_mc(&empty_bar[stage]);
          if (++stage == C2_STAGES) { stage = 0; phase ^= 1; }
        }
        tc5_commit_mc(&tmem_full[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..5 (both CTAs, own 128 rows)
    const int quad = warp & 3;
    float* stg = reinterpret_cast<float*>(smem + C2_EPI_OFF) + (warp - 2) * (32 * 36);
    float* sbias = reinterpret_cast<float*>(smem + C2_BIAS_OFF) + (warp - 2) * (2 * C2_BN);
    float* sscale = sbias + C2_BN;
    const int c4 = (lane & 7) * 4, rsub = lane >> 3;
    int it = 0;
    for (int t = cluster_id; t < num_tiles; t += n_clusters, ++it) {
      int m_blk, n_blk;
      tile_coords(t, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      __syncwarp();
      for (int j = lane; j < C2_BN; j += 32) {
        const int col = n_blk * C2_BN + j;
        sbias[j] = (epi.bias && col < N) ? __ldg(epi.bias + col) : 0.f;
        if constexpr (KIND == E_F32_ACC || KIND == E_GENERIC)
          sscale[j] = (epi.colscale && col < N) ? __ldg(epi.colscale + col) : 1.f;
      }
      __syncwarp();
      mbar_wait(&tmem_full[acc], acc_phase);
      tc5_fence_after();
      const int row_base = m_blk * 2 * C2_BM + rank * C2_BM + quad * 32;
      const uint32_t taddr0 = tmem_base + acc * C2_BN + ((uint32_t)(quad * 32) << 16);
      const int n_chunks = min(C2_BN / 32, (N - n_blk * C2_BN + 31) / 32);
      uint32_t r[32];
      tc5_ld_32x32(taddr0, r);
      float4 prev_next[8];
      if constexpr (KIND == E_F32_ACC) {
        const int col_first = n_blk * C2_BN + c4;
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
        const int col = n_blk * C2_BN + c * 32 + c4;
        const bool full4 = col + 3 < N;
        tc5_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stg + lane * 36 + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        __syncwarp();
        if (c + 1 < n_chunks) {
          tc5_ld_32x32(taddr0 + (c + 1) * 32, r);
        } else {
          // the accumulator has been read out completely (tcgen05.wait::ld above): hand it back to the MMA warp now,
          // the stores of this last chunk overlap the next tile's MMAs
          tc5_fence_before();
          if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
        }
        const float4 bb = *reinterpret_cast<const float4*>(sbias + c * 32 + c4);
        float s4[4] = {1.f, 1.f, 1.f, 1.f};
        if constexpr (KIND == E_F32_ACC || KIND == E_GENERIC) {
          const float4 sc = *reinterpret_cast<const float4*>(sscale + c * 32 + c4);
          s4[0] = sc.x; s4[1] = sc.y; s4[2] = sc.z; s4[3] = sc.w;
        }
        float4 prev[8];
        if constexpr (KIND == E_F32_ACC) {
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
        __syncwarp();
      }
    }
  }

  tc5_fence_before();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still signal / read
  if (warp == 1) {
    tc5_fence_after();
    tc5_dealloc2(tmem_base, TMEM_COLS);
  }
}

template <int KIND>
static int launch_2cta(const GemmCall& g, cudaStream_t stream, int num_sms) {
  CUtensorMap ta, tb, ta8, tb8;
  const int Ka = g.awrap > 0 ? g.awrap : g.K;
  const int Kb = g.kwrap > 0 ? g.kwrap : g.K;
  if (make_tmap_k64(&ta, g.A, g.M, Ka, g.lda, C2_BM) != M5_OK) return M5_ERR_CUDA;
  if (make_tmap_k64(&tb, g.W, g.N, Kb, g.ldw, C2_BN / 2) != M5_OK) return M5_ERR_CUDA;
  ta8 = ta; tb8 = tb;
  if (g.A8) {
    if (!g.W8 || g.K8 <= 0 || g.K8 % 128 != 0 || g.lda8 % 16 != 0 || g.ldw8 % 16 != 0) return M5_ERR_ARG;
    if (make_tmap_u8_k128(&ta8, g.A8, g.M, g.K8, g.lda8, C2_BM) != M5_OK) return M5_ERR_CUDA;
    if (make_tmap_u8_k128(&tb8, g.W8, g.N, g.K8, g.ldw8, C2_BN / 2) != M5_OK) return M5_ERR_CUDA;
  }
  static DeviceOnce once;   // one per template instantiation
  unsigned long long bit;
  if (once.needed(bit)) {
    if (cudaFuncSetAttribute(gemm_tc5_2cta_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_TOTAL) != cudaSuccess)
      return M5_ERR_CUDA;
    once.done(bit);
  }
  const int m_tiles = (g.M + 2 * C2_BM - 1) / (2 * C2_BM);
  const int n_tiles = (g.N + C2_BN - 1) / C2_BN;
  const int clusters = min(num_sms / 2, m_tiles * n_tiles);
  GemmEpi e;
  e.bias = g.bias; e.colscale = g.colscale; e.out = g.out; e.out_lo = g.out_lo; e.lo_from_col = g.lo_from_col; e.ldc = g.ldc;
  e.out_lo8 = g.out_lo8; e.ldc8 = g.ldc8;
  e.mode = g.mode; e.act = g.act; e.accumulate = g.accumulate;
  gemm_tc5_2cta_kernel<KIND><<<2 * clusters, C2_THREADS, C2_TOTAL, stream>>>(ta, tb, ta8, tb8, g.M, g.N, g.K, g.kwrap, g.awrap,
                                                                               g.A8 ? g.K8 : 0, e);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

int gemm_tc5_2cta(const GemmCall& g, cudaStream_t stream, int num_sms) {
  switch (gemm_epi_kind(g)) {
    case E_F32: return launch_2cta<E_F32>(g, stream, num_sms);
    case E_F32_ACC: return launch_2cta<E_F32_ACC>(g, stream, num_sms);
    case E_F16: return launch_2cta<E_F16>(g, stream, num_sms);
    case E_SWIGLU: return launch_2cta<E_SWIGLU>(g, stream, num_sms);
    case E_F16_SPLIT: return launch_2cta<E_F16_SPLIT>(g, stream, num_sms);
    case E_SWIGLU_SPLIT: return launch_2cta<E_SWIGLU_SPLIT>(g, stream, num_sms);
    case E_SWIGLU_SPLIT8: return launch_2cta<E_SWIGLU_SPLIT8>(g, stream, num_sms);
    default: return launch_2cta<E_GENERIC>(g, stream, num_sms);
  }
}

}  // namespace m5
