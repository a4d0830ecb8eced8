// Flash attention on the 5th-generation tensor cores (tcgen05 + TMEM) for the NAR transformer: bidirectional
// softmax(Q K^T / 8) V over packed variable-length sequences, head_dim 64 (reference: nn.MultiheadAttention inside
// nn.TransformerEncoder/DecoderLayer, mars5/model.py:179-204,339-341).
//
// One CTA = 128 queries of one (sequence, head); keys/values stream in tiles of 128.
//   warp 0     TMA producer: Q once, then K_j / V_j tiles (cp.async.bulk.tensor, 128B swizzle) into single buffers that
//              are refilled as soon as the MMA that read them has committed
//   warp 1     MMA issuer:  S_j = Q K_j^T   (UMMA 128x128x16, both operands K-major)         -> TMEM cols [0,128)
//                           O_j = P_j V_j   (UMMA 128x64x16, A = P_j in smem, B = V_j MN-major) -> TMEM cols [128,192)
//   warps 2-5  softmax, one query row per thread: two passes over S_j in TMEM (row max, then exp2 / row sum), P_j written
//              to shared memory as fp16 in the K-major 128B-swizzle layout the UMMA A-descriptor expects; O_j is read
//              back from TMEM and accumulated in registers with the usual online-softmax rescale (no TMEM read-modify-
//              write); final O / l is stored as fp16.
// 80 KB of shared memory and 256 TMEM columns per CTA -> two CTAs per SM, so one CTA's exponentials overlap the other's
// MMAs without an intra-CTA ping-pong.
#include <cuda.h>

#include "m5_internal.h"
#include "ptx.cuh"

namespace m5 {

static constexpr int AT_BQ = 128, AT_BK = 128, AT_HD = 64, AT_THREADS = 192;
static constexpr int AT_Q_BYTES = AT_BQ * AT_HD * 2;     // 16 KB
static constexpr int AT_K_BYTES = AT_BK * AT_HD * 2;     // 16 KB
static constexpr int AT_V_BYTES = AT_BK * AT_HD * 2;     // 16 KB
static constexpr int AT_P_BYTES = AT_BQ * AT_BK * 2;     // 32 KB (two 128x64 K-major blocks)
static constexpr int AT_OFF_Q = 0, AT_OFF_K = AT_Q_BYTES, AT_OFF_V = AT_OFF_K + AT_K_BYTES, AT_OFF_P = AT_OFF_V + AT_V_BYTES;
static constexpr int AT_OFF_BAR = AT_OFF_P + AT_P_BYTES;
static constexpr int AT_SMEM = AT_OFF_BAR + 128 + 1024;

// Instruction descriptor: fp16 x fp16 -> fp32, A K-major, B major selectable.
__host__ __device__ constexpr uint32_t at_idesc(uint32_t M, uint32_t N, uint32_t b_mn_major) {
  return (1u << 4) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// MN-major operand stored as [k rows][64 elements = 128 B] with the 128B swizzle (exactly what TMA writes for a
// [rows, 64] box): 8-row groups 1024 B apart along K, a single 64-wide atom along MN.
M5_DEVINL uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 16;  // leading byte offset (between 64-wide MN atoms; only one atom is used)
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset between 8-row groups along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

M5_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnTc5Params {
  const int* q_start; const int* q_len; const int* k_start; const int* k_len;
  __half* O; int ldo;
  float scale_log2;
};

__global__ void __launch_bounds__(AT_THREADS, 2)
flash_tc5_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, AttnTc5Params p) {
  const int seq = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
  const int q_len = p.q_len[seq], k_len = p.k_len[seq];
  const int q0 = qt * AT_BQ;
  if (q0 >= q_len) return;
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = at_smem_raw + ((1024u - (smem_u32(at_smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + AT_OFF_BAR);
  uint64_t *q_full = bar, *k_full = bar + 1, *v_full = bar + 2, *k_free = bar + 3, *v_free = bar + 4, *s_ready = bar + 5,
           *p_ready = bar + 6, *pv_done = bar + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (k_len + AT_BK - 1) / AT_BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(v_full, 1); mbar_init(k_free, 1); mbar_init(v_free, 1);
    mbar_init(s_ready, 1); mbar_init(p_ready, 4); mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tc5_alloc(tmem_slot, 256);
  tc5_fence_before();
  __syncthreads();
  tc5_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      const int qrow = p.q_start[seq] + q0, krow = p.k_start[seq];
      mbar_arrive_expect_tx(q_full, AT_Q_BYTES);
      tma_load_2d(smem + AT_OFF_Q, &tmap_q, q_full, head * AT_HD, qrow);
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(k_free, ph ^ 1);
        mbar_arrive_expect_tx(k_full, AT_K_BYTES);
        tma_load_2d(smem + AT_OFF_K, &tmap_k, k_full, head * AT_HD, krow + j * AT_BK);
        mbar_wait(v_free, ph ^ 1);
        mbar_arrive_expect_tx(v_full, AT_V_BYTES);
        tma_load_2d(smem + AT_OFF_V, &tmap_v, v_full, head * AT_HD, krow + j * AT_BK);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = at_idesc(AT_BQ, AT_BK, 0);   // S = Q K^T : B (= K tile) K-major
      constexpr uint32_t idesc_o = at_idesc(AT_BQ, AT_HD, 1);   // O = P V   : B (= V tile) MN-major
      const uint64_t dq = umma_desc_k_sw128(smem_u32(smem + AT_OFF_Q));
      const uint64_t dk = umma_desc_k_sw128(smem_u32(smem + AT_OFF_K));
      const uint64_t dp = umma_desc_k_sw128(smem_u32(smem + AT_OFF_P));
      const uint64_t dv = umma_desc_mn_sw128(smem_u32(smem + AT_OFF_V));
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t ph = j & 1;
        // S_j (the previous S was fully consumed before p_ready[j-1] was signalled)
        mbar_wait(k_full, ph);
        tc5_fence_after();
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k) tc5_mma_f16(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        tc5_commit(k_free);
        tc5_commit(s_ready);
        // O_j = P_j V_j
        mbar_wait(v_full, ph);
        mbar_wait(p_ready, ph);
        tc5_fence_after();
#pragma unroll
        for (int k = 0; k < AT_BK / 16; ++k) {
          // A: 16 keys = 32 bytes inside the 128-byte row of P block (k / 4); B: 16 key rows = 2048 bytes of the V tile
          const uint64_t da = dp + (uint64_t)((k >> 2) * (AT_BQ * 128 >> 4)) + 2 * (k & 3);
          const uint64_t db = dv + (uint64_t)(k * (16 * 128 >> 4));
          tc5_mma_f16(tmem_o, da, db, idesc_o, k != 0);
        }
        tc5_commit(v_free);
        tc5_commit(pv_done);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / output warps: thread = query row
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                     // row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    float o_acc[AT_HD];
#pragma unroll
    for (int i = 0; i < AT_HD; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* sp = smem + AT_OFF_P;
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t ph = j & 1;
      const int kvalid = min(AT_BK, k_len - j * AT_BK);
      mbar_wait(s_ready, ph);
      tc5_fence_after();
      const bool full_tile = kvalid == AT_BK;  // warp-uniform: only the last tile of a sequence needs key masking
      // pass 1: row max (the TMEM load of chunk c+1 is in flight while chunk c is reduced)
      float mx = -INFINITY;
      {
        uint32_t ra[32], rb[32];
        tc5_ld_32x32(tmem_s + lane_off, ra);
#pragma unroll
        for (int c = 0; c < AT_BK / 32; c += 2) {
          tc5_wait_ld();
          tc5_ld_32x32(tmem_s + lane_off + (c + 1) * 32, rb);
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (full_tile || c * 32 + i < kvalid) mx = fmaxf(mx, __uint_as_float(ra[i]));
          tc5_wait_ld();
          if (c + 2 < AT_BK / 32) tc5_ld_32x32(tmem_s + lane_off + (c + 2) * 32, ra);
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (full_tile || (c + 1) * 32 + i < kvalid) mx = fmaxf(mx, __uint_as_float(rb[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      // O_{j-1} is complete: fold it into the register accumulator, then rescale to the new maximum
      if (j > 0) {
        mbar_wait(pv_done, ph ^ 1);
        tc5_fence_after();
#pragma unroll
        for (int c = 0; c < AT_HD / 32; ++c) {
          uint32_t r[32];
          tc5_ld_32x32(tmem_o + lane_off + c * 32, r);
          tc5_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] += __uint_as_float(r[i]);
        }
      }
      const float corr = ex2_approx(m_run - m_safe);  // m_run = -inf -> 0
      if (corr != 1.0f) {
#pragma unroll
        for (int i = 0; i < AT_HD; ++i) o_acc[i] *= corr;
      }
      l_run *= corr;
      m_run = m_new;
      // pass 2: P = exp2(S * scale - m), row sum, fp16 into the swizzled K-major layout
      float rs = 0.f;
      {
        uint32_t r[32];
        tc5_ld_32x32(tmem_s + lane_off, r);
#pragma unroll
        for (int c = 0; c < AT_BK / 32; ++c) {
          tc5_wait_ld();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_safe));
            float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -m_safe));
            if (!full_tile) {
              if (c * 32 + i >= kvalid) p0 = 0.f;
              if (c * 32 + i + 1 >= kvalid) p1 = 0.f;
            }
            rs += p0 + p1;
            pk[i >> 1] = pack_half2(p0, p1);
          }
          if (c + 1 < AT_BK / 32) tc5_ld_32x32(tmem_s + lane_off + (c + 1) * 32, r);  // in flight during the smem stores
          // 32 columns = 4 chunks of 16 bytes; chunk index inside the 64-column block: (c & 1) * 4 + q
          uint8_t* blk = sp + (c >> 1) * (AT_BQ * 128) + row * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = ((c & 1) * 4 + q) ^ (row & 7);
            *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
      }
      l_run += rs;
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc5_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    // last O tile
    if (n_tiles > 0) {
      mbar_wait(pv_done, (n_tiles - 1) & 1);
      tc5_fence_after();
#pragma unroll
      for (int c = 0; c < AT_HD / 32; ++c) {
        uint32_t r[32];
        tc5_ld_32x32(tmem_o + lane_off + c * 32, r);
        tc5_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] += __uint_as_float(r[i]);
      }
    }
    if (q0 + row < q_len) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      __half* og = p.O + (size_t)(p.q_start[seq] + q0 + row) * p.ldo + head * AT_HD;
#pragma unroll
      for (int i = 0; i < AT_HD; i += 8) {
        uint4 w;
        w.x = pack_half2(o_acc[i] * inv, o_acc[i + 1] * inv);
        w.y = pack_half2(o_acc[i + 2] * inv, o_acc[i + 3] * inv);
        w.z = pack_half2(o_acc[i + 4] * inv, o_acc[i + 5] * inv);
        w.w = pack_half2(o_acc[i + 6] * inv, o_acc[i + 7] * inv);
        *reinterpret_cast<uint4*>(og + i) = w;
      }
    }
  }
  tc5_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc5_fence_after();
    tc5_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn at_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
static int at_tmap(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = at_encode_fn();
  if (!fn) return M5_ERR_CUDA;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {AT_HD, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? M5_OK : M5_ERR_CUDA;
}

int flash_attn_tc5(const AttnCall& c, cudaStream_t stream) {
  if (c.n_seqs <= 0 || c.max_q <= 0) return M5_OK;
  if (c.causal || c.q_rows <= 0 || c.k_rows <= 0) return M5_ERR_ARG;
  if ((c.ldq | c.ldk | c.ldv | c.ldo) % 8 != 0) return M5_ERR_ARG;
  CUtensorMap tq, tk, tv;
  const uint64_t cols = (uint64_t)c.n_heads * AT_HD;
  if (at_tmap(&tq, c.Q, c.q_rows, cols, c.ldq, AT_BQ) != M5_OK) return M5_ERR_CUDA;
  if (at_tmap(&tk, c.K, c.k_rows, cols, c.ldk, AT_BK) != M5_OK) return M5_ERR_CUDA;
  if (at_tmap(&tv, c.V, c.k_rows, cols, c.ldv, AT_BK) != M5_OK) return M5_ERR_CUDA;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(flash_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess) return M5_ERR_CUDA;
    attr_set = true;
  }
  AttnTc5Params p;
  p.q_start = c.q_start; p.q_len = c.q_len; p.k_start = c.k_start; p.k_len = c.k_len; p.O = c.O; p.ldo = c.ldo;
  p.scale_log2 = c.scale * 1.4426950408889634f;
  dim3 grid((c.max_q + AT_BQ - 1) / AT_BQ, c.n_heads, c.n_seqs);
  flash_tc5_kernel<<<grid, AT_THREADS, AT_SMEM, stream>>>(tq, tk, tv, p);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

}  // namespace m5
