// Flash attention on the 5th-generation tensor cores (tcgen05 + TMEM) for the NAR transformer: bidirectional
// softmax(Q K^T / 8) V over packed variable-length sequences, head_dim 64 (reference: nn.MultiheadAttention inside
// nn.TransformerEncoder/DecoderLayer, mars5/model.py:179-204,339-341).
//
// One CTA = 128 queries of one (sequence, head); keys/values stream in tiles of 64.
//   warp 0     TMA producer: Q once, then K_j / V_j tiles (cp.async.bulk.tensor, 128B swizzle) into a 4-deep K ring and
//              a 3-deep V ring; K runs two tiles ahead of V
//   warp 1     MMA issuer:  S_j = Q K_j^T   (UMMA 128x64x16, both operands K-major)  -> TMEM score slot j % 3
//                           O  += P_j V_j   (UMMA 128x64x16, A = P_j in smem, B = V_j MN-major) -> TMEM columns 192..255
//              S_{j+2} is issued before PV_j, so the scores of the next two tiles are computed while the softmax warps
//              still work on tile j: the S -> softmax -> P -> PV chain of a CTA has no tensor-core round trip in it
//   warps 2-5  softmax, one query row per thread, ONE pass over S_j ("lazy" reference point):
//              * O stays resident in TMEM for the whole key loop (PV accumulates in place);
//              * a tile is exponentiated against the row maximum already in use while its own maximum is tracked on the
//                side; only when some row of the warp grew by more than 2^8 (or on the first tile) is the exact path
//                taken: row maximum first, then O (tcgen05.ld -> mul -> tcgen05.st) and l are moved to the new
//                reference point.  P <= 2^8 fits fp16 and O / l does not depend on the reference point, so the result
//                is the same softmax(QK^T/8) V;
//              * sums and maxima use independent accumulators (3-input max), P_j is written to shared memory as fp16 in
//                the K-major 128B-swizzle layout the UMMA A-descriptor expects (double buffered).
// 104 KB of shared memory and 256 TMEM columns per CTA -> two CTAs per SM.
//
// History (profiles/README.md): the first version (128-key tiles, two passes over S, O folded into registers every
// tile) ran at 413 TFLOP/s on the NAR shape; one-pass lazy softmax with O in TMEM 551; S issued ahead of PV 620; this
// 64-key ring 610 with the tensor round trip gone -- what is left is per-tile instruction overhead and MUFU.EX2.
#include <cuda.h>
#include <cuda_fp8.h>

#include "m5_internal.h"
#include "ptx.cuh"

namespace m5 {

static constexpr int AT_BQ = 128, AT_BK = 64, AT_HD = 64, AT_THREADS = 192;
static constexpr int AT_KLEAD = 2, AT_SST = 3;           // K lead over V (tiles), S slots
static constexpr int AT_Q_BYTES = AT_BQ * AT_HD * 2;     // 16 KB
static constexpr int AT_KV_BYTES = AT_BK * AT_HD * 2;    //  8 KB
static constexpr int AT_P_BYTES = AT_BQ * AT_BK * 2;     // 16 KB (one 128-row K-major block)
static constexpr int AT_TMEM_COLS = 256, AT_TMEM_O = AT_SST * AT_BK;   // S slots at 0/64/128, O at 192
static_assert(AT_TMEM_O + AT_HD <= AT_TMEM_COLS, "TMEM budget");
// SPLIT = values (and, with KPAIR, keys) arrive as fp16 (hi, lo) pairs ("mixed" NAR numerics): a ring stage
// holds the hi tile followed by the lo tile, S_j = Q K_hi^T + Q K_lo^T and O += P V_hi + P V_lo accumulate in the same TMEM
// columns, and O is written as a (hi, lo) pair.  The rings are shallower (2 stages of 16 KB) so that two CTAs still fit one SM.
// SPLIT && !KPAIR ("mixed8k"): keys are single fp16 values -- one S pass instead of two, 4 K stages of 8 KB; their rounding is
// averaged over the ~2k keys of a decoder sequence (tools/precision_budget_mixed8.py: 1.3e-5 rms on the logits).
template <bool SPLIT, bool KPAIR>
struct AtCfg {
  static_assert(SPLIT || !KPAIR, "key pairs only together with value pairs");
  static constexpr int KST = KPAIR ? 2 : 4, VST = SPLIT ? 2 : 3;        // K ring, V ring
  static constexpr int KSTAGE = KPAIR ? 2 * AT_KV_BYTES : AT_KV_BYTES, VSTAGE = SPLIT ? 2 * AT_KV_BYTES : AT_KV_BYTES;
  static constexpr int OFF_Q = 0, OFF_K = AT_Q_BYTES, OFF_V = OFF_K + KST * KSTAGE, OFF_P = OFF_V + VST * VSTAGE,
                       OFF_BAR = OFF_P + 2 * AT_P_BYTES;
  static constexpr int SMEM = OFF_BAR + 256;
  // mbarrier slots
  static constexpr int B_QFULL = 0, B_KFULL = 1, B_KFREE = B_KFULL + KST, B_VFULL = B_KFREE + KST, B_VFREE = B_VFULL + VST,
                       B_SREADY = B_VFREE + VST, B_PREADY = B_SREADY + AT_SST, B_PVDONE = B_PREADY + 2, B_COUNT = B_PVDONE + 2;
  static_assert(B_COUNT * 8 + 8 <= 256, "barrier block");
  static_assert(2 * (SMEM + 1024) <= 228 * 1024, "two CTAs per SM");
};

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
M5_DEVINL float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

struct AttnTc5Params {
  const int* q_start; const int* q_len; const int* k_start; const int* k_len;
  __half* O; __half* Olo; int ldo;
  uint8_t* Olo8; int ldo8;   // SPLIT: lo half of the output as e5m2 scaled by 2^-2 instead of fp16 (mixed8 numerics)
  float scale_log2;
};

// 32 score columns of this thread's row -> P (fp16, swizzled smem), partial row sums, running tile maximum.
// FULL = no key of the chunk is beyond the sequence (every tile but possibly the last).
template <bool FULL>
M5_DEVINL void softmax_chunk(const uint32_t (&r)[32], int c, int kvalid, float scale, float m_used, float (&rs)[4],
                             float (&tm)[2], uint8_t* sp, int row) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const int q = i >> 1;
    float s0 = __uint_as_float(r[i]), s1 = __uint_as_float(r[i + 1]);
    float p0 = ex2_approx(fmaf(s0, scale, -m_used));
    float p1 = ex2_approx(fmaf(s1, scale, -m_used));
    if (!FULL) {
      if (c * 32 + i >= kvalid) { p0 = 0.f; s0 = -INFINITY; }
      if (c * 32 + i + 1 >= kvalid) { p1 = 0.f; s1 = -INFINITY; }
    }
    rs[q & 3] += p0 + p1;
    tm[q & 1] = fmax3(tm[q & 1], s0, s1);
    pk[q] = pack_half2(p0, p1);
  }
  // 32 columns = 4 chunks of 16 bytes of the 128-byte row; 16-byte chunk index XOR (row & 7) = 128B swizzle
  uint8_t* blk = sp + row * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = (c * 4 + q) ^ (row & 7);
    *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
}

template <bool FULL>
M5_DEVINL float chunk_max(const uint32_t (&r)[32], int c, int kvalid, float a) {
  float b = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float s0 = __uint_as_float(r[i]), s1 = __uint_as_float(r[i + 1]), s2 = __uint_as_float(r[i + 2]), s3 = __uint_as_float(r[i + 3]);
    if (!FULL) {
      if (c * 32 + i >= kvalid) s0 = -INFINITY;
      if (c * 32 + i + 1 >= kvalid) s1 = -INFINITY;
      if (c * 32 + i + 2 >= kvalid) s2 = -INFINITY;
      if (c * 32 + i + 3 >= kvalid) s3 = -INFINITY;
    }
    a = fmax3(a, s0, s1);
    b = fmax3(b, s2, s3);
  }
  return fmaxf(a, b);
}

// ring position that advances by one tile without integer division
struct Ring {
  int slot = 0;
  uint32_t phase = 0;
  M5_DEVINL void next(int depth) {
    if (++slot == depth) { slot = 0; phase ^= 1u; }
  }
};

template <bool SPLIT, bool KPAIR>
__global__ void __launch_bounds__(AT_THREADS, 2)
flash_tc5_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_klo,
                 const __grid_constant__ CUtensorMap tmap_vlo, AttnTc5Params p) {
  using C = AtCfg<SPLIT, KPAIR>;
  constexpr int AT_KST = C::KST, AT_VST = C::VST, AT_KSTAGE = C::KSTAGE, AT_VSTAGE = C::VSTAGE;
  constexpr int AT_OFF_Q = C::OFF_Q, AT_OFF_K = C::OFF_K, AT_OFF_V = C::OFF_V, AT_OFF_P = C::OFF_P, AT_OFF_BAR = C::OFF_BAR;
  constexpr int B_QFULL = C::B_QFULL, B_KFULL = C::B_KFULL, B_KFREE = C::B_KFREE, B_VFULL = C::B_VFULL, B_VFREE = C::B_VFREE,
                B_SREADY = C::B_SREADY, B_PREADY = C::B_PREADY, B_PVDONE = C::B_PVDONE, B_COUNT = C::B_COUNT;
  const int seq = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
  const int q_len = p.q_len[seq], k_len = p.k_len[seq];
  const int q0 = qt * AT_BQ;
  if (q0 >= q_len) return;
  // no static shared memory in this kernel: the dynamic window starts at the CTA's shared base, which is 1024-byte aligned
  // (checked below); the 128B-swizzle tiles need that alignment
  extern __shared__ __align__(1024) uint8_t at_smem_raw[];
  uint8_t* smem = at_smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) { printf("m5: flash_tc5 shared memory base is not 1 KiB aligned\n"); __trap(); }
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + AT_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + B_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (k_len + AT_BK - 1) / AT_BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    for (int i = 0; i < B_COUNT; ++i) mbar_init(bar + i, (i == B_PREADY || i == B_PREADY + 1) ? 4 : 1);
    fence_barrier_init();
  }
  if (warp == 1) tc5_alloc(tmem_slot, AT_TMEM_COLS);
  tc5_fence_before();
  __syncthreads();
  tc5_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + AT_TMEM_O;

  if (warp == 0) {
    if (lane == 0) {
      const int qrow = p.q_start[seq] + q0, krow = p.k_start[seq];
      mbar_arrive_expect_tx(bar + B_QFULL, AT_Q_BYTES);
      tma_load_2d(smem + AT_OFF_Q, &tmap_q, bar + B_QFULL, head * AT_HD, qrow);
      // K runs AT_KLEAD tiles ahead of V: S_{j+2} is issued two tiles before PV_j, and a V slot only frees when its PV
      // has committed -- a strictly alternating K_j, V_j order would hold K_{j+2} back behind V_{j+1}.
      Ring kr, vr;
      for (int i = 0; i < n_tiles + AT_KLEAD; ++i) {
        if (i < n_tiles) {
          mbar_wait(bar + B_KFREE + kr.slot, kr.phase ^ 1);
          mbar_arrive_expect_tx(bar + B_KFULL + kr.slot, AT_KSTAGE);
          tma_load_2d(smem + AT_OFF_K + kr.slot * AT_KSTAGE, &tmap_k, bar + B_KFULL + kr.slot, head * AT_HD, krow + i * AT_BK);
          if (KPAIR)
            tma_load_2d(smem + AT_OFF_K + kr.slot * AT_KSTAGE + AT_KV_BYTES, &tmap_klo, bar + B_KFULL + kr.slot, head * AT_HD, krow + i * AT_BK);
          kr.next(AT_KST);
        }
        if (i >= AT_KLEAD) {
          mbar_wait(bar + B_VFREE + vr.slot, vr.phase ^ 1);
          mbar_arrive_expect_tx(bar + B_VFULL + vr.slot, AT_VSTAGE);
          tma_load_2d(smem + AT_OFF_V + vr.slot * AT_VSTAGE, &tmap_v, bar + B_VFULL + vr.slot, head * AT_HD,
                      krow + (i - AT_KLEAD) * AT_BK);
          if (SPLIT)
            tma_load_2d(smem + AT_OFF_V + vr.slot * AT_VSTAGE + AT_KV_BYTES, &tmap_vlo, bar + B_VFULL + vr.slot, head * AT_HD,
                        krow + (i - AT_KLEAD) * AT_BK);
          vr.next(AT_VST);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = at_idesc(AT_BQ, AT_BK, 0);   // S = Q K^T : B (= K tile) K-major
      constexpr uint32_t idesc_o = at_idesc(AT_BQ, AT_HD, 1);   // O += P V  : B (= V tile) MN-major
      const uint64_t dq = umma_desc_k_sw128(smem_u32(smem + AT_OFF_Q));
      Ring kr, sr;   // position of the next S tile to issue
      auto issue_s = [&]() {
        mbar_wait(bar + B_KFULL + kr.slot, kr.phase);
        tc5_fence_after();
        const uint64_t dk = umma_desc_k_sw128(smem_u32(smem + AT_OFF_K + kr.slot * AT_KSTAGE));
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k) tc5_mma_f16(tmem_base + sr.slot * AT_BK, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        if (KPAIR) {   // + Q K_lo^T: the lo tile sits AT_KV_BYTES behind the hi tile of the stage
          const uint64_t dkl = dk + (uint64_t)(AT_KV_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < AT_HD / 16; ++k) tc5_mma_f16(tmem_base + sr.slot * AT_BK, dq + 2 * k, dkl + 2 * k, idesc_s, 1);
        }
        tc5_commit(bar + B_KFREE + kr.slot);
        tc5_commit(bar + B_SREADY + sr.slot);
        kr.next(AT_KST);
        sr.next(AT_SST);
      };
      mbar_wait(bar + B_QFULL, 0);
      if (n_tiles > 0) issue_s();
      if (n_tiles > 1) issue_s();
      Ring vr, pr;
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 2 < n_tiles) issue_s();   // its slot was last read by tile j-1: released by p_ready(j-1), waited below
        mbar_wait(bar + B_PREADY + pr.slot, pr.phase);
        mbar_wait(bar + B_VFULL + vr.slot, vr.phase);
        tc5_fence_after();
        const uint64_t dp = umma_desc_k_sw128(smem_u32(smem + AT_OFF_P + pr.slot * AT_P_BYTES));
        const uint64_t dv = umma_desc_mn_sw128(smem_u32(smem + AT_OFF_V + vr.slot * AT_VSTAGE));
#pragma unroll
        for (int k = 0; k < AT_BK / 16; ++k)   // A: 16 keys = 32 B inside the 128-byte P row; B: 16 key rows = 2048 B of V
          tc5_mma_f16(tmem_o, dp + 2 * k, dv + (uint64_t)(k * (16 * 128 >> 4)), idesc_o, (j != 0) || (k != 0));
        if (SPLIT) {
          const uint64_t dvl = dv + (uint64_t)(AT_KV_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < AT_BK / 16; ++k) tc5_mma_f16(tmem_o, dp + 2 * k, dvl + (uint64_t)(k * (16 * 128 >> 4)), idesc_o, 1);
        }
        tc5_commit(bar + B_VFREE + vr.slot);
        tc5_commit(bar + B_PVDONE + pr.slot);
        vr.next(AT_VST);
        pr.next(2);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / output warps: thread = query row
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                     // row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t to = tmem_o + lane_off;
    const float scale = p.scale_log2;
    float m_used = -INFINITY, l_run = 0.f;
    uint32_t ra[32], rb[32];
    Ring sr, pr;   // score slot / P buffer of tile j
    for (int j = 0; j < n_tiles; ++j) {
      const int kvalid = min(AT_BK, k_len - j * AT_BK);
      const bool full_tile = kvalid == AT_BK;             // warp-uniform: only the last tile needs key masking
      const uint32_t ts = tmem_base + sr.slot * AT_BK + lane_off;
      uint8_t* sp = smem + AT_OFF_P + pr.slot * AT_P_BYTES;
      mbar_wait(bar + B_SREADY + sr.slot, sr.phase);
      if (j >= 2) mbar_wait(bar + B_PVDONE + pr.slot, pr.phase ^ 1);   // PV_{j-2} no longer reads this P buffer
      tc5_fence_after();
      bool exact = (j == 0);
      float rs[4];
      for (;;) {
        tc5_ld_32x32(ts, ra);
        tc5_ld_32x32(ts + 32, rb);
        tc5_wait_ld();
        if (exact) {
          // exact path: row maximum first, then move the reference point (and O, l with it)
          float mx = full_tile ? chunk_max<true>(ra, 0, kvalid, -INFINITY) : chunk_max<false>(ra, 0, kvalid, -INFINITY);
          mx = full_tile ? chunk_max<true>(rb, 1, kvalid, mx) : chunk_max<false>(rb, 1, kvalid, mx);
          float m_new = fmaxf(m_used, mx * scale);
          if (m_new == -INFINITY) m_new = 0.f;
          if (j > 0) {
            const float corr = ex2_approx(m_used - m_new);
            // PV_{j-1} (other P buffer) has committed: tile j-1 is phase (j-1)>>1 of that barrier
            mbar_wait(bar + B_PVDONE + (pr.slot ^ 1), pr.slot ? pr.phase : (pr.phase ^ 1));
            tc5_fence_after();
            tc5_ld_32x32(to, ra);
            tc5_ld_32x32(to + 32, rb);
            tc5_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              ra[i] = __float_as_uint(__uint_as_float(ra[i]) * corr);
              rb[i] = __float_as_uint(__uint_as_float(rb[i]) * corr);
            }
            tc5_st_32x32(to, ra);
            tc5_st_32x32(to + 32, rb);
            tc5_wait_st();
            l_run *= corr;
            tc5_ld_32x32(ts, ra);       // the score registers were used as scratch: read S_j again
            tc5_ld_32x32(ts + 32, rb);
            tc5_wait_ld();
          }
          m_used = m_new;
        }
        rs[0] = rs[1] = rs[2] = rs[3] = 0.f;
        float tm[2] = {-INFINITY, -INFINITY};
        if (full_tile) {
          softmax_chunk<true>(ra, 0, kvalid, scale, m_used, rs, tm, sp, row);
          softmax_chunk<true>(rb, 1, kvalid, scale, m_used, rs, tm, sp, row);
        } else {
          softmax_chunk<false>(ra, 0, kvalid, scale, m_used, rs, tm, sp, row);
          softmax_chunk<false>(rb, 1, kvalid, scale, m_used, rs, tm, sp, row);
        }
        if (exact) break;
        // stale reference point: acceptable while no row of this warp outgrew it by more than 2^8
        const bool grew = fmaf(fmaxf(tm[0], tm[1]), scale, -m_used) > 8.f;
        if (!__any_sync(0xffffffffu, grew)) break;
        exact = true;   // redo this tile on the exact path (S is still in TMEM, P has not been published)
      }
      l_run += (rs[0] + rs[1]) + (rs[2] + rs[3]);
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc5_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + B_PREADY + pr.slot);
      sr.next(AT_SST);
      pr.next(2);
    }
    if (n_tiles > 0) {
      const int last = n_tiles - 1;   // tile `last` is phase last>>1 of PV barrier last&1; PVs commit in order
      mbar_wait(bar + B_PVDONE + (last & 1), (last >> 1) & 1);
      tc5_fence_after();
      tc5_ld_32x32(to, ra);
      tc5_ld_32x32(to + 32, rb);
      tc5_wait_ld();
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) ra[i] = rb[i] = 0u;
    }
    if (q0 + row < q_len) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      const size_t ooff = (size_t)(p.q_start[seq] + q0 + row) * p.ldo + head * AT_HD;
      __half* og = p.O + ooff;
      auto store32 = [&](const uint32_t (&r)[32], int base) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint32_t h[4], l[4];
          uint32_t l8[2] = {0u, 0u};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v0 = __uint_as_float(r[i + 2 * e]) * inv, v1 = __uint_as_float(r[i + 2 * e + 1]) * inv;
            h[e] = pack_half2(v0, v1);
            if (SPLIT) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h[e]));
              l[e] = pack_half2(v0 - f.x, v1 - f.y);
              if (p.Olo8) {
                const uint32_t b0 = (uint32_t)__nv_cvt_float_to_fp8((v0 - f.x) * 0.25f, __NV_SATFINITE, __NV_E5M2);
                const uint32_t b1 = (uint32_t)__nv_cvt_float_to_fp8((v1 - f.y) * 0.25f, __NV_SATFINITE, __NV_E5M2);
                l8[e >> 1] |= (b0 | (b1 << 8)) << (16 * (e & 1));
              }
            }
          }
          *reinterpret_cast<uint4*>(og + base + i) = make_uint4(h[0], h[1], h[2], h[3]);
          if (SPLIT) {
            if (p.Olo8) *reinterpret_cast<uint2*>(p.Olo8 + (size_t)(p.q_start[seq] + q0 + row) * p.ldo8 + head * AT_HD + base + i) = make_uint2(l8[0], l8[1]);
            else *reinterpret_cast<uint4*>(p.Olo + ooff + base + i) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
      };
      store32(ra, 0);
      store32(rb, 32);
    }
  }
  tc5_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc5_fence_after();
    tc5_dealloc(tmem_base, AT_TMEM_COLS);
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
  const bool split = c.Vlo != nullptr;   // values (and keys, when Klo is given) as (hi, lo) pairs, O written as a pair; Q and P stay single fp16
  const bool kpair = split && c.Klo != nullptr;
  if ((split && !(c.Olo || c.Olo8)) || (c.Klo && !c.Vlo)) return M5_ERR_ARG;
  CUtensorMap tq, tk, tv, tkl, tvl;
  const uint64_t cols = (uint64_t)c.n_heads * AT_HD;
  if (at_tmap(&tq, c.Q, c.q_rows, cols, c.ldq, AT_BQ) != M5_OK) return M5_ERR_CUDA;
  if (at_tmap(&tk, c.K, c.k_rows, cols, c.ldk, AT_BK) != M5_OK) return M5_ERR_CUDA;
  if (at_tmap(&tv, c.V, c.k_rows, cols, c.ldv, AT_BK) != M5_OK) return M5_ERR_CUDA;
  if (at_tmap(&tkl, kpair ? c.Klo : c.K, c.k_rows, cols, c.ldk, AT_BK) != M5_OK) return M5_ERR_CUDA;
  if (at_tmap(&tvl, split ? c.Vlo : c.V, c.k_rows, cols, c.ldv, AT_BK) != M5_OK) return M5_ERR_CUDA;
  static DeviceOnce once;
  unsigned long long bit;
  if (once.needed(bit)) {
    if (cudaFuncSetAttribute(flash_tc5_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtCfg<false, false>::SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(flash_tc5_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtCfg<true, true>::SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(flash_tc5_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtCfg<true, false>::SMEM) != cudaSuccess)
      return M5_ERR_CUDA;
    once.done(bit);
  }
  AttnTc5Params p;
  p.q_start = c.q_start; p.q_len = c.q_len; p.k_start = c.k_start; p.k_len = c.k_len; p.O = c.O; p.Olo = c.Olo; p.ldo = c.ldo;
  p.Olo8 = c.Olo8; p.ldo8 = c.ldo8;
  p.scale_log2 = c.scale * 1.4426950408889634f;
  dim3 grid((c.max_q + AT_BQ - 1) / AT_BQ, c.n_heads, c.n_seqs);
  if (kpair) flash_tc5_kernel<true, true><<<grid, AT_THREADS, AtCfg<true, true>::SMEM, stream>>>(tq, tk, tv, tkl, tvl, p);
  else if (split) flash_tc5_kernel<true, false><<<grid, AT_THREADS, AtCfg<true, false>::SMEM, stream>>>(tq, tk, tv, tkl, tvl, p);
  else flash_tc5_kernel<false, false><<<grid, AT_THREADS, AtCfg<false, false>::SMEM, stream>>>(tq, tk, tv, tkl, tvl, p);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

}  // namespace m5
