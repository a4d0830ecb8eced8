// Fused, persistent AR decode step: ONE cooperative kernel runs all transformer layers of a KV-cached CodecLM step for
// B <= 32 utterances (reference loop body: mars5/ar_generate.py:62-71 -> model.py:95-141 -> nn_future.py:235-274,
// 297-333,369-398), followed by the fused sampler kernel (sampler.cu).
//
// Grid = one CTA per SM (256 threads, 8 warps), resident for the whole step; phases are separated by a device-wide
// barrier (5 per layer):
//   P0        x = embed(last token) (+ per-row sum-of-squares partials)
//   per layer:
//     P1  QKV projection      X = fp16(RMSNorm(x) * gamma) staged per CTA, weights streamed once through a 16-deep per-warp
//                             cp.async ring (128 KB in flight per SM) into mma.sync fragments, split-K partials, last CTA of
//                             a row tile reduces -> qkv fp32
//     P2  attention           one (utterance, head, 256-key split) per WARP: q / new k get RoPE from the fp32 qkv row, the new
//                             k, v are appended to the fp16 KV cache, the cached keys / values arrive as 32-key TMA tiles
//                             (cp.async.bulk.tensor, 4-deep per-warp mbarrier ring in shared memory, 6 attention warps per
//                             CTA), blocked online softmax in fp32; per-split (m, l, acc) partials only when the context is cut
//     P3  WO projection       X = merged attention output (split partials merged while staging); epilogue x += ...,
//                             and the row sums of squares the next RMSNorm needs (per 128-column tile, fixed order)
//     P4  W1|W3 + SwiGLU      X = fp16(RMSNorm(x) * gamma); epilogue silu(w1 x) * (w3 x) -> g fp16
//     P5  W2 projection       X = g; epilogue x += ..., sums of squares
//   final     vocabulary projection of RMSNorm(x) -> logits fp32
// The first weight fragments (or KV tiles) of the NEXT phase are requested before a CTA waits at the barrier, so HBM
// keeps streaming across phase boundaries.  Everything a later phase reads that an earlier phase wrote goes through L2
// (ld.global.cg / cp.async.cg / TMA), never through the non-coherent L1.
// Algorithmic bytes per step: the fp16 weights once (1.37 GB) + every cached K/V once.
#include <cuda.h>

#include "ar_decode.h"
#include "sampler_body.cuh"
#include "ptx.cuh"

namespace m5 {

static constexpr int AD_THREADS = 256, AD_WARPS = 8;
static constexpr int AD_ROWS = 128;      // weight rows per GEMM work item (16 per warp)
static constexpr int AD_WST = 16;        // 32-column weight chunks (1 KB each) a warp keeps in flight through cp.async
static constexpr int AD_W_BYTES = AD_WARPS * AD_WST * 1024;                  // weight staging: 128 KB, [warp][slot][2][lane][16 B]
static constexpr int AD_KT = 32;         // keys per TMA tile
static constexpr int AD_NST = 4;         // TMA stages per attention warp (3 tiles = 24 KB in flight while one is consumed)
static constexpr int AD_AWARPS = 6;      // warps of a CTA that run attention items: 6 x 4 stages x 8 KB = the 192 KB ring
static constexpr int AD_STAGE_BYTES = 2 * AD_KT * 128;                      // K tile + V tile
static constexpr int AD_PART = 68;       // floats per split partial: m, l, pad, pad, acc[64] (16-byte aligned rows)
static constexpr int AD_SMEM_KV = AD_AWARPS * AD_NST * AD_STAGE_BYTES;       // 192 KB (the GEMM phases alias it: W staging + X)
static constexpr int AD_RED = 4;        // split-K partials the reducing CTA keeps in flight per round (16 measured slower: profiles/r2_ar_decode_timeline.txt)
static constexpr int AD_MAX_KSLICE = 896;                                    // activation slice: 32 rows x (2 * 896 + 64) B = 58 KB
static_assert(AD_W_BYTES + 32 * (AD_MAX_KSLICE * 2 + 64) <= AD_SMEM_KV, "weight staging + activation slice must fit the KV ring region");
static constexpr int AD_SMEM = AD_SMEM_KV + 1024;                            // + mbarriers, tickets

M5_DEVINL uint4 ad_ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
M5_DEVINL unsigned ad_ld_relaxed(const unsigned* p) {   // L2 poll without the L1 invalidation an acquire load implies
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Device-wide barrier for a co-resident grid (cooperative launch): monotonic arrival counter, zeroed by the host before
// every launch.  `epoch` is the number of arrivals that completes the barrier this CTA is about to join.
M5_DEVINL void grid_sync(unsigned* bar, unsigned& epoch, unsigned long long* prof = nullptr) {
  __syncthreads();
  if (prof && blockIdx.x == 0 && threadIdx.x == 0) prof[0] = global_timer_ns();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    if (ad_ld_relaxed(bar) < epoch) {
      const uint64_t t0 = global_timer_ns();
      uint32_t spins = 0;
      while (ad_ld_relaxed(bar) < epoch) {
        if ((++spins & 0x3FF) == 0 && global_timer_ns() - t0 > 2000000000ull) {
          printf("m5: ar_decode grid barrier timed out (block %d, epoch %u)\n", blockIdx.x, epoch);
          __trap();
        }
      }
    }
    __threadfence();
    if (prof && blockIdx.x == 0) prof[1] = global_timer_ns();
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ GEMM phases
// item -> (row tile, K slice); rows of W this lane streams
struct GemmItem {
  int tile, ks, n0, kbase;
  const __half *w0, *w1;
};
M5_DEVINL GemmItem gemm_item(const ArGemm& g, int item, int warp, int lane) {
  GemmItem it;
  it.tile = item / g.ksplit; it.ks = item - it.tile * g.ksplit;
  it.n0 = it.tile * AD_ROWS + warp * 16;
  it.kbase = it.ks * g.kslice;
  const int gq = lane >> 2, t = lane & 3;
  const int r0 = min(it.n0 + gq, g.N - 1), r1 = min(it.n0 + gq + 8, g.N - 1);
  it.w0 = g.W + (size_t)r0 * g.K + it.kbase + 8 * t;
  it.w1 = g.W + (size_t)r1 * g.K + it.kbase + 8 * t;
  return it;
}
// Weight streaming: every lane copies ITS OWN two 16-byte fragment pieces of chunk c (rows n0+g and n0+g+8, columns
// 32c + 8t ..) with cp.async.cg into a lane-private slot of shared memory, one commit group per chunk.  AD_WST chunks
// (16 KB per warp, 128 KB per SM) are in flight -- more than registers could hold -- and the first ones are requested
// before the preceding grid barrier.
M5_DEVINL void w_issue(const GemmItem& it, uint8_t* wst, int lane, int c) {
  uint8_t* dst = wst + (c & (AD_WST - 1)) * 1024 + lane * 16;
  cp_async16(dst, it.w0 + c * 32, true);
  cp_async16(dst + 512, it.w1 + c * 32, true);
}
M5_DEVINL void w_prologue(const ArGemm& g, const GemmItem& it, uint8_t* wst, int lane) {
  const int chunks = g.kslice / 32;
#pragma unroll
  for (int u = 0; u < AD_WST; ++u) {
    if (u < chunks) w_issue(it, wst, lane, u);
    cp_async_commit();   // (possibly empty) group u: the wait_group arithmetic below stays uniform
  }
}

enum { X_NORM = 0, X_ATTN = 1, X_F16 = 2 };        // how the activation slice of a GEMM phase is produced
enum { EPI_STORE = 0, EPI_RESID = 1, EPI_SWIGLU = 2 };

// Stages X[:, kbase : kbase + kslice] as fp16 rows with a 64-byte skew (conflict-free 16-byte reads).
template <int NT, int XSRC>
M5_DEVINL void stage_x(const ArDecodeParams& p, const ArGemm& g, int kbase, const float* gamma, uint8_t* xs, float* s_scale) {
  constexpr int BT = 8 * NT;
  const int tid = threadIdx.x;
  const int xstride = g.kslice * 2 + 64;
  if constexpr (XSRC == X_NORM) {
    // RMSNorm (nn_future.py:301-312): x * rsqrt(mean(x^2) + eps) * weight, the mean from the per-tile partial sums
    if (tid < BT) {
      float ss = 0.f;
      if (tid < p.B)
        for (int t = 0; t < p.ssq_tiles; ++t) ss += __ldcg(p.ssq + t * 32 + tid);
      s_scale[tid] = rsqrtf(ss / (float)p.D + p.eps);
    }
    __syncthreads();
    // warp w stages rows w, w + 8, ...; a lane covers the float4 columns lane, lane + 32, ... of the slice.  Every load of a
    // pair of rows is requested before the first one is used: the slice arrives in ONE or two L2 round trips instead of one
    // per float4 (a loop of dependent load -> convert -> store iterations cost 12 - 24 serial round trips per GEMM phase)
    const int warp = tid >> 5, lane = tid & 31;
    const int quads = g.kslice / 4;
    constexpr int MAXC = (AD_MAX_KSLICE / 4 + 31) / 32;   // float4 columns per lane
    constexpr int RB = 2;                                  // rows in flight per warp
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 gm[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int q = lane + 32 * c;
      gm[c] = q < quads ? __ldg(reinterpret_cast<const float4*>(gamma + kbase) + q) : z4;
    }
    for (int r0 = warp; r0 < BT; r0 += AD_WARPS * RB) {
      float4 v[RB][MAXC];
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        const int row = r0 + rr * AD_WARPS;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          const int q = lane + 32 * c;
          v[rr][c] = (row < p.B && q < quads) ? __ldcg(reinterpret_cast<const float4*>(p.x + (size_t)row * p.D + kbase) + q) : z4;
        }
      }
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        const int row = r0 + rr * AD_WARPS;
        if (row >= BT) continue;
        const float sc = s_scale[row];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          const int q = lane + 32 * c;
          if (q >= quads) continue;
          uint2 o = make_uint2(0u, 0u);
          if (row < p.B) {
            o.x = pack_half2((v[rr][c].x * sc) * gm[c].x, (v[rr][c].y * sc) * gm[c].y);
            o.y = pack_half2((v[rr][c].z * sc) * gm[c].z, (v[rr][c].w * sc) * gm[c].w);
          }
          *reinterpret_cast<uint2*>(xs + (size_t)row * xstride + q * 8) = o;
        }
      }
    }
  } else {
    const int vec_per_row = g.kslice / 8;
    for (int i = tid; i < BT * vec_per_row; i += AD_THREADS) {
      const int row = i / vec_per_row, v = i - row * vec_per_row;
      const bool ok = row < p.B;
      const __half* src = (XSRC == X_ATTN) ? p.att16 : p.g16;   // merged attention output / gated FFN activations
      cp_async16(xs + (size_t)row * xstride + v * 16, ok ? (src + (size_t)row * g.K + kbase + v * 8) : src, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
  }
  __syncthreads();
}

// One GEMM phase: this CTA's work items (item = cta, cta + grid, ...).  `pre` holds the first weight fragments of the
// first item when `have_pre` (requested before the preceding grid barrier).
template <int NT, int XSRC, int EPI>
M5_DEVINL void gemm_phase(const ArDecodeParams& p, const ArGemm& g, const float* gamma, float* out_f32, int ldo, uint8_t* smem,
                          bool have_pre) {
  constexpr int BT = 8 * NT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, t = lane & 3;
  uint8_t* wst = smem + warp * (AD_WST * 1024);                          // this warp's weight slots
  uint8_t* xs = smem + AD_W_BYTES;                                        // activation slice of the item
  float* s_scale = reinterpret_cast<float*>(smem + AD_SMEM_KV + 512);   // [32]
  int* s_ticket = reinterpret_cast<int*>(smem + AD_SMEM_KV + 512 + 128);
  const int n_items = g.tiles * g.ksplit;
  const int chunks = g.kslice / 32;
  const int xstride = g.kslice * 2 + 64;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const GemmItem it = gemm_item(g, item, warp, lane);
    if (!have_pre) w_prologue(g, it, wst, lane);
    have_pre = false;
    stage_x<NT, XSRC>(p, g, it.kbase, gamma, xs, s_scale);
    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    // chunk c is the (c+1)-th oldest commit group still tracked: all but the newest AD_WST - 1 groups have landed
#pragma unroll 4
    for (int c = 0; c < chunks; ++c) {
      cp_async_wait<AD_WST - 1>();
      const uint8_t* src = wst + (c & (AD_WST - 1)) * 1024 + lane * 16;
      const uint4 wa = *reinterpret_cast<const uint4*>(src), wb = *reinterpret_cast<const uint4*>(src + 512);
      if (c + AD_WST < chunks) w_issue(it, wst, lane, c + AD_WST);   // refill the slot just read (lane-private data)
      cp_async_commit();
      const uint32_t a1[4] = {wa.x, wb.x, wa.y, wb.y};
      const uint32_t a2[4] = {wa.z, wb.z, wa.w, wb.w};
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint4 x = *reinterpret_cast<const uint4*>(xs + (size_t)(nt * 8 + gq) * xstride + (c * 32 + 8 * t) * 2);
        mma_16816(acc[nt], a1, x.x, x.y);
        mma_16816(acc[nt], a2, x.z, x.w);
      }
    }
    cp_async_wait<0>();
    // ---- partial tile -> scratch[item][b][row]
    float* part = p.scratch + (size_t)item * (32 * AD_ROWS);
    const int rl = warp * 16 + gq;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int b0 = nt * 8 + 2 * t;
      __stcg(part + (size_t)b0 * AD_ROWS + rl, acc[nt][0]);
      __stcg(part + (size_t)(b0 + 1) * AD_ROWS + rl, acc[nt][1]);
      __stcg(part + (size_t)b0 * AD_ROWS + rl + 8, acc[nt][2]);
      __stcg(part + (size_t)(b0 + 1) * AD_ROWS + rl + 8, acc[nt][3]);
    }
    __threadfence();
    __syncthreads();   // also: every warp is done with the staged X before the next item overwrites it
    if (tid == 0) *s_ticket = atomicAdd(p.counters + it.tile, 1);
    __syncthreads();
    if (*s_ticket != g.ksplit - 1) continue;
    // ---- last CTA of this row tile: ordered reduction over the K slices + epilogue
    __threadfence();
    if (tid == 0) p.counters[it.tile] = 0;
    const float* base = p.scratch + (size_t)(it.tile * g.ksplit) * (32 * AD_ROWS);
    const int nrow0 = it.tile * AD_ROWS;
    if constexpr (EPI == EPI_SWIGLU) {
      // a thread owns IT (batch row, column pair) outputs; the partial sums of ALL of them are requested before the first is
      // used (AD_RED slices per round), then summed in slice order (deterministic)
      constexpr int IT = BT * (AD_ROWS / 2) / AD_THREADS;
      float a[IT], c[IT];
      bool ok[IT];
#pragma unroll
      for (int u = 0; u < IT; ++u) {
        const int i = tid + u * AD_THREADS;
        const int b = i / (AD_ROWS / 2), pr = i - b * (AD_ROWS / 2);
        ok[u] = b < p.B && nrow0 + 2 * pr + 1 < g.N;
        a[u] = c[u] = 0.f;
      }
      for (int s0 = 0; s0 < g.ksplit; s0 += AD_RED) {
        float2 v2[IT][AD_RED];
#pragma unroll
        for (int u = 0; u < IT; ++u) {
          const int i = tid + u * AD_THREADS;
          const int b = i / (AD_ROWS / 2), pr = i - b * (AD_ROWS / 2);
#pragma unroll
          for (int e = 0; e < AD_RED; ++e)
            v2[u][e] = (ok[u] && s0 + e < g.ksplit)
                           ? __ldcg(reinterpret_cast<const float2*>(base + (size_t)(s0 + e) * (32 * AD_ROWS) + (size_t)b * AD_ROWS + 2 * pr))
                           : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < IT; ++u) {
#pragma unroll
          for (int e = 0; e < AD_RED; ++e) { a[u] += v2[u][e].x; c[u] += v2[u][e].y; }
        }
      }
#pragma unroll
      for (int u = 0; u < IT; ++u) {
        const int i = tid + u * AD_THREADS;
        const int b = i / (AD_ROWS / 2), pr = i - b * (AD_ROWS / 2);
        if (ok[u]) p.g16[(size_t)b * (g.N / 2) + ((nrow0 + 2 * pr) >> 1)] = __float2half_rn((a[u] / (1.f + __expf(-a[u]))) * c[u]);
      }
    } else {
      // warp w owns batch rows w, w + 8, ...; a lane owns 4 consecutive weight rows (= output columns) of the tile.  The
      // residual segments and the partial sums of ALL rows of the warp are requested together (AD_RED slices per round), the
      // sums run in slice order (deterministic): ksplit / AD_RED L2 round trips per tile instead of one chain per batch row
      constexpr int RW = BT / AD_WARPS;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 acc[RW], rpre[RW];
      const bool vec = nrow0 + 4 * lane + 3 < g.N && (ldo & 3) == 0;
#pragma unroll
      for (int u = 0; u < RW; ++u) {
        const int b = warp + u * AD_WARPS;
        acc[u] = z4; rpre[u] = z4;
        if constexpr (EPI == EPI_RESID) {
          if (b < p.B && vec) rpre[u] = __ldcg(reinterpret_cast<const float4*>(out_f32 + (size_t)b * ldo + nrow0 + 4 * lane));
        }
      }
      for (int s0 = 0; s0 < g.ksplit; s0 += AD_RED) {
        float4 q[RW][AD_RED];
#pragma unroll
        for (int u = 0; u < RW; ++u) {
          const int b = warp + u * AD_WARPS;
#pragma unroll
          for (int e = 0; e < AD_RED; ++e)
            q[u][e] = (b < p.B && s0 + e < g.ksplit)
                          ? __ldcg(reinterpret_cast<const float4*>(base + (size_t)(s0 + e) * (32 * AD_ROWS) + (size_t)b * AD_ROWS + 4 * lane))
                          : z4;
        }
#pragma unroll
        for (int u = 0; u < RW; ++u) {
#pragma unroll
          for (int e = 0; e < AD_RED; ++e) { acc[u].x += q[u][e].x; acc[u].y += q[u][e].y; acc[u].z += q[u][e].z; acc[u].w += q[u][e].w; }
        }
      }
#pragma unroll
      for (int u = 0; u < RW; ++u) {
        const int b = warp + u * AD_WARPS;
        if (b >= p.B) continue;
        float4 v = acc[u];
        const float4 r_pre = rpre[u];
        const int n = nrow0 + 4 * lane;
        float* o = out_f32 + (size_t)b * ldo + n;
        float sq = 0.f;
        if (n + 3 < g.N && (ldo & 3) == 0) {
          if constexpr (EPI == EPI_RESID) {
            const float4 r = r_pre;
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            sq = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          }
          __stcg(reinterpret_cast<float4*>(o), v);
        } else {
          const float e[4] = {v.x, v.y, v.z, v.w};
          for (int j = 0; j < 4; ++j) {
            if (n + j < g.N) {
              float w = e[j];
              if constexpr (EPI == EPI_RESID) { w += __ldcg(o + j); sq += w * w; }
              __stcg(o + j, w);
            }
          }
        }
        if constexpr (EPI == EPI_RESID) {
          sq = warp_sum(sq);
          if (lane == 0) __stcg(p.ssq + it.tile * 32 + b, sq);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention phase
struct AttnItem {
  int b, h, s, k0, n_cache, has_new, row0;   // n_cache: cached keys of this split; row0: first cache row of the split
};
M5_DEVINL bool attn_item(const ArDecodeParams& p, int layer, int idx, AttnItem& it) {
  const int per_b = p.H * p.n_split;
  it.b = idx / per_b;
  const int r = idx - it.b * per_b;
  it.h = r / p.n_split;
  it.s = r - it.h * p.n_split;
  if (p.done && p.done[it.b]) return false;
  const int L = p.kv_len[it.b];               // positions 0 .. L-2 are cached, L-1 is the token of this step
  it.k0 = it.s * p.split_keys;
  if (it.k0 >= L) return false;
  it.n_cache = min(p.split_keys, L - 1 - it.k0);
  it.has_new = (L - 1 < it.k0 + p.split_keys) ? 1 : 0;
  it.row0 = (layer * p.B + it.b) * p.Wc + it.k0;
  return true;
}
// walks the (item, tile) sequence of one warp for one layer
struct TileIter {
  int idx, t, n_tiles, row0, h;
  bool valid;
};
M5_DEVINL void tile_seek(const ArDecodeParams& p, int layer, int total, int stride, TileIter& ti) {
  ti.valid = false;
  while (ti.idx < total) {
    AttnItem it;
    if (attn_item(p, layer, ti.idx, it) && it.n_cache > 0) {
      ti.n_tiles = (it.n_cache + AD_KT - 1) / AD_KT; ti.row0 = it.row0; ti.h = it.h; ti.t = 0; ti.valid = true;
      return;
    }
    ti.idx += stride;
  }
}
M5_DEVINL void tile_next(const ArDecodeParams& p, int layer, int total, int stride, TileIter& ti) {
  if (++ti.t < ti.n_tiles) return;
  ti.idx += stride;
  tile_seek(p, layer, total, stride, ti);
}
M5_DEVINL void tile_issue(const CUtensorMap* tk, const CUtensorMap* tv, uint8_t* ring, uint64_t* bars, uint32_t seq, const TileIter& ti) {
  const int st = seq % AD_NST;
  uint8_t* dst = ring + st * AD_STAGE_BYTES;
  mbar_arrive_expect_tx(bars + st, AD_STAGE_BYTES);
  tma_load_2d(dst, tk, bars + st, ti.h * 64, ti.row0 + ti.t * AD_KT);
  tma_load_2d(dst + AD_KT * 128, tv, bars + st, ti.h * 64, ti.row0 + ti.t * AD_KT);
}

struct AttnWarpState {
  uint32_t issued, consumed;   // tiles since kernel start (stage = seq % NST, parity = (seq / NST) & 1)
  TileIter prod;
};

M5_DEVINL void attn_prefetch(const ArDecodeParams& p, const CUtensorMap* tk, const CUtensorMap* tv, int layer, uint8_t* ring,
                             uint64_t* bars, AttnWarpState& st) {
  if ((threadIdx.x >> 5) >= AD_AWARPS) return;
  const int lane = threadIdx.x & 31, gw = blockIdx.x * AD_AWARPS + (threadIdx.x >> 5);
  const int total = p.B * p.H * p.n_split, stride = gridDim.x * AD_AWARPS;
  st.prod.idx = gw;
  tile_seek(p, layer, total, stride, st.prod);
  while (st.prod.valid && st.issued - st.consumed < AD_NST) {
    if (lane == 0) tile_issue(tk, tv, ring, bars, st.issued, st.prod);
    ++st.issued;
    tile_next(p, layer, total, stride, st.prod);
  }
}

M5_DEVINL void attn_phase(const ArDecodeParams& p, const ArLayerDev& lw, const CUtensorMap* tk, const CUtensorMap* tv, int layer,
                          uint8_t* ring, uint64_t* bars, AttnWarpState& st) {
  if ((threadIdx.x >> 5) >= AD_AWARPS) return;   // warps 6, 7 own no KV ring: they wait at the CTA barrier that follows
  const int lane = threadIdx.x & 31, gw = blockIdx.x * AD_AWARPS + (threadIdx.x >> 5);
  const int sub = lane & 7, grp = lane >> 3;
  const int total = p.B * p.H * p.n_split, stride = gridDim.x * AD_AWARPS;
  const int D = p.D;
  const float sl2 = 0.125f * 1.4426950408889634f;
  (void)lw;
  for (int idx = gw; idx < total; idx += stride) {
    AttnItem it;
    if (!attn_item(p, layer, idx, it)) continue;
    const int L = p.kv_len[it.b];
    const float* qkv = p.qkv + (size_t)it.b * 3 * D + it.h * 64 + sub * 8;
    // q of this head: fp16 projection output, RoPE at position L-1 in fp32, rounded to fp16 (nn_future.py:186-191)
    float q[8], cs[4], sn[4];
    {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(qkv)), c = __ldcg(reinterpret_cast<const float4*>(qkv + 4));
      const float r[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float ang = (float)(L - 1) * __ldg(p.inv_freq + sub * 4 + i);
        sincosf(ang, &sn[i], &cs[i]);
        const float2 h = __half22float2(__floats2half2_rn(r[2 * i], r[2 * i + 1]));
        const float2 o = __half22float2(__floats2half2_rn(h.x * cs[i] - h.y * sn[i], h.x * sn[i] + h.y * cs[i]));
        q[2 * i] = o.x; q[2 * i + 1] = o.y;
      }
    }
    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // partial dot product of this lane's 8 dims with one key row; the 8 lanes of a key group are summed by the caller
    auto dot8 = [&](const uint4& kv) {
      const __half2* kh = reinterpret_cast<const __half2*>(&kv);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(kh[i]); s += q[2 * i] * f.x + q[2 * i + 1] * f.y; }
      return s;
    };
    const int n_tiles = (it.n_cache + AD_KT - 1) / AD_KT;
    for (int t = 0; t < n_tiles; ++t) {
      const int sg = st.consumed % AD_NST;
      mbar_wait(bars + sg, (st.consumed / AD_NST) & 1);
      const uint8_t* sk = ring + sg * AD_STAGE_BYTES;
      const uint8_t* sv = sk + AD_KT * 128;
      const int kvalid = it.n_cache - t * AD_KT;   // keys of this tile that exist (>= 1)
      // blocked online softmax over the tile: this lane's key group sees keys u * 4 + grp, u = 0 .. 7.  All 8 scores first
      // (independent dot products and shuffles), ONE new running maximum, one rescale of (l, acc), then the 8 weighted
      // value rows -- instead of 8 serially dependent (max, rescale, accumulate) updates
      float sc[AD_KT / 4];
#pragma unroll
      for (int u = 0; u < AD_KT / 4; ++u) sc[u] = dot8(*reinterpret_cast<const uint4*>(sk + (u * 4 + grp) * 128 + sub * 16));
#pragma unroll
      for (int u = 0; u < AD_KT / 4; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 1);
#pragma unroll
      for (int u = 0; u < AD_KT / 4; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 2);
      float mt = -INFINITY;
#pragma unroll
      for (int u = 0; u < AD_KT / 4; ++u) {
        sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 4);
        sc[u] = (u * 4 + grp < kvalid) ? sc[u] * sl2 : -INFINITY;
        mt = fmaxf(mt, sc[u]);
      }
      const float mn = fmaxf(m, mt);
      if (mn > -INFINITY) {   // (a group may own no valid key of a short tile while it is still empty)
        const float c = exp2f(m - mn);
        float ps = 0.f, pa[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pa[i] = 0.f;
#pragma unroll
        for (int u = 0; u < AD_KT / 4; ++u) {
          // rows of the tile beyond the context hold whatever the arena held (possibly NaN / Inf bit patterns): they must be
          // skipped, not multiplied by a zero weight
          if (u * 4 + grp < kvalid) {
            const float pe = exp2f(sc[u] - mn);
            ps += pe;
            const uint4 vv = *reinterpret_cast<const uint4*>(sv + (u * 4 + grp) * 128 + sub * 16);
            const __half2* vh = reinterpret_cast<const __half2*>(&vv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __half22float2(vh[i]);
              pa[2 * i] += pe * f.x;
              pa[2 * i + 1] += pe * f.y;
            }
          }
        }
        l = l * c + ps;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = acc[i] * c + pa[i];
        m = mn;
      }
      __syncwarp();            // every lane has read the stage: it may be refilled
      ++st.consumed;
      if (st.prod.valid) {
        if (lane == 0) tile_issue(tk, tv, ring, bars, st.issued, st.prod);
        ++st.issued;
        tile_next(p, layer, total, stride, st.prod);
      }
    }
    if (it.has_new) {
      // K / V of the token fed in this step: rounded to fp16 like the projection output, K with RoPE; appended to the cache
      // (nn_future.py:248-252) and folded into the softmax straight from registers
      const float4 ka = __ldcg(reinterpret_cast<const float4*>(qkv + D)), kb = __ldcg(reinterpret_cast<const float4*>(qkv + D + 4));
      const float4 va = __ldcg(reinterpret_cast<const float4*>(qkv + 2 * D)), vb = __ldcg(reinterpret_cast<const float4*>(qkv + 2 * D + 4));
      const float kr[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
      uint4 kn, vn;
      uint32_t* kw = reinterpret_cast<uint32_t*>(&kn);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 h = __half22float2(__floats2half2_rn(kr[2 * i], kr[2 * i + 1]));
        kw[i] = pack_half2(h.x * cs[i] - h.y * sn[i], h.x * sn[i] + h.y * cs[i]);
      }
      vn = make_uint4(pack_half2(va.x, va.y), pack_half2(va.z, va.w), pack_half2(vb.x, vb.y), pack_half2(vb.z, vb.w));
      if (grp == 0) {
        const size_t coff = ((size_t)(layer * p.B + it.b) * p.Wc + (L - 1)) * D + it.h * 64 + sub * 8;
        *reinterpret_cast<uint4*>(p.kc + coff) = kn;
        *reinterpret_cast<uint4*>(p.vc + coff) = vn;
      }
      float s = dot8(kn);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (grp == 0) {
        s *= sl2;
        const float mn = fmaxf(m, s);
        const float c = exp2f(m - mn), pe = exp2f(s - mn);
        l = l * c + pe;
        const __half2* vh = reinterpret_cast<const __half2*>(&vn);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __half22float2(vh[i]);
          acc[2 * i] = acc[2 * i] * c + pe * f.x;
          acc[2 * i + 1] = acc[2 * i + 1] * c + pe * f.y;
        }
        m = mn;
      }
    }
    // merge the 4 key groups of the warp (fixed order), write the split partial (m, l, acc[64])
    float M = fmaxf(fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8)), fmaxf(__shfl_xor_sync(0xffffffffu, m, 16), __shfl_xor_sync(0xffffffffu, m, 24)));
    const float c = (m == -INFINITY) ? 0.f : exp2f(m - M);
    l *= c;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= c;
    l += __shfl_xor_sync(0xffffffffu, l, 8);
    l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    float* sp0 = p.attn_part + ((size_t)(it.b * p.H + it.h) * p.n_split) * AD_PART;
    float* sp = sp0 + (size_t)it.s * AD_PART;
    const bool single = (L + p.split_keys - 1) / p.split_keys == 1;
    if (lane == 0 && !single) { __stcg(sp, M); __stcg(sp + 1, l); }
    if (grp == 0 && !single) {
      __stcg(reinterpret_cast<float4*>(sp + 4 + sub * 8), make_float4(acc[0], acc[1], acc[2], acc[3]));
      __stcg(reinterpret_cast<float4*>(sp + 4 + sub * 8 + 4), make_float4(acc[4], acc[5], acc[6], acc[7]));
    }
    // the warp that completes the last split of this (utterance, head) merges all of them (in split order: deterministic)
    // and writes the fp16 attention output row the WO projection stages; the others move on to their next item
    const int ns = (L + p.split_keys - 1) / p.split_keys;
    if (ns == 1) {
      // the whole context of this (utterance, head) was one work item: normalise and write the fp16 output row directly
      const float inv = l > 0.f ? 1.f / l : 0.f;
      if (grp == 0) {
        *reinterpret_cast<uint4*>(p.att16 + (size_t)it.b * D + it.h * 64 + sub * 8) =
            make_uint4(pack_half2(acc[0] * inv, acc[1] * inv), pack_half2(acc[2] * inv, acc[3] * inv),
                       pack_half2(acc[4] * inv, acc[5] * inv), pack_half2(acc[6] * inv, acc[7] * inv));
      }
      continue;
    }
    __threadfence();
    __syncwarp();
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(p.attn_tickets + it.b * p.H + it.h, 1);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket == ns - 1) {
      __threadfence();
      if (lane == 0) p.attn_tickets[it.b * p.H + it.h] = 0;
      // lanes 0 .. ns-1 fetch (m, l) of one split each, everybody gets the maximum and the scaled weights by shuffle
      float ms = -INFINITY, ls = 0.f;
      if (lane < ns) { ms = __ldcg(sp0 + (size_t)lane * AD_PART); ls = __ldcg(sp0 + (size_t)lane * AD_PART + 1); }
      float Mx = ms;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) Mx = fmaxf(Mx, __shfl_xor_sync(0xffffffffu, Mx, o));
      const float cw = (ms == -INFINITY) ? 0.f : exp2f(ms - Mx);
      float Ls = 0.f, ox = 0.f, oy = 0.f;
      for (int s2 = 0; s2 < ns; s2 += 4) {          // 4 splits' rows in flight per round
        float2 a[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          a[e] = (s2 + e < ns) ? __ldcg(reinterpret_cast<const float2*>(sp0 + (size_t)(s2 + e) * AD_PART + 4 + 2 * lane)) : make_float2(0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float c = __shfl_sync(0xffffffffu, cw, min(s2 + e, 31));
          const float le = __shfl_sync(0xffffffffu, ls, min(s2 + e, 31));
          if (s2 + e < ns) { Ls += le * c; ox += a[e].x * c; oy += a[e].y * c; }
        }
      }
      const float inv = Ls > 0.f ? 1.f / Ls : 0.f;
      *reinterpret_cast<uint32_t*>(p.att16 + (size_t)it.b * D + it.h * 64 + 2 * lane) = pack_half2(ox * inv, oy * inv);
    }
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int NT>
__global__ void __launch_bounds__(AD_THREADS, 1)
ar_decode_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                 const __grid_constant__ ArDecodeParams p) {
  extern __shared__ __align__(1024) uint8_t ad_smem[];
  uint8_t* smem = ad_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AD_SMEM_KV) + warp * AD_NST;   // this warp's stage barriers
  uint8_t* ring = smem + (warp < AD_AWARPS ? warp : 0) * (AD_NST * AD_STAGE_BYTES);
  if (lane == 0 && warp < AD_AWARPS) {
    for (int s = 0; s < AD_NST; ++s) mbar_init(bars + s, 1);
    fence_barrier_init();
  }
  if (tid == 0) { tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); }
  __syncthreads();
  unsigned epoch = 0;
  AttnWarpState ast;
  ast.issued = ast.consumed = 0;
  ast.prod.valid = false;
  uint8_t* wst = smem + warp * (AD_WST * 1024);

  // ---- P0: x = embed[last token] (nn.Embedding is not autocast: fp32 value of the fp16-exact weight) + sums of squares
  {
    const int tiles = p.ssq_tiles;   // ceil(D / 128)
    for (int tl = blockIdx.x; tl < tiles; tl += gridDim.x) {
      for (int b = warp; b < p.B; b += AD_WARPS) {
        const int tok = p.ids[(size_t)b * p.ids_stride + p.tok_len[b] - 1];
        const int c = tl * 128 + lane * 4;
        float2 a = make_float2(0.f, 0.f), bb = make_float2(0.f, 0.f);
        if (c < p.D) {
          const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p.embed + (size_t)tok * p.D + c));
          a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)); bb = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
          __stcg(reinterpret_cast<float4*>(p.x + (size_t)b * p.D + c), make_float4(a.x, a.y, bb.x, bb.y));
        }
        const float sq = warp_sum(a.x * a.x + a.y * a.y + bb.x * bb.x + bb.y * bb.y);
        if (lane == 0) __stcg(p.ssq + tl * 32 + b, sq);
      }
    }
  }
  {
    const ArLayerDev& l0 = p.layers[0];
    ArGemm g = p.g_qkv; g.W = l0.wqkv;
    if ((int)blockIdx.x < g.tiles * g.ksplit) w_prologue(g, gemm_item(g, blockIdx.x, warp, lane), wst, lane);
  }
  unsigned long long* pf = p.prof;   // advances by 2 stamps per barrier
  grid_sync(p.gbar, epoch, pf);
  if (pf) pf += 2;

  for (int layer = 0; layer < p.n_layers; ++layer) {
    const ArLayerDev& lw = p.layers[layer];
    ArGemm g;
    // ---- P1: qkv = Wqkv . rmsnorm(x)
    g = p.g_qkv; g.W = lw.wqkv;
    gemm_phase<NT, X_NORM, EPI_STORE>(p, g, lw.attn_norm, p.qkv, 3 * p.D, smem, (int)blockIdx.x < g.tiles * g.ksplit);
    __syncthreads();
    attn_prefetch(p, &tmap_k, &tmap_v, layer, ring, bars, ast);
    grid_sync(p.gbar, epoch, pf);
    if (pf) pf += 2;
    // ---- P2: attention over the cache + the new token
    attn_phase(p, lw, &tmap_k, &tmap_v, layer, ring, bars, ast);
    __syncthreads();   // the weight slots alias OTHER warps' KV rings: every warp of the CTA must be through with its tiles
    g = p.g_wo; g.W = lw.wo;
    const bool pre_wo = (int)blockIdx.x < g.tiles * g.ksplit;
    if (pre_wo) w_prologue(g, gemm_item(g, blockIdx.x, warp, lane), wst, lane);
    grid_sync(p.gbar, epoch, pf);
    if (pf) pf += 2;
    // ---- P3: x += Wo . attn
    gemm_phase<NT, X_ATTN, EPI_RESID>(p, g, nullptr, p.x, p.D, smem, pre_wo);
    g = p.g_w13; g.W = lw.w13;
    const bool pre_13 = (int)blockIdx.x < g.tiles * g.ksplit;
    if (pre_13) w_prologue(g, gemm_item(g, blockIdx.x, warp, lane), wst, lane);
    grid_sync(p.gbar, epoch, pf);
    if (pf) pf += 2;
    // ---- P4: g = silu(W1 . h) * (W3 . h), h = rmsnorm(x)
    gemm_phase<NT, X_NORM, EPI_SWIGLU>(p, g, lw.ffn_norm, nullptr, 0, smem, pre_13);
    g = p.g_w2; g.W = lw.w2;
    const bool pre_2 = (int)blockIdx.x < g.tiles * g.ksplit;
    if (pre_2) w_prologue(g, gemm_item(g, blockIdx.x, warp, lane), wst, lane);
    grid_sync(p.gbar, epoch, pf);
    if (pf) pf += 2;
    // ---- P5: x += W2 . g
    gemm_phase<NT, X_F16, EPI_RESID>(p, g, nullptr, p.x, p.D, smem, pre_2);
    if (layer + 1 < p.n_layers) { g = p.g_qkv; g.W = p.layers[layer + 1].wqkv; }
    else { g = p.g_out; }
    if ((int)blockIdx.x < g.tiles * g.ksplit) w_prologue(g, gemm_item(g, blockIdx.x, warp, lane), wst, lane);
    grid_sync(p.gbar, epoch, pf);
    if (pf) pf += 2;
  }
  // ---- logits = Wout . rmsnorm(x)
  gemm_phase<NT, X_NORM, EPI_STORE>(p, p.g_out, p.final_norm, p.logits, p.V, smem, (int)blockIdx.x < p.g_out.tiles * p.g_out.ksplit);
  if (pf && blockIdx.x == 0 && tid == 0) pf[0] = pf[1] = global_timer_ns();
  // ---- sampler (ar_generate.py:73-135): warp the logits of row b, draw the token, append it / stop the row
  if (p.fuse_sample) {
    grid_sync(p.gbar, epoch, nullptr);
    if ((int)blockIdx.x < p.B) ar_sample_row(p.sample, blockIdx.x, smem);
  }
}

// ------------------------------------------------------------------------------------------------ host
static void pick(ArGemm& g, int N, int K, int grid) {
  g.N = N; g.K = K; g.tiles = (N + AD_ROWS - 1) / AD_ROWS;
  const int kb = K / 64;
  int best = 1;
  double best_u = -1e9;
  for (int s = 1; s <= kb && s <= 16; ++s) {
    if (kb % s) continue;
    const int ksl = K / s;
    if (ksl < 128 || ksl > AD_MAX_KSLICE) continue;
    const int items = g.tiles * s;
    // SM utilisation of the phase minus a price per extra K slice (partial-sum traffic, one staging + ticket per item)
    const double u = (double)items / ((double)((items + grid - 1) / grid) * grid) - 0.02 * s;
    if (u > best_u) { best_u = u; best = s; }
  }
  if (best_u < -1e8) {   // K too small / too large for the limits above: smallest admissible split
    for (int s = 1; s <= kb; ++s) if (kb % s == 0 && K / s <= AD_MAX_KSLICE) { best = s; break; }
  }
  g.ksplit = best; g.kslice = K / best;
}

int ar_decode_plan(ArDecodeParams& p, int num_sms) {
  if (p.B <= 0 || p.B > 32 || p.D % 64 != 0 || p.F % 64 != 0 || p.D != p.H * 64) return M5_ERR_ARG;
  pick(p.g_qkv, 3 * p.D, p.D, num_sms);
  pick(p.g_wo, p.D, p.D, num_sms);
  pick(p.g_w13, 2 * p.F, p.D, num_sms);
  pick(p.g_w2, p.D, p.F, num_sms);
  pick(p.g_out, p.V, p.D, num_sms);
  if (p.g_wo.kslice % 64 != 0) return M5_ERR_ARG;   // whole heads per K slice (attention merge while staging)
  p.ssq_tiles = (p.D + 127) / 128;
  return M5_OK;
}
size_t ar_decode_scratch_floats(const ArDecodeParams& p) {
  int mx = 0;
  for (const ArGemm* g : {&p.g_qkv, &p.g_wo, &p.g_w13, &p.g_w2, &p.g_out}) mx = std::max(mx, g->tiles * g->ksplit);
  return (size_t)mx * 32 * AD_ROWS;
}
int ar_decode_max_tiles(const ArDecodeParams& p) {
  int mx = 0;
  for (const ArGemm* g : {&p.g_qkv, &p.g_wo, &p.g_w13, &p.g_w2, &p.g_out}) mx = std::max(mx, g->tiles);
  return mx;
}
// Keys per attention work item: one item per warp and (utterance, head) when B * H alone keeps most warps busy (no split
// partials, no merge); otherwise the context is cut so that about 0.7 * (warps of the grid) items exist, >= 256 keys each.
int ar_decode_split_keys(int B, int H, int max_kv, int num_sms) {
  const int warps = num_sms * AD_AWARPS;
  const int target = std::max(1, (int)(0.7 * warps) / std::max(1, B * H));
  int keys = (max_kv + target - 1) / target;
  keys = std::max(256, ((keys + AD_KT - 1) / AD_KT) * AD_KT);
  return keys;
}
int ar_decode_splits_for(int max_kv, int split_keys) { return (max_kv + split_keys - 1) / split_keys; }
size_t ar_decode_attn_floats(int B, int H, int n_split) { return (size_t)B * H * n_split * AD_PART; }

typedef CUresult (*AdEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int ad_tmap(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols) {
  static AdEncodeFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) != cudaSuccess || !q) return M5_ERR_CUDA;
    fn = reinterpret_cast<AdEncodeFn>(q);
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {64, AD_KT};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? M5_OK : M5_ERR_CUDA;
}

bool ar_decode_can_fuse_sampler(int V, int top_k) {
  static_assert(SP_THREADS == AD_THREADS, "the sampler phase runs on the decode kernel's CTA");
  return sample_smem_bytes(V, sample_cap(V, top_k)) <= (size_t)AD_SMEM_KV;
}

int ar_decode_launch(const ArDecodeParams& p, int num_sms, cudaStream_t stream) {
  CUtensorMap tk, tv;
  const uint64_t rows = (uint64_t)p.n_layers * p.B * p.Wc;
  if (ad_tmap(&tk, p.kc, rows, p.D) != M5_OK || ad_tmap(&tv, p.vc, rows, p.D) != M5_OK) return M5_ERR_CUDA;
  void (*kern)(CUtensorMap, CUtensorMap, ArDecodeParams) =
      p.B <= 8 ? ar_decode_kernel<1> : (p.B <= 16 ? ar_decode_kernel<2> : ar_decode_kernel<4>);
  static DeviceOnce once;
  unsigned long long bit;
  if (once.needed(bit)) {
    if (cudaFuncSetAttribute(ar_decode_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AD_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(ar_decode_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AD_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(ar_decode_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, AD_SMEM) != cudaSuccess)
      return M5_ERR_CUDA;
    once.done(bit);
  }
  if (cudaMemsetAsync(p.gbar, 0, sizeof(unsigned), stream) != cudaSuccess) return M5_ERR_CUDA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(num_sms); cfg.blockDim = dim3(AD_THREADS); cfg.dynamicSmemBytes = AD_SMEM; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the device-wide barrier cannot deadlock
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, tk, tv, p) == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

}  // namespace m5
