// Attention kernels of the hot path (head_dim = 64 everywhere: nn_future.py:149, model.py nhead=16 x 64).
//
//  flash_attn_kernel : packed variable-length softmax(QK^T/8)V with online softmax; fp16 operands on the
//                      mma.sync tensor path, fp32 softmax state.  Serves AR prefill (causal, nn_future.py:254-272),
//                      the AR/NAR speaker encoders, the NAR encoder / decoder self-attention and the NAR
//                      cross-attention (model.py:339-341 -> nn.MultiheadAttention).
//  decode_attn_*     : single-query attention over the fp16 KV cache (AR decode step, nn_future.py:257-272),
//                      split over the key axis so that B*H*n_split CTAs stream the cache at HBM rate.
#include "m5_internal.h"
#include "ptx.cuh"

namespace m5 {

static constexpr int FA_BQ = 64;   // queries per CTA (16 per warp)
static constexpr int FA_BK = 64;   // keys per tile
static constexpr int FA_THREADS = 128;
static constexpr int HD = 64;

// smem tile [rows][64] fp16, 16-byte chunks XOR-swizzled by (row & 7) -> conflict-free ldmatrix.
__device__ __forceinline__ int sw_off(int row, int chunk) { return row * HD + ((chunk ^ (row & 7)) << 3); }

__device__ __forceinline__ void load_tile_async(__half* dst, const __half* src, int ld, int rows_valid, int tid) {
  // 64 rows x 8 chunks = 512 chunks, 4 per thread
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = tid + i * FA_THREADS;
    const int row = c >> 3, chunk = c & 7;
    const bool ok = row < rows_valid;
    cp_async16(dst + sw_off(row, chunk), ok ? (src + (size_t)row * ld + chunk * 8) : src, ok);
  }
}

// PREC: every fp16 operand carries a low half (x = hi + lo): S = Qhi Khi + Qlo Khi + Qhi Klo, O = Phi Vhi + Plo Vhi + Phi Vlo
// (three MMAs per product, fp32-class accuracy); the output is written as (hi, lo) as well.
template <bool PREC>
__global__ void __launch_bounds__(FA_THREADS)
flash_attn_kernel(AttnCall p) {
  const int seq = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
  const int q_len = p.q_len[seq], k_len = p.k_len[seq];
  const int q0 = qt * FA_BQ;
  if (q0 >= q_len) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  extern __shared__ __align__(128) uint8_t fa_smem[];
  constexpr int TILE = FA_BQ * HD;  // halves per 64x64 tile
  __half* sQ = reinterpret_cast<__half*>(fa_smem);
  __half* sK = sQ + TILE;            // [2][TILE]
  __half* sV = sK + 2 * TILE;        // [2][TILE]
  __half* sQl = sV + 2 * TILE;       // PREC only
  __half* sKl = sQl + TILE;
  __half* sVl = sKl + 2 * TILE;

  const size_t qoff = (size_t)(p.q_start[seq] + q0) * p.ldq + head * HD;
  const size_t koff = (size_t)p.k_start[seq] * p.ldk + head * HD;
  const size_t voff = (size_t)p.k_start[seq] * p.ldv + head * HD;
  const int q_valid = min(FA_BQ, q_len - q0);
  const int causal_off = k_len - q_len;
  int k_end = k_len;  // keys this CTA must visit
  if (p.causal) k_end = min(k_len, q0 + q_valid + causal_off);
  const int n_tiles = (k_end + FA_BK - 1) / FA_BK;

  auto load_kv = [&](int buf, int kn) {
    const int rows = min(FA_BK, k_end - kn);
    load_tile_async(sK + buf * TILE, p.K + koff + (size_t)kn * p.ldk, p.ldk, rows, tid);
    load_tile_async(sV + buf * TILE, p.V + voff + (size_t)kn * p.ldv, p.ldv, rows, tid);
    if constexpr (PREC) {
      load_tile_async(sKl + buf * TILE, p.Klo + koff + (size_t)kn * p.ldk, p.ldk, rows, tid);
      load_tile_async(sVl + buf * TILE, p.Vlo + voff + (size_t)kn * p.ldv, p.ldv, rows, tid);
    }
  };
  load_tile_async(sQ, p.Q + qoff, p.ldq, q_valid, tid);
  if constexpr (PREC) load_tile_async(sQl, p.Qlo + qoff, p.ldq, q_valid, tid);
  if (n_tiles > 0) load_kv(0, 0);
  cp_async_commit();

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
  const float sl2 = p.scale * 1.4426950408889634f;
  uint32_t qf[4][4], qfl[PREC ? 4 : 1][4];

  for (int kt = 0; kt < n_tiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_tiles) {
      load_kv(buf ^ 1, (kt + 1) * FA_BK);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
      // Q fragments: rows warp*16 .. +15, 4 k-steps of 16
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int row = warp * 16 + (lane & 15);
        const int chunk = kk * 2 + (lane >> 4);
        ldmatrix_x4(qf[kk], sQ + sw_off(row, chunk));
        if constexpr (PREC) ldmatrix_x4(qfl[kk], sQl + sw_off(row, chunk));
      }
    }
    // ---- S = Q K^T (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // pairs of 8-key n-tiles
        uint32_t b[4];
        const int row = jp * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int chunk = kk * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(b, sK + buf * TILE + sw_off(row, chunk));
        mma_16816(s[2 * jp], qf[kk], b[0], b[1]);
        mma_16816(s[2 * jp + 1], qf[kk], b[2], b[3]);
        if constexpr (PREC) {
          mma_16816(s[2 * jp], qfl[kk], b[0], b[1]);
          mma_16816(s[2 * jp + 1], qfl[kk], b[2], b[3]);
          uint32_t bl[4];
          ldmatrix_x4(bl, sKl + buf * TILE + sw_off(row, chunk));
          mma_16816(s[2 * jp], qf[kk], bl[0], bl[1]);
          mma_16816(s[2 * jp + 1], qf[kk], bl[2], bl[3]);
        }
      }
    }
    // ---- mask + online softmax
    const int kbase = kt * FA_BK;
    const int qrow0 = q0 + warp * 16 + g;  // row for regs 0,1 ; +8 for regs 2,3
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kbase + j * 8 + 2 * t + (e & 1);
        const int qr = qrow0 + ((e >> 1) << 3);
        bool ok = key < k_len;
        if (p.causal) ok = ok && (key <= qr + causal_off);
        const float v = ok ? s[j][e] * sl2 : -INFINITY;
        s[j][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mnew[r] = fmaxf(m_i[r], mx[r]);
      const float msafe = (mnew[r] == -INFINITY) ? 0.f : mnew[r];
      corr[r] = exp2f(m_i[r] - msafe);  // m_i = -inf -> 0
      m_i[r] = mnew[r];
      mnew[r] = msafe;
      l_i[r] *= corr[r];
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f(s[j][e] - mnew[e >> 1]);
        s[j][e] = pv;
        rs[e >> 1] += pv;
      }
      o[j][0] *= corr[0]; o[j][1] *= corr[0]; o[j][2] *= corr[1]; o[j][3] *= corr[1];
    }
    l_i[0] += rs[0];
    l_i[1] += rs[1];
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 16 keys per step
      uint32_t a[4], al[4];
      a[0] = pack_half2(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_half2(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_half2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_half2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      if constexpr (PREC) {
        auto lo2 = [](float x, float y, uint32_t hi) {
          const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
          return pack_half2(x - h.x, y - h.y);
        };
        al[0] = lo2(s[2 * kk][0], s[2 * kk][1], a[0]);
        al[1] = lo2(s[2 * kk][2], s[2 * kk][3], a[1]);
        al[2] = lo2(s[2 * kk + 1][0], s[2 * kk + 1][1], a[2]);
        al[3] = lo2(s[2 * kk + 1][2], s[2 * kk + 1][3], a[3]);
      }
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // pairs of 8-wide output column tiles
        uint32_t b[4];
        const int row = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int chunk = jp * 2 + (lane >> 4);
        ldmatrix_x4_trans(b, sV + buf * TILE + sw_off(row, chunk));
        mma_16816(o[2 * jp], a, b[0], b[1]);
        mma_16816(o[2 * jp + 1], a, b[2], b[3]);
        if constexpr (PREC) {
          mma_16816(o[2 * jp], al, b[0], b[1]);
          mma_16816(o[2 * jp + 1], al, b[2], b[3]);
          uint32_t bl[4];
          ldmatrix_x4_trans(bl, sVl + buf * TILE + sw_off(row, chunk));
          mma_16816(o[2 * jp], a, bl[0], bl[1]);
          mma_16816(o[2 * jp + 1], a, bl[2], bl[3]);
        }
      }
    }
    __syncthreads();
  }
  // ---- finalize
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
  }
  const float inv0 = l_i[0] > 0.f ? 1.f / l_i[0] : 0.f;
  const float inv1 = l_i[1] > 0.f ? 1.f / l_i[1] : 0.f;
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  const size_t ooff = (size_t)(p.q_start[seq] + q0) * p.ldo + head * HD;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = j * 8 + 2 * t;
    const float v00 = o[j][0] * inv0, v01 = o[j][1] * inv0, v10 = o[j][2] * inv1, v11 = o[j][3] * inv1;
    const uint32_t h0 = pack_half2(v00, v01), h1 = pack_half2(v10, v11);
    if (r0 < q_valid) *reinterpret_cast<uint32_t*>(p.O + ooff + (size_t)r0 * p.ldo + col) = h0;
    if (r1 < q_valid) *reinterpret_cast<uint32_t*>(p.O + ooff + (size_t)r1 * p.ldo + col) = h1;
    if constexpr (PREC) {
      const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
      if (r0 < q_valid) *reinterpret_cast<uint32_t*>(p.Olo + ooff + (size_t)r0 * p.ldo + col) = pack_half2(v00 - f0.x, v01 - f0.y);
      if (r1 < q_valid) *reinterpret_cast<uint32_t*>(p.Olo + ooff + (size_t)r1 * p.ldo + col) = pack_half2(v10 - f1.x, v11 - f1.y);
    }
  }
}

int flash_attn(const AttnCall& c, cudaStream_t stream) {
  if (c.n_seqs <= 0 || c.max_q <= 0) return M5_OK;
  if ((c.ldq | c.ldk | c.ldv) % 8 != 0 || c.ldo % 2 != 0) return M5_ERR_ARG;
  dim3 grid((c.max_q + FA_BQ - 1) / FA_BQ, c.n_heads, c.n_seqs);
  const bool prec = c.Qlo != nullptr;
  if (prec && (!c.Klo || !c.Vlo || !c.Olo)) return M5_ERR_ARG;
  const size_t smem = (size_t)(prec ? 10 : 5) * FA_BQ * HD * sizeof(__half);
  static DeviceOnce once;
  unsigned long long bit;
  if (once.needed(bit)) {
    cudaFuncSetAttribute(flash_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 10 * FA_BQ * HD * 2);
    once.done(bit);
  }
  if (prec) flash_attn_kernel<true><<<grid, FA_THREADS, smem, stream>>>(c);
  else flash_attn_kernel<false><<<grid, FA_THREADS, smem, stream>>>(c);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// Decode attention: one query per (row b, head h).  Grid (n_split, H, B); split s owns keys [s*DA_KEYS, (s+1)*DA_KEYS)
// (CTAs beyond the row's length exit at once, so the grid can be sized for the longest possible context and captured
// in a CUDA graph).  8 lanes x 16-byte loads cover one 128-byte K (or V) row, a warp covers 4 keys per instruction and
// keeps 4 such instructions in flight for K and for V; partial (max, sum, acc[64]) per split are merged by a second
// kernel in a fixed order (deterministic).
static constexpr int DA_THREADS = 128;
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {  // the KV cache is read once per step: keep it out of L1
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
static constexpr int DA_KEYS = 256;   // keys per CTA

__global__ void __launch_bounds__(DA_THREADS)
decode_attn_split_kernel(DecodeAttnCall p) {
  pdl_launch_dependents();
  pdl_wait();
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  if (p.done && p.done[b]) return;
  const int L = p.kv_len[b];
  const int k0 = split * DA_KEYS, k1 = min(L, k0 + DA_KEYS);
  if (k0 >= L) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = lane & 7, grp = lane >> 3;  // 8 lanes per key row, 4 keys per warp instruction
  const int D = p.H * HD;
  float q[8];
  {
    const uint4 qv = *reinterpret_cast<const uint4*>(p.q + (size_t)b * D + h * HD + sub * 8);
    const __half2* qh = reinterpret_cast<const __half2*>(&qv);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(qh[i]); q[2 * i] = f.x; q[2 * i + 1] = f.y; }
  }
  const float sl2 = 0.125f * 1.4426950408889634f;
  const __half* kb = p.kc + ((size_t)b * p.W) * D + h * HD + sub * 8;
  const __half* vb = p.vc + ((size_t)b * p.W) * D + h * HD + sub * 8;
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  // each warp iteration covers 16 keys: key = kbase + u*4 + grp
  for (int kbase = k0 + warp * 16; kbase < k1; kbase += 4 * 16) {
    uint4 kv[4], vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int key = kbase + u * 4 + grp;
      if (key < k1) {
        kv[u] = ldg_stream16(kb + (size_t)key * D);
        vv[u] = ldg_stream16(vb + (size_t)key * D);
      } else {
        kv[u] = make_uint4(0, 0, 0, 0);
        vv[u] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int key = kbase + u * 4 + grp;
      const __half2* kh = reinterpret_cast<const __half2*>(&kv[u]);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(kh[i]); s += q[2 * i] * f.x + q[2 * i + 1] * f.y; }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (key < k1) {
        s *= sl2;
        const float mn = fmaxf(m, s);
        const float c = exp2f(m - mn), pe = exp2f(s - mn);
        l = l * c + pe;
        const __half2* vh = reinterpret_cast<const __half2*>(&vv[u]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __half22float2(vh[i]);
          acc[2 * i] = acc[2 * i] * c + pe * f.x;
          acc[2 * i + 1] = acc[2 * i + 1] * c + pe * f.y;
        }
        m = mn;
      }
    }
  }
  // merge the 4 key groups of the warp, then the 4 warps through smem
  __shared__ float sm[16], sl[16], sa[16][HD];
  const int slot = warp * 4 + grp;
  if (sub == 0) { sm[slot] = m; sl[slot] = l; }
#pragma unroll
  for (int i = 0; i < 8; ++i) sa[slot][sub * 8 + i] = acc[i];
  __syncthreads();
  if (tid < HD) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 16; ++w) M = fmaxf(M, sm[w]);
    float Ls = 0.f, o = 0.f;
    if (M > -INFINITY) {
#pragma unroll
      for (int w = 0; w < 16; ++w) {
        const float c = exp2f(sm[w] - M);
        Ls += sl[w] * c;
        o += sa[w][tid] * c;
      }
    }
    float* sp = p.scratch + ((size_t)(b * p.H + h) * p.n_split + split) * (HD + 2);
    if (tid == 0) { sp[0] = M; sp[1] = Ls; }
    sp[2 + tid] = o;
  }
}

__global__ void decode_attn_merge_kernel(DecodeAttnCall p) {
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;  // 32 threads
  if (p.done && p.done[b]) return;
  const float* sp = p.scratch + ((size_t)(b * p.H + h) * p.n_split) * (HD + 2);
  const int ns = min(p.n_split, (p.kv_len[b] + DA_KEYS - 1) / DA_KEYS);  // splits that own at least one key
  float M = -INFINITY;
  for (int s = 0; s < ns; ++s) M = fmaxf(M, sp[s * (HD + 2)]);
  float L = 0.f, ox = 0.f, oy = 0.f;
  for (int s = 0; s < ns; ++s) {
    const float* q = sp + s * (HD + 2);
    if (q[0] == -INFINITY) continue;
    const float c = exp2f(q[0] - M);
    L += q[1] * c;
    ox += q[2 + 2 * lane] * c;
    oy += q[2 + 2 * lane + 1] * c;
  }
  const float inv = L > 0.f ? 1.f / L : 0.f;
  *reinterpret_cast<__half2*>(p.out + (size_t)b * p.H * HD + h * HD + 2 * lane) = __floats2half2_rn(ox * inv, oy * inv);
}

size_t decode_attn_scratch_bytes(int B, int H, int n_split) { return (size_t)B * H * n_split * (HD + 2) * sizeof(float); }
int decode_attn_splits_for(int max_kv) { return (max_kv + DA_KEYS - 1) / DA_KEYS; }

int decode_attn(const DecodeAttnCall& c, cudaStream_t stream) {
  if (c.B <= 0) return M5_OK;
  dim3 grid(c.n_split, c.H, c.B);
  if (launch_k(decode_attn_split_kernel, grid, dim3(DA_THREADS), 0, stream, c) != cudaSuccess) return M5_ERR_CUDA;
  dim3 g2(c.H, c.B);
  return launch_k(decode_attn_merge_kernel, g2, dim3(32), 0, stream, c) == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

}  // namespace m5
