// Device body of the fused categorical sampler (see sampler.cu for the chain it implements, step by step).
#pragma once
#include "sampler.h"

#include "philox.cuh"
#include "ptx.cuh"

namespace m5 {

__device__ __forceinline__ uint32_t fkey(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t stream, uint32_t step, uint32_t idx) {
  uint32_t o[4];
  philox4x32((uint32_t)seed, (uint32_t)(seed >> 32), idx, step, (uint32_t)stream, (uint32_t)(stream >> 32), o);
  return ((o[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
}

static constexpr int SP_THREADS = 256;

// The whole chain for row b by one CTA of SP_THREADS threads; sp_smem: sample_smem_bytes(V, cap) bytes, 16-byte aligned.
// Shared by ar_sample_kernel (one CTA per row) and by the last phase of the persistent decode kernel (ar_decode.cu).
static __device__ __noinline__ void ar_sample_row(const SampleCall& p, const int b, uint8_t* sp_smem) {
  const int tid = threadIdx.x;
  const int V = p.V;
  float* z = reinterpret_cast<float*>(sp_smem);            // [V]
  float* sv = z + ((V + 3) & ~3);                          // [cap] sorted values
  int* si = reinterpret_cast<int*>(sv + p.cap);            // [cap] their vocabulary ids
  __shared__ int hist[256];
  __shared__ int s_cnt;
  __shared__ uint32_t s_prefix;
  __shared__ int s_krem;
  __shared__ float s_red[SP_THREADS / 32];
  __shared__ int s_redi[SP_THREADS / 32];
  __shared__ float s_bcast[2];
  __shared__ int s_ncut;

  if (p.done && p.done[b]) return;
  const int n_gen = p.n_gen[b];
  const int len = p.tok_len ? p.tok_len[b] : 0;
  const float* lg = p.logits + (size_t)b * p.ld_logits;
  const float NEG = -INFINITY;

  // ---- steps 1-4
  for (int i = tid; i < V; i += SP_THREADS) z[i] = lg[i];
  if (p.logits_dump && n_gen < p.dump_steps) {
    float* d = p.logits_dump + ((size_t)b * p.dump_steps + n_gen) * V;
    for (int i = tid; i < V; i += SP_THREADS) d[i] = lg[i];
  }
  __syncthreads();
  if (n_gen > 1) {
    // counts over the last `window` generated ids: the first occurrence (scanning back) applies the full penalty
    const int w = min(p.cfg.penalty_window, n_gen);
    const int* h = p.hist + (size_t)b * p.hist_stride + (p.hist_is_ids ? (len - w) : (n_gen - w));
    for (int i = tid; i < w; i += SP_THREADS) {
      const int id = h[i];
      bool first = true;
      int c = 0;
      for (int j = 0; j < w; ++j) {
        if (h[j] == id) {
          if (j < i) { first = false; break; }
          ++c;
        }
      }
      if (first) {
        float v = z[id];
        v = v - (float)c * p.cfg.alpha_frequency;
        v = v - p.cfg.alpha_presence;
        z[id] = v;
      }
    }
  }
  __syncthreads();
  const int mask_below = p.text_vocab - 1;
  const int eos = p.cfg.eos_id;
  const int est = p.n_phones ? p.n_phones[b] : -1;
  const bool force_mode = p.cfg.force_len > 0;
  float eos_mod = 0.f;
  if (est >= 0 && n_gen <= est) {
    const int pen = max(est - n_gen, 1);
    eos_mod = (float)((double)p.cfg.eos_penalty_factor * pow((double)pen, (double)p.cfg.eos_penalty_decay));
  }
  for (int i = tid; i < V; i += SP_THREADS) {
    float v = z[i];
    if (i < mask_below) v = NEG;
    if (i == eos) {
      if (est >= 0 && n_gen <= est) v -= eos_mod;
      if (force_mode && n_gen < p.cfg.force_len) v = NEG;
    }
    z[i] = v / p.cfg.temperature;
  }
  __syncthreads();

  // ---- step 5: k-th largest via radix select on order-preserving keys
  float kth = NEG;
  const int k = p.cfg.top_k > 0 ? min(max(p.cfg.top_k, 1), V) : 0;
  if (k > 0) {
    if (tid == 0) { s_prefix = 0; s_krem = k; }
    __syncthreads();
    for (int pass = 3; pass >= 0; --pass) {
      hist[tid] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t himask = pass == 3 ? 0u : (0xFFFFFFFFu << ((pass + 1) * 8));
      for (int i = tid; i < V; i += SP_THREADS) {
        const uint32_t key = fkey(z[i]);
        if ((key & himask) == (prefix & himask)) atomicAdd(&hist[(key >> (pass * 8)) & 255], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int rem = s_krem, bin = 255;
        for (; bin > 0; --bin) {
          if (hist[bin] >= rem) break;
          rem -= hist[bin];
        }
        s_krem = rem;
        s_prefix = prefix | ((uint32_t)bin << (pass * 8));
      }
      __syncthreads();
    }
    const uint32_t kk = s_prefix;
    kth = __uint_as_float((kk & 0x80000000u) ? (kk & 0x7FFFFFFFu) : ~kk);
  }
  // survivors: finite and >= kth
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  for (int i = tid; i < V; i += SP_THREADS) {
    const float v = z[i];
    if (v > NEG && !(v < kth)) {
      const int slot = atomicAdd(&s_cnt, 1);
      if (slot < p.cap) { sv[slot] = v; si[slot] = i; }
    }
  }
  __syncthreads();
  const int ns = min(s_cnt, p.cap);
  int np2 = 1;
  while (np2 < ns) np2 <<= 1;
  for (int i = ns + tid; i < np2; i += SP_THREADS) { sv[i] = NEG; si[i] = 0x7FFFFFFF; }
  __syncthreads();
  // ---- step 6: bitonic sort, descending by value (ties: lower id first, deterministic)
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < np2 / 2; i += SP_THREADS) {
        const int lo = (i / stride) * stride * 2 + (i % stride);
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const float a = sv[lo], c = sv[hi];
        const int ia = si[lo], ic = si[hi];
        const bool a_first = (a > c) || (a == c && ia < ic);
        if (a_first != desc) { sv[lo] = c; sv[hi] = a; si[lo] = ic; si[hi] = ia; }
      }
      __syncthreads();
    }
  }
  // softmax over sorted survivors + cumulative sum; find how many to keep
  const float vmax = ns > 0 ? sv[0] : 0.f;
  float part = 0.f;
  for (int i = tid; i < ns; i += SP_THREADS) part += expf(sv[i] - vmax);
  part = warp_sum(part);
  if ((tid & 31) == 0) s_red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < SP_THREADS / 32; ++w) s += s_red[w];
    s_bcast[0] = s;
    int keep = ns;
    if (p.cfg.top_p < 1.0f) {
      float cum = 0.f;
      keep = 0;
      for (int i = 0; i < ns; ++i) {
        // token i is removed iff i >= 1 and cumsum_{i-1} > top_p
        if (i >= 1 && cum > p.cfg.top_p) break;
        cum += expf(sv[i] - vmax) / s;
        keep = i + 1;
      }
    }
    s_ncut = keep;
  }
  __syncthreads();
  const int keep = s_ncut;
  // ---- step 6b: typical-p over the `keep` survivors (ar_generate.py:93)
  if (p.cfg.typical_p > 0.f && p.cfg.typical_p <= 0.999f && keep > 0) {
    float* ts = reinterpret_cast<float*>(si + p.cap);   // [cap] scores, later sorted ascending
    float* tp = ts + p.cap;                              // [cap] probabilities carried along
    // log_softmax over the survivors, entropy
    part = 0.f;
    for (int i = tid; i < keep; i += SP_THREADS) part += expf(sv[i] - vmax);
    part = warp_sum(part);
    if ((tid & 31) == 0) s_red[tid >> 5] = part;
    __syncthreads();
    float ssum = 0.f;
    for (int w = 0; w < SP_THREADS / 32; ++w) ssum += s_red[w];
    const float lse_t = logf(ssum);
    __syncthreads();
    float ent = 0.f;
    for (int i = tid; i < keep; i += SP_THREADS) {
      const float nl = (sv[i] - vmax) - lse_t;
      ent -= nl * expf(nl);
    }
    ent = warp_sum(ent);
    if ((tid & 31) == 0) s_red[tid >> 5] = ent;
    __syncthreads();
    float H = 0.f;
    for (int w = 0; w < SP_THREADS / 32; ++w) H += s_red[w];
    __syncthreads();
    int kp2 = 1;
    while (kp2 < keep) kp2 <<= 1;
    for (int i = tid; i < kp2; i += SP_THREADS) {
      if (i < keep) {
        const float nl = (sv[i] - vmax) - lse_t;
        ts[i] = fabsf(-nl - H);
        tp[i] = expf(nl);
      } else { ts[i] = INFINITY; tp[i] = 0.f; }
    }
    __syncthreads();
    for (int size = 2; size <= kp2; size <<= 1) {      // bitonic sort, ascending by score
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = tid; i < kp2 / 2; i += SP_THREADS) {
          const int lo = (i / stride) * stride * 2 + (i % stride);
          const int hi = lo + stride;
          const bool asc = ((lo & size) == 0);
          const float a = ts[lo], c = ts[hi];
          if ((a > c) == asc) { ts[lo] = c; ts[hi] = a; const float q = tp[lo]; tp[lo] = tp[hi]; tp[hi] = q; }
        }
        __syncthreads();
      }
    }
    if (tid == 0) {
      float cum = 0.f;
      int last = 0;
      for (int i = 0; i < keep; ++i) {   // last_ind = #(cumulative_probs < mass)
        cum += tp[i];
        if (cum < p.cfg.typical_p) last = i + 1;
      }
      s_bcast[0] = ts[min(last, keep - 1)];
    }
    __syncthreads();
    const float thr = s_bcast[0];
    for (int i = tid; i < keep; i += SP_THREADS) {
      const float nl = (sv[i] - vmax) - lse_t;
      if (fabsf(-nl - H) > thr) sv[i] = NEG;   // removed: exp(-inf) = 0 in the final softmax, never sampled
    }
    __syncthreads();
  }
  // ---- step 7: log_softmax over the kept set, sample argmax p_i / e_i
  part = 0.f;
  for (int i = tid; i < keep; i += SP_THREADS) part += expf(sv[i] - vmax);
  part = warp_sum(part);
  __syncthreads();
  if ((tid & 31) == 0) s_red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < SP_THREADS / 32; ++w) s += s_red[w];
    s_bcast[1] = logf(s);
  }
  __syncthreads();
  const float lse = s_bcast[1];
  if (p.out_logprobs) {
    float* o = p.out_logprobs + (size_t)b * V;
    for (int i = tid; i < V; i += SP_THREADS) o[i] = NEG;
    __syncthreads();
    for (int i = tid; i < keep; i += SP_THREADS) o[si[i]] = (sv[i] - vmax) - lse;
  }
  const uint64_t utt = p.utt_ids ? (uint64_t)p.utt_ids[b] : (uint64_t)b;
  float best = -1.f;
  int best_id = 0x7FFFFFFF;
  for (int i = tid; i < keep; i += SP_THREADS) {
    const int id = si[i];
    const float pr = expf((sv[i] - vmax) - lse);
    float e;
    if (p.noise) e = p.noise[((size_t)b * p.noise_steps + min(n_gen, p.noise_steps - 1)) * V + id];
    else e = -logf(philox_uniform(p.seed, utt, (uint32_t)n_gen, (uint32_t)id));
    const float q = pr / e;
    if (q > best || (q == best && id < best_id)) { best = q; best_id = id; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_id, o);
    if (ob > best || (ob == best && oi < best_id)) { best = ob; best_id = oi; }
  }
  if ((tid & 31) == 0) { s_red[tid >> 5] = best; s_redi[tid >> 5] = best_id; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < SP_THREADS / 32; ++w)
      if (s_red[w] > best || (s_red[w] == best && s_redi[w] < best_id)) { best = s_red[w]; best_id = s_redi[w]; }
    int tok = best_id;
    if (force_mode && n_gen >= p.cfg.force_len) tok = eos;
    if (p.out_tok) p.out_tok[b] = tok;
    if (p.ids) {
      // loop mode: append or stop
      bool stop = (tok == eos);
      if (!stop) {
        p.ids[(size_t)b * p.hist_stride + len] = tok;
        p.tok_len[b] = len + 1;
        p.kv_len[b] = len + 2;  // cache positions after the next decode step appends this token
        p.n_gen[b] = n_gen + 1;
        if (len + 1 >= p.cfg.max_len) stop = true;
      }
      if (stop) {
        p.done[b] = 1;
        atomicAdd(p.n_done, 1);
      }
    }
  }
}

}  // namespace m5
