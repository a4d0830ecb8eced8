// Fused multinomial-diffusion posterior + Gumbel-argmax sampling (mars5/diffuser.py:359-393), one warp per
// (sequence position, codebook).  For the K = 1025 classes of one categorical:
//
//   l   = (w*cond + (1-w)*uncond) / temp                               classifier-free guidance, diffuser.py:361-366
//   a   = log_softmax(l)                                               diffuser.py:367
//   A   = t>0 ? logaddexp(a + log_cumprod_alpha[t-1], log_1_min_cumprod_alpha[t-1] - ln K) : a    q_pred, :161-174,190-197
//   Bq  = logaddexp(log_onehot_eps(x_t) + log_alpha[t], log_1_min_alpha[t] - ln K)                q_pred_one_timestep, :118-134
//   p   = (A + Bq) - logsumexp(A + Bq)                                 q_posterior, :201-206
//   x   = argmax_k ( p_k - log(-log(clamp(u_k,1e-7)).clamp(1e-7)) )    log_sample_categorical, :219-228
//
// The logits are read once (128-bit loads); nothing of size K is written back -- only the sampled code.
// nar_renoise handles the known region: q_sample(x_known, t) (diffuser.py:230-236,386-390), the mask merge (:393) and
// the clean-L0 override (:467-468).
#include <math.h>

#include "philox.cuh"
#include "ptx.cuh"
#include "sampler.h"

namespace m5 {


__device__ __forceinline__ float log_add_exp(float a, float b) {
  const float m = fmaxf(a, b);
  return m + logf(expf(a - m) + expf(b - m));
}
__device__ __forceinline__ float gumbel_from_u(float u) {
  u = fmaxf(u, 1e-7f);
  return -logf(fmaxf(-logf(u), 1e-7f));
}
__device__ __forceinline__ void uniforms4(uint64_t seed, uint64_t utt, uint32_t pos, uint32_t tag, uint32_t grp, float out[4]) {
  uint32_t o[4];
  // key = seed, counter = (group, tag, position, utterance): (seed 1, utt 0) and (seed 0, utt 1) are different streams
  philox4x32((uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x4d415253u, grp, tag, pos, (uint32_t)utt ^ ((uint32_t)(utt >> 32) * 0x9E3779B9u), o);
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = (o[i] >> 8) * (1.0f / 16777216.0f);  // [0,1) like torch.rand
}

static constexpr int KMAX_GROUPS = 9;  // ceil(ceil(1025/4)/32)

__global__ void __launch_bounds__(256) nar_posterior_kernel(PosteriorCall p, float log_eps, float ln_k) {
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int n_items = p.q < 0 ? p.R * p.Q : p.R;
  if (warp_global >= n_items) return;
  int lrow, q;
  if (p.q < 0) { lrow = warp_global / p.Q; q = warp_global - lrow * p.Q; }
  else { lrow = warp_global; q = p.q; }
  const size_t loff = (p.q < 0 ? (size_t)warp_global : (size_t)lrow) * p.ld;
  const float* pc = p.cond + loff;
  const float* pu = p.uncond ? p.uncond + loff : nullptr;
  const int xrow = p.row_map ? p.row_map[lrow] : lrow;
  const int xt = p.x_t[(size_t)xrow * p.Q + q];
  const int K = p.K;
  const int groups = (K + 3) >> 2;
  const bool vec_ok = (p.ld & 3) == 0;
  const float w = p.guidance_w, w1 = 1.0f - p.guidance_w;

  float l[KMAX_GROUPS][4];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < KMAX_GROUPS; ++i) {
    const int gidx = lane + 32 * i;
    const int k0 = gidx * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) l[i][e] = -INFINITY;
    if (gidx < groups) {
      float c[4], u[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec_ok && k0 + 3 < K) {
        const float4 cv = *reinterpret_cast<const float4*>(pc + k0);
        c[0] = cv.x; c[1] = cv.y; c[2] = cv.z; c[3] = cv.w;
        if (pu) { const float4 uv = *reinterpret_cast<const float4*>(pu + k0); u[0] = uv.x; u[1] = uv.y; u[2] = uv.z; u[3] = uv.w; }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          c[e] = (k0 + e < K) ? pc[k0 + e] : 0.f;
          if (pu) u[e] = (k0 + e < K) ? pu[k0 + e] : 0.f;
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (k0 + e < K) {
          float v = c[e];
          if (pu) v = w * c[e] + w1 * u[e];
          v = v / p.x0_temp;
          l[i][e] = v;
          mx = fmaxf(mx, v);
        }
      }
    }
  }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < KMAX_GROUPS; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) s += (l[i][e] > -INFINITY) ? expf(l[i][e] - mx) : 0.f;
  s = warp_sum(s);
  const float lse = logf(s);
  const float c1 = p.log_1m_cum_tm1 - ln_k;
  const float c2 = p.log_1m_alpha_t - ln_k;
  const float b_same = log_add_exp(0.0f + p.log_alpha_t, c2);
  const float b_diff = log_add_exp(log_eps + p.log_alpha_t, c2);
  float umax = -INFINITY;
#pragma unroll
  for (int i = 0; i < KMAX_GROUPS; ++i) {
    const int k0 = (lane + 32 * i) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (k0 + e < K && l[i][e] > -INFINITY) {
        const float a = (l[i][e] - mx) - lse;
        const float A = p.t > 0 ? log_add_exp(a + p.log_cum_tm1, c1) : a;
        const float v = A + ((k0 + e == xt) ? b_same : b_diff);
        l[i][e] = v;
        umax = fmaxf(umax, v);
      }
    }
  }
  umax = warp_max(umax);
  float s2 = 0.f;
#pragma unroll
  for (int i = 0; i < KMAX_GROUPS; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) s2 += (l[i][e] > -INFINITY) ? expf(l[i][e] - umax) : 0.f;
  s2 = warp_sum(s2);
  const float lse2 = umax + logf(s2);
  // ---- Gumbel-argmax
  const uint64_t utt = p.row_utt ? (uint64_t)p.row_utt[xrow] : 0ull;
  const uint32_t pos = p.row_pos ? (uint32_t)p.row_pos[xrow] : (uint32_t)xrow;
  const uint32_t tag = ((uint32_t)p.t << 8) | ((uint32_t)p.draw << 4) | (uint32_t)q;
  const float* un = p.u ? p.u + ((size_t)(p.u_rows_are_x ? xrow : lrow) * p.Q + q) * K : nullptr;
  float best = -INFINITY;
  int best_k = 0x7FFFFFFF;
#pragma unroll
  for (int i = 0; i < KMAX_GROUPS; ++i) {
    const int gidx = lane + 32 * i;
    const int k0 = gidx * 4;
    if (gidx < groups) {
      float u4[4];
      if (un) {
#pragma unroll
        for (int e = 0; e < 4; ++e) u4[e] = (k0 + e < K) ? un[k0 + e] : 0.5f;
      } else {
        uniforms4(p.seed, utt, pos, tag, (uint32_t)gidx, u4);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (k0 + e < K) {
          const float v = gumbel_from_u(u4[e]) + (l[i][e] - lse2);
          if (v > best || (v == best && k0 + e < best_k)) { best = v; best_k = k0 + e; }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
    if (ob > best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
  }
  if (lane == 0) p.x_out[(size_t)xrow * p.Q + q] = best_k;
}

int nar_posterior(const PosteriorCall& c, cudaStream_t stream) {
  const int n_items = c.q < 0 ? c.R * c.Q : c.R;
  if (n_items <= 0) return M5_OK;
  if (c.K > KMAX_GROUPS * 128) return M5_ERR_ARG;
  const int wpb = 8;
  nar_posterior_kernel<<<(n_items + wpb - 1) / wpb, wpb * 32, 0, stream>>>(c, logf(1e-7f), (float)log((double)c.K));
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------ known region
__global__ void __launch_bounds__(256) nar_renoise_kernel(RenoiseCall p, float log_eps, float ln_k) {
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (item >= p.R * p.Q) return;
  const int row = item / p.Q, q = item - row * p.Q;
  const bool known = p.forward || p.known[item] != 0;
  int result;
  if (!known) {
    result = p.x[item];
  } else if (p.t == 0 && !p.forward) {
    result = p.x_known[item];
  } else {
    const int xk = p.x_known[item];
    const int K = p.K;
    const float c1 = p.log_1m_cum_t - ln_k;
    const float v_same = log_add_exp(0.0f + p.log_cum_t, c1);
    const float v_diff = log_add_exp(log_eps + p.log_cum_t, c1);
    const uint64_t utt = p.row_utt ? (uint64_t)p.row_utt[row] : 0ull;
    const uint32_t pos = p.row_pos ? (uint32_t)p.row_pos[row] : (uint32_t)row;
    const uint32_t tag = ((uint32_t)p.t << 8) | ((p.forward ? 2u : 1u) << 4) | (uint32_t)q;
    const float* un = p.u ? p.u + (size_t)item * K : nullptr;
    float best = -INFINITY;
    int best_k = 0x7FFFFFFF;
    const int groups = (K + 3) >> 2;
    for (int gidx = lane; gidx < groups; gidx += 32) {
      const int k0 = gidx * 4;
      float u4[4];
      if (un) {
#pragma unroll
        for (int e = 0; e < 4; ++e) u4[e] = (k0 + e < K) ? un[k0 + e] : 0.5f;
      } else {
        uniforms4(p.seed, utt, pos, tag, (uint32_t)gidx, u4);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (k0 + e < K) {
          const float v = gumbel_from_u(u4[e]) + ((k0 + e == xk) ? v_same : v_diff);
          if (v > best || (v == best && k0 + e < best_k)) { best = v; best_k = k0 + e; }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
      if (ob > best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
    }
    result = best_k;
  }
  if (q == 0 && p.q0_override) result = p.x_q0[row];
  if (lane == 0) p.x[item] = result;
}

int nar_renoise(const RenoiseCall& c, cudaStream_t stream) {
  const int n = c.R * c.Q;
  if (n <= 0) return M5_OK;
  const int wpb = 8;
  nar_renoise_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, stream>>>(c, logf(1e-7f), (float)log((double)c.K));
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

}  // namespace m5
