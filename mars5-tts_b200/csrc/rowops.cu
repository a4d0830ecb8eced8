// Row-wise (HBM-bound) kernels: LayerNorm / RMSNorm producing the fp16 GEMM operand, RoPE + KV-cache append,
// embedding gathers with sinusoidal position terms.  One warp per row, 128-bit loads, fp32 statistics.
#include "m5_internal.h"
#include "ptx.cuh"
#include <cuda_fp8.h>

#include "rowops.h"

namespace m5 {

// ------------------------------------------------------------------------------------------------ norms
// RMSNorm: nn_future.py:301-312 (x * rsqrt(mean(x^2) + eps) * weight, fp32).  LayerNorm: torch F.layer_norm
// (biased variance, eps inside the sqrt), used with eps 4e-5 (model.py:13) and 1e-5 (heads, model.py:237).
__global__ void __launch_bounds__(256) norm_rows_kernel(NormCall p) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.M) return;
  const int src = p.row_map ? p.row_map[row] : row;
  const float* x = p.x + (size_t)src * p.ldx;
  const int D = p.D;
  float s = 0.f, ss = 0.f;
  for (int i = lane * 4; i < D; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    s += v.x + v.y + v.z + v.w;
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  float mean = 0.f, rstd;
  if (p.rms) {
    rstd = rsqrtf(ss / D + p.eps);
  } else {
    mean = s / D;
    // two-pass variance for accuracy (row is L1/L2 resident)
    float vs = 0.f;
    for (int i = lane * 4; i < D; i += 128) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
      vs += a * a + b * b + c * c + d * d;
    }
    vs = warp_sum(vs);
    rstd = rsqrtf(vs / D + p.eps);
  }
  for (int i = lane * 4; i < D; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    float y[4] = {(v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd};
    if (p.gamma) {
      const float4 gm = *reinterpret_cast<const float4*>(p.gamma + i);
      y[0] *= gm.x; y[1] *= gm.y; y[2] *= gm.z; y[3] *= gm.w;
    }
    if (p.beta) {
      const float4 bt = *reinterpret_cast<const float4*>(p.beta + i);
      y[0] += bt.x; y[1] += bt.y; y[2] += bt.z; y[3] += bt.w;
    }
    if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + (size_t)row * p.ldo + i) = make_float4(y[0], y[1], y[2], y[3]);
    if (p.out) {
      const __half h0 = __float2half_rn(y[0]), h1 = __float2half_rn(y[1]), h2 = __float2half_rn(y[2]),
                   h3 = __float2half_rn(y[3]);
      __half2* o = reinterpret_cast<__half2*>(p.out + (size_t)row * p.ldo + i);
      o[0] = __halves2half2(h0, h1);
      o[1] = __halves2half2(h2, h3);
      if (p.out_lo) {
        __half2* ol = reinterpret_cast<__half2*>(p.out_lo + (size_t)row * p.ldo + i);
        ol[0] = __floats2half2_rn(y[0] - __half2float(h0), y[1] - __half2float(h1));
        ol[1] = __floats2half2_rn(y[2] - __half2float(h2), y[3] - __half2float(h3));
      }
      if (p.out_lo8) {
        const __half hh[4] = {h0, h1, h2, h3};
        for (int e = 0; e < 4; ++e)
          p.out_lo8[(size_t)row * p.ldo8 + i + e] = (uint8_t)__nv_cvt_float_to_fp8((y[e] - __half2float(hh[e])) * 0.25f, __NV_SATFINITE, __NV_E5M2);
      }
    }
  }
}

// Register-resident variant for the row widths of the hot path (D = 128 * NV: 1024 NAR, 1536 AR, 384 vocoder): the row is
// loaded ONCE (NV float4 per lane, all loads in flight together), statistics and outputs come from registers.  Element
// assignment and summation order equal norm_rows_kernel's (lane-strided float4, warp butterfly), so results are bit-identical.
template <int NV>
__global__ void __launch_bounds__(256) norm_rows_reg_kernel(NormCall p) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.M) return;
  const int src = p.row_map ? p.row_map[row] : row;
  const float* x = p.x + (size_t)src * p.ldx;
  constexpr int D = 128 * NV;
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = *reinterpret_cast<const float4*>(x + lane * 4 + 128 * j);
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    s += v[j].x + v[j].y + v[j].z + v[j].w;
    ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  float mean = 0.f, rstd;
  if (p.rms) {
    rstd = rsqrtf(ss / D + p.eps);
  } else {
    mean = s / D;
    float vs = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      vs += a * a + b * b + c * c + d * d;
    }
    vs = warp_sum(vs);
    rstd = rsqrtf(vs / D + p.eps);
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane * 4 + 128 * j;
    float y[4] = {(v[j].x - mean) * rstd, (v[j].y - mean) * rstd, (v[j].z - mean) * rstd, (v[j].w - mean) * rstd};
    if (p.gamma) {
      const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma + i));
      y[0] *= gm.x; y[1] *= gm.y; y[2] *= gm.z; y[3] *= gm.w;
    }
    if (p.beta) {
      const float4 bt = __ldg(reinterpret_cast<const float4*>(p.beta + i));
      y[0] += bt.x; y[1] += bt.y; y[2] += bt.z; y[3] += bt.w;
    }
    if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + (size_t)row * p.ldo + i) = make_float4(y[0], y[1], y[2], y[3]);
    if (p.out) {
      const __half h0 = __float2half_rn(y[0]), h1 = __float2half_rn(y[1]), h2 = __float2half_rn(y[2]), h3 = __float2half_rn(y[3]);
      *reinterpret_cast<uint2*>(p.out + (size_t)row * p.ldo + i) =
          make_uint2((uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16), (uint32_t)__half_as_ushort(h2) | ((uint32_t)__half_as_ushort(h3) << 16));
      if (p.out_lo) {
        const __half2 l0 = __floats2half2_rn(y[0] - __half2float(h0), y[1] - __half2float(h1));
        const __half2 l1 = __floats2half2_rn(y[2] - __half2float(h2), y[3] - __half2float(h3));
        *reinterpret_cast<uint2*>(p.out_lo + (size_t)row * p.ldo + i) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
      }
      if (p.out_lo8) {   // lo halves as e5m2 scaled by 2^-2 (mixed8 numerics: the fp8 pass of the consuming GEMM)
        const __half hh[4] = {h0, h1, h2, h3};
        uint32_t w = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          w |= (uint32_t)__nv_cvt_float_to_fp8((y[e] - __half2float(hh[e])) * 0.25f, __NV_SATFINITE, __NV_E5M2) << (8 * e);
        *reinterpret_cast<uint32_t*>(p.out_lo8 + (size_t)row * p.ldo8 + i) = w;
      }
    }
  }
}

// Few rows (AR decode: M = batch): one CTA per row so that the whole row is in flight at once (latency-bound case).
__global__ void __launch_bounds__(256) norm_row_cta_kernel(NormCall p) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x, tid = threadIdx.x;
  const int src = p.row_map ? p.row_map[row] : row;
  const float* x = p.x + (size_t)src * p.ldx;
  const int D = p.D, nv = D >> 2;
  __shared__ float red[2][8];
  float4 v[2];
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = tid + j * 256;
    v[j] = i < nv ? *reinterpret_cast<const float4*>(x + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += v[j].x + v[j].y + v[j].z + v[j].w;
    ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
  }
  s = warp_sum(s); ss = warp_sum(ss);
  if ((tid & 31) == 0) { red[0][tid >> 5] = s; red[1][tid >> 5] = ss; }
  __syncthreads();
  s = 0.f; ss = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { s += red[0][w]; ss += red[1][w]; }
  float mean = 0.f, rstd;
  if (p.rms) {
    rstd = rsqrtf(ss / D + p.eps);
  } else {
    mean = s / D;
    float vs = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (tid + j * 256 < nv) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        vs += a * a + b * b + c * c + d * d;
      }
    }
    vs = warp_sum(vs);
    __syncthreads();
    if ((tid & 31) == 0) red[0][tid >> 5] = vs;
    __syncthreads();
    vs = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) vs += red[0][w];
    rstd = rsqrtf(vs / D + p.eps);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = 4 * (tid + j * 256);
    if (i >= D) continue;
    float y[4] = {(v[j].x - mean) * rstd, (v[j].y - mean) * rstd, (v[j].z - mean) * rstd, (v[j].w - mean) * rstd};
    if (p.gamma) { const float4 gm = *reinterpret_cast<const float4*>(p.gamma + i); y[0] *= gm.x; y[1] *= gm.y; y[2] *= gm.z; y[3] *= gm.w; }
    if (p.beta) { const float4 bt = *reinterpret_cast<const float4*>(p.beta + i); y[0] += bt.x; y[1] += bt.y; y[2] += bt.z; y[3] += bt.w; }
    if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + (size_t)row * p.ldo + i) = make_float4(y[0], y[1], y[2], y[3]);
    if (p.out) {
      const __half h0 = __float2half_rn(y[0]), h1 = __float2half_rn(y[1]), h2 = __float2half_rn(y[2]), h3 = __float2half_rn(y[3]);
      __half2* o = reinterpret_cast<__half2*>(p.out + (size_t)row * p.ldo + i);
      o[0] = __halves2half2(h0, h1);
      o[1] = __halves2half2(h2, h3);
      if (p.out_lo) {
        __half2* ol = reinterpret_cast<__half2*>(p.out_lo + (size_t)row * p.ldo + i);
        ol[0] = __floats2half2_rn(y[0] - __half2float(h0), y[1] - __half2float(h1));
        ol[1] = __floats2half2_rn(y[2] - __half2float(h2), y[3] - __half2float(h3));
      }
    }
  }
}

int norm_rows(const NormCall& c, cudaStream_t stream) {
  if (c.M <= 0) return M5_OK;
  if (c.D % 4 != 0 || c.ldx % 4 != 0 || c.ldo % 4 != 0) return M5_ERR_ARG;
  if (c.M <= 64 && c.D <= 2048) {
    return launch_k(norm_row_cta_kernel, dim3(c.M), dim3(256), 0, stream, c) == cudaSuccess ? M5_OK : M5_ERR_CUDA;
  }
  const int wpb = 8;
  const dim3 grid((c.M + wpb - 1) / wpb), block(wpb * 32);
  const bool aligned = ((size_t)c.x % 16 == 0) && (!c.out || (size_t)c.out % 8 == 0) && (!c.out_lo || (size_t)c.out_lo % 8 == 0) &&
                       (!c.out_f32 || (size_t)c.out_f32 % 16 == 0);
  if (aligned && c.D == 1024) norm_rows_reg_kernel<8><<<grid, block, 0, stream>>>(c);
  else if (aligned && c.D == 1536) norm_rows_reg_kernel<12><<<grid, block, 0, stream>>>(c);
  else if (aligned && c.D == 384) norm_rows_reg_kernel<3><<<grid, block, 0, stream>>>(c);
  else norm_rows_kernel<<<grid, block, 0, stream>>>(c);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------ fp32 -> fp16 rows
__global__ void cast_rows_kernel(const float* x, int ldx, __half* o, __half* olo, int ldo, int M, int D,
                                 const int* row_map) {
  const size_t n = (size_t)M * (D / 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / (D / 4)), c = (int)(i % (D / 4)) * 4;
    const int src = row_map ? row_map[row] : row;
    const float4 v = *reinterpret_cast<const float4*>(x + (size_t)src * ldx + c);
    const __half h0 = __float2half_rn(v.x), h1 = __float2half_rn(v.y), h2 = __float2half_rn(v.z), h3 = __float2half_rn(v.w);
    __half2* op = reinterpret_cast<__half2*>(o + (size_t)row * ldo + c);
    op[0] = __halves2half2(h0, h1);
    op[1] = __halves2half2(h2, h3);
    if (olo) {
      __half2* ol = reinterpret_cast<__half2*>(olo + (size_t)row * ldo + c);
      ol[0] = __floats2half2_rn(v.x - __half2float(h0), v.y - __half2float(h1));
      ol[1] = __floats2half2_rn(v.z - __half2float(h2), v.w - __half2float(h3));
    }
  }
}
int cast_rows(const float* x, int ldx, __half* o, __half* olo, int ldo, int M, int D, const int* row_map,
              cudaStream_t stream) {
  if (M <= 0) return M5_OK;
  const size_t n = (size_t)M * (D / 4);
  const int blocks = (int)min((size_t)148 * 8, (n + 255) / 256);
  cast_rows_kernel<<<blocks, 256, 0, stream>>>(x, ldx, o, olo, ldo, M, D, row_map);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------ RoPE + KV append
// Interleaved-pair RoPE (nn_future.py:166-198): (x[2i] + j x[2i+1]) * exp(j * pos * theta_i), theta_i = 10000^(-2i/64),
// computed in fp32 and rounded back to fp16 (apply_rotary_emb's .type_as).  q is rotated in place inside the packed
// qkv buffer; k (rotated) and v are scattered into the cache at [seq, pos] (nn_future.py:248-252; the window never
// wraps because max_len < sliding_window, ar_generate.py:57).
__global__ void rope_kv_kernel(__half* qkv, int ld, int n_rows, int H, const int* row_seq, const int* row_pos,
                               __half* kc, __half* vc, int W, const float* inv_freq) {
  const int D = H * 64;
  const int pairs = D / 2;
  const size_t total = (size_t)n_rows * pairs;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / pairs), pr = (int)(i % pairs);
    const int pos = row_pos[row], seq = row_seq[row];
    const int fi = pr & 31;
    // angle in fp32 exactly as torch.outer(t, freqs).float() then polar (cos, sin of the fp32 product)
    const float ang = (float)pos * inv_freq[fi];
    float sn, cs;
    sincosf(ang, &sn, &cs);
    __half2* qp = reinterpret_cast<__half2*>(qkv + (size_t)row * ld) + pr;
    const float2 q = __half22float2(*qp);
    *qp = __floats2half2_rn(q.x * cs - q.y * sn, q.x * sn + q.y * cs);
    const __half2* kp = reinterpret_cast<const __half2*>(qkv + (size_t)row * ld + D) + pr;
    const float2 k = __half22float2(*kp);
    const size_t coff = ((size_t)seq * W + pos) * D;
    reinterpret_cast<__half2*>(kc + coff)[pr] = __floats2half2_rn(k.x * cs - k.y * sn, k.x * sn + k.y * cs);
    reinterpret_cast<__half2*>(vc + coff)[pr] = reinterpret_cast<const __half2*>(qkv + (size_t)row * ld + 2 * D)[pr];
  }
}
int rope_kv(__half* qkv, int ld, int n_rows, int H, const int* row_seq, const int* row_pos, __half* kc, __half* vc,
            int W, const float* inv_freq, cudaStream_t stream) {
  if (n_rows <= 0) return M5_OK;
  const size_t total = (size_t)n_rows * H * 32;
  const int blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
  rope_kv_kernel<<<blocks, 256, 0, stream>>>(qkv, ld, n_rows, H, row_seq, row_pos, kc, vc, W, inv_freq);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

// Decode variant: the QKV projection of the skinny GEMM is fp32 [B, 3D]; round to fp16 first (autocast Linear output),
// rotate, write q as fp16 [B, D] and append k, v at position pos[b] = len[b] - 1.
__global__ void rope_kv_decode_kernel(const float* qkv, int B, int H, const int* len, __half* qout, __half* kc,
                                      __half* vc, int W, const float* inv_freq, const int* active) {
  pdl_launch_dependents();
  pdl_wait();
  const int D = H * 64, pairs = D / 2;
  const int total = B * pairs;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / pairs, pr = i % pairs;
    if (active && !active[b]) continue;
    const int pos = len[b] - 1;
    const float ang = (float)pos * inv_freq[pr & 31];
    float sn, cs;
    sincosf(ang, &sn, &cs);
    const float* r = qkv + (size_t)b * 3 * D;
    const float2 q = __half22float2(__floats2half2_rn(r[2 * pr], r[2 * pr + 1]));
    const float2 k = __half22float2(__floats2half2_rn(r[D + 2 * pr], r[D + 2 * pr + 1]));
    reinterpret_cast<__half2*>(qout + (size_t)b * D)[pr] = __floats2half2_rn(q.x * cs - q.y * sn, q.x * sn + q.y * cs);
    const size_t coff = ((size_t)b * W + pos) * D;
    reinterpret_cast<__half2*>(kc + coff)[pr] = __floats2half2_rn(k.x * cs - k.y * sn, k.x * sn + k.y * cs);
    reinterpret_cast<__half2*>(vc + coff)[pr] = __floats2half2_rn(r[2 * D + 2 * pr], r[2 * D + 2 * pr + 1]);
  }
}
int rope_kv_decode(const float* qkv, int B, int H, const int* len, __half* qout, __half* kc, __half* vc, int W,
                   const float* inv_freq, const int* active, cudaStream_t stream) {
  const int total = B * H * 32;
  return launch_k(rope_kv_decode_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, qkv, B, H, len, qout, kc, vc, W, inv_freq, active) == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------ embeddings
// SinePositionalEmbedding (nn_future.py:35-83): x + alpha * pe[pos], pe[:,0::2]=sin(pos*div), pe[:,1::2]=cos(pos*div),
// div_i = exp(2i * -(ln 1e4 / D)).  `pe` tables are precomputed on the host with torch-identical fp32 arithmetic.
//
// ChunkedEmbedding (model.py:147-159): concat over Q codebooks of emb_q[code_q] (dim D/Q each).
// out[row] = (row is identity slot ? identity : chunked(codes[row])) + alpha * pe[pos] (+ add_vec)
__global__ void chunked_embed_kernel(EmbedCall p) {
  const int row = blockIdx.x;
  const int src = p.code_row[row];
  const int pos = p.pos[row];
  const int D = p.D, dq = D / p.Q;
  float* o = p.out + (size_t)row * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float v;
    if (src < 0) {
      v = p.identity[c];
    } else {
      const int qi = c / dq;
      const int code = p.codes[(size_t)src * p.Q + qi];
      v = __half2float(p.tables[((size_t)qi * p.n_codes + code) * dq + (c - qi * dq)]);
    }
    if (p.pe) v += p.alpha * p.pe[(size_t)pos * D + c];
    if (p.add_vec) v += p.add_vec[c];
    o[c] = v;
  }
}
int chunked_embed(const EmbedCall& c, cudaStream_t stream) {
  if (c.n_rows <= 0) return M5_OK;
  chunked_embed_kernel<<<c.n_rows, 256, 0, stream>>>(c);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

// Token embedding rows: out[row] = (tok >= 0 ? table[tok] : vec_rows[-tok-1]) (+ alpha*pe[pos]) (+ add_vec)
__global__ void token_embed_kernel(TokEmbedCall p) {
  const int row = blockIdx.x;
  const int tok = p.tok[row];
  const int D = p.D;
  float* o = p.out + (size_t)row * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float v = tok >= 0 ? __half2float(p.table[(size_t)tok * D + c]) : p.vec_rows[(size_t)(-tok - 1) * D + c];
    if (p.pe) v += p.alpha * p.pe[(size_t)p.pos[row] * D + c];
    if (p.add_vec) v += p.add_vec[c];
    o[c] = v;
  }
}
int token_embed(const TokEmbedCall& c, cudaStream_t stream) {
  if (c.n_rows <= 0) return M5_OK;
  token_embed_kernel<<<c.n_rows, 256, 0, stream>>>(c);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

// Gather rows: out[i] = x[idx[i]]  (fp32, D % 4 == 0)
__global__ void gather_rows_kernel(const float* x, int ldx, const int* idx, float* out, int ldo, int n, int D) {
  const int row = blockIdx.x;
  const float* s = x + (size_t)idx[row] * ldx;
  float* o = out + (size_t)row * ldo;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) *reinterpret_cast<float4*>(o + c) = *reinterpret_cast<const float4*>(s + c);
}
int gather_rows(const float* x, int ldx, const int* idx, float* out, int ldo, int n, int D, cudaStream_t stream) {
  if (n <= 0) return M5_OK;
  gather_rows_kernel<<<n, 128, 0, stream>>>(x, ldx, idx, out, ldo, n, D);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

}  // namespace m5
