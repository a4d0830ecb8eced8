// Encodec 24 kHz encoder + residual vector quantiser on the device: the step right BEFORE the hot
// path, `self.codec.encode(ref_audio[None])` at inference.py:233 with EncodecModel.encodec_model_24khz() at 6 kbps
// (inference.py:87-88) -> (8, T) codes per reference clip, T = ceil(samples / 320).
//
// Third-party algorithm (package `encodec`, not in /root/reference): SEANetEncoder (reflect-padded CAUSAL convolutions --
// encodec_model_24khz is the causal model: the whole padding sits on the left -- 1 -> 32 -> 64 -> 128 -> 256 -> 512 channels
// with strides 2, 4, 5, 8, one residual block per scale, ELU), a 2-layer LSTM with skip connection, a 512 -> 128 convolution,
// then 8 greedy nearest-neighbour stages over 1024 x 128 codebooks.  oracle/encodec_oracle.py restates it with file citations
// and is pinned against the independent `transformers` implementation of the same model (tests/test_encodec_hf_cpu.py);
// parity with the released WEIGHTS needs the real package.
//
// A batch of B clips is processed together; activations are channel-major [C][sum_b L_b] with the clips side by side on the
// time axis (each clip keeps its own reflect padding).  All arithmetic is fp32 on the CUDA cores: the output is an arg-max
// over distances, so the tensor-core fp16 paths of the hot loop are not appropriate here, and the work (~18 GFLOP per 6 s
// clip, once per utterance) is three orders of magnitude below one synthesis step.
//   enc_conv_kernel   direct convolution, 16 output channels x 128 output times per CTA, input window and weights staged
//                     in shared memory, optional ELU on the input, optional residual add
//   enc_lstm_kernel   one CTA per clip and layer: the input projections of all steps come from enc_conv (k = 1), the
//                     recurrent matrix (4H x H fp32) is streamed from L2 every step, h lives in shared memory
//   enc_rvq_kernel    one CTA per frame: 8 x (1024 distances, arg-max with the lowest index on ties, residual update)
#include <math.h>

#include <string>
#include <vector>

#include "ctx.h"

namespace m5 {

static constexpr int EC_TT = 128;   // output times per CTA
static constexpr int EC_TC = 16;    // output channels per CTA
static constexpr int EC_CI = 8;     // input channels per staging round

// value of the reflect-padded clip at padded index j (SConv1d / pad1d of encodec.modules.conv, restated in the oracle): the
// clip is zero-extended by `ext` samples first when it is not longer than the larger pad (reflection needs pad < length)
__device__ __forceinline__ float enc_padded(const float* x, int len, int left, int ext, int j, bool elu) {
  const int Lp = len + ext;
  int s = j - left;
  if (s < 0) s = -s;
  else if (s >= Lp) s = 2 * (Lp - 1) - s;
  float v = (s >= 0 && s < len) ? x[s] : 0.f;
  if (elu) v = v > 0.f ? v : expm1f(v);
  return v;
}

// y[co][out_off[b] + t] = bias[co] + sum_{ci, kk} w[co][ci][kk] * pad(elu?(x))[ci][t * stride + kk]   (+ resid)
__global__ void __launch_bounds__(EC_TT) enc_conv_kernel(const float* x, const int* in_off, const int* in_len, long long ld_in,
                                                         const float* w, const float* bias, float* y, const int* out_off,
                                                         const int* out_len, long long ld_out, int Ci, int Co, int k, int stride,
                                                         int pre_elu, const float* resid) {
  extern __shared__ float ec_smem[];
  const int b = blockIdx.z, t0 = blockIdx.x * EC_TT, co0 = blockIdx.y * EC_TC, tid = threadIdx.x;
  const int Lin = in_len[b], Lout = out_len[b];
  if (t0 >= Lout) return;
  const int total = k - stride;                       // padding_total (dilation 1 everywhere in the encoder)
  const int right = 0, left = total;                  // causal SConv1d (encodec_model_24khz): all of it on the left
  const int extra = (Lout - 1) * stride + (k - total) - Lin;   // get_extra_padding_for_conv1d
  const int max_pad = max(left, right + extra);
  const int ext = Lin <= max_pad ? max_pad - Lin + 1 : 0;
  const int win = (EC_TT - 1) * stride + k;
  float* xs = ec_smem;                                // [EC_CI][win]
  float* ws = ec_smem + EC_CI * win;                  // [EC_TC][EC_CI][k]
  float acc[EC_TC];
#pragma unroll
  for (int c = 0; c < EC_TC; ++c) acc[c] = 0.f;
  const int t = t0 + tid;
  for (int ci0 = 0; ci0 < Ci; ci0 += EC_CI) {
    const int nci = min(EC_CI, Ci - ci0);
    for (int i = tid; i < nci * win; i += EC_TT) {
      const int ci = i / win, j = i - ci * win;
      const int pj = t0 * stride + j;                 // index into the padded clip
      float v = 0.f;
      if (pj < Lin + left + right + extra) v = enc_padded(x + (size_t)(ci0 + ci) * ld_in + in_off[b], Lin, left, ext, pj, pre_elu != 0);
      xs[ci * win + j] = v;
    }
    for (int i = tid; i < EC_TC * nci * k; i += EC_TT) {
      const int c = i / (nci * k), r = i - c * (nci * k), ci = r / k, kk = r - ci * k;
      ws[(c * EC_CI + ci) * k + kk] = (co0 + c < Co) ? w[((size_t)(co0 + c) * Ci + ci0 + ci) * k + kk] : 0.f;
    }
    __syncthreads();
    if (t < Lout) {
      for (int ci = 0; ci < nci; ++ci) {
        const float* xr = xs + ci * win + tid * stride;
        for (int kk = 0; kk < k; ++kk) {
          const float v = xr[kk];
#pragma unroll
          for (int c = 0; c < EC_TC; ++c) acc[c] = fmaf(ws[(c * EC_CI + ci) * k + kk], v, acc[c]);
        }
      }
    }
    __syncthreads();
  }
  if (t >= Lout) return;
#pragma unroll
  for (int c = 0; c < EC_TC; ++c) {
    if (co0 + c >= Co) break;
    const size_t o = (size_t)(co0 + c) * ld_out + out_off[b] + t;
    float v = acc[c] + (bias ? bias[co0 + c] : 0.f);
    if (resid) v += resid[o];
    y[o] = v;
  }
}

// nn.LSTM layer over time for one clip per CTA.  pre [4H][ld] holds W_ih x_t + b_ih for every step (gate order i, f, g, o);
// out [H][ld] (+ skip when `skip` != null: SLSTM's y = lstm(x) + x, added after the LAST layer only).
__global__ void __launch_bounds__(1024) enc_lstm_kernel(const float* pre, const float* w_hh, const float* b_hh, float* out,
                                                        const float* skip, const int* off, const int* len, long long ld, int H) {
  extern __shared__ float el_smem[];
  float* h = el_smem;            // [H]
  float* g = el_smem + H;        // [4H] gate pre-activations of the step
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  const int T = len[b], o0 = off[b];
  for (int i = tid; i < H; i += blockDim.x) h[i] = 0.f;
  float c_reg = 0.f;             // cell state of unit `tid` (threads < H)
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    // g = pre[:, t] + W_hh h + b_hh : one warp per gate row, lanes stride over the H columns (coalesced weight reads)
    for (int r = warp; r < 4 * H; r += nwarp) {
      const float* wr = w_hh + (size_t)r * H;
      float s = 0.f;
      for (int j = lane; j < H; j += 32) s = fmaf(__ldg(wr + j), h[j], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) g[r] = s + pre[(size_t)r * ld + o0 + t] + b_hh[r];
    }
    __syncthreads();
    if (tid < H) {
      const float ig = 1.f / (1.f + expf(-g[tid])), fg = 1.f / (1.f + expf(-g[H + tid]));
      const float gg = tanhf(g[2 * H + tid]), og = 1.f / (1.f + expf(-g[3 * H + tid]));
      c_reg = fg * c_reg + ig * gg;
      const float hn = og * tanhf(c_reg);
      h[tid] = hn;
      const size_t o = (size_t)tid * ld + o0 + t;
      out[o] = hn + (skip ? skip[o] : 0.f);
    }
    __syncthreads();
  }
}

// Residual vector quantisation of one frame per CTA (core_vq.py: dist = -(|x|^2 - 2 x.e + |e|^2), arg-max, first index on ties)
__global__ void __launch_bounds__(256) enc_rvq_kernel(const float* emb, long long ld, int n_frames, int dim, int bins, int n_q,
                                                      const float* const* codebooks, const float* const* e_sq, int* codes) {
  extern __shared__ float rq_smem[];
  float* r = rq_smem;                      // [dim] residual
  __shared__ float s_best[8];
  __shared__ int s_idx[8];
  __shared__ float s_xsq;
  const int f = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (f >= n_frames) return;
  for (int d = tid; d < dim; d += blockDim.x) r[d] = emb[(size_t)d * ld + f];
  __syncthreads();
  for (int q = 0; q < n_q; ++q) {
    const float* cb = codebooks[q];
    if (warp == 0) {
      float s = 0.f;
      for (int d = lane; d < dim; d += 32) s = fmaf(r[d], r[d], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) s_xsq = s;
    }
    __syncthreads();
    float best = -INFINITY;
    int bi = 0x7FFFFFFF;
    for (int c = tid; c < bins; c += blockDim.x) {
      const float* e = cb + (size_t)c * dim;
      float dot = 0.f;
      for (int d = 0; d < dim; ++d) dot = fmaf(r[d], __ldg(e + d), dot);
      const float dist = -(s_xsq - 2.f * dot + e_sq[q][c]);
      if (dist > best || (dist == best && c < bi)) { best = dist; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { s_best[warp] = best; s_idx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w2 = 1; w2 < (int)(blockDim.x >> 5); ++w2)
        if (s_best[w2] > best || (s_best[w2] == best && s_idx[w2] < bi)) { best = s_best[w2]; bi = s_idx[w2]; }
      s_idx[0] = bi;
      codes[(size_t)f * n_q + q] = bi;
    }
    __syncthreads();
    const int pick = s_idx[0];
    for (int d = tid; d < dim; d += blockDim.x) r[d] -= __ldg(cb + (size_t)pick * dim + d);
    __syncthreads();
  }
}

__global__ void enc_rowsq_kernel(const float* cb, int bins, int dim, float* out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= bins) return;
  float s = 0.f;
  for (int d = 0; d < dim; ++d) s = fmaf(cb[(size_t)c * dim + d], cb[(size_t)c * dim + d], s);
  out[c] = s;
}

struct EncLayer { const float* w; const float* b; int Ci, Co, k, stride; };

}  // namespace m5

using namespace m5;

extern "C" int m5_encodec_encode(m5_ctx* ctx, int32_t B, const float* wav, const int32_t* n_samples, int32_t mem, int32_t n_q,
                                 int32_t* codes_out) {
  if (!ctx || B <= 0 || !wav || !n_samples || !codes_out || n_q <= 0) return M5_ERR_ARG;
  ctx->last_error.clear();
  cudaSetDevice(ctx->device);
  // ---- weights (names: weights.repack_encodec)
  auto getw = [&](const std::string& n, int64_t* numel) -> const float* {
    const m5_tensor* t = find_weight(ctx, n);
    if (!t) return nullptr;
    if (numel) *numel = t->numel;
    return reinterpret_cast<const float*>(t->ptr);
  };
  int64_t n0 = 0, nfin = 0, ncb = 0;
  const float* c0w = getw("enc.c0.w", &n0);
  if (!c0w) return M5_ERR_MISSING_WEIGHT;
  const int F = (int)(n0 / 7), H = 16 * F;
  const float* finw = getw("enc.final.w", &nfin);
  const float* cb0 = getw("enc.cb0", &ncb);
  if (!finw || !cb0 || F <= 0) return M5_ERR_MISSING_WEIGHT;
  const int dim = (int)(nfin / ((int64_t)H * 7)), bins = (int)(ncb / dim);
  if (dim <= 0 || bins <= 0 || n_q > 32) return ctx->fail(M5_ERR_ARG, "encodec: inconsistent weight shapes / n_q");
  static const int ratios[4] = {2, 4, 5, 8};
  // ---- lengths per scale
  std::vector<std::vector<int>> len(6, std::vector<int>(B)), off(6, std::vector<int>(B + 1, 0));
  for (int b = 0; b < B; ++b) {
    if (n_samples[b] <= 0) return ctx->fail(M5_ERR_ARG, "encodec: empty clip");
    len[0][b] = n_samples[b];
    for (int s = 0; s < 4; ++s) len[s + 1][b] = (len[s][b] + ratios[s] - 1) / ratios[s];
  }
  for (int s = 0; s < 5; ++s)
    for (int b = 0; b < B; ++b) off[s][b + 1] = off[s][b] + len[s][b];
  const long long L0 = off[0][B], LT = off[4][B];
  // ---- workspace: two ping-pong activation buffers sized for the widest stage + LSTM buffers
  size_t widest = 0;
  for (int s = 0; s < 5; ++s) widest = std::max(widest, (size_t)(F << s) * (size_t)off[s][B]);
  Arena ar(ctx);
  const size_t bytes = (3 * widest + (size_t)L0 + (size_t)LT * (4 * H + 3 * H + dim)) * 4 + (size_t)LT * n_q * 4 + (size_t)n_q * bins * 4 +
                       (size_t)(12 * (B + 1)) * 4 + (1 << 20);
  M5_TRY(ar.reserve(bytes));
  float* d_wav = const_cast<float*>(wav);
  if (mem == M5_MEM_HOST) {
    d_wav = ar.get<float>(L0);
    if (!d_wav) return ctx->fail(M5_ERR_NOMEM, "arena too small (encodec)");
    M5_CUDA(cudaMemcpyAsync(d_wav, wav, (size_t)L0 * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  int* d_off[5]; int* d_len[5];
  for (int s = 0; s < 5; ++s) {
    d_off[s] = ar.get<int>(B + 1); d_len[s] = ar.get<int>(B);
    if (!d_off[s] || !d_len[s]) return ctx->fail(M5_ERR_NOMEM, "arena too small (encodec)");
    cudaMemcpyAsync(d_off[s], off[s].data(), (B + 1) * 4, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_len[s], len[s].data(), B * 4, cudaMemcpyHostToDevice, ctx->stream);
  }
  float* bufA = ar.get<float>(widest); float* bufB = ar.get<float>(widest); float* bufC = ar.get<float>(widest);
  float* pre = ar.get<float>((size_t)LT * 4 * H); float* l0o = ar.get<float>((size_t)LT * H); float* l1o = ar.get<float>((size_t)LT * H);
  float* emb = ar.get<float>((size_t)LT * dim);
  int* d_codes = codes_out;
  if (mem == M5_MEM_HOST) d_codes = ar.get<int>((size_t)LT * n_q);
  float* esq = ar.get<float>((size_t)n_q * bins);
  const float** d_cbp = reinterpret_cast<const float**>(ar.get<double>(n_q)); const float** d_esqp = reinterpret_cast<const float**>(ar.get<double>(n_q));
  if (!bufC || !emb || !d_codes || !esq || !d_cbp || !d_esqp || !l1o) return ctx->fail(M5_ERR_NOMEM, "arena too small (encodec)");

  auto conv = [&](const float* x, int s_in, long long ld_in, const std::string& name, int Ci, int Co, int k, int stride, bool elu,
                  float* y, int s_out, long long ld_out, const float* resid) -> int {
    const float* w = getw(name + ".w", nullptr); const float* bs = getw(name + ".b", nullptr);
    if (!w || !bs) return M5_ERR_MISSING_WEIGHT;
    int mx = 0;
    for (int b = 0; b < B; ++b) mx = std::max(mx, len[s_out][b]);
    dim3 grid((mx + EC_TT - 1) / EC_TT, (Co + EC_TC - 1) / EC_TC, B);
    const size_t smem = ((size_t)EC_CI * ((EC_TT - 1) * stride + k) + (size_t)EC_TC * EC_CI * k) * 4;
    enc_conv_kernel<<<grid, EC_TT, smem, ctx->stream>>>(x, d_off[s_in], d_len[s_in], ld_in, w, bs, y, d_off[s_out], d_len[s_out], ld_out,
                                                        Ci, Co, k, stride, elu ? 1 : 0, resid);
    ctx->launches++;
    return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
  };
  static DeviceOnce once;
  unsigned long long bit;
  if (once.needed(bit)) {
    cudaFuncSetAttribute(enc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    once.done(bit);
  }
  // ---- SEANet encoder
  float* x = bufA; float* y1 = bufB; float* y2 = bufC;
  M5_TRY(conv(d_wav, 0, L0, "enc.c0", 1, F, 7, 1, false, x, 0, L0, nullptr));
  int C = F;
  for (int s = 0; s < 4; ++s) {
    const long long ld = off[s][B];
    const std::string r = "enc.r" + std::to_string(s);
    M5_TRY(conv(x, s, ld, r + ".a", C, C / 2, 3, 1, true, y1, s, ld, nullptr));          // ELU -> k3, dim -> dim/2
    M5_TRY(conv(x, s, ld, r + ".s", C, C, 1, 1, false, y2, s, ld, nullptr));             // shortcut k1 (true_skip = False)
    M5_TRY(conv(y1, s, ld, r + ".b", C / 2, C, 1, 1, true, y2, s, ld, y2));              // ELU -> k1, + shortcut
    M5_TRY(conv(y2, s, ld, "enc.d" + std::to_string(s), C, 2 * C, 2 * ratios[s], ratios[s], true, x, s + 1, off[s + 1][B], nullptr));
    C *= 2;
  }
  // ---- SLSTM (2 layers, skip) on [H][LT]
  for (int layer = 0; layer < 2; ++layer) {
    const std::string l = "enc.lstm" + std::to_string(layer);
    const float* xin = layer == 0 ? x : l0o;
    M5_TRY(conv(xin, 4, LT, l + ".ih", H, 4 * H, 1, 1, false, pre, 4, LT, nullptr));    // W_ih x_t + b_ih for all t
    const float* whh = getw(l + ".hh.w", nullptr); const float* bhh = getw(l + ".hh.b", nullptr);
    if (!whh || !bhh) return M5_ERR_MISSING_WEIGHT;
    const int threads = std::min(1024, std::max(64, ((H + 31) / 32) * 32));
    enc_lstm_kernel<<<B, threads, (size_t)5 * H * 4, ctx->stream>>>(pre, whh, bhh, layer == 0 ? l0o : l1o, layer == 1 ? x : nullptr,
                                                                     d_off[4], d_len[4], LT, H);
    ctx->launches++;
    if (H > threads) return ctx->fail(M5_ERR_ARG, "encodec: LSTM width exceeds one CTA");
  }
  M5_TRY(conv(l1o, 4, LT, "enc.final", H, dim, 7, 1, true, emb, 4, LT, nullptr));
  // ---- RVQ
  std::vector<const float*> h_cb(n_q), h_esq(n_q);
  for (int q = 0; q < n_q; ++q) {
    h_cb[q] = getw("enc.cb" + std::to_string(q), nullptr);
    if (!h_cb[q]) return M5_ERR_MISSING_WEIGHT;
    h_esq[q] = esq + (size_t)q * bins;
    enc_rowsq_kernel<<<(bins + 255) / 256, 256, 0, ctx->stream>>>(h_cb[q], bins, dim, esq + (size_t)q * bins);
  }
  M5_CUDA(cudaMemcpyAsync(d_cbp, h_cb.data(), n_q * sizeof(float*), cudaMemcpyHostToDevice, ctx->stream));
  M5_CUDA(cudaMemcpyAsync(d_esqp, h_esq.data(), n_q * sizeof(float*), cudaMemcpyHostToDevice, ctx->stream));
  enc_rvq_kernel<<<(unsigned)LT, 256, (size_t)dim * 4, ctx->stream>>>(emb, LT, (int)LT, dim, bins, n_q, d_cbp, d_esqp, d_codes);
  ctx->launches += 1 + n_q;
  if (mem == M5_MEM_HOST) M5_CUDA(cudaMemcpyAsync(codes_out, d_codes, (size_t)LT * n_q * 4, cudaMemcpyDeviceToHost, ctx->stream));
  M5_CUDA(cudaStreamSynchronize(ctx->stream));
  if (cudaGetLastError() != cudaSuccess) return ctx->fail(M5_ERR_CUDA, "encodec kernels failed");
  return M5_OK;
}
