// m5_vocode: Mars5TTS.vocode (inference.py:160-172) = vocos.codes_to_features + VocosBackbone + ISTFTHead, batched
// over B utterances packed along the frame axis (convolutions are zero-padded at every utterance edge, as a bs=1 run).
// Algorithm restated from vocos 0.1.0; the kernels are in vocos.cu, the dense layers run on the
// tcgen05 GEMM with three-term split-fp16 operands (fp32-class accuracy).
#include <algorithm>
#include <vector>

#include "layers.h"
#include "sampler.h"
#include "vocos.h"

namespace m5 {

struct VocLayerW { const float *dw_w, *dw_b, *n_scale, *n_shift, *pw1_b, *pw2_b, *gamma; const __half *pw1_w, *pw2_w; };
struct VocWeights {
  const float* codebook; const __half* embed_w; const float* embed_b; const float *n_scale, *n_shift;
  std::vector<VocLayerW> layers;
  const float *fln_w, *fln_b; const __half* head_w; const float* head_b;
  const float2 *w1280, *w640; const float* window;
};

static int load_voc(m5_ctx* ctx, VocWeights& w) {
  const m5_model_cfg& c = ctx->cfg;
#define GETW(dst, T, name) do { dst = W<T>(ctx, name); if (!(dst)) return M5_ERR_MISSING_WEIGHT; } while (0)
  GETW(w.codebook, float, "voc.codebook"); GETW(w.embed_w, __half, "voc.embed_w"); GETW(w.embed_b, float, "voc.embed_b");
  GETW(w.n_scale, float, "voc.norm_scale"); GETW(w.n_shift, float, "voc.norm_shift");
  w.layers.resize(c.voc_layers);
  for (int i = 0; i < c.voc_layers; ++i) {
    const std::string p = "voc.l" + std::to_string(i) + ".";
    VocLayerW& l = w.layers[i];
    GETW(l.dw_w, float, p + "dw_w"); GETW(l.dw_b, float, p + "dw_b"); GETW(l.n_scale, float, p + "norm_scale");
    GETW(l.n_shift, float, p + "norm_shift"); GETW(l.pw1_w, __half, p + "pw1_w"); GETW(l.pw1_b, float, p + "pw1_b");
    GETW(l.pw2_w, __half, p + "pw2_w"); GETW(l.pw2_b, float, p + "pw2_b"); GETW(l.gamma, float, p + "gamma");
  }
  GETW(w.fln_w, float, "voc.final_ln_w"); GETW(w.fln_b, float, "voc.final_ln_b");
  GETW(w.head_w, __half, "voc.head_w"); GETW(w.head_b, float, "voc.head_b");
  GETW(w.w1280, float2, "voc.w1280"); GETW(w.w640, float2, "voc.w640"); GETW(w.window, float, "voc.window");
#undef GETW
  return M5_OK;
}

// LayerNorm (optionally adaptive scale/shift) -> [hi | lo] fp16 halves, row stride 2*D
static int ln_split(m5_ctx* ctx, const float* x, int rows, int D, const float* g, const float* b, float eps, __half* out) {
  NormCall n;
  n.x = x; n.M = rows; n.D = D; n.ldx = D; n.gamma = g; n.beta = b; n.eps = eps; n.out = out; n.out_lo = out + D; n.ldo = 2 * D;
  return run_norm(ctx, n);
}
// A = [hi | lo] (row stride 2*K) times W3 = [W_hi | W_hi | W_lo] ([N, 3*K])
static GemmCall lin3(const __half* A, int rows, int K, const __half* W3, int N, const float* bias) {
  GemmCall g;
  g.A = A; g.W = W3; g.M = rows; g.N = N; g.K = 3 * K; g.lda = 2 * K; g.ldw = 3 * K; g.awrap = 2 * K; g.bias = bias;
  return g;
}

static bool g_istft_ready = false;

static int istft_run(m5_ctx* ctx, const VocWeights& w, const float* spec, int ld, int B, const std::vector<int>& nf,
                     float* frames, int* d_frame0, int* d_nframes, float* wav) {
  if (!g_istft_ready) {
    if (istft_setup_constants() != M5_OK) return ctx->fail(M5_ERR_CUDA, "istft constants");
    g_istft_ready = true;
  }
  int N = 0, mx = 0;
  for (int b = 0; b < B; ++b) { N += nf[b]; mx = std::max(mx, nf[b]); }
  if (istft_frames(spec, ld, N, w.w1280, w.w640, w.window, frames, ctx->stream) != M5_OK)
    return ctx->fail(M5_ERR_CUDA, "istft_frames failed");
  if (istft_ola(frames, w.window, d_frame0, d_nframes, B, mx, wav, ctx->stream) != M5_OK)
    return ctx->fail(M5_ERR_CUDA, "istft_ola failed");
  ctx->launches += 2;
  return M5_OK;
}

}  // namespace m5

using namespace m5;

extern "C" {

// Silence trim on the device, right behind the overlap-add (inference.py:304-305 -> mars5/trim.py:110-177): the waveform is
// still in HBM, so the frame powers are one more pass over it instead of a host loop over a copied buffer.
//   trim_power_kernel : one warp per 2048-sample frame (hop 512) of the reflect-padded waveform, mean square in double
//   trim_bounds_kernel: one CTA per utterance: loudest frame, then the first / last frame within top_db of it --
//                       10 log10(max(1e-10, p_f)) - 10 log10(max(1e-10, max_f p_f)) > -top_db, evaluated in double exactly
//                       like the host entry point m5_trim_bounds (trim.cu)
__global__ void trim_power_kernel(const float* wav, const long long* off, const int* frame0, int B, int frame_length, int hop,
                                  double* power) {
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int total = frame0[B];
  if (gw >= total) return;
  int b = 0;
  while (b + 1 < B && frame0[b + 1] <= gw) ++b;
  const int f = gw - frame0[b];
  const float* y = wav + off[b];
  const long long n = off[b + 1] - off[b], pad = frame_length / 2;
  double s = 0.0;
  for (int i = lane; i < frame_length; i += 32) {
    long long j = (long long)f * hop + i - pad;   // index into the unpadded signal; reflect without repeating the edge
    if (j < 0) j = -j;
    else if (j >= n) j = 2 * (n - 1) - j;
    const double v = (double)y[j];
    s += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) power[gw] = s / (double)frame_length;
}
__global__ void trim_bounds_kernel(const double* power, const long long* off, const int* frame0, int hop, double top_db,
                                   long long* start, long long* end) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* p = power + frame0[b];
  const int nfr = frame0[b + 1] - frame0[b];
  __shared__ double s_max[256];
  __shared__ int s_first[256], s_last[256];
  double mx = 0.0;
  for (int f = tid; f < nfr; f += blockDim.x) mx = fmax(mx, p[f]);
  s_max[tid] = mx;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) { if (tid < o) s_max[tid] = fmax(s_max[tid], s_max[tid + o]); __syncthreads(); }
  const double ref_db = 10.0 * log10(fmax(1e-10, s_max[0]));
  int first = 0x7FFFFFFF, last = -1;
  for (int f = tid; f < nfr; f += blockDim.x) {
    const double db = 10.0 * log10(fmax(1e-10, p[f])) - ref_db;
    if (db > -top_db) { first = min(first, f); last = max(last, f); }
  }
  s_first[tid] = first; s_last[tid] = last;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (tid < o) { s_first[tid] = min(s_first[tid], s_first[tid + o]); s_last[tid] = max(s_last[tid], s_last[tid + o]); }
    __syncthreads();
  }
  if (tid == 0) {
    const long long n = off[b + 1] - off[b];
    if (s_last[0] < 0) { start[b] = 0; end[b] = 0; }   // "the signal only contains zeros"
    else { start[b] = (long long)s_first[0] * hop; end[b] = min(n, (long long)(s_last[0] + 1) * hop); }
  }
}

static int vocode_impl(m5_ctx* ctx, int32_t B, const int32_t* codes, const int32_t* n_frames, int32_t bandwidth_id, int32_t mem,
                       float* wav_out, bool do_trim, float top_db, int frame_length, int hop_length, int64_t* t_start, int64_t* t_end) {
  if (!ctx || B <= 0) return M5_ERR_ARG;
  ctx->last_error.clear();
  cudaSetDevice(ctx->device);
  const m5_model_cfg& c = ctx->cfg;
  if (c.voc_nfft != 1280 || c.voc_hop != 320) return ctx->fail(M5_ERR_ARG, "iSTFT kernel is specialised for n_fft 1280 / hop 320");
  if (bandwidth_id < 0 || bandwidth_id >= c.voc_n_bw) return ctx->fail(M5_ERR_ARG, "bandwidth_id out of range");
  VocWeights w;
  M5_TRY(load_voc(ctx, w));
  const int C = c.voc_feat, D = c.voc_dim, I = c.voc_inter, Q = c.n_quant, NB = c.voc_nfft + 2;
  std::vector<int> nf(n_frames, n_frames + B), fpos, flen, f0(B);
  int N = 0;
  for (int b = 0; b < B; ++b) {
    f0[b] = N;
    for (int i = 0; i < nf[b]; ++i) { fpos.push_back(i); flen.push_back(nf[b]); }
    N += nf[b];
  }
  if (N == 0) return M5_OK;
  const int ldspec = (NB + 3) & ~3;
  Arena ar(ctx);
  size_t bytes = (size_t)N * (C * 4 + 2 * 7 * C * 2 + 2 * D * 4 + 2 * D * 2 + 2 * I * 2 + ldspec * 4 + 1280 * 4 + Q * 4 + 8 + 320 * 4) +
                 (size_t(16) << 20);
  M5_TRY(ar.reserve(bytes));
  const int* d_codes = codes;
  if (mem == M5_MEM_HOST) {
    int* d = ar.get<int>((size_t)N * Q);
    cudaMemcpyAsync(d, codes, (size_t)N * Q * 4, cudaMemcpyHostToDevice, ctx->stream);
    d_codes = d;
  }
  int* d_fpos = ar.get<int>(N); int* d_flen = ar.get<int>(N); int* d_f0 = ar.get<int>(B); int* d_nf = ar.get<int>(B);
  cudaMemcpyAsync(d_fpos, fpos.data(), (size_t)N * 4, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(d_flen, flen.data(), (size_t)N * 4, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(d_f0, f0.data(), (size_t)B * 4, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(d_nf, nf.data(), (size_t)B * 4, cudaMemcpyHostToDevice, ctx->stream);
  float* feat = ar.get<float>((size_t)N * C);
  __half* col16 = ar.get<__half>((size_t)N * 2 * 7 * C);
  float* x = ar.get<float>((size_t)N * D);
  float* y = ar.get<float>((size_t)N * D);
  __half* h16 = ar.get<__half>((size_t)N * 2 * D);
  __half* g16 = ar.get<__half>((size_t)N * 2 * I);
  float* spec = ar.get<float>((size_t)N * ldspec);
  float* frames = ar.get<float>((size_t)N * 1280);
  float* d_wav = wav_out;
  if (mem == M5_MEM_HOST) d_wav = ar.get<float>((size_t)N * 320);
  if (!frames || !d_wav) return ctx->fail(M5_ERR_NOMEM, "arena too small (vocode)");
  const float eps = 1e-6f;
  // codes_to_features + embed conv (k=7) as im2col GEMM
  if (voc_features(d_codes, w.codebook, feat, N, Q, C, c.voc_codebook, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "voc_features");
  if (voc_im2col(feat, d_fpos, d_flen, col16, N, C, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "voc_im2col");
  ctx->launches += 2;
  GemmCall ge = lin3(col16, N, 7 * C, w.embed_w, D, w.embed_b);
  ge.out = y; ge.ldc = D; ge.mode = M5_OUT_F32;
  M5_TRY(run_gemm(ctx, ge));
  // backbone.norm: AdaLayerNorm(bandwidth_id) -> x (fp32 residual stream)
  NormCall n0;
  n0.x = y; n0.M = N; n0.D = D; n0.ldx = D; n0.gamma = w.n_scale + (size_t)bandwidth_id * D; n0.beta = w.n_shift + (size_t)bandwidth_id * D;
  n0.eps = eps; n0.out_f32 = x; n0.ldo = D;
  M5_TRY(run_norm(ctx, n0));
  for (int l = 0; l < c.voc_layers; ++l) {
    const VocLayerW& lw = w.layers[l];
    if (voc_dwconv(x, lw.dw_w, lw.dw_b, d_fpos, d_flen, y, N, D, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "voc_dwconv");
    ctx->launches++;
    M5_TRY(ln_split(ctx, y, N, D, lw.n_scale + (size_t)bandwidth_id * D, lw.n_shift + (size_t)bandwidth_id * D, eps, h16));
    GemmCall g1 = lin3(h16, N, D, lw.pw1_w, I, lw.pw1_b);
    g1.act = M5_ACT_GELU; g1.mode = M5_OUT_F16_SPLIT; g1.out = g16; g1.out_lo = g16 + I; g1.ldc = 2 * I;
    M5_TRY(run_gemm(ctx, g1));
    GemmCall g2 = lin3(g16, N, I, lw.pw2_w, D, lw.pw2_b);
    g2.colscale = lw.gamma; g2.mode = M5_OUT_F32; g2.out = x; g2.ldc = D; g2.accumulate = 1;
    M5_TRY(run_gemm(ctx, g2));
  }
  M5_TRY(ln_split(ctx, x, N, D, w.fln_w, w.fln_b, eps, h16));
  GemmCall gh = lin3(h16, N, D, w.head_w, NB, w.head_b);
  gh.out = spec; gh.ldc = ldspec; gh.mode = M5_OUT_F32;
  M5_TRY(run_gemm(ctx, gh));
  M5_TRY(istft_run(ctx, w, spec, ldspec, B, nf, frames, d_f0, d_nf, d_wav));
  std::vector<long long> h_bounds;
  long long* d_bounds = nullptr;
  if (do_trim) {
    // frames of every utterance: 1 + (n + 2 * (frame_length / 2) - frame_length) / hop   (centered, reflect padded)
    std::vector<long long> off(B + 1, 0);
    std::vector<int> fr0(B + 1, 0);
    for (int b = 0; b < B; ++b) {
      const long long n = (long long)nf[b] * 320;
      if (n <= frame_length / 2) return ctx->fail(M5_ERR_ARG, "silence trim: a waveform must be longer than frame_length / 2 samples");
      off[b + 1] = off[b] + n;
      fr0[b + 1] = fr0[b] + (int)(1 + (n + 2 * (frame_length / 2) - frame_length) / hop_length);
    }
    long long* d_off = ar.get<long long>(B + 1); int* d_fr0 = ar.get<int>(B + 1);
    double* d_pow = ar.get<double>(fr0[B]);
    d_bounds = ar.get<long long>((size_t)2 * B);
    if (!d_off || !d_fr0 || !d_pow || !d_bounds) return ctx->fail(M5_ERR_NOMEM, "arena too small (trim)");
    cudaMemcpyAsync(d_off, off.data(), (B + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_fr0, fr0.data(), (B + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    trim_power_kernel<<<(fr0[B] + 7) / 8, 256, 0, ctx->stream>>>(d_wav, d_off, d_fr0, B, frame_length, hop_length, d_pow);
    trim_bounds_kernel<<<B, 256, 0, ctx->stream>>>(d_pow, d_off, d_fr0, hop_length, (double)top_db, d_bounds, d_bounds + B);
    ctx->launches += 2;
    h_bounds.resize((size_t)2 * B);
    M5_CUDA(cudaMemcpyAsync(h_bounds.data(), d_bounds, (size_t)2 * B * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (mem == M5_MEM_HOST) M5_CUDA(cudaMemcpyAsync(wav_out, d_wav, (size_t)N * 320 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  M5_CUDA(cudaStreamSynchronize(ctx->stream));
  if (do_trim)
    for (int b = 0; b < B; ++b) { t_start[b] = h_bounds[b]; t_end[b] = h_bounds[B + b]; }
  return M5_OK;
}

int m5_vocode(m5_ctx* ctx, int32_t B, const int32_t* codes, const int32_t* n_frames, int32_t bandwidth_id, int32_t mem,
              float* wav_out) {
  return vocode_impl(ctx, B, codes, n_frames, bandwidth_id, mem, wav_out, false, 0.f, 0, 0, nullptr, nullptr);
}

int m5_vocode_trim(m5_ctx* ctx, int32_t B, const int32_t* codes, const int32_t* n_frames, int32_t bandwidth_id, int32_t mem,
                   float top_db, int32_t frame_length, int32_t hop_length, float* wav_out, int64_t* start, int64_t* end) {
  if (!start || !end || frame_length < 2 || hop_length < 1 || top_db < 0.f) return M5_ERR_ARG;
  return vocode_impl(ctx, B, codes, n_frames, bandwidth_id, mem, wav_out, true, top_db, frame_length, hop_length, start, end);
}

// ------------------------------------------------------------------------------------------------ debug entry points
int m5_dbg_istft(m5_ctx* ctx, const float* spec, int32_t B, const int32_t* n_frames_host, float* wav) {
  if (!ctx || B <= 0) return M5_ERR_ARG;
  ctx->last_error.clear();
  VocWeights w;
  w.w1280 = W<float2>(ctx, "voc.w1280"); w.w640 = W<float2>(ctx, "voc.w640"); w.window = W<float>(ctx, "voc.window");
  if (!w.w1280 || !w.w640 || !w.window) return M5_ERR_MISSING_WEIGHT;
  std::vector<int> nf(n_frames_host, n_frames_host + B), f0(B);
  int N = 0;
  for (int b = 0; b < B; ++b) { f0[b] = N; N += nf[b]; }
  Arena ar(ctx);
  M5_TRY(ar.reserve((size_t)N * 1280 * 4 + 8 * B + (1 << 20)));
  int* d_f0 = ar.get<int>(B); int* d_nf = ar.get<int>(B);
  cudaMemcpyAsync(d_f0, f0.data(), B * 4, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(d_nf, nf.data(), B * 4, cudaMemcpyHostToDevice, ctx->stream);
  float* frames = ar.get<float>((size_t)N * 1280);
  M5_TRY(istft_run(ctx, w, spec, 1282, B, nf, frames, d_f0, d_nf, wav));
  M5_CUDA(cudaStreamSynchronize(ctx->stream));
  return M5_OK;
}

int m5_dbg_skinny(m5_ctx* ctx, const void* X, const void* Wt, int32_t B, int32_t N, int32_t K, float* out_f32, void* out_f16,
                  int32_t ldc, int32_t swiglu, int32_t accumulate) {
  if (!ctx) return M5_ERR_ARG;
  SkinnyCall s;
  s.X = (const __half*)X; s.W = (const __half*)Wt; s.B = B; s.N = N; s.K = K; s.out_f32 = out_f32; s.out_f16 = (__half*)out_f16;
  s.ldc = ldc; s.swiglu = swiglu; s.accumulate = accumulate; s.scratch = ctx->skinny_scratch; s.counters = ctx->skinny_counters;
  int r = gemm_skinny(s, ctx->stream, ctx->num_sms);
  if (r != M5_OK) return ctx->fail(r, "gemm_skinny failed");
  ctx->launches++;
  return M5_OK;
}

int m5_dbg_sample(m5_ctx* ctx, const float* logits, int32_t B, int32_t V, const m5_ar_cfg* cfg, int32_t text_vocab,
                  const int32_t* hist, int32_t hist_stride, const int32_t* n_gen, const int32_t* n_phones,
                  const float* noise, uint64_t seed, int32_t* out_tok, float* out_logprobs) {
  if (!ctx || !cfg) return M5_ERR_ARG;
  SampleCall sc;
  sc.logits = logits; sc.ld_logits = V; sc.B = B; sc.V = V; sc.text_vocab = text_vocab; sc.cfg = *cfg; sc.hist = hist;
  sc.hist_stride = hist_stride; sc.hist_is_ids = 0; sc.n_gen = const_cast<int*>(n_gen); sc.n_phones = n_phones;
  sc.noise = noise; sc.noise_steps = 1; sc.seed = seed; sc.out_tok = out_tok; sc.out_logprobs = out_logprobs;
  if (ar_sample(sc, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "ar_sample failed");
  ctx->launches++;
  return M5_OK;
}

int m5_dbg_posterior(m5_ctx* ctx, const float* cond, const float* uncond, int32_t R, int32_t t, const float* sched6,
                     float guidance_w, float x0_temp, const int32_t* x_t, const int32_t* x_known, const uint8_t* mask,
                     const float* u_unknown, const float* u_known, uint64_t seed, int32_t* x_out) {
  if (!ctx) return M5_ERR_ARG;
  if (!sched6) return M5_ERR_ARG;
  const float* sp = sched6;  // HOST: log_alpha[t], log_1_min_alpha[t], log_cumprod[t-1], log_1_min_cumprod[t-1], log_cumprod[t], log_1_min_cumprod[t]
  const int Q = ctx->cfg.n_quant, K = ctx->cfg.n_classes;
  M5_CUDA(cudaMemcpyAsync(x_out, x_t, (size_t)R * Q * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  PosteriorCall pc;
  pc.cond = cond; pc.uncond = (guidance_w != 1.0f) ? uncond : nullptr; pc.ld = K; pc.R = R; pc.K = K; pc.Q = Q; pc.q = -1;
  pc.guidance_w = guidance_w; pc.x0_temp = x0_temp; pc.log_alpha_t = sp[0]; pc.log_1m_alpha_t = sp[1];
  pc.log_cum_tm1 = sp[2]; pc.log_1m_cum_tm1 = sp[3]; pc.t = t;
  pc.x_t = x_t; pc.x_out = x_out; pc.u = u_unknown; pc.u_rows_are_x = 1; pc.seed = seed;
  if (nar_posterior(pc, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "nar_posterior failed");
  RenoiseCall rc;
  rc.R = R; rc.Q = Q; rc.K = K; rc.x_known = x_known; rc.known = mask; rc.x = x_out; rc.x_q0 = nullptr;
  rc.log_cum_t = sp[4]; rc.log_1m_cum_t = sp[5]; rc.t = t; rc.q0_override = 0; rc.u = u_known; rc.seed = seed;
  if (nar_renoise(rc, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "nar_renoise failed");
  ctx->launches += 2;
  return M5_OK;
}

}  // extern "C"
