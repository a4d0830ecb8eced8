// AR path: CodecLM.forward (mars5/model.py:95-141) + MistralTransformer (mars5/nn_future.py:336-398) and the decode loop
// ar_generate (mars5/ar_generate.py:15-165), batched over B independent utterances.
//
//   speaker encoder  : once per utterance (the reference recomputes it every step with an identical result,
//                      model.py:109-127)
//   prefill          : packed causal pass over [spk slot, prompt ids] with the tcgen05 GEMM + flash attention; K (after
//                      RoPE) and V land in the fp16 KV cache; only the last row of each sequence is projected to logits
//   decode           : one CUDA graph per step (26 x {RMSNorm, skinny QKV GEMM, RoPE+append, split-KV attention,
//                      skinny WO, RMSNorm, skinny W1|W3+SwiGLU, skinny W2} + final norm + vocabulary GEMM + fused
//                      sampler).  All per-row state (ids, lengths, done flags) lives on the device; the host only polls
//                      an "all rows done" counter every cfg.sync_every steps.
// Numerics mirror the reference's GPU path (fp16 GEMM operands / fp16 KV cache / fp32 residual stream).
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "ar_decode.h"
#include "layers.h"
#include "ptx.cuh"
#include "sampler.h"

namespace m5 {

struct ArLayerW { const float *attn_norm, *ffn_norm; const __half *wqkv, *wo, *w13, *w2; };
struct ArWeights {
  const __half* embed; const __half* output; const float* norm;
  std::vector<ArLayerW> layers;
  const __half* spk_tables; const float* spk_identity;
  std::vector<EncLayerW> spk; const float *spk_nw, *spk_nb;
  const float* pe;
  const float* inv_freq;
};

static int load_ar(m5_ctx* ctx, ArWeights& w) {
  const m5_model_cfg& c = ctx->cfg;
#define GETW(dst, T, name) do { dst = W<T>(ctx, name); if (!(dst)) return M5_ERR_MISSING_WEIGHT; } while (0)
  GETW(w.embed, __half, "ar.embed"); GETW(w.output, __half, "ar.output"); GETW(w.norm, float, "ar.norm");
  w.layers.resize(c.ar_layers);
  for (int i = 0; i < c.ar_layers; ++i) {
    const std::string p = "ar.l" + std::to_string(i) + ".";
    GETW(w.layers[i].attn_norm, float, p + "attn_norm"); GETW(w.layers[i].ffn_norm, float, p + "ffn_norm");
    GETW(w.layers[i].wqkv, __half, p + "wqkv"); GETW(w.layers[i].wo, __half, p + "wo");
    GETW(w.layers[i].w13, __half, p + "w13"); GETW(w.layers[i].w2, __half, p + "w2");
  }
  GETW(w.spk_tables, __half, "ar.spk.tables"); GETW(w.spk_identity, float, "ar.spk.identity");
  w.spk.resize(c.ar_spk_layers);
  for (int i = 0; i < c.ar_spk_layers; ++i) M5_TRY(load_enc_layer(ctx, "ar.spk.l" + std::to_string(i) + ".", w.spk[i]));
  GETW(w.spk_nw, float, "ar.spk.norm_w"); GETW(w.spk_nb, float, "ar.spk.norm_b");
  GETW(w.pe, float, "tab.pe_ar");
  const float* f = W<float>(ctx, "tab.rope_inv_freq");
  w.inv_freq = f ? f : ctx->rope_inv_freq;
  ctx->last_error.clear();
#undef GETW
  return M5_OK;
}

template <typename T>
static T* upload(m5_ctx* ctx, Arena& ar, const std::vector<T>& v) {
  T* d = ar.get<T>(v.size() ? v.size() : 1);
  if (d && !v.empty()) cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
  return d;
}
static const int* to_dev_ints(m5_ctx* ctx, Arena& ar, const int32_t* src, size_t n, int mem) {
  if (mem == M5_MEM_DEVICE || !src) return src;
  int* d = ar.get<int>(n ? n : 1);
  if (d && n) cudaMemcpyAsync(d, src, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
  return d;
}

// construct_padding_mask(spk_reference[:, :, 0], 1024) (model.py:119-125, utils.py:41-42): every key from the first pad
// code onwards is hidden; the identity slot is always visible.
__global__ void spk_klen_kernel(const int* codes, int Q, const int* code_off, const int* spk_len, int pad, int* klen, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int k = spk_len[b];
  for (int i = 0; i < spk_len[b]; ++i)
    if (codes[(size_t)(code_off[b] + i) * Q] == pad) { k = i; break; }
  klen[b] = 1 + k;
}

// Writes the prompt into the id buffer and initialises the per-row state.
__global__ void ar_init_state_kernel(const int* prompt, const int* p_off, const int* p_len, int B, int stride, int* ids,
                                     int* tok_len, int* kv_len, int* n_gen, int* done, int* n_done) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < p_len[b]; i += blockDim.x) ids[(size_t)b * stride + i] = prompt[p_off[b] + i];
  if (threadIdx.x == 0) {
    tok_len[b] = p_len[b];
    kv_len[b] = p_len[b] + 1;  // spk slot + prompt tokens are cached after the prefill
    n_gen[b] = 0;
    done[b] = 0;
    if (b == 0) *n_done = 0;
  }
}

// Embedding of the most recent token of every row -> x [B, D] fp32 (nn.Embedding is not autocast: fp32 value of the
// fp16-exact weight).
__global__ void ar_embed_last_kernel(const int* ids, int stride, const int* tok_len, const __half* table, int D, float* x) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int tok = ids[(size_t)b * stride + tok_len[b] - 1];
  for (int c = threadIdx.x; c < D; c += blockDim.x) x[(size_t)b * D + c] = __half2float(table[(size_t)tok * D + c]);
}

static int rms_to_f16(m5_ctx* ctx, const float* x, int rows, int D, const float* g, float eps, __half* out, const int* row_map) {
  NormCall n;
  n.x = x; n.M = rows; n.D = D; n.ldx = D; n.gamma = g; n.eps = eps; n.rms = 1; n.out = out; n.ldo = D; n.row_map = row_map;
  return run_norm(ctx, n);
}

struct ArPlan {
  int B = 0;
  std::vector<int> P, Pf, p_off, c_off;
  SeqSet spk, seq;
  int *spk_code_row, *spk_pos, *spk_first, *spk_klen;
  int *tok, *pos, *row_seq, *last_row;
};

static int ar_build_plan(m5_ctx* ctx, Arena& ar, ArPlan& p, int B, const int* prompt_len, const int* spk_len) {
  p.B = B;
  p.P.assign(prompt_len, prompt_len + B); p.Pf.assign(spk_len, spk_len + B);
  p.p_off.resize(B); p.c_off.resize(B);
  std::vector<int> code_row, spos, sstart, slen, sfirst, tok, pos, rseq, qstart, qlen, last;
  int r = 0, co = 0, mx = 0;
  for (int b = 0; b < B; ++b) {
    if (1 + p.Pf[b] > ctx->cfg.max_pos || p.Pf[b] < 0 || p.P[b] <= 0)
      return ctx->fail(M5_ERR_ARG, "utterance " + std::to_string(b) + ": speaker reference of " + std::to_string(p.Pf[b]) +
                                       " frames exceeds m5_model_cfg.max_pos = " + std::to_string(ctx->cfg.max_pos) + " (or empty prompt)");
    p.c_off[b] = co;
    sstart.push_back(r); slen.push_back(1 + p.Pf[b]); sfirst.push_back(r);
    code_row.push_back(-1); spos.push_back(0);
    for (int i = 0; i < p.Pf[b]; ++i) { code_row.push_back(co + i); spos.push_back(i + 1); }
    co += p.Pf[b]; r += 1 + p.Pf[b]; mx = std::max(mx, 1 + p.Pf[b]);
  }
  p.spk.n = B; p.spk.rows = r; p.spk.max_len = mx;
  int pr = 0, po = 0, pmx = 0;
  for (int b = 0; b < B; ++b) {
    p.p_off[b] = po;
    qstart.push_back(pr); qlen.push_back(1 + p.P[b]);
    tok.push_back(-(b + 1)); pos.push_back(0); rseq.push_back(b);
    for (int i = 0; i < p.P[b]; ++i) { tok.push_back(po + i); pos.push_back(i + 1); rseq.push_back(b); }
    po += p.P[b]; pr += 1 + p.P[b]; pmx = std::max(pmx, 1 + p.P[b]);
    last.push_back(pr - 1);
  }
  p.seq.n = B; p.seq.rows = pr; p.seq.max_len = pmx;
  for (int b = 0; b < B; ++b) { p.seq.self_pairs += (double)(1 + p.P[b]) * (1 + p.P[b]); p.spk.self_pairs += (double)(1 + p.Pf[b]) * (1 + p.Pf[b]); }
  p.spk_code_row = upload(ctx, ar, code_row); p.spk_pos = upload(ctx, ar, spos); p.spk_first = upload(ctx, ar, sfirst);
  int* d_sstart = upload(ctx, ar, sstart); int* d_slen = upload(ctx, ar, slen);
  p.spk.start = d_sstart; p.spk.len = d_slen;
  p.spk_klen = ar.get<int>(B);
  p.tok = upload(ctx, ar, tok); p.pos = upload(ctx, ar, pos); p.row_seq = upload(ctx, ar, rseq);
  p.last_row = upload(ctx, ar, last);
  int* d_qstart = upload(ctx, ar, qstart); int* d_qlen = upload(ctx, ar, qlen);
  p.seq.start = d_qstart; p.seq.len = d_qlen;
  if (!p.last_row || !d_qlen || !p.spk_klen) return ctx->fail(M5_ERR_NOMEM, "arena too small for AR plan");
  return M5_OK;
}

__global__ void resolve_tokens2_kernel(int* tok, const int* ids, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && tok[i] >= 0) tok[i] = ids[tok[i]];
}

// speaker vectors [B, D] (model.py:109-127)
static int ar_speaker(m5_ctx* ctx, const ArWeights& w, ArPlan& p, const int* d_codes, const int* d_coff, const int* d_spklen,
                      float* spk_x, float* spk_vec, const BlockScratch& bs) {
  const m5_model_cfg& c = ctx->cfg;
  spk_klen_kernel<<<(p.B + 127) / 128, 128, 0, ctx->stream>>>(d_codes, c.n_quant, d_coff, d_spklen, c.n_classes - 1, p.spk_klen, p.B);
  ctx->launches++;
  p.spk.klen = p.spk_klen;
  EmbedCall e;
  e.codes = d_codes; e.code_row = p.spk_code_row; e.pos = p.spk_pos; e.tables = w.spk_tables; e.identity = w.spk_identity;
  e.pe = w.pe; e.alpha = c.ar_pos_alpha; e.n_rows = p.spk.rows; e.D = c.ar_dim; e.Q = c.n_quant; e.n_codes = c.n_classes;
  e.out = spk_x;
  if (chunked_embed(e, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "chunked_embed(ar spk) failed");
  ctx->launches++;
  for (int l = 0; l < c.ar_spk_layers; ++l)
    M5_TRY(encoder_layer(ctx, spk_x, p.spk, w.spk[l], c.ar_dim, c.ar_heads, c.ar_spk_ff, c.ln_eps, M5_NUM_PRECISE, bs));
  NormCall n;
  n.x = spk_x; n.M = p.B; n.D = c.ar_dim; n.ldx = c.ar_dim; n.gamma = w.spk_nw; n.beta = w.spk_nb; n.eps = c.ln_eps;
  n.out_f32 = spk_vec; n.ldo = c.ar_dim; n.row_map = p.spk_first;
  return run_norm(ctx, n);
}

// Packed causal pass over all prompt rows.  K (after RoPE) and V are written to the cache [layer][seq][pos][D]; the
// attention of the pass itself reads them back from there (sequence b's keys start at cache row b*Wc).
// layer_stride == 0 reuses one scratch cache layer (m5_ar_forward).  x is [rows, D] fp32, updated in place.
static int ar_prefill_trunk(m5_ctx* ctx, const ArWeights& w, const ArPlan& p, float* x, __half* kc, __half* vc, int Wc,
                            size_t layer_stride, const int* kc_start, const BlockScratch& bs) {
  const m5_model_cfg& c = ctx->cfg;
  const int D = c.ar_dim, F = c.ar_hidden, rows = p.seq.rows;
  for (int l = 0; l < c.ar_layers; ++l) {
    const ArLayerW& lw = w.layers[l];
    M5_TRY(rms_to_f16(ctx, x, rows, D, lw.attn_norm, c.ar_norm_eps, bs.h16, nullptr));
    GemmCall gq;
    gq.A = bs.h16; gq.W = lw.wqkv; gq.M = rows; gq.N = 3 * D; gq.K = D; gq.lda = D; gq.ldw = D; gq.out = bs.qkv16;
    gq.ldc = 3 * D; gq.mode = M5_OUT_F16;
    M5_TRY(run_gemm(ctx, gq));
    __half* kcl = kc + (size_t)l * layer_stride;
    __half* vcl = vc + (size_t)l * layer_stride;
    if (rope_kv(bs.qkv16, 3 * D, rows, c.ar_heads, p.row_seq, p.pos, kcl, vcl, Wc, w.inv_freq, ctx->stream) != M5_OK)
      return ctx->fail(M5_ERR_CUDA, "rope_kv failed");
    ctx->launches++;
    AttnCall a;
    a.Q = bs.qkv16; a.ldq = 3 * D; a.K = kcl; a.V = vcl; a.ldk = a.ldv = D; a.O = bs.att16; a.ldo = D;
    a.n_heads = c.ar_heads; a.n_seqs = p.B; a.max_q = p.seq.max_len; a.q_start = p.seq.start; a.q_len = p.seq.len;
    a.k_start = kc_start; a.k_len = p.seq.len; a.causal = 1;
    a.flops_hint = 128.0 * c.ar_heads * p.seq.self_pairs;
    M5_TRY(run_attn(ctx, a));
    GemmCall go;
    go.A = bs.att16; go.W = lw.wo; go.M = rows; go.N = D; go.K = D; go.lda = D; go.ldw = D; go.out = x; go.ldc = D;
    go.mode = M5_OUT_F32; go.accumulate = 1;
    M5_TRY(run_gemm(ctx, go));
    M5_TRY(rms_to_f16(ctx, x, rows, D, lw.ffn_norm, c.ar_norm_eps, bs.h16, nullptr));
    GemmCall g1;
    g1.A = bs.h16; g1.W = lw.w13; g1.M = rows; g1.N = 2 * F; g1.K = D; g1.lda = D; g1.ldw = D; g1.out = bs.g16; g1.ldc = F;
    g1.mode = M5_OUT_SWIGLU_F16;
    M5_TRY(run_gemm(ctx, g1));
    GemmCall g2;
    g2.A = bs.g16; g2.W = lw.w2; g2.M = rows; g2.N = D; g2.K = F; g2.lda = F; g2.ldw = F; g2.out = x; g2.ldc = D;
    g2.mode = M5_OUT_F32; g2.accumulate = 1;
    M5_TRY(run_gemm(ctx, g2));
  }
  return M5_OK;
}

static int run_skinny(m5_ctx* ctx, SkinnyCall s) {
  s.scratch = ctx->skinny_scratch; s.counters = ctx->skinny_counters;
  int r = gemm_skinny(s, ctx->stream, ctx->num_sms);
  if (r != M5_OK) return ctx->fail(r, "gemm_skinny failed (N=" + std::to_string(s.N) + " K=" + std::to_string(s.K) + ")");
  ctx->launches++;
  return M5_OK;
}

struct ArState {
  int *ids, *tok_len, *kv_len, *n_gen, *done, *n_done;
  float *x, *qkv32, *logits;
  __half *h16, *q16, *att16, *g16;
  float* attn_scratch;
  int n_split;
};

// One decode step for all B rows (enqueued on ctx->stream; captured into a CUDA graph by the caller).
static int ar_decode_step(m5_ctx* ctx, const ArWeights& w, int B, const ArState& st, __half* kc, __half* vc, int Wc,
                          SampleCall& sc) {
  const m5_model_cfg& c = ctx->cfg;
  const int D = c.ar_dim, F = c.ar_hidden, V = c.ar_vocab;
  const size_t layer_stride = (size_t)B * Wc * D;
  if (launch_k(ar_embed_last_kernel, dim3(B), dim3(256), 0, ctx->stream, (const int*)st.ids, sc.hist_stride, (const int*)st.tok_len, w.embed, D, st.x) != cudaSuccess)
    return ctx->fail(M5_ERR_CUDA, "ar_embed_last launch failed");
  ctx->launches++;
  for (int l = 0; l < c.ar_layers; ++l) {
    const ArLayerW& lw = w.layers[l];
    M5_TRY(rms_to_f16(ctx, st.x, B, D, lw.attn_norm, c.ar_norm_eps, st.h16, nullptr));
    SkinnyCall q;
    q.X = st.h16; q.W = lw.wqkv; q.B = B; q.N = 3 * D; q.K = D; q.out_f32 = st.qkv32; q.ldc = 3 * D;
    M5_TRY(run_skinny(ctx, q));
    __half* kcl = kc + (size_t)l * layer_stride;
    __half* vcl = vc + (size_t)l * layer_stride;
    if (rope_kv_decode(st.qkv32, B, c.ar_heads, st.kv_len, st.q16, kcl, vcl, Wc, w.inv_freq, nullptr, ctx->stream) != M5_OK)
      return ctx->fail(M5_ERR_CUDA, "rope_kv_decode failed");
    ctx->launches++;
    DecodeAttnCall a;
    a.q = st.q16; a.kc = kcl; a.vc = vcl; a.B = B; a.H = c.ar_heads; a.W = Wc; a.kv_len = st.kv_len; a.done = st.done;
    a.out = st.att16; a.scratch = st.attn_scratch; a.n_split = st.n_split;
    if (decode_attn(a, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "decode_attn failed");
    ctx->launches += 2;
    SkinnyCall o;
    o.X = st.att16; o.W = lw.wo; o.B = B; o.N = D; o.K = D; o.out_f32 = st.x; o.ldc = D; o.accumulate = 1;
    M5_TRY(run_skinny(ctx, o));
    M5_TRY(rms_to_f16(ctx, st.x, B, D, lw.ffn_norm, c.ar_norm_eps, st.h16, nullptr));
    SkinnyCall f1;
    f1.X = st.h16; f1.W = lw.w13; f1.B = B; f1.N = 2 * F; f1.K = D; f1.out_f16 = st.g16; f1.ldc = F; f1.swiglu = 1;
    M5_TRY(run_skinny(ctx, f1));
    SkinnyCall f2;
    f2.X = st.g16; f2.W = lw.w2; f2.B = B; f2.N = D; f2.K = F; f2.out_f32 = st.x; f2.ldc = D; f2.accumulate = 1;
    M5_TRY(run_skinny(ctx, f2));
  }
  M5_TRY(rms_to_f16(ctx, st.x, B, D, w.norm, c.ar_norm_eps, st.h16, nullptr));
  SkinnyCall lo;
  lo.X = st.h16; lo.W = w.output; lo.B = B; lo.N = V; lo.K = D; lo.out_f32 = st.logits; lo.ldc = V;
  M5_TRY(run_skinny(ctx, lo));
  if (ar_sample(sc, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "ar_sample failed");
  ctx->launches++;
  return M5_OK;
}

}  // namespace m5

using namespace m5;

extern "C" {

int m5_ar_forward(m5_ctx* ctx, int32_t B, const int32_t* prompt_ids, const int32_t* prompt_len,
                  const int32_t* spk_codes, const int32_t* spk_len, int32_t mem, float* logits_out) {
  if (!ctx || B <= 0) return M5_ERR_ARG;
  ctx->last_error.clear();
  cudaSetDevice(ctx->device);
  ArWeights w;
  M5_TRY(load_ar(ctx, w));
  const m5_model_cfg& c = ctx->cfg;
  const int D = c.ar_dim, V = c.ar_vocab, Q = c.n_quant;
  size_t n_ids = 0, n_codes = 0; int maxP = 0;
  for (int b = 0; b < B; ++b) { n_ids += prompt_len[b]; n_codes += spk_len[b]; maxP = std::max(maxP, prompt_len[b]); }
  const int rows = (int)n_ids + B, spk_rows = (int)n_codes + B, Wc = maxP + 1;
  const int big = std::max(rows, spk_rows);
  const int ffmax = std::max(c.ar_hidden, c.ar_spk_ff);
  Arena ar(ctx);
  size_t bytes = block_scratch_bytes(big, 0, D, ffmax) + (size_t)spk_rows * D * 4 + (size_t)B * D * 4 + (size_t)rows * D * 4 +
                 (size_t)2 * B * Wc * D * 2 + (size_t)rows * (V + 4) * 4 * 2 + (n_ids + n_codes * Q) * 4 + (size_t)(rows + spk_rows) * 64 + (size_t(32) << 20);
  M5_TRY(ar.reserve(bytes));
  ArPlan p;
  M5_TRY(ar_build_plan(ctx, ar, p, B, prompt_len, spk_len));
  const int* d_ids = to_dev_ints(ctx, ar, prompt_ids, n_ids, mem);
  const int* d_codes = to_dev_ints(ctx, ar, spk_codes, n_codes * Q, mem);
  int* d_coff = upload(ctx, ar, p.c_off); int* d_spklen = upload(ctx, ar, p.Pf);
  std::vector<int> kcs(B); for (int b = 0; b < B; ++b) kcs[b] = b * Wc;
  int* d_kcs = upload(ctx, ar, kcs);
  float* spk_x = ar.get<float>((size_t)spk_rows * D); float* spk_vec = ar.get<float>((size_t)B * D);
  float* x = ar.get<float>((size_t)rows * D);
  __half* kc = ar.get<__half>((size_t)B * Wc * D); __half* vc = ar.get<__half>((size_t)B * Wc * D);
  const int ldV = (V + 3) & ~3;
  float* lg = ar.get<float>((size_t)rows * ldV);
  BlockScratch bs;
  block_scratch_carve(ar, bs, big, 0, D, ffmax);
  if (!d_ids || !d_codes || !d_kcs || !lg || !bs.g16) return ctx->fail(M5_ERR_NOMEM, "arena too small (ar_forward)");
  resolve_tokens2_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>(p.tok, d_ids, rows);
  M5_TRY(ar_speaker(ctx, w, p, d_codes, d_coff, d_spklen, spk_x, spk_vec, bs));
  TokEmbedCall te;
  te.tok = p.tok; te.pos = p.pos; te.table = w.embed; te.vec_rows = spk_vec; te.n_rows = rows; te.D = D; te.out = x;
  if (token_embed(te, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "token_embed failed");
  ctx->launches++;
  M5_TRY(ar_prefill_trunk(ctx, w, p, x, kc, vc, Wc, 0, d_kcs, bs));
  M5_TRY(rms_to_f16(ctx, x, rows, D, w.norm, c.ar_norm_eps, bs.h16, nullptr));
  GemmCall g;
  g.A = bs.h16; g.W = w.output; g.M = rows; g.N = V; g.K = D; g.lda = D; g.ldw = D; g.out = lg; g.ldc = ldV; g.mode = M5_OUT_F32;
  M5_TRY(run_gemm(ctx, g));
  // strip the speaker slot of every sequence (model.py:138-139)
  size_t out_off = 0; int row0 = 0;
  for (int b = 0; b < B; ++b) {
    M5_CUDA(cudaMemcpy2DAsync(logits_out + out_off, (size_t)V * sizeof(float), lg + (size_t)(row0 + 1) * ldV, (size_t)ldV * sizeof(float), (size_t)V * sizeof(float), (size_t)p.P[b],
                            mem == M5_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
    out_off += (size_t)p.P[b] * V; row0 += 1 + p.P[b];
  }
  M5_CUDA(cudaStreamSynchronize(ctx->stream));
  return M5_OK;
}

int m5_ar_generate(m5_ctx* ctx, int32_t B, const int32_t* prompt_ids, const int32_t* prompt_len,
                   const int32_t* spk_codes, const int32_t* spk_len, const int32_t* n_phones_gen,
                   const m5_ar_cfg* cfg, int32_t mem, const float* noise, int32_t noise_steps, uint64_t seed,
                   const int64_t* utt_ids, int32_t* out_ids, int32_t* out_len, int32_t* hit_maxlen,
                   float* logits_dump, int32_t dump_steps) {
  if (!ctx || !cfg || B <= 0) return M5_ERR_ARG;
  ctx->last_error.clear();
  if (B > 32) return ctx->fail(M5_ERR_ARG, "m5_ar_generate keeps at most 32 utterances in flight per call (B = " + std::to_string(B) +
                                               "); split the batch (Mars5TTS.tts_batch does)");
  cudaSetDevice(ctx->device);
  ArWeights w;
  M5_TRY(load_ar(ctx, w));
  const m5_model_cfg& c = ctx->cfg;
  const int D = c.ar_dim, V = c.ar_vocab, Q = c.n_quant, F = c.ar_hidden, L = c.ar_layers;
  const int max_len = cfg->max_len;
  // The reference attends through a 3000-position sliding window (ar_generate.py:57, nn_future.py:381-392); the cache here
  // never rotates, which is identical as long as spk slot + max_len tokens fit the window.
  if (max_len + 1 > 3000)
    return ctx->fail(M5_ERR_ARG, "max_len " + std::to_string(max_len) + " exceeds the reference's 3000-position sliding window "
                                 "(ar_generate.py:57): longer generations are not implemented");
  size_t n_ids = 0, n_codes = 0; int maxP = 0, minP = 1 << 30;
  for (int b = 0; b < B; ++b) {
    n_ids += prompt_len[b]; n_codes += spk_len[b];
    maxP = std::max(maxP, prompt_len[b]); minP = std::min(minP, prompt_len[b]);
    if (prompt_len[b] >= max_len) return ctx->fail(M5_ERR_ARG, "prompt is not shorter than max_len");
  }
  const int rows = (int)n_ids + B, spk_rows = (int)n_codes + B;
  const int Wc = max_len + 1;  // spk slot + every token that can ever be fed (sliding window 3000 never wraps, ar_generate.py:57)
  const int big = std::max(rows, spk_rows);
  const int ffmax = std::max(F, c.ar_spk_ff);
  const int n_split = decode_attn_splits_for(Wc);
  Arena ar(ctx);
  const size_t cache_bytes = (size_t)2 * L * B * Wc * D * 2;
  const size_t noise_bytes = (noise && mem == M5_MEM_HOST) ? (size_t)B * noise_steps * V * 4 : 0;
  const size_t dump_bytes = (logits_dump && mem == M5_MEM_HOST) ? (size_t)B * dump_steps * V * 4 : 0;
  size_t bytes = block_scratch_bytes(big, 0, D, ffmax) + (size_t)spk_rows * D * 4 + (size_t)B * D * 4 + (size_t)rows * D * 4 + cache_bytes +
                 noise_bytes + dump_bytes + (size_t)B * (max_len + 64) * 4 + (size_t)B * (3 * D + V + 4 * D + F) * 4 +
                 decode_attn_scratch_bytes(B, c.ar_heads, n_split) + (n_ids + n_codes * Q) * 4 + (size_t)(rows + spk_rows) * 64 + (size_t(32) << 20) +
                 ar_decode_attn_floats(B, c.ar_heads, ar_decode_splits_for(Wc, 256)) * 4 + (size_t)(4 * ctx->num_sms + 64) * 32 * 128 * 4;
  M5_TRY(ar.reserve(bytes));
  ArPlan p;
  M5_TRY(ar_build_plan(ctx, ar, p, B, prompt_len, spk_len));
  const int* d_ids = to_dev_ints(ctx, ar, prompt_ids, n_ids, mem);
  const int* d_codes = to_dev_ints(ctx, ar, spk_codes, n_codes * Q, mem);
  int* d_coff = upload(ctx, ar, p.c_off); int* d_spklen = upload(ctx, ar, p.Pf);
  int* d_poff = upload(ctx, ar, p.p_off); int* d_plen = upload(ctx, ar, p.P);
  std::vector<int> kcs(B); for (int b = 0; b < B; ++b) kcs[b] = b * Wc;
  int* d_kcs = upload(ctx, ar, kcs);
  std::vector<int> nph(B, -1); if (n_phones_gen) nph.assign(n_phones_gen, n_phones_gen + B);
  int* d_nph = n_phones_gen ? upload(ctx, ar, nph) : nullptr;
  int64_t* d_utt = nullptr;
  if (utt_ids) { std::vector<int64_t> u(utt_ids, utt_ids + B); d_utt = upload(ctx, ar, u); }
  float* spk_x = ar.get<float>((size_t)spk_rows * D); float* spk_vec = ar.get<float>((size_t)B * D);
  float* x = ar.get<float>((size_t)rows * D);
  __half* kc = ar.get<__half>((size_t)L * B * Wc * D); __half* vc = ar.get<__half>((size_t)L * B * Wc * D);
  ArState st;
  st.ids = ar.get<int>((size_t)B * max_len); st.tok_len = ar.get<int>(B); st.kv_len = ar.get<int>(B); st.n_gen = ar.get<int>(B);
  st.done = ar.get<int>(B); st.n_done = ar.get<int>(1);
  st.x = ar.get<float>((size_t)B * D); st.qkv32 = ar.get<float>((size_t)B * 3 * D); st.logits = ar.get<float>((size_t)B * V);
  st.h16 = ar.get<__half>((size_t)B * std::max(D, F)); st.q16 = ar.get<__half>((size_t)B * D); st.att16 = ar.get<__half>((size_t)B * D);
  st.g16 = ar.get<__half>((size_t)B * F);
  st.attn_scratch = ar.get<float>(decode_attn_scratch_bytes(B, c.ar_heads, n_split) / 4);
  st.n_split = n_split;
  // fused persistent decode kernel (ar_decode.cu); M5_AR_DECODE_LEGACY=1 keeps the per-kernel graph of round 1 for A/B runs
  static const bool legacy_decode = getenv("M5_AR_DECODE_LEGACY") != nullptr;
  ArDecodeParams dp;
  dp.B = B; dp.n_layers = L; dp.D = D; dp.F = F; dp.H = c.ar_heads; dp.V = V; dp.Wc = Wc; dp.eps = c.ar_norm_eps;
  dp.final_norm = w.norm; dp.embed = w.embed; dp.inv_freq = w.inv_freq;
  dp.ids = st.ids; dp.ids_stride = max_len; dp.tok_len = st.tok_len; dp.kv_len = st.kv_len; dp.done = st.done;
  dp.x = st.x; dp.qkv = st.qkv32; dp.g16 = st.g16; dp.att16 = st.att16; dp.logits = st.logits; dp.kc = kc; dp.vc = vc;
  M5_TRY(ar_decode_plan(dp, ctx->num_sms));
  dp.g_out.W = w.output;
  dp.split_keys = ar_decode_split_keys(B, c.ar_heads, Wc, ctx->num_sms);
  dp.n_split = ar_decode_splits_for(Wc, dp.split_keys);
  {
    std::vector<ArLayerDev> hl(L);
    for (int l = 0; l < L; ++l) hl[l] = {w.layers[l].attn_norm, w.layers[l].ffn_norm, w.layers[l].wqkv, w.layers[l].wo, w.layers[l].w13, w.layers[l].w2};
    dp.layers = upload(ctx, ar, hl);
    dp.ssq = ar.get<float>((size_t)dp.ssq_tiles * 32);
    dp.attn_part = ar.get<float>(ar_decode_attn_floats(B, c.ar_heads, dp.n_split));
    dp.scratch = ar.get<float>(ar_decode_scratch_floats(dp));
    dp.counters = ar.get<int>(ar_decode_max_tiles(dp) + 1);
    dp.attn_tickets = ar.get<int>((size_t)B * c.ar_heads);
    if (dp.attn_tickets) cudaMemsetAsync(dp.attn_tickets, 0, (size_t)B * c.ar_heads * sizeof(int), ctx->stream);
    dp.gbar = reinterpret_cast<unsigned*>(ar.get<int>(4));
    if (!dp.layers || !dp.ssq || !dp.attn_part || !dp.scratch || !dp.counters || !dp.gbar || !dp.attn_tickets)
      return ctx->fail(M5_ERR_NOMEM, "arena too small (fused decode buffers)");
    cudaMemsetAsync(dp.counters, 0, (ar_decode_max_tiles(dp) + 1) * sizeof(int), ctx->stream);
    if (getenv("M5_AR_PROFILE")) {   // per-phase timeline of CTA 0 of the LAST decode step, printed to stderr after the loop
      dp.prof = reinterpret_cast<unsigned long long*>(ar.get<double>((size_t)2 * (5 * L + 2)));
      if (dp.prof) cudaMemsetAsync(dp.prof, 0, sizeof(double) * 2 * (5 * L + 2), ctx->stream);
    }
  }
  const float* d_noise = noise;
  if (noise && mem == M5_MEM_HOST) {
    float* dn = ar.get<float>((size_t)B * noise_steps * V);
    if (dn) cudaMemcpyAsync(dn, noise, (size_t)B * noise_steps * V * 4, cudaMemcpyHostToDevice, ctx->stream);
    d_noise = dn;
  }
  float* d_dump = logits_dump;
  if (logits_dump && mem == M5_MEM_HOST) d_dump = ar.get<float>((size_t)B * dump_steps * V);
  BlockScratch bs;
  block_scratch_carve(ar, bs, big, 0, D, ffmax);
  if (!d_ids || !d_codes || !d_kcs || !st.attn_scratch || !bs.g16 || (noise && !d_noise) || (logits_dump && !d_dump))
    return ctx->fail(M5_ERR_NOMEM, "arena too small (ar_generate)");
  resolve_tokens2_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>(p.tok, d_ids, rows);
  ar_init_state_kernel<<<B, 128, 0, ctx->stream>>>(d_ids, d_poff, d_plen, B, max_len, st.ids, st.tok_len, st.kv_len, st.n_gen,
                                                   st.done, st.n_done);
  ctx->launches += 2;
  // ---- speaker vectors + prefill
  M5_TRY(ar_speaker(ctx, w, p, d_codes, d_coff, d_spklen, spk_x, spk_vec, bs));
  TokEmbedCall te;
  te.tok = p.tok; te.pos = p.pos; te.table = w.embed; te.vec_rows = spk_vec; te.n_rows = rows; te.D = D; te.out = x;
  if (token_embed(te, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "token_embed failed");
  ctx->launches++;
  M5_TRY(ar_prefill_trunk(ctx, w, p, x, kc, vc, Wc, (size_t)B * Wc * D, d_kcs, bs));
  M5_TRY(rms_to_f16(ctx, x, B, D, w.norm, c.ar_norm_eps, st.h16, p.last_row));
  SkinnyCall lo;
  lo.X = st.h16; lo.W = w.output; lo.B = B; lo.N = V; lo.K = D; lo.out_f32 = st.logits; lo.ldc = V;
  M5_TRY(run_skinny(ctx, lo));
  SampleCall sc;
  sc.logits = st.logits; sc.ld_logits = V; sc.B = B; sc.V = V; sc.text_vocab = c.ar_text_vocab; sc.cfg = *cfg;
  sc.hist = st.ids; sc.hist_stride = max_len; sc.hist_is_ids = 1; sc.n_gen = st.n_gen; sc.n_phones = d_nph;
  sc.noise = d_noise; sc.noise_steps = noise_steps; sc.seed = seed; sc.utt_ids = d_utt;
  sc.logits_dump = d_dump; sc.dump_steps = dump_steps;
  sc.ids = st.ids; sc.tok_len = st.tok_len; sc.kv_len = st.kv_len; sc.done = st.done; sc.n_done = st.n_done;
  if (ar_sample(sc, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "ar_sample failed");
  ctx->launches++;
  // ---- decode loop: capture one step, replay
  const int max_steps = max_len - minP - 1;  // the first token was sampled by the prefill
  const int sync_every = cfg->sync_every > 0 ? cfg->sync_every : 16;
  cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
  int64_t step_launches = 0;
  if (max_steps > 0) {
    const int64_t before = ctx->launches;
    M5_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    static const bool no_pdl = getenv("M5_DISABLE_PDL") != nullptr;
    g_use_pdl = !no_pdl;  // kernels of the step overlap their prologues (weight prefetch) with their predecessor's tail
    int rc;
    if (legacy_decode) {
      rc = ar_decode_step(ctx, w, B, st, kc, vc, Wc, sc);
    } else {
      g_use_pdl = false;
      // the sampler is the kernel's last phase (one launch per decode step); M5_AR_SAMPLE_SEPARATE=1 keeps it a kernel of
      // its own behind the decode kernel (debugging aid)
      static const bool sample_separate = getenv("M5_AR_SAMPLE_SEPARATE") != nullptr;
      dp.fuse_sample = (!sample_separate && ar_decode_can_fuse_sampler(V, cfg->top_k)) ? 1 : 0;
      dp.sample = sc;   // sc.cap was filled by the prefill's ar_sample call above
      rc = ar_decode_launch(dp, ctx->num_sms, ctx->stream);
      if (rc != M5_OK) ctx->fail(rc, std::string("ar_decode_launch failed: ") + cudaGetErrorString(cudaGetLastError()));
      ctx->launches++;
      if (!dp.fuse_sample) {
        if (rc == M5_OK && ar_sample(sc, ctx->stream) != M5_OK) rc = ctx->fail(M5_ERR_CUDA, "ar_sample failed");
        ctx->launches++;
      }
    }
    g_use_pdl = false;
    cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    step_launches = ctx->launches - before;
    ctx->launches = before;
    if (rc != M5_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return ctx->fail(M5_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
    M5_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  }
  int h_done = 0, steps_run = 0;
  cudaEvent_t pe = prof_begin(ctx);   // kind 2: the whole decode loop (bench.py's AR HBM roofline)
  for (int s = 0; s < max_steps; ++s) {
    cudaError_t le = cudaGraphLaunch(exec, ctx->stream);
    if (le != cudaSuccess) { cudaGraphExecDestroy(exec); cudaGraphDestroy(graph); return ctx->fail(M5_ERR_CUDA, std::string("graph launch: ") + cudaGetErrorString(le)); }
    ctx->launches += step_launches;
    steps_run = s + 1;
    if ((s + 1) % sync_every == 0 || s + 1 == max_steps) {
      cudaMemcpyAsync(&h_done, st.n_done, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
      cudaError_t se = cudaStreamSynchronize(ctx->stream);
      if (se != cudaSuccess) { cudaGraphExecDestroy(exec); cudaGraphDestroy(graph); return ctx->fail(M5_ERR_CUDA, std::string("decode loop: ") + cudaGetErrorString(se)); }
      if (h_done >= B) break;
    }
  }
  if (pe) prof_end(ctx, pe, 2, 0.0, 0.0, steps_run);
  if (dp.prof && steps_run > 0 && !legacy_decode) {
    std::vector<unsigned long long> ts((size_t)2 * (5 * L + 2));
    cudaMemcpyAsync(ts.data(), dp.prof, ts.size() * 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    // stamp pairs: [0] after P0, then per layer after P1 (qkv), P2 (attn), P3 (wo), P4 (w13), P5 (w2); last = end of logits
    const char* nm[5] = {"qkv", "attn", "wo", "w13", "w2"};
    double work[5] = {0, 0, 0, 0, 0}, wait[5] = {0, 0, 0, 0, 0};
    for (int l = 0; l < L; ++l)
      for (int ph = 0; ph < 5; ++ph) {
        const size_t i = 1 + (size_t)l * 5 + ph;
        work[ph] += (double)(ts[2 * i] - ts[2 * (i - 1) + 1]);      // barrier passed -> own work done
        wait[ph] += (double)(ts[2 * i + 1] - ts[2 * i]);            // own work done -> barrier passed
      }
    fprintf(stderr, "m5 ar_decode profile (CTA 0, last step, B=%d, per layer avg us):", B);
    for (int ph = 0; ph < 5; ++ph) fprintf(stderr, "  %s work %.1f wait %.1f", nm[ph], work[ph] / L / 1e3, wait[ph] / L / 1e3);
    const size_t last = 1 + (size_t)5 * L;
    fprintf(stderr, "  | P0 %.1f  logits %.1f  total %.1f us\n", (double)(ts[1] - ts[0]) / 1e3 + 0.0, (double)(ts[2 * last] - ts[2 * (last - 1) + 1]) / 1e3,
            (double)(ts[2 * last] - ts[0]) / 1e3);
  }
  const bool prof_rec = pe && !ctx->prof_pending.empty() && ctx->prof_pending.back().kind == 2;
  if (exec) cudaGraphExecDestroy(exec);
  if (graph) cudaGraphDestroy(graph);
  // ---- results
  std::vector<int> h_len(B);
  M5_CUDA(cudaMemcpyAsync(h_len.data(), st.tok_len, B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  M5_CUDA(cudaMemcpyAsync(out_ids, st.ids, (size_t)B * max_len * sizeof(int),
                          mem == M5_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
  if (logits_dump && mem == M5_MEM_HOST)
    M5_CUDA(cudaMemcpyAsync(logits_dump, d_dump, (size_t)B * dump_steps * V * 4, cudaMemcpyDeviceToHost, ctx->stream));
  M5_CUDA(cudaStreamSynchronize(ctx->stream));
  if (prof_rec) {
    // algorithmic bytes of the decode steps: every step reads the fp16 weights once, and for every row
    // still running the K/V of all cached tokens (4*layers*D bytes each) plus the new token's K/V and embedding row
    const double w_ar = 2.0 * ((double)L * (4.0 * D * D + 3.0 * (double)F * D) + (2.0 * L + 1) * D + (double)V * D);
    const double kvb = 4.0 * L * D;
    double bytes = (double)steps_run * w_ar, flops = 0.0;
    for (int b = 0; b < B; ++b) {
      const int gen = h_len[b] - p.P[b];                    // tokens appended (the first one came from the prefill)
      const int act = std::max(0, std::min(steps_run, gen)); // decode steps in which row b was still running
      const double l0 = p.P[b] + 2;                         // spk slot + prompt + the token being fed at decode step 0
      bytes += kvb * (act * l0 + 0.5 * act * (act - 1.0)) + act * (kvb + 2.0 * D);
      flops += act * 2.0 * ((double)L * (4.0 * D * D + 3.0 * (double)F * D) + (double)V * D) + kvb * (act * l0 + 0.5 * act * (act - 1.0));
    }
    ctx->prof_pending.back().bytes = bytes;
    ctx->prof_pending.back().flops = flops;
  }
  for (int b = 0; b < B; ++b) {
    if (out_len) out_len[b] = h_len[b];
    if (hit_maxlen) hit_maxlen[b] = h_len[b] >= max_len - 1 ? 1 : 0;  // ar_generate.py:160-162
  }
  return M5_OK;
}

}  // extern "C"
