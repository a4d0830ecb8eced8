// NAR path: ResidualTransformer.forward (mars5/model.py:264-343) and the multinomial-DDPM reverse loop
// perform_simple_inference / reverse_diffusion (mars5/diffuser.py:345-472), batched over B independent utterances as
// packed variable-length sequences (every row equals a bs=1 reference run; no padding tokens exist).
//
// Structure per reverse step:   encoder (8 layers over 1+text rows, cond and uncond stacked)  ->  decoder (16 layers
// over S rows, cond and uncond stacked)  ->  final LayerNorm on the rows that are still unknown  ->  per-codebook head
// (LayerNorm + 1024->1025 GEMM)  ->  fused CFG + posterior + Gumbel-argmax (posterior.cu)  ->  known-region re-noise.
// Step-invariant work is hoisted: the speaker encoder runs once per utterance (its input does not depend on t; the
// unconditional speaker vector sees only the identity token and is therefore one length-1 sequence), and the timestep
// MLPs (model.py:206-215) are evaluated for all T steps in one GEMM pair up front.  Codebook 0 is never sampled from
// the model (mask m[...,0]=True, diffuser.py:419-420), so its head is skipped in the loop.
#include <math.h>

#include <algorithm>
#include <vector>

#include "layers.h"
#include "philox.cuh"
#include "sampler.h"

namespace m5 {

struct NarWeights {
  const __half* text_embed; const __half* ref_tables; const __half* res_tables; const float* spk_identity;
  std::vector<EncLayerW> spk, enc; std::vector<DecLayerW> dec;
  const float *spk_nw, *spk_nb, *enc_nw, *enc_nb, *dec_nw, *dec_nb;
  const __half *te_w0, *te_w1, *td_w0, *td_w1; const float *te_b0, *te_b1, *td_b0, *td_b1;
  const float* t_emb_table;  // [n_t, nar_dim] sinusoidal timestep embeddings (model.py:18-35), torch-computed
  int n_t = 0;
  const float *head_lnw[16], *head_lnb[16], *head_b[16]; const __half* head_w[16];
  const float* pe;  // [max_pos, D]
  float alpha_pos, alpha_cond, alpha_ref;
};

static int load_nar(m5_ctx* ctx, NarWeights& w) {
  const m5_model_cfg& c = ctx->cfg;
#define GETW(dst, T, name) do { dst = W<T>(ctx, name); if (!(dst)) return M5_ERR_MISSING_WEIGHT; } while (0)
  GETW(w.text_embed, __half, "nar.text_embed"); GETW(w.ref_tables, __half, "nar.ref.tables");
  GETW(w.res_tables, __half, "nar.res.tables"); GETW(w.spk_identity, float, "nar.spk_identity");
  w.spk.resize(c.nar_spk_layers); w.enc.resize(c.nar_enc_layers); w.dec.resize(c.nar_dec_layers);
  for (int i = 0; i < c.nar_spk_layers; ++i) M5_TRY(load_enc_layer(ctx, "nar.spk.l" + std::to_string(i) + ".", w.spk[i]));
  for (int i = 0; i < c.nar_enc_layers; ++i) M5_TRY(load_enc_layer(ctx, "nar.enc.l" + std::to_string(i) + ".", w.enc[i]));
  for (int i = 0; i < c.nar_dec_layers; ++i) M5_TRY(load_dec_layer(ctx, "nar.dec.l" + std::to_string(i) + ".", w.dec[i]));
  GETW(w.spk_nw, float, "nar.spk.norm_w"); GETW(w.spk_nb, float, "nar.spk.norm_b");
  GETW(w.enc_nw, float, "nar.enc.norm_w"); GETW(w.enc_nb, float, "nar.enc.norm_b");
  GETW(w.dec_nw, float, "nar.dec.norm_w"); GETW(w.dec_nb, float, "nar.dec.norm_b");
  GETW(w.te_w0, __half, "nar.t_enc.w0"); GETW(w.te_b0, float, "nar.t_enc.b0");
  GETW(w.te_w1, __half, "nar.t_enc.w1"); GETW(w.te_b1, float, "nar.t_enc.b1");
  GETW(w.td_w0, __half, "nar.t_dec.w0"); GETW(w.td_b0, float, "nar.t_dec.b0");
  GETW(w.td_w1, __half, "nar.t_dec.w1"); GETW(w.td_b1, float, "nar.t_dec.b1");
  const m5_tensor* tt = find_weight(ctx, "tab.t_emb");
  if (!tt) return M5_ERR_MISSING_WEIGHT;
  w.t_emb_table = (const float*)tt->ptr; w.n_t = (int)(tt->numel / c.nar_dim);
  for (int q = 0; q < c.n_quant; ++q) {
    const std::string p = "nar.head." + std::to_string(q) + ".";
    GETW(w.head_lnw[q], float, p + "ln_w"); GETW(w.head_lnb[q], float, p + "ln_b");
    GETW(w.head_w[q], __half, p + "w"); GETW(w.head_b[q], float, p + "b");
  }
  GETW(w.pe, float, "tab.pe_nar");
  w.alpha_pos = c.nar_pos_alpha; w.alpha_cond = c.nar_cond_alpha; w.alpha_ref = c.nar_ref_alpha;
#undef GETW
  return M5_OK;
}

// ------------------------------------------------------------------------------------------------ planning
struct NarPlan {
  int B = 0, npass = 1;
  std::vector<int> n_text, Pf, Nx, S, x_off;  // per utterance
  int Rx = 0;                                  // total x rows (sum S)
  int sum_Nx = 0;
  // device index arrays
  int *spk_code_row = nullptr, *spk_pos = nullptr, *spk_start = nullptr, *spk_len = nullptr, *spk_first = nullptr;
  int *enc_tok = nullptr, *enc_pos = nullptr, *enc_start = nullptr, *enc_len = nullptr;
  int *dec_xrow = nullptr, *dec_pos = nullptr, *dec_start = nullptr, *dec_len = nullptr;
  int *lg_decrow = nullptr, *lg_xrow = nullptr;   // rows that need logits: decoder row / x row (per pass block)
  int *x_pos = nullptr; int64_t* x_utt = nullptr;  // per x row
  SeqSet spk, enc, dec;
  int n_lg = 0;  // logits rows per pass
};

template <typename T>
static T* upload(m5_ctx* ctx, Arena& ar, const std::vector<T>& v) {
  T* d = ar.get<T>(v.size() ? v.size() : 1);
  if (d && !v.empty()) cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
  return d;
}

// Builds all index arrays. `deep`: x = cat(prompt, generated) (diffuser.py:423-436); `logits_all`: every row needs logits.
static int build_plan(m5_ctx* ctx, Arena& ar, NarPlan& p, int B, const int* c_text_len, const int* c_codes_len,
                      const int* x_len, bool deep, int npass, bool uncond_only, bool logits_all, const int64_t* utt_ids) {
  p.B = B; p.npass = npass;
  p.n_text.assign(c_text_len, c_text_len + B);
  p.Pf.assign(c_codes_len, c_codes_len + B);
  p.Nx.assign(x_len, x_len + B);
  p.S.resize(B); p.x_off.resize(B);
  p.Rx = 0; p.sum_Nx = 0;
  const int max_pos = ctx->cfg.max_pos;
  for (int b = 0; b < B; ++b) {
    p.S[b] = (deep ? p.Pf[b] : 0) + p.Nx[b];
    // the sinusoidal tables (tab.pe_nar) hold max_pos rows; the reference grows its table on demand
    // (SinePositionalEmbedding.extend_pe, nn_future.py:51-76) -- here a longer sequence is an error, not a silent overrun
    if (p.S[b] > max_pos || 1 + p.n_text[b] > max_pos || 1 + p.Pf[b] > max_pos || p.Nx[b] < 0 || p.Pf[b] < 0 || p.n_text[b] < 0)
      return ctx->fail(M5_ERR_ARG, "utterance " + std::to_string(b) + ": sequence of " + std::to_string(std::max(p.S[b], std::max(1 + p.n_text[b], 1 + p.Pf[b]))) +
                                       " positions exceeds m5_model_cfg.max_pos = " + std::to_string(max_pos) + " (rows of the positional tables)");
    p.x_off[b] = p.Rx;
    p.Rx += p.S[b];
    p.sum_Nx += p.Nx[b];
  }
  // speaker batch: B conditional sequences (identity + Pf codes) + one unconditional length-1 sequence
  std::vector<int> code_row, pos, start, len, first;
  int codes_off = 0, r = 0, mx = 0;
  for (int b = 0; b < B; ++b) {
    start.push_back(r); len.push_back(1 + p.Pf[b]); first.push_back(r);
    code_row.push_back(-1); pos.push_back(0);
    for (int i = 0; i < p.Pf[b]; ++i) { code_row.push_back(codes_off + i); pos.push_back(i + 1); }
    codes_off += p.Pf[b];
    r += 1 + p.Pf[b];
    mx = std::max(mx, 1 + p.Pf[b]);
  }
  start.push_back(r); len.push_back(1); first.push_back(r); code_row.push_back(-1); pos.push_back(0); r += 1;
  p.spk.n = B + 1; p.spk.rows = r; p.spk.max_len = mx;
  for (int v : len) p.spk.self_pairs += (double)v * v;
  p.spk_code_row = upload(ctx, ar, code_row); p.spk_pos = upload(ctx, ar, pos);
  p.spk_start = upload(ctx, ar, start); p.spk_len = upload(ctx, ar, len); p.spk_first = upload(ctx, ar, first);
  p.spk.start = p.spk_start; p.spk.len = p.spk_len;
  // encoder / decoder batches, pass-major
  std::vector<int> etok, epos, estart, elen, dx, dpos, dstart, dlen, lgd, lgx;
  int er = 0, dr = 0, emx = 0, dmx = 0;
  for (int ps = 0; ps < npass; ++ps) {
    const bool uncond = uncond_only || ps == 1;
    int text_off = 0;
    for (int b = 0; b < B; ++b) {
      estart.push_back(er); elen.push_back(1 + p.n_text[b]);
      etok.push_back(-((uncond ? B : b) + 1)); epos.push_back(0);
      for (int i = 0; i < p.n_text[b]; ++i) { etok.push_back(text_off + i); epos.push_back(i + 1); }  // index, resolved below
      text_off += p.n_text[b];
      er += 1 + p.n_text[b];
      emx = std::max(emx, 1 + p.n_text[b]);
      dstart.push_back(dr); dlen.push_back(p.S[b]);
      for (int i = 0; i < p.S[b]; ++i) {
        dx.push_back(p.x_off[b] + i); dpos.push_back(i);
        if (ps == 0 && (logits_all || i >= (deep ? p.Pf[b] : 0))) { lgd.push_back(dr + i); lgx.push_back(p.x_off[b] + i); }
      }
      dr += p.S[b];
      dmx = std::max(dmx, p.S[b]);
    }
  }
  p.n_lg = (int)lgd.size();
  if (npass == 2) {  // second block of logits rows points at the unconditional decoder rows
    const int half = dr / 2;
    for (int i = 0; i < p.n_lg; ++i) lgd.push_back(lgd[i] + half);
  }
  p.enc.n = npass * B; p.enc.rows = er; p.enc.max_len = emx;
  p.dec.n = npass * B; p.dec.rows = dr; p.dec.max_len = dmx;
  for (size_t i = 0; i < elen.size(); ++i) {
    p.enc.self_pairs += (double)elen[i] * elen[i];
    p.dec.self_pairs += (double)dlen[i] * dlen[i];
    p.dec.cross_pairs += (double)dlen[i] * elen[i];
  }
  p.enc_tok = upload(ctx, ar, etok); p.enc_pos = upload(ctx, ar, epos);
  p.enc_start = upload(ctx, ar, estart); p.enc_len = upload(ctx, ar, elen);
  p.dec_xrow = upload(ctx, ar, dx); p.dec_pos = upload(ctx, ar, dpos);
  p.dec_start = upload(ctx, ar, dstart); p.dec_len = upload(ctx, ar, dlen);
  p.lg_decrow = upload(ctx, ar, lgd); p.lg_xrow = upload(ctx, ar, lgx);
  p.enc.start = p.enc_start; p.enc.len = p.enc_len; p.dec.start = p.dec_start; p.dec.len = p.dec_len;
  std::vector<int> xp; std::vector<int64_t> xu;
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < p.S[b]; ++i) { xp.push_back(i); xu.push_back(utt_ids ? utt_ids[b] : (int64_t)b); }
  p.x_pos = upload(ctx, ar, xp); p.x_utt = upload(ctx, ar, xu);
  if (!p.spk_code_row || !p.enc_tok || !p.dec_xrow || !p.lg_decrow || !p.x_pos || !p.x_utt)
    return ctx->fail(M5_ERR_NOMEM, "arena too small for NAR plan");
  return M5_OK;
}

// enc_tok holds text INDICES (into the packed c_text array) for non-negative entries; turn them into token ids.
__global__ void resolve_tokens_kernel(int* tok, const int* c_text, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && tok[i] >= 0) tok[i] = c_text[tok[i]];
}

struct NarBuffers {
  float* spk_x = nullptr; float* spk_vec = nullptr;  // [spk rows, D], [B+1, D]
  float* t_enc = nullptr; float* t_dec = nullptr;    // [T, D]
  float* xe = nullptr; __half* mem16 = nullptr;      // encoder stream / memory
  float* xd = nullptr;                               // decoder stream
  float* dn = nullptr;                               // final-norm rows needing logits [npass*n_lg, D]
  __half* hd16 = nullptr;                            // head LN output
  float* logits = nullptr;                           // [npass*n_lg, ldl]
  int ldl = 0;
  BlockScratch bs;
};

static size_t nar_bytes(const m5_model_cfg& c, const NarPlan& p, int T, bool precise) {
  const size_t D = c.nar_dim;
  const int big = std::max(std::max(p.dec.rows, p.enc.rows), std::max(p.spk.rows, T));
  size_t b = 0;
  b += (size_t)p.spk.rows * D * 4 + (size_t)(p.B + 1) * D * 4 + 2 * (size_t)T * D * 4;
  b += (size_t)p.enc.rows * D * 4 + (size_t)p.enc.rows * 2 * D * 2;
  b += (size_t)p.dec.rows * D * 4;
  b += (size_t)p.npass * p.n_lg * D * 4 + (size_t)p.npass * p.n_lg * 2 * D * 2;
  b += (size_t)p.npass * p.n_lg * (c.n_classes + 7) * 4;
  b += block_scratch_bytes(big, p.enc.rows, (int)D, c.nar_ff);
  (void)precise;
  return b + (size_t(64) << 20);
}

static int carve(m5_ctx* ctx, Arena& ar, const NarPlan& p, int T, NarBuffers& nb) {
  const m5_model_cfg& c = ctx->cfg;
  const size_t D = c.nar_dim;
  nb.spk_x = ar.get<float>((size_t)p.spk.rows * D);
  nb.spk_vec = ar.get<float>((size_t)(p.B + 1) * D);
  nb.t_enc = ar.get<float>((size_t)T * D);
  nb.t_dec = ar.get<float>((size_t)T * D);
  nb.xe = ar.get<float>((size_t)p.enc.rows * D);
  nb.mem16 = ar.get<__half>((size_t)p.enc.rows * 2 * D);
  nb.xd = ar.get<float>((size_t)p.dec.rows * D);
  nb.dn = ar.get<float>((size_t)p.npass * p.n_lg * D);
  nb.hd16 = ar.get<__half>((size_t)p.npass * p.n_lg * 2 * D);
  nb.ldl = (c.n_classes + 3) & ~3;
  nb.logits = ar.get<float>((size_t)p.npass * p.n_lg * nb.ldl);
  const int big = std::max(std::max(p.dec.rows, p.enc.rows), std::max(p.spk.rows, T));
  block_scratch_carve(ar, nb.bs, big, p.enc.rows, (int)D, c.nar_ff);
  if (!nb.bs.kv16 || !nb.logits) return ctx->fail(M5_ERR_NOMEM, "arena too small for NAR buffers");
  return M5_OK;
}

// speaker vectors (model.py:298-310): once per call
static int nar_speaker(m5_ctx* ctx, const NarWeights& w, const NarPlan& p, const int* c_codes_dev, NarBuffers& nb) {
  const m5_model_cfg& c = ctx->cfg;
  EmbedCall e;
  e.codes = c_codes_dev; e.code_row = p.spk_code_row; e.pos = p.spk_pos; e.tables = w.ref_tables;
  e.identity = w.spk_identity; e.pe = w.pe; e.alpha = w.alpha_ref; e.n_rows = p.spk.rows; e.D = c.nar_dim;
  e.Q = c.n_quant; e.n_codes = c.n_classes; e.out = nb.spk_x;
  if (chunked_embed(e, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "chunked_embed(spk) failed");
  ctx->launches++;
  for (int l = 0; l < c.nar_spk_layers; ++l)
    M5_TRY(encoder_layer(ctx, nb.spk_x, p.spk, w.spk[l], c.nar_dim, c.nar_heads, c.nar_ff, c.ln_eps, M5_NUM_PRECISE, nb.bs));
  NormCall n;
  n.x = nb.spk_x; n.M = p.B + 1; n.D = c.nar_dim; n.ldx = c.nar_dim; n.gamma = w.spk_nw; n.beta = w.spk_nb;
  n.eps = c.ln_eps; n.out_f32 = nb.spk_vec; n.ldo = c.nar_dim; n.row_map = p.spk_first;
  return run_norm(ctx, n);
}

// timestep MLPs for steps 0..T-1 (model.py:315-317)
static int nar_timestep_tables(m5_ctx* ctx, const NarWeights& w, int T, NarBuffers& nb) {
  const int D = ctx->cfg.nar_dim;
  if (T > w.n_t) return ctx->fail(M5_ERR_ARG, "T exceeds the timestep-embedding table");
  __half* a16 = nb.bs.h16;  // [T, 2D]
  if (cast_rows(w.t_emb_table, D, a16, a16 + D, 2 * D, T, D, nullptr, ctx->stream) != M5_OK)
    return ctx->fail(M5_ERR_CUDA, "cast_rows failed");
  ctx->launches++;
  for (int which = 0; which < 2; ++which) {
    GemmCall g1;
    g1.A = a16; g1.W = which ? w.td_w0 : w.te_w0; g1.M = T; g1.N = D; g1.K = 2 * D; g1.lda = 2 * D; g1.ldw = D;
    g1.kwrap = D; g1.bias = which ? w.td_b0 : w.te_b0; g1.act = M5_ACT_SILU; g1.mode = M5_OUT_F16_SPLIT;
    g1.out = nb.bs.g16; g1.out_lo = nb.bs.g16 + D; g1.ldc = 2 * D;
    M5_TRY(run_gemm(ctx, g1));
    GemmCall g2;
    g2.A = nb.bs.g16; g2.W = which ? w.td_w1 : w.te_w1; g2.M = T; g2.N = D; g2.K = 2 * D; g2.lda = 2 * D; g2.ldw = D;
    g2.kwrap = D; g2.bias = which ? w.td_b1 : w.te_b1; g2.mode = M5_OUT_F32; g2.out = which ? nb.t_dec : nb.t_enc; g2.ldc = D;
    M5_TRY(run_gemm(ctx, g2));
  }
  return M5_OK;
}

// One model evaluation at timestep t for every sequence of the plan; leaves the logits of the rows in p.lg_* in
// nb.logits (codebook q -> logits for that head only; caller loops over q).
static int nar_trunk(m5_ctx* ctx, const NarWeights& w, const NarPlan& p, const int* x_dev, int t, int mode,
                     NarBuffers& nb) {
  const m5_model_cfg& c = ctx->cfg;
  const int D = c.nar_dim;
  const bool precise = mode != M5_NUM_FAST;   // are GEMM activations carried as (hi, lo) pairs?
  TokEmbedCall te;
  te.tok = p.enc_tok; te.pos = p.enc_pos; te.table = w.text_embed; te.vec_rows = nb.spk_vec; te.pe = w.pe;
  te.alpha = w.alpha_cond; te.add_vec = nb.t_enc + (size_t)t * D; te.n_rows = p.enc.rows; te.D = D; te.out = nb.xe;
  if (token_embed(te, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "token_embed failed");
  ctx->launches++;
  for (int l = 0; l < c.nar_enc_layers; ++l)
    M5_TRY(encoder_layer(ctx, nb.xe, p.enc, w.enc[l], D, c.nar_heads, c.nar_ff, c.ln_eps, mode, nb.bs));
  NormCall en;
  en.x = nb.xe; en.M = p.enc.rows; en.D = D; en.ldx = D; en.gamma = w.enc_nw; en.beta = w.enc_nb; en.eps = c.ln_eps;
  en.out = nb.mem16; en.out_lo = precise ? nb.mem16 + D : nullptr; en.ldo = precise ? 2 * D : D;
  M5_TRY(run_norm(ctx, en));
  EmbedCall de;
  de.codes = x_dev; de.code_row = p.dec_xrow; de.pos = p.dec_pos; de.tables = w.res_tables; de.identity = w.spk_identity;
  de.pe = w.pe; de.alpha = w.alpha_pos; de.add_vec = nb.t_dec + (size_t)t * D; de.n_rows = p.dec.rows; de.D = D;
  de.Q = c.n_quant; de.n_codes = c.n_classes; de.out = nb.xd;
  if (chunked_embed(de, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "chunked_embed(dec) failed");
  ctx->launches++;
  for (int l = 0; l < c.nar_dec_layers; ++l)
    M5_TRY(decoder_layer(ctx, nb.xd, p.dec, nb.mem16, p.enc, w.dec[l], D, c.nar_heads, c.nar_ff, c.ln_eps, mode, nb.bs));
  NormCall dn;
  dn.x = nb.xd; dn.M = p.npass * p.n_lg; dn.D = D; dn.ldx = D; dn.gamma = w.dec_nw; dn.beta = w.dec_nb; dn.eps = c.ln_eps;
  dn.out_f32 = nb.dn; dn.ldo = D; dn.row_map = p.lg_decrow;
  return run_norm(ctx, dn);
}

static int nar_head(m5_ctx* ctx, const NarWeights& w, const NarPlan& p, int q, int mode, NarBuffers& nb) {
  const bool precise = mode != M5_NUM_FAST;
  const m5_model_cfg& c = ctx->cfg;
  const int D = c.nar_dim, R = p.npass * p.n_lg;
  NormCall hn;
  hn.x = nb.dn; hn.M = R; hn.D = D; hn.ldx = D; hn.gamma = w.head_lnw[q]; hn.beta = w.head_lnb[q]; hn.eps = c.head_ln_eps;
  hn.out = nb.hd16; hn.out_lo = precise ? nb.hd16 + D : nullptr; hn.ldo = precise ? 2 * D : D;
  M5_TRY(run_norm(ctx, hn));
  GemmCall g;
  g.A = nb.hd16; g.W = w.head_w[q]; g.M = R; g.N = c.n_classes; g.K = precise ? 2 * D : D; g.lda = g.K; g.ldw = D;
  g.kwrap = precise ? D : 0; g.bias = w.head_b[q]; g.mode = M5_OUT_F32; g.out = nb.logits; g.ldc = nb.ldl;
  return run_gemm(ctx, g);
}

// scatter logits [R, ldl] of head q into out [R, Q, K]
__global__ void scatter_logits_kernel(const float* lg, int ldl, float* out, int R, int Q, int K, int q) {
  const size_t n = (size_t)R * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / K), k = (int)(i % K);
    out[((size_t)r * Q + q) * K + k] = lg[(size_t)r * ldl + k];
  }
}

// initial state (diffuser.py:404-438)
__global__ void nar_init_state_kernel(int Rx, int Q, const int* row_b, const int* row_i, const int* Pf, const int* x_off_nx,
                                      const int* c_off, const int* c_codes, const int* x_l0, const int* x_init, int deep,
                                      int K, uint64_t seed, const int64_t* x_utt, int* x, int* x_known, uint8_t* known,
                                      int* x_q0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Rx * Q) return;
  const int r = idx / Q, q = idx - r * Q;
  const int b = row_b[r], i = row_i[r];
  const int pf = deep ? Pf[b] : 0;
  int xv, kv; uint8_t m;
  if (i < pf) {
    const int code = c_codes[(size_t)(c_off[b] + i) * Q + q];
    xv = code; kv = code; m = 1;
    if (q == 0) x_q0[r] = code;
  } else {
    const int j = x_off_nx[b] + (i - pf);
    const int l0 = x_l0[j];
    if (q == 0) { xv = l0; kv = l0; m = 1; x_q0[r] = l0; }
    else {
      if (x_init) xv = x_init[(size_t)j * Q + q];
      else {
        uint32_t o[4];
        philox4x32((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(i - pf), (uint32_t)q, (uint32_t)x_utt[r], 0x494e4954u, o);
        xv = (int)(o[0] % (uint32_t)K);
      }
      kv = 0; m = 0;
    }
  }
  x[idx] = xv; x_known[idx] = kv; known[idx] = m;
}

// cosine schedule tables (diffuser.py:76-109), fp64 -> fp32 exactly as the reference
struct Schedule { std::vector<float> log_alpha, log_1m_alpha, log_cum, log_1m_cum; };
static void make_schedule(int T, Schedule& s) {
  // torch.linspace / cos are evaluated in fp32 by the reference before the .to(float64)
  std::vector<float> ac(T + 1);
  const float sft = 0.008f;
  for (int i = 0; i <= T; ++i) {
    // torch.linspace(0, T, T+1) yields exact integers here
    const float x = (float)i;
    const float v = cosf(((x / (float)T) + sft) / (1 + sft) * 3.14159265358979323846f * 0.5f);
    ac[i] = v * v;
  }
  const float a0 = ac[0];
  for (int i = 0; i <= T; ++i) ac[i] = ac[i] / a0;
  s.log_alpha.resize(T); s.log_1m_alpha.resize(T); s.log_cum.resize(T); s.log_1m_cum.resize(T);
  double cum = 0.0;
  for (int i = 0; i < T; ++i) {
    float a = ac[i + 1] / ac[i];
    a = fminf(fmaxf(a, 0.001f), 1.0f);
    const float sa = sqrtf(a);
    const double la = log((double)sa);
    cum += la;
    auto l1m = [](double v) { double e = 1.0 - exp(v); if (e < 1e-30) e = 1e-30; return log(e); };
    s.log_alpha[i] = (float)la; s.log_1m_alpha[i] = (float)l1m(la);
    s.log_cum[i] = (float)cum; s.log_1m_cum[i] = (float)l1m(cum);
  }
}

static const int* to_dev_ints(m5_ctx* ctx, Arena& ar, const int32_t* src, size_t n, int mem) {
  if (mem == M5_MEM_DEVICE || !src) return src;
  int* d = ar.get<int>(n ? n : 1);
  if (d && n) cudaMemcpyAsync(d, src, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
  return d;
}

}  // namespace m5

using namespace m5;

extern "C" {

int m5_nar_forward(m5_ctx* ctx, int32_t B, const int32_t* c_text, const int32_t* c_text_len, const int32_t* c_codes,
                   const int32_t* c_codes_len, const int32_t* x, const int32_t* x_len, int32_t t, int32_t drop_cond,
                   int32_t precise, int32_t mem, float* logits_out) {
  if (!ctx || B <= 0 || precise < 0 || precise > M5_NUM_MIXED8K) return M5_ERR_ARG;
  ctx->last_error.clear();
  cudaSetDevice(ctx->device);
  NarWeights w;
  M5_TRY(load_nar(ctx, w));
  const m5_model_cfg& c = ctx->cfg;
  // plan sizes first (host only) to size the arena
  size_t n_text = 0, n_codes = 0, n_x = 0;
  for (int b = 0; b < B; ++b) { n_text += c_text_len[b]; n_codes += c_codes_len[b]; n_x += x_len[b]; }
  NarPlan tmp;  // dry sizes
  tmp.B = B; tmp.npass = 1; tmp.spk.rows = (int)n_codes + B + 1; tmp.enc.rows = (int)n_text + B; tmp.dec.rows = (int)n_x;
  tmp.n_lg = (int)n_x;
  const int T = t + 1;
  Arena ar(ctx);
  const size_t idx_bytes = (tmp.spk.rows * 2 + tmp.enc.rows * 2 + tmp.dec.rows * 4 + 8 * B + 64) * sizeof(int) * 2 +
                           (size_t)n_x * 16 + (n_text + n_codes * c.n_quant + n_x * c.n_quant) * 4 + (1 << 16);
  M5_TRY(ar.reserve(nar_bytes(c, tmp, T, precise) + idx_bytes + (size_t)n_x * c.n_quant * c.n_classes * 4 * (mem == M5_MEM_HOST)));
  NarPlan p;
  M5_TRY(build_plan(ctx, ar, p, B, c_text_len, c_codes_len, x_len, false, 1, drop_cond != 0, true, nullptr));
  const int* d_text = to_dev_ints(ctx, ar, c_text, n_text, mem);
  const int* d_codes = to_dev_ints(ctx, ar, c_codes, n_codes * c.n_quant, mem);
  const int* d_x = to_dev_ints(ctx, ar, x, n_x * c.n_quant, mem);
  float* d_out = logits_out;
  if (mem == M5_MEM_HOST) d_out = ar.get<float>(n_x * c.n_quant * c.n_classes);
  if (!d_text || !d_codes || !d_x || !d_out) return ctx->fail(M5_ERR_NOMEM, "arena too small (inputs)");
  resolve_tokens_kernel<<<(p.enc.rows + 255) / 256, 256, 0, ctx->stream>>>(p.enc_tok, d_text, p.enc.rows);
  NarBuffers nb;
  M5_TRY(carve(ctx, ar, p, T, nb));
  M5_TRY(nar_speaker(ctx, w, p, d_codes, nb));
  M5_TRY(nar_timestep_tables(ctx, w, T, nb));
  M5_TRY(nar_trunk(ctx, w, p, d_x, t, precise, nb));
  for (int q = 0; q < c.n_quant; ++q) {
    M5_TRY(nar_head(ctx, w, p, q, precise, nb));
    scatter_logits_kernel<<<148 * 4, 256, 0, ctx->stream>>>(nb.logits, nb.ldl, d_out, p.n_lg, c.n_quant, c.n_classes, q);
    ctx->launches++;
  }
  if (mem == M5_MEM_HOST)
    M5_CUDA(cudaMemcpyAsync(logits_out, d_out, n_x * c.n_quant * c.n_classes * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  M5_CUDA(cudaStreamSynchronize(ctx->stream));
  return M5_OK;
}

int m5_nar_infer(m5_ctx* ctx, int32_t B, const int32_t* c_text, const int32_t* c_text_len, const int32_t* c_codes,
                 const int32_t* c_codes_len, const int32_t* x_l0, const int32_t* x_len, const m5_nar_cfg* cfg,
                 int32_t mem, const int32_t* x_init, const float* noise, uint64_t seed, const int64_t* utt_ids,
                 int32_t* out_codes) {
  if (!ctx || !cfg || B <= 0) return M5_ERR_ARG;
  if (cfg->precise < 0 || cfg->precise > M5_NUM_MIXED8K)
    return ctx->fail(M5_ERR_ARG, "m5_nar_cfg.precise must be 0 (fast), 1 (precise), 2 (mixed), 3 (mixed8) or 4 (mixed8k)");
  ctx->last_error.clear();
  cudaSetDevice(ctx->device);
  NarWeights w;
  M5_TRY(load_nar(ctx, w));
  const m5_model_cfg& c = ctx->cfg;
  const int T = cfg->T, Q = c.n_quant, K = c.n_classes;
  const bool deep = cfg->deep_clone != 0;
  const int precise = cfg->precise;   // M5_NUM_*
  const bool cfg_on = cfg->guidance_w != 1.0f;  // diffuser.py:361
  const int npass = cfg_on ? 2 : 1;
  size_t n_text = 0, n_codes = 0, n_x = 0, Rx = 0;
  for (int b = 0; b < B; ++b) {
    n_text += c_text_len[b]; n_codes += c_codes_len[b]; n_x += x_len[b];
    Rx += x_len[b] + (deep ? c_codes_len[b] : 0);
  }
  NarPlan tmp;
  tmp.B = B; tmp.npass = npass; tmp.spk.rows = (int)n_codes + B + 1; tmp.enc.rows = npass * ((int)n_text + B);
  tmp.dec.rows = npass * (int)Rx; tmp.n_lg = (int)n_x;
  Arena ar(ctx);
  const size_t idx_bytes = ((size_t)tmp.spk.rows * 2 + tmp.enc.rows * 2 + (size_t)tmp.dec.rows * 4 + 8 * B + 64) * sizeof(int) * 2 +
                           Rx * (16 + 4 * Q * 3 + Q + 16) + (n_text + n_codes * Q + n_x * (Q + 1) * 2) * 4 + (1 << 16);
  const size_t noise_bytes = (noise && mem == M5_MEM_HOST) ? (size_t)2 * Rx * Q * K * 4 : 0;
  M5_TRY(ar.reserve(nar_bytes(c, tmp, T, precise) + idx_bytes + noise_bytes));
  NarPlan p;
  M5_TRY(build_plan(ctx, ar, p, B, c_text_len, c_codes_len, x_len, deep, npass, false, false, utt_ids));
  const int* d_text = to_dev_ints(ctx, ar, c_text, n_text, mem);
  const int* d_codes = to_dev_ints(ctx, ar, c_codes, n_codes * Q, mem);
  const int* d_l0 = to_dev_ints(ctx, ar, x_l0, n_x, mem);
  const int* d_xinit = to_dev_ints(ctx, ar, x_init, n_x * Q, mem);
  if (!d_text || !d_codes || !d_l0) return ctx->fail(M5_ERR_NOMEM, "arena too small (inputs)");
  resolve_tokens_kernel<<<(p.enc.rows + 255) / 256, 256, 0, ctx->stream>>>(p.enc_tok, d_text, p.enc.rows);
  // state
  std::vector<int> row_b, row_i, xoffnx(B), coff(B);
  int a1 = 0, a2 = 0;
  for (int b = 0; b < B; ++b) {
    xoffnx[b] = a1; coff[b] = a2; a1 += p.Nx[b]; a2 += p.Pf[b];
    for (int i = 0; i < p.S[b]; ++i) { row_b.push_back(b); row_i.push_back(i); }
  }
  int* d_row_b = upload(ctx, ar, row_b); int* d_row_i = upload(ctx, ar, row_i);
  int* d_xoffnx = upload(ctx, ar, xoffnx); int* d_coff = upload(ctx, ar, coff); int* d_Pf = upload(ctx, ar, p.Pf);
  int* d_x = ar.get<int>(Rx * Q); int* d_xk = ar.get<int>(Rx * Q); uint8_t* d_m = ar.get<uint8_t>(Rx * Q);
  int* d_q0 = ar.get<int>(Rx);
  float* d_noise = nullptr;
  if (noise && mem == M5_MEM_HOST) d_noise = ar.get<float>((size_t)2 * Rx * Q * K);
  if (!d_row_b || !d_x || !d_xk || !d_m || !d_q0) return ctx->fail(M5_ERR_NOMEM, "arena too small (state)");
  nar_init_state_kernel<<<((int)Rx * Q + 255) / 256, 256, 0, ctx->stream>>>(
      (int)Rx, Q, d_row_b, d_row_i, d_Pf, d_xoffnx, d_coff, d_codes, d_l0, d_xinit, deep ? 1 : 0, K, seed, p.x_utt, d_x, d_xk,
      d_m, d_q0);
  ctx->launches++;
  NarBuffers nb;
  M5_TRY(carve(ctx, ar, p, T, nb));
  M5_TRY(nar_speaker(ctx, w, p, d_codes, nb));
  M5_TRY(nar_timestep_tables(ctx, w, T, nb));
  Schedule sch;
  if (cfg->schedule) {
    const float* sp = cfg->schedule;
    sch.log_alpha.assign(sp, sp + T); sch.log_1m_alpha.assign(sp + T, sp + 2 * T);
    sch.log_cum.assign(sp + 2 * T, sp + 3 * T); sch.log_1m_cum.assign(sp + 3 * T, sp + 4 * T);
  } else {
    make_schedule(T, sch);
  }
  // RePaint schedule (get_schedule, diffuser.py:318-333); with jump_len = jump_n_sample = 1 it is simply T-1 ... 0
  const int jl = std::max(1, cfg->jump_len), jn = std::max(1, cfg->jump_n_sample);
  if ((jl > 1 || jn > 1) && cfg->scaled_forward)
    return ctx->fail(M5_ERR_ARG, "RePaint jumps with enable_kevin_scaled_inference: the reference's q_pred_one_timestep_scaled "
                                 "(diffuser.py:136-159) raises for every sequence length != 8; pass scaled_forward = 0");
  std::vector<int> times;
  {
    std::vector<int> jumps(T + 1, 0);
    for (int j = 0; j < T - jl; j += jl) jumps[j] = jn - 1;
    int t = T;
    while (t >= 1) {
      t -= 1;
      times.push_back(t);
      if (jumps[t] > 0) {
        jumps[t] -= 1;
        for (int i = 0; i < jl; ++i) { t += 1; times.push_back(t); }
      }
    }
    times.push_back(-1);
  }
  for (size_t step = 0; step + 1 < times.size(); ++step) {
    const int t = times[step], t_cur = times[step + 1];
    const float* u0 = nullptr; const float* u1 = nullptr;
    if (noise) {
      const size_t per = (size_t)Rx * Q * K;
      if (mem == M5_MEM_HOST) {
        M5_CUDA(cudaMemcpyAsync(d_noise, noise + step * 2 * per, 2 * per * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        u0 = d_noise; u1 = d_noise + per;
      } else { u0 = noise + step * 2 * per; u1 = u0 + per; }
    }
    RenoiseCall rc;
    rc.R = (int)Rx; rc.Q = Q; rc.K = K; rc.x = d_x; rc.x_q0 = d_q0; rc.t = t;
    rc.q0_override = (cfg->q0_override_steps < t) ? 1 : 0;  // retain_quant0 (diffuser.py:467-468)
    rc.seed = seed; rc.row_utt = p.x_utt; rc.row_pos = p.x_pos;
    if (t_cur > t) {
      // forward step x_t -> x_{t+1} ~ q(x_{t+1} | x_t) for EVERY entry (forward_diffusion, diffuser.py:336-342,462-465)
      rc.forward = 1; rc.x_known = d_x; rc.known = d_m; rc.u = u0;
      rc.log_cum_t = sch.log_alpha[t]; rc.log_1m_cum_t = sch.log_1m_alpha[t];
      if (nar_renoise(rc, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "nar_renoise(forward) failed");
      ctx->launches++;
      continue;
    }
    M5_TRY(nar_trunk(ctx, w, p, d_x, t, precise, nb));
    for (int q = 1; q < Q; ++q) {  // codebook 0 is always known
      M5_TRY(nar_head(ctx, w, p, q, precise, nb));
      PosteriorCall pc;
      pc.cond = nb.logits; pc.uncond = cfg_on ? nb.logits + (size_t)p.n_lg * nb.ldl : nullptr; pc.ld = nb.ldl;
      pc.R = p.n_lg; pc.K = K; pc.Q = Q; pc.q = q; pc.guidance_w = cfg->guidance_w; pc.x0_temp = cfg->x0_temp;
      pc.log_alpha_t = sch.log_alpha[t]; pc.log_1m_alpha_t = sch.log_1m_alpha[t];
      pc.log_cum_tm1 = sch.log_cum[t > 0 ? t - 1 : 0]; pc.log_1m_cum_tm1 = sch.log_1m_cum[t > 0 ? t - 1 : 0];
      pc.t = t; pc.row_map = p.lg_xrow; pc.x_t = d_x; pc.x_out = d_x; pc.u = u0; pc.u_rows_are_x = 1; pc.seed = seed;
      pc.row_utt = p.x_utt; pc.row_pos = p.x_pos; pc.draw = 0;
      if (nar_posterior(pc, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "nar_posterior failed");
      ctx->launches++;
    }
    rc.x_known = d_xk; rc.known = d_m; rc.u = u1;
    rc.log_cum_t = sch.log_cum[t]; rc.log_1m_cum_t = sch.log_1m_cum[t];
    if (nar_renoise(rc, ctx->stream) != M5_OK) return ctx->fail(M5_ERR_CUDA, "nar_renoise failed");
    ctx->launches++;
  }
  // crop the prompt (diffuser.py:471) and return [sum Nx, Q]
  for (int b = 0, off = 0; b < B; ++b) {
    const int pf = deep ? p.Pf[b] : 0;
    const int* src = d_x + (size_t)(p.x_off[b] + pf) * Q;
    M5_CUDA(cudaMemcpyAsync(out_codes + (size_t)off * Q, src, (size_t)p.Nx[b] * Q * sizeof(int),
                            mem == M5_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
    off += p.Nx[b];
  }
  M5_CUDA(cudaStreamSynchronize(ctx->stream));
  return M5_OK;
}

}  // extern "C"
