#include "layers.h"

namespace m5 {

int run_gemm(m5_ctx* ctx, const GemmCall& g) {
  cudaEvent_t pe = prof_begin(ctx);
  int r = gemm_tc5(g, ctx->stream, ctx->num_sms);
  if (pe) {
    const double out_b = (g.mode == M5_OUT_F32) ? 4.0 * (g.accumulate ? 2 : 1) : 2.0;
    const double wk = g.kwrap > 0 ? g.kwrap : g.K, ak = g.awrap > 0 ? g.awrap : g.K;
    prof_end(ctx, pe, 0, 2.0 * g.M * (double)g.N * g.K, 2.0 * g.M * ak + 2.0 * g.N * wk + out_b * g.M * (double)g.N);
  }
  if (r != M5_OK) return ctx->fail(r, "gemm_tc5 failed (M=" + std::to_string(g.M) + " N=" + std::to_string(g.N) + " K=" +
                                          std::to_string(g.K) + "): " + cudaGetErrorString(cudaGetLastError()));
  ctx->launches += 1;
  return M5_OK;
}
int run_norm(m5_ctx* ctx, const NormCall& n) {
  int r = norm_rows(n, ctx->stream);
  if (r != M5_OK) return ctx->fail(r, "norm_rows failed");
  ctx->launches += 1;
  return M5_OK;
}
int run_attn(m5_ctx* ctx, const AttnCall& a) {
  cudaEvent_t pe = prof_begin(ctx);
  const bool tc5 = a.impl == 2 || (a.impl == 0 && !a.causal && a.q_rows > 0 && a.k_rows > 0 && a.max_q >= 256);
  int r = tc5 ? flash_attn_tc5(a, ctx->stream) : flash_attn(a, ctx->stream);
  if (pe) prof_end(ctx, pe, 1, a.flops_hint, 0.0);
  if (r != M5_OK) return ctx->fail(r, "flash_attn failed");
  ctx->launches += 1;
  return M5_OK;
}

#define NEEDW(field, T, name)                                   \
  do {                                                          \
    w.field = W<T>(ctx, prefix + name);                         \
    if (!w.field) return M5_ERR_MISSING_WEIGHT;                 \
  } while (0)

int load_enc_layer(m5_ctx* ctx, const std::string& prefix, EncLayerW& w) {
  NEEDW(n1w, float, "n1w"); NEEDW(n1b, float, "n1b"); NEEDW(n2w, float, "n2w"); NEEDW(n2b, float, "n2b");
  NEEDW(in_w, __half, "in_w"); NEEDW(in_b, float, "in_b"); NEEDW(out_w, __half, "out_w"); NEEDW(out_b, float, "out_b");
  NEEDW(wv, __half, "wv"); NEEDW(w2, __half, "w2"); NEEDW(b2, float, "b2");
  return M5_OK;
}
int load_dec_layer(m5_ctx* ctx, const std::string& prefix, DecLayerW& w) {
  NEEDW(n1w, float, "n1w"); NEEDW(n1b, float, "n1b"); NEEDW(n2w, float, "n2w"); NEEDW(n2b, float, "n2b");
  NEEDW(n3w, float, "n3w"); NEEDW(n3b, float, "n3b");
  NEEDW(sa_in_w, __half, "sa_in_w"); NEEDW(sa_in_b, float, "sa_in_b");
  NEEDW(sa_out_w, __half, "sa_out_w"); NEEDW(sa_out_b, float, "sa_out_b");
  NEEDW(ca_q_w, __half, "ca_q_w"); NEEDW(ca_q_b, float, "ca_q_b");
  NEEDW(ca_kv_w, __half, "ca_kv_w"); NEEDW(ca_kv_b, float, "ca_kv_b");
  NEEDW(ca_out_w, __half, "ca_out_w"); NEEDW(ca_out_b, float, "ca_out_b");
  NEEDW(wv, __half, "wv"); NEEDW(w2, __half, "w2"); NEEDW(b2, float, "b2");
  return M5_OK;
}

static size_t al(size_t n) { return (n + 255) & ~size_t(255); }
size_t block_scratch_bytes(int rows, int mem_rows, int D, int ff) {
  return al((size_t)rows * 2 * D * 2) + al((size_t)rows * 3 * D * 2) + al((size_t)rows * D * 2) +
         al((size_t)rows * 2 * ff * 2) + al((size_t)mem_rows * 2 * D * 2) + 4096;
}
void block_scratch_carve(Arena& a, BlockScratch& s, int rows, int mem_rows, int D, int ff) {
  s.h16 = a.get<__half>((size_t)rows * 2 * D);
  s.qkv16 = a.get<__half>((size_t)rows * 3 * D);
  s.att16 = a.get<__half>((size_t)rows * D);
  s.g16 = a.get<__half>((size_t)rows * 2 * ff);
  s.kv16 = mem_rows > 0 ? a.get<__half>((size_t)mem_rows * 2 * D) : nullptr;
}

// h16 <- LayerNorm(x) as fp16 (hi | lo halves side by side when precise)
static int ln_to_f16(m5_ctx* ctx, const float* x, int rows, int D, const float* g, const float* b, float eps,
                     bool precise, __half* h16) {
  NormCall n;
  n.x = x; n.M = rows; n.D = D; n.ldx = D; n.gamma = g; n.beta = b; n.eps = eps;
  n.out = h16; n.out_lo = precise ? h16 + D : nullptr; n.ldo = precise ? 2 * D : D;
  return run_norm(ctx, n);
}
static GemmCall lin(const __half* A, int rows, int K, bool split, const __half* Wt, int N, const float* bias) {
  GemmCall g;
  g.A = A; g.W = Wt; g.M = rows; g.N = N;
  g.K = split ? 2 * K : K; g.lda = split ? 2 * K : K; g.ldw = K; g.kwrap = split ? K : 0; g.bias = bias;
  return g;
}

// shared FFN tail: x += W2 * swiglu(WV * LN(x)) + b2
static int ffn_block(m5_ctx* ctx, float* x, int rows, int D, int ff, const float* nw, const float* nb, float eps,
                     const __half* wv, const __half* w2, const float* b2, bool precise, const BlockScratch& s) {
  M5_TRY(ln_to_f16(ctx, x, rows, D, nw, nb, eps, precise, s.h16));
  GemmCall g1 = lin(s.h16, rows, D, precise, wv, 2 * ff, nullptr);
  g1.out = s.g16;
  g1.out_lo = precise ? s.g16 + ff : nullptr;
  g1.ldc = precise ? 2 * ff : ff;
  g1.mode = precise ? M5_OUT_SWIGLU_F16_SPLIT : M5_OUT_SWIGLU_F16;
  M5_TRY(run_gemm(ctx, g1));
  GemmCall g2 = lin(s.g16, rows, ff, precise, w2, D, b2);
  g2.out = x; g2.ldc = D; g2.mode = M5_OUT_F32; g2.accumulate = 1;
  M5_TRY(run_gemm(ctx, g2));
  return M5_OK;
}

int encoder_layer(m5_ctx* ctx, float* x, const SeqSet& seqs, const EncLayerW& w, int D, int H, int ff, float eps,
                  bool precise, const BlockScratch& s) {
  const int rows = seqs.rows;
  // x = x + out_proj(MHA(LN1(x)))
  M5_TRY(ln_to_f16(ctx, x, rows, D, w.n1w, w.n1b, eps, precise, s.h16));
  GemmCall gq = lin(s.h16, rows, D, precise, w.in_w, 3 * D, w.in_b);
  gq.out = s.qkv16; gq.ldc = 3 * D; gq.mode = M5_OUT_F16;
  M5_TRY(run_gemm(ctx, gq));
  AttnCall a;
  a.Q = s.qkv16; a.K = s.qkv16 + D; a.V = s.qkv16 + 2 * D; a.ldq = a.ldk = a.ldv = 3 * D;
  a.O = s.att16; a.ldo = D; a.n_heads = H; a.n_seqs = seqs.n; a.max_q = seqs.max_len;
  a.q_start = seqs.start; a.q_len = seqs.len; a.k_start = seqs.start; a.k_len = seqs.klen ? seqs.klen : seqs.len;
  a.flops_hint = 256.0 * H * seqs.self_pairs; a.q_rows = rows; a.k_rows = rows;
  M5_TRY(run_attn(ctx, a));
  GemmCall go = lin(s.att16, rows, D, false, w.out_w, D, w.out_b);
  go.out = x; go.ldc = D; go.mode = M5_OUT_F32; go.accumulate = 1;
  M5_TRY(run_gemm(ctx, go));
  // x = x + linear2(swiglu(LN2(x)))
  return ffn_block(ctx, x, rows, D, ff, w.n2w, w.n2b, eps, w.wv, w.w2, w.b2, precise, s);
}

int decoder_layer(m5_ctx* ctx, float* x, const SeqSet& seqs, const __half* mem16, const SeqSet& mem_seqs,
                  const DecLayerW& w, int D, int H, int ff, float eps, bool precise, const BlockScratch& s) {
  const int rows = seqs.rows;
  // self attention
  M5_TRY(ln_to_f16(ctx, x, rows, D, w.n1w, w.n1b, eps, precise, s.h16));
  GemmCall gq = lin(s.h16, rows, D, precise, w.sa_in_w, 3 * D, w.sa_in_b);
  gq.out = s.qkv16; gq.ldc = 3 * D; gq.mode = M5_OUT_F16;
  M5_TRY(run_gemm(ctx, gq));
  AttnCall a;
  a.Q = s.qkv16; a.K = s.qkv16 + D; a.V = s.qkv16 + 2 * D; a.ldq = a.ldk = a.ldv = 3 * D;
  a.O = s.att16; a.ldo = D; a.n_heads = H; a.n_seqs = seqs.n; a.max_q = seqs.max_len;
  a.q_start = seqs.start; a.q_len = seqs.len; a.k_start = seqs.start; a.k_len = seqs.len;
  a.flops_hint = 256.0 * H * seqs.self_pairs; a.q_rows = rows; a.k_rows = rows;
  M5_TRY(run_attn(ctx, a));
  GemmCall go = lin(s.att16, rows, D, false, w.sa_out_w, D, w.sa_out_b);
  go.out = x; go.ldc = D; go.mode = M5_OUT_F32; go.accumulate = 1;
  M5_TRY(run_gemm(ctx, go));
  // cross attention over the encoder memory
  M5_TRY(ln_to_f16(ctx, x, rows, D, w.n2w, w.n2b, eps, precise, s.h16));
  GemmCall gcq = lin(s.h16, rows, D, precise, w.ca_q_w, D, w.ca_q_b);
  gcq.out = s.qkv16; gcq.ldc = D; gcq.mode = M5_OUT_F16;
  M5_TRY(run_gemm(ctx, gcq));
  GemmCall gkv = lin(mem16, mem_seqs.rows, D, precise, w.ca_kv_w, 2 * D, w.ca_kv_b);
  gkv.out = s.kv16; gkv.ldc = 2 * D; gkv.mode = M5_OUT_F16;
  M5_TRY(run_gemm(ctx, gkv));
  AttnCall c;
  c.Q = s.qkv16; c.ldq = D; c.K = s.kv16; c.V = s.kv16 + D; c.ldk = c.ldv = 2 * D;
  c.O = s.att16; c.ldo = D; c.n_heads = H; c.n_seqs = seqs.n; c.max_q = seqs.max_len;
  c.q_start = seqs.start; c.q_len = seqs.len; c.k_start = mem_seqs.start; c.k_len = mem_seqs.len;
  c.flops_hint = 256.0 * H * seqs.cross_pairs; c.q_rows = rows; c.k_rows = mem_seqs.rows;
  M5_TRY(run_attn(ctx, c));
  GemmCall gco = lin(s.att16, rows, D, false, w.ca_out_w, D, w.ca_out_b);
  gco.out = x; gco.ldc = D; gco.mode = M5_OUT_F32; gco.accumulate = 1;
  M5_TRY(run_gemm(ctx, gco));
  return ffn_block(ctx, x, rows, D, ff, w.n3w, w.n3b, eps, w.wv, w.w2, w.b2, precise, s);
}

}  // namespace m5
