#include "layers.h"

namespace m5 {

int run_gemm(m5_ctx* ctx, const GemmCall& g) {
  cudaEvent_t pe = prof_begin(ctx);
  int r = gemm_tc5(g, ctx->stream, ctx->num_sms);
  if (pe) {
    const double out_b = (g.mode == M5_OUT_F32) ? 4.0 * (g.accumulate ? 2 : 1) : 2.0;
    const double wk = g.kwrap > 0 ? g.kwrap : g.K, ak = g.awrap > 0 ? g.awrap : g.K;
    // fp8 lo pass (mixed8): counted at half weight -- "fp16-equivalent" tensor work, the fp8 UMMA runs at twice the fp16 rate --
    // so that the class figure stays comparable with the measured bf16 peak
    const double k_eq = (double)g.K + (g.A8 ? 0.5 * g.K8 : 0.0);
    prof_end(ctx, pe, 0, 2.0 * g.M * (double)g.N * k_eq,
             2.0 * g.M * ak + 2.0 * g.N * wk + (g.A8 ? (double)g.K8 * (g.M + g.N) : 0.0) + out_b * g.M * (double)g.N);
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
  // optional fp8 copies (weights.repack writes them for the NAR decoder): absent -> mixed8 falls back to fp16 pairs
  auto opt8 = [&](const char* name) -> const uint8_t* {
    auto it = ctx->weights.find(prefix + name);
    return it == ctx->weights.end() ? nullptr : reinterpret_cast<const uint8_t*>(it->second.ptr);
  };
  w.sa_in_w8 = opt8("sa_in_w8"); w.sa_out_w8 = opt8("sa_out_w8"); w.ca_out_w8 = opt8("ca_out_w8"); w.wv8 = opt8("wv8"); w.w28 = opt8("w28");
  return M5_OK;
}

static size_t al(size_t n) { return (n + 255) & ~size_t(255); }
size_t block_scratch_bytes(int rows, int mem_rows, int D, int ff) {
  return al((size_t)rows * 2 * D * 2) + 2 * al((size_t)rows * 3 * D * 2) + al((size_t)rows * 2 * D * 2) +
         al((size_t)rows * 2 * ff * 2) + 2 * al((size_t)mem_rows * 2 * D * 2) + 2 * al((size_t)rows * D) + al((size_t)rows * ff) + 8192;
}
void block_scratch_carve(Arena& a, BlockScratch& s, int rows, int mem_rows, int D, int ff) {
  s.h16 = a.get<__half>((size_t)rows * 2 * D);
  s.qkv16 = a.get<__half>((size_t)rows * 3 * D);
  s.qkv16_lo = a.get<__half>((size_t)rows * 3 * D);
  s.att16 = a.get<__half>((size_t)rows * 2 * D);
  s.g16 = a.get<__half>((size_t)rows * 2 * ff);
  s.kv16 = mem_rows > 0 ? a.get<__half>((size_t)mem_rows * 2 * D) : nullptr;
  s.kv16_lo = mem_rows > 0 ? a.get<__half>((size_t)mem_rows * 2 * D) : nullptr;
  s.h8 = a.get<uint8_t>((size_t)rows * D);
  s.att8 = a.get<uint8_t>((size_t)rows * D);
  s.g8 = a.get<uint8_t>((size_t)rows * ff);
}

// h16 <- LayerNorm(x) as fp16 (hi | lo halves side by side when precise)
static int ln_to_f16(m5_ctx* ctx, const float* x, int rows, int D, const float* g, const float* b, float eps,
                     bool precise, __half* h16, uint8_t* h8 = nullptr) {
  NormCall n;
  n.x = x; n.M = rows; n.D = D; n.ldx = D; n.gamma = g; n.beta = b; n.eps = eps;
  n.out = h16; n.out_lo = (precise && !h8) ? h16 + D : nullptr; n.ldo = precise ? 2 * D : D;
  n.out_lo8 = h8; n.ldo8 = D;   // mixed8: hi halves keep their [rows, 2D] slots, lo halves go to the byte buffer
  return run_norm(ctx, n);
}
// fp8 lo pass of a GEMM whose fp16 hi halves sit in A (row stride 2K): A8 [rows, K] e5m2, W8 [N, K] e4m3
static void use_f8(GemmCall& g, int K, const uint8_t* A8, const uint8_t* W8) {
  g.K = K; g.kwrap = 0;            // hi halves only on the fp16 side (lda stays 2K)
  g.A8 = A8; g.W8 = W8; g.K8 = K; g.lda8 = K; g.ldw8 = K;
}
static GemmCall lin(const __half* A, int rows, int K, bool split, const __half* Wt, int N, const float* bias) {
  GemmCall g;
  g.A = A; g.W = Wt; g.M = rows; g.N = N;
  g.K = split ? 2 * K : K; g.lda = split ? 2 * K : K; g.ldw = K; g.kwrap = split ? K : 0; g.bias = bias;
  return g;
}

// shared FFN tail: x += W2 * swiglu(WV * LN(x)) + b2
static int ffn_block(m5_ctx* ctx, float* x, int rows, int D, int ff, const float* nw, const float* nb, float eps,
                     const __half* wv, const __half* w2, const float* b2, int mode, const BlockScratch& s,
                     const uint8_t* wv8 = nullptr, const uint8_t* w28 = nullptr) {
  const bool split = mode != M5_NUM_FAST;
  const bool f8 = wv8 && w28;   // decided by decoder_layer (mixed8, long sequences, CTA-pair-eligible shapes)
  M5_TRY(ln_to_f16(ctx, x, rows, D, nw, nb, eps, split, s.h16, f8 ? s.h8 : nullptr));
  GemmCall g1 = lin(s.h16, rows, D, split, wv, 2 * ff, nullptr);
  g1.out = s.g16;
  g1.out_lo = split ? s.g16 + ff : nullptr;
  g1.ldc = split ? 2 * ff : ff;
  g1.mode = split ? M5_OUT_SWIGLU_F16_SPLIT : M5_OUT_SWIGLU_F16;
  if (f8) { use_f8(g1, D, s.h8, wv8); g1.out_lo8 = s.g8; g1.ldc8 = ff; }
  M5_TRY(run_gemm(ctx, g1));
  GemmCall g2 = lin(s.g16, rows, ff, split, w2, D, b2);
  g2.out = x; g2.ldc = D; g2.mode = M5_OUT_F32; g2.accumulate = 1;
  if (f8) use_f8(g2, ff, s.g8, w28);
  M5_TRY(run_gemm(ctx, g2));
  return M5_OK;
}

// x += out_proj(softmax(q k^T / 8) v): shared by self- and cross-attention.  Queries come from h16 (already normalised),
// keys/values either from the same projection (kv_src == nullptr) or from the encoder memory.
static int attn_block(m5_ctx* ctx, float* x, const SeqSet& seqs, int D, int H, int mode, const BlockScratch& s,
                      const __half* in_w, const float* in_b, const __half* out_w, const float* out_b,
                      const __half* mem16, const SeqSet* mem_seqs, const __half* kv_w, const float* kv_b,
                      const uint8_t* in_w8 = nullptr, const uint8_t* out_w8 = nullptr, bool f8 = false) {
  const int rows = seqs.rows;
  const bool cross = mem16 != nullptr;
  const bool split = mode != M5_NUM_FAST;
  // mixed: sequences long enough for the tcgen05 kernel keep Q (and P) single fp16 and carry K, V, O as pairs; shorter
  // ones (text encoder, tiny inputs) run the fully split mma.sync kernel
  // (below ~1k rows the probabilities' fp16 rounding is averaged over too few keys: uncond pass at S = 300 measured 7e-4)
  const bool tc5_split = num_is_mixed(mode) && seqs.max_len >= 1024;
  // mixed8k: the keys of the (long) decoder self-attention stay single fp16 -- one S pass; cross-attention (137 keys) keeps pairs
  const bool k_single = mode == M5_NUM_MIXED8K && tc5_split && !cross;
  const bool q_pair = split && !tc5_split;   // does the attention kernel consume low halves of Q?
  // f8 (decided by the caller, decoder_layer): fp8 lo pass in the projections around the tcgen05 pair attention; for self-
  // attention the caller's LayerNorm wrote the lo halves to s.h8 (and NOT to the fp16 lo slots)
  if (f8 && (!tc5_split || !out_w8 || (!cross && !in_w8))) return ctx->fail(M5_ERR_STATE, "mixed8 preconditions violated");
  const int qn = cross ? D : 3 * D;  // width of the projection of h16
  GemmCall gq = lin(s.h16, rows, D, split && !(cross && tc5_split), in_w, qn, in_b);
  if (cross && tc5_split) gq.lda = 2 * D;  // hi halves only: the rounding of a cross-attention query is 2e-5 rms on the logits
  gq.out = s.qkv16; gq.ldc = qn;
  const bool q_out_pair = cross ? q_pair : split;   // the fused QKV projection always needs the K / V low halves
  gq.mode = q_out_pair ? M5_OUT_F16_SPLIT : M5_OUT_F16; gq.out_lo = q_out_pair ? s.qkv16_lo : nullptr;
  // the tcgen05 pair attention reads no lo halves of the queries (columns < D), with single keys none below 2 D either
  if (!cross && tc5_split) gq.lo_from_col = k_single ? 2 * D : D;
  if (f8 && !cross) use_f8(gq, D, s.h8, in_w8);
  M5_TRY(run_gemm(ctx, gq));
  AttnCall a;
  a.Q = s.qkv16; a.ldq = qn; a.Qlo = q_pair ? s.qkv16_lo : nullptr;
  if (cross) {
    GemmCall gkv = lin(mem16, mem_seqs->rows, D, split, kv_w, 2 * D, kv_b);
    gkv.out = s.kv16; gkv.ldc = 2 * D; gkv.mode = split ? M5_OUT_F16_SPLIT : M5_OUT_F16; gkv.out_lo = split ? s.kv16_lo : nullptr;
    M5_TRY(run_gemm(ctx, gkv));
    a.K = s.kv16; a.V = s.kv16 + D; a.ldk = a.ldv = 2 * D;
    if (split) { a.Klo = s.kv16_lo; a.Vlo = s.kv16_lo + D; }
    a.k_start = mem_seqs->start; a.k_len = mem_seqs->len; a.k_rows = mem_seqs->rows;
    a.flops_hint = 256.0 * H * seqs.cross_pairs;
  } else {
    a.K = s.qkv16 + D; a.V = s.qkv16 + 2 * D; a.ldk = a.ldv = 3 * D;
    if (split) { a.Klo = k_single ? nullptr : s.qkv16_lo + D; a.Vlo = s.qkv16_lo + 2 * D; }
    a.k_start = seqs.start; a.k_len = seqs.klen ? seqs.klen : seqs.len; a.k_rows = rows;
    a.flops_hint = 256.0 * H * seqs.self_pairs;
  }
  if (tc5_split) a.flops_hint *= k_single ? 1.5 : 2.0;   // S and PV each run two UMMA passes (hi and lo tiles); S one with single keys
  a.O = s.att16; a.ldo = split ? 2 * D : D; a.Olo = split ? s.att16 + D : nullptr;
  if (f8) { a.Olo = nullptr; a.Olo8 = s.att8; a.ldo8 = D; }
  a.n_heads = H; a.n_seqs = seqs.n; a.max_q = seqs.max_len; a.q_start = seqs.start; a.q_len = seqs.len; a.q_rows = rows;
  if (q_pair) a.impl = 1;        // the fully split path lives in the mma.sync kernel
  else if (tc5_split) a.impl = 2;
  M5_TRY(run_attn(ctx, a));
  GemmCall go = lin(s.att16, rows, D, split, out_w, D, out_b);
  go.out = x; go.ldc = D; go.mode = M5_OUT_F32; go.accumulate = 1;
  if (f8) use_f8(go, D, s.att8, out_w8);
  return run_gemm(ctx, go);
}

int encoder_layer(m5_ctx* ctx, float* x, const SeqSet& seqs, const EncLayerW& w, int D, int H, int ff, float eps,
                  int mode, const BlockScratch& s) {
  const int rows = seqs.rows;
  const bool split = mode != M5_NUM_FAST;
  // x = x + out_proj(MHA(LN1(x)))
  M5_TRY(ln_to_f16(ctx, x, rows, D, w.n1w, w.n1b, eps, split, s.h16));
  M5_TRY(attn_block(ctx, x, seqs, D, H, mode, s, w.in_w, w.in_b, w.out_w, w.out_b, nullptr, nullptr, nullptr, nullptr));
  // x = x + linear2(swiglu(LN2(x)))
  return ffn_block(ctx, x, rows, D, ff, w.n2w, w.n2b, eps, w.wv, w.w2, w.b2, mode, s);
}

int decoder_layer(m5_ctx* ctx, float* x, const SeqSet& seqs, const __half* mem16, const SeqSet& mem_seqs,
                  const DecLayerW& w, int D, int H, int ff, float eps, int mode, const BlockScratch& s) {
  const int rows = seqs.rows;
  const bool split = mode != M5_NUM_FAST;
  // mixed8 applies where the tcgen05 pair attention does (long sequences) and the shapes run on the CTA-pair GEMM
  const bool f8 = num_has_f8(mode) && seqs.max_len >= 1024 && w.sa_in_w8 && w.sa_out_w8 && w.ca_out_w8 && w.wv8 && w.w28 &&
                  D % 128 == 0 && ff % 128 == 0 && gemm_f8lo_eligible(rows, D, ctx->num_sms);
  M5_TRY(ln_to_f16(ctx, x, rows, D, w.n1w, w.n1b, eps, split, s.h16, f8 ? s.h8 : nullptr));
  M5_TRY(attn_block(ctx, x, seqs, D, H, mode, s, w.sa_in_w, w.sa_in_b, w.sa_out_w, w.sa_out_b, nullptr, nullptr, nullptr, nullptr,
                    w.sa_in_w8, w.sa_out_w8, f8));
  // cross-attention: the query projection reads the hi halves only (its rounding is 2e-5 rms on the logits)
  M5_TRY(ln_to_f16(ctx, x, rows, D, w.n2w, w.n2b, eps, split, s.h16));
  M5_TRY(attn_block(ctx, x, seqs, D, H, mode, s, w.ca_q_w, w.ca_q_b, w.ca_out_w, w.ca_out_b, mem16, &mem_seqs, w.ca_kv_w, w.ca_kv_b,
                    nullptr, w.ca_out_w8, f8));
  return ffn_block(ctx, x, rows, D, ff, w.n3w, w.n3b, eps, w.wv, w.w2, w.b2, mode, s, f8 ? w.wv8 : nullptr, f8 ? w.w28 : nullptr);
}

}  // namespace m5
