// Host-side composition of kernels into the transformer blocks shared by the AR speaker encoder, the NAR speaker
// encoder, the NAR encoder (nn.TransformerEncoderLayer, norm_first, FNNSwiGLU activation with linear1 = Identity;
// model.py:56-67,179-199) and the NAR decoder (nn.TransformerDecoderLayer, model.py:186-193).
#pragma once
#include <string>

#include "ctx.h"
#include "rowops.h"

namespace m5 {

// A packed batch of sequences living in DEVICE int arrays (+ host copies of the extents).
struct SeqSet {
  int n = 0;          // sequences
  int rows = 0;       // total rows
  int max_len = 0;    // longest sequence
  const int* start = nullptr;  // device [n]
  const int* len = nullptr;    // device [n]
  const int* klen = nullptr;   // device [n] optional: visible keys per sequence (key-padding mask), default = len
  double self_pairs = 0;       // sum_s len_s^2 (profiling: attention FLOPs)
  double cross_pairs = 0;      // sum_s len_s * mem_len_s (decoder sets only)
};

struct EncLayerW {
  const float *n1w, *n1b, *n2w, *n2b;
  const __half* in_w; const float* in_b;    // [3D, D]
  const __half* out_w; const float* out_b;  // [D, D]
  const __half* wv;                         // [2*ff, D] rows interleaved (W_j, V_j)
  const __half* w2; const float* b2;        // [D, ff]
};
struct DecLayerW {
  const float *n1w, *n1b, *n2w, *n2b, *n3w, *n3b;
  const __half* sa_in_w; const float* sa_in_b; const __half* sa_out_w; const float* sa_out_b;
  const __half* ca_q_w; const float* ca_q_b;      // [D, D]
  const __half* ca_kv_w; const float* ca_kv_b;    // [2D, D]
  const __half* ca_out_w; const float* ca_out_b;
  const __half* wv; const __half* w2; const float* b2;
  // optional e4m3 copies (x 2^+2) for the fp8 lo pass (mixed8): null when the tensor table has none
  const uint8_t *sa_in_w8 = nullptr, *sa_out_w8 = nullptr, *ca_out_w8 = nullptr, *wv8 = nullptr, *w28 = nullptr;
};
int load_enc_layer(m5_ctx* ctx, const std::string& prefix, EncLayerW& w);
int load_dec_layer(m5_ctx* ctx, const std::string& prefix, DecLayerW& w);

// Scratch shared by the blocks (sized for `rows` rows by the caller).
struct BlockScratch {
  __half* h16 = nullptr;    // [rows, 2*D]   normalised activations (hi | lo when precise)
  __half* qkv16 = nullptr;  // [rows, 3*D]
  __half* att16 = nullptr;  // [rows, 2*D]   attention output (hi | lo when precise)
  __half* g16 = nullptr;    // [rows, 2*ff]  gated activations (hi | lo when precise)
  __half* kv16 = nullptr;   // [mem_rows, 2*D] cross-attention keys/values
  __half* qkv16_lo = nullptr;  // [rows, 3*D]     low halves of q/k/v (precise mode)
  __half* kv16_lo = nullptr;   // [mem_rows, 2*D]
  uint8_t* h8 = nullptr;       // [rows, D]   mixed8: e5m2 lo halves of the normalised activations
  uint8_t* att8 = nullptr;     // [rows, D]   ... of the attention output
  uint8_t* g8 = nullptr;       // [rows, ff]  ... of the gated FFN activations
};
size_t block_scratch_bytes(int rows, int mem_rows, int D, int ff);
void block_scratch_carve(Arena& a, BlockScratch& s, int rows, int mem_rows, int D, int ff);

// NAR numerics (m5_nar_cfg.precise):
//   M5_NUM_FAST    every GEMM / attention operand is one fp16 value (fp32 accumulate)
//   M5_NUM_PRECISE every activation operand is an fp16 (hi, lo) pair; attention in the 3-term mma.sync kernel
//   M5_NUM_MIXED   GEMM activations, keys and values are (hi, lo) pairs, queries and probabilities single fp16; attention on
//                  tcgen05 (split-KV kernel) -- the cheapest setting that holds 1e-3 max-abs on the logits
//   M5_NUM_MIXED8  like MIXED, but in the big decoder GEMMs the lo half of every activation pair is an e5m2 value (x 2^-2)
//                  multiplied against an e4m3 copy of the weights (x 2^+2) by a kind::f8f6f4 UMMA pass into the same TMEM
//                  accumulator: the correction term needs ~4 bits, the fp8 pass runs at twice the fp16 rate
//   M5_NUM_MIXED8K like MIXED8, but the keys of the decoder SELF-attention are single fp16 values (values stay pairs): their
//                  rounding is averaged over the ~2k keys of a decoder sequence (1.3e-5 rms on the logits in the CPU emulation,
//                  tools/precision_budget_mixed8.py) and S = Q K^T needs one UMMA pass instead of two
enum { M5_NUM_FAST = 0, M5_NUM_PRECISE = 1, M5_NUM_MIXED = 2, M5_NUM_MIXED8 = 3, M5_NUM_MIXED8K = 4 };
inline bool num_is_mixed(int mode) { return mode == M5_NUM_MIXED || mode == M5_NUM_MIXED8 || mode == M5_NUM_MIXED8K; }
inline bool num_has_f8(int mode) { return mode == M5_NUM_MIXED8 || mode == M5_NUM_MIXED8K; }

// x (fp32 [rows, D]) is updated in place.
int encoder_layer(m5_ctx* ctx, float* x, const SeqSet& seqs, const EncLayerW& w, int D, int H, int ff, float eps,
                  int mode, const BlockScratch& s);
// mem16: encoder output after its final LayerNorm as fp16 [mem rows, D*(precise?2:1)]
int decoder_layer(m5_ctx* ctx, float* x, const SeqSet& seqs, const __half* mem16, const SeqSet& mem_seqs,
                  const DecLayerW& w, int D, int H, int ff, float eps, int mode, const BlockScratch& s);

// convenience wrappers that count launches
int run_gemm(m5_ctx* ctx, const GemmCall& g);
int run_norm(m5_ctx* ctx, const NormCall& n);
int run_attn(m5_ctx* ctx, const AttnCall& a);

}  // namespace m5
