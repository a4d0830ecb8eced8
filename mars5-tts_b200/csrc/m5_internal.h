// Internal declarations shared by the kernels and the C-ABI layer of libmars5_b200.so.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/mars5_b200.h"

namespace m5 {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE attribute: a launcher keeps one bit per device (a second
// context on another GPU of the same process must opt in again).  Setting the attribute twice from racing threads is harmless.
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  bool needed(unsigned long long& bit) const {
    int dev = 0;
    cudaGetDevice(&dev);
    bit = 1ull << (dev & 63);
    return (mask.load(std::memory_order_acquire) & bit) == 0;
  }
  void done(unsigned long long bit) { mask.fetch_or(bit, std::memory_order_release); }
};

// ---- tcgen05 GEMM (gemm_tc5.cu) ----------------------------------------------------------------
enum { M5_OUT_F32 = 0, M5_OUT_F16 = 1, M5_OUT_SWIGLU_F16 = 2, M5_OUT_F16_SPLIT = 3, M5_OUT_SWIGLU_F16_SPLIT = 4 };
enum { M5_ACT_NONE = 0, M5_ACT_GELU = 1, M5_ACT_SILU = 2 };

struct GemmCall {
  const __half* A = nullptr;  // [M, K] row stride lda
  const __half* W = nullptr;  // [N, K (or kwrap)] row stride ldw
  int M = 0, N = 0, K = 0;
  int lda = 0, ldw = 0;
  int kwrap = 0;  // >0: W k-coordinate wraps modulo kwrap (A = [hi | lo] against one copy of W)
  int awrap = 0;  // >0: A k-coordinate wraps modulo awrap (A = [hi | lo] against W = [W_hi | W_hi | W_lo])
  const float* bias = nullptr;
  const float* colscale = nullptr;
  void* out = nullptr;
  void* out_lo = nullptr;
  int lo_from_col = 0;  // M5_OUT_F16_SPLIT: lo halves are written for columns >= lo_from_col only (a multiple of 4)
  int ldc = 0;
  int mode = M5_OUT_F32;
  int act = M5_ACT_NONE;
  int accumulate = 0;
  int force_bn = 0;
  // fp8 "lo" pass (mixed8 numerics): after the K fp16 columns of A (the hi halves) the SAME accumulator receives
  // A8[M, K8] (e5m2, the lo halves scaled by 2^-2) x W8[N, K8] (e4m3, the weights scaled by 2^+2) through
  // tcgen05.mma kind::f8f6f4 at twice the fp16 rate.  CTA-pair kernel only (gemm_f8lo_eligible); K8 % 128 == 0.
  const uint8_t* A8 = nullptr; const uint8_t* W8 = nullptr;
  int K8 = 0, lda8 = 0, ldw8 = 0;   // row strides in bytes
  uint8_t* out_lo8 = nullptr; int ldc8 = 0;   // SwiGLU pair output with the lo half as e5m2 (scaled 2^-2)
};
int gemm_tc5(const GemmCall& g, cudaStream_t stream, int num_sms);
bool gemm_f8lo_eligible(int M, int N, int num_sms);   // does this shape run on the CTA-pair kernel?

// ---- skinny (M <= 32) weight-streaming GEMM for the AR decode step (gemm_skinny.cu) ------------
struct SkinnyCall {
  const __half* X = nullptr;  // [B, K] fp16 activations, row stride K
  const __half* W = nullptr;  // [N, K] fp16
  int B = 0, N = 0, K = 0;
  float* out_f32 = nullptr;  // [B, ldc]; accumulate => out += result
  __half* out_f16 = nullptr;  // swiglu mode: [B, N/2]
  int ldc = 0;
  int swiglu = 0;
  int accumulate = 0;
  float* scratch = nullptr;  // split-K partial tiles (ctx-owned, gemm_skinny_scratch_bytes)
  int* counters = nullptr;   // one ticket per 128-row tile, zero between launches
};
size_t gemm_skinny_scratch_bytes(int num_sms);
int gemm_skinny(const SkinnyCall& c, cudaStream_t stream, int num_sms);

// ---- row kernels (rowops.cu) -------------------------------------------------------------------
// y = LayerNorm(x) (affine optional) or RMSNorm(x); x fp32 [M, D] -> fp16 [M, ldo] (+ optional lo halves and fp32 copy).
struct NormCall {
  const float* x = nullptr;
  int M = 0, D = 0, ldx = 0;
  const float* gamma = nullptr;
  const float* beta = nullptr;  // null for RMSNorm
  float eps = 1e-5f;
  int rms = 0;
  __half* out = nullptr;
  __half* out_lo = nullptr;  // optional
  uint8_t* out_lo8 = nullptr;  // optional: lo half as e5m2 scaled by 2^-2 (mixed8 numerics), row stride ldo8 bytes
  int ldo8 = 0;
  float* out_f32 = nullptr;  // optional
  int ldo = 0;
  const int* row_map = nullptr;  // optional: output row i reads input row row_map[i]
};
int norm_rows(const NormCall& c, cudaStream_t stream);

// ---- attention (attention.cu) -------------------------------------------------------------------
// Packed variable-length batch. Sequence s has queries at rows [q_start[s], q_start[s]+q_len[s]) of Q and keys at
// rows [k_start[s], k_start[s]+k_len[s]) of K/V. head_dim is 64. Row strides in elements.
struct AttnCall {
  const __half* Q = nullptr; const __half* K = nullptr; const __half* V = nullptr;
  int ldq = 0, ldk = 0, ldv = 0;
  __half* O = nullptr; int ldo = 0;
  int n_heads = 0, n_seqs = 0, max_q = 0;
  const int* q_start = nullptr; const int* q_len = nullptr;
  const int* k_start = nullptr; const int* k_len = nullptr;
  int causal = 0;  // key j visible to query i iff j <= i + (k_len - q_len)
  float scale = 0.125f;
  double flops_hint = 0.0;  // 4 * 64 * heads * sum_s(q_len*k_len) (halved when causal); profiling only
  int q_rows = 0, k_rows = 0;  // total rows of the Q and K/V buffers (TMA bounds); required by the tcgen05 kernel
  int impl = 0;                // 0 auto, 1 mma.sync kernel, 2 tcgen05 kernel
  // split-precision mode (mma.sync kernel only): low halves of Q/K/V (same strides) and of the output
  const __half* Qlo = nullptr; const __half* Klo = nullptr; const __half* Vlo = nullptr; __half* Olo = nullptr;
  uint8_t* Olo8 = nullptr; int ldo8 = 0;   // tcgen05 pair kernel: lo half of the output as e5m2 scaled by 2^-2 instead of fp16
};
int flash_attn(const AttnCall& c, cudaStream_t stream);      // mma.sync (causal prefill, tiny problems)
int flash_attn_tc5(const AttnCall& c, cudaStream_t stream);  // tcgen05 / TMEM (attention_tc5.cu)

// Single-query attention over the fp16 KV cache of the AR model (decode step).
struct DecodeAttnCall {
  const __half* q = nullptr;  // [B, H*64]
  const __half* kc = nullptr; const __half* vc = nullptr;  // cache for this layer: [B, W, H*64]
  int B = 0, H = 0, W = 0;
  const int* kv_len = nullptr;  // [B] number of valid cached positions (including the current token)
  const int* done = nullptr;    // [B] optional: rows that already stopped are skipped
  __half* out = nullptr;        // [B, H*64]
  float* scratch = nullptr;     // split workspace
  int n_split = 1;
};
int decode_attn(const DecodeAttnCall& c, cudaStream_t stream);
size_t decode_attn_scratch_bytes(int B, int H, int n_split);
int decode_attn_splits_for(int max_kv);  // n_split needed so that every key of a context of max_kv is covered

}  // namespace m5
