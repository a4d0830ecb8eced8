#pragma once
#include "m5_internal.h"

namespace m5 {

int cast_rows(const float* x, int ldx, __half* o, __half* olo, int ldo, int M, int D, const int* row_map,
              cudaStream_t stream);
int rope_kv(__half* qkv, int ld, int n_rows, int H, const int* row_seq, const int* row_pos, __half* kc, __half* vc,
            int W, const float* inv_freq, cudaStream_t stream);
int rope_kv_decode(const float* qkv, int B, int H, const int* len, __half* qout, __half* kc, __half* vc, int W,
                   const float* inv_freq, const int* active, cudaStream_t stream);

struct EmbedCall {
  const int* codes = nullptr;     // [n_code_rows, Q]
  const int* code_row = nullptr;  // [n_rows] index into codes, or -1 for the identity slot
  const int* pos = nullptr;       // [n_rows] position for the sinusoidal term
  const __half* tables = nullptr; // [Q, n_codes, D/Q] fp16 (checkpoint weights are fp16-exact)
  const float* identity = nullptr;  // [D]
  const float* pe = nullptr;        // [max_pos, D] or null
  const float* add_vec = nullptr;   // [D] or null (timestep embedding)
  float alpha = 1.f;
  int n_rows = 0, D = 0, Q = 8, n_codes = 1025;
  float* out = nullptr;  // [n_rows, D]
};
int chunked_embed(const EmbedCall& c, cudaStream_t stream);

struct TokEmbedCall {
  const int* tok = nullptr;  // [n_rows]; negative => vec_rows[-tok-1]
  const int* pos = nullptr;
  const __half* table = nullptr;   // [vocab, D]
  const float* vec_rows = nullptr; // [*, D] fp32 substitute rows (speaker vectors)
  const float* pe = nullptr;
  const float* add_vec = nullptr;
  float alpha = 1.f;
  int n_rows = 0, D = 0;
  float* out = nullptr;
};
int token_embed(const TokEmbedCall& c, cudaStream_t stream);
int gather_rows(const float* x, int ldx, const int* idx, float* out, int ldo, int n, int D, cudaStream_t stream);

}  // namespace m5
