#pragma once
#include "m5_internal.h"

namespace m5 {

// AR sampler (sampler.cu). Loop mode (ids != null) appends into the per-row id buffer and maintains the row state;
// debug mode only reports out_tok / out_logprobs.
struct SampleCall {
  const float* logits = nullptr;  // [B, ld_logits] fp32
  int ld_logits = 0;
  int B = 0, V = 0, text_vocab = 0;
  m5_ar_cfg cfg{};
  // history of generated ids: either the id buffer itself (hist_is_ids: row b = ids[b*hist_stride ...], generated
  // tokens are the last n_gen of the first tok_len entries) or a plain [B, hist_stride] array of generated ids.
  const int* hist = nullptr;
  int hist_stride = 0;
  int hist_is_ids = 0;
  int* n_gen = nullptr;          // [B]
  const int* n_phones = nullptr; // [B] or null
  const float* noise = nullptr;  // [B, noise_steps, V] Exp(1) draws or null
  int noise_steps = 0;
  uint64_t seed = 0;
  const int64_t* utt_ids = nullptr;
  int* out_tok = nullptr;        // [B] optional
  float* out_logprobs = nullptr; // [B, V] optional
  float* logits_dump = nullptr;  // [B, dump_steps, V] optional
  int dump_steps = 0;
  // loop-mode state
  int* ids = nullptr;      // [B, hist_stride]
  int* tok_len = nullptr;  // [B]
  int* kv_len = nullptr;   // [B]
  int* done = nullptr;     // [B]
  int* n_done = nullptr;   // [1]
  int cap = 0;             // filled by ar_sample
};
int ar_sample(SampleCall& c, cudaStream_t stream);
int sample_cap(int V, int top_k);              // survivor slots the sort works on
size_t sample_smem_bytes(int V, int cap);      // dynamic shared memory of one sampler CTA

// NAR posterior + sampling (posterior.cu): one warp per (row, codebook).
struct PosteriorCall {
  const float* cond = nullptr;    // [R, ld] logits of the conditional pass for ONE codebook, or [R, Q, ld] when q_stride>0
  const float* uncond = nullptr;
  int ld = 0;                     // row stride (floats) of the logits
  int R = 0;                      // rows (sequence positions)
  int K = 1025;                   // classes
  int Q = 8;                      // codebooks in x
  int q = 0;                      // codebook handled by this call (-1: all Q, logits laid out [R, Q, ld])
  float guidance_w = 3.f, x0_temp = 0.7f;
  // schedule scalars for this step (fp32 table entries, diffuser.py:92-95)
  float log_alpha_t = 0, log_1m_alpha_t = 0, log_cum_tm1 = 0, log_1m_cum_tm1 = 0;
  int t = 0;
  const int* row_map = nullptr;   // optional: logits row i corresponds to x row row_map[i]
  const int* x_t = nullptr;       // [Rx, Q] current codes
  int* x_out = nullptr;           // [Rx, Q]
  const float* u = nullptr;       // optional uniforms [R, Q, K] (parity) laid out like the reference's rand_like
  int u_rows_are_x = 0;           // u indexed by x row (1) or logits row (0)
  uint64_t seed = 0;
  const int64_t* row_utt = nullptr;  // [Rx] utterance id per x row (Philox stream), optional
  const int* row_pos = nullptr;      // [Rx] position inside the utterance (Philox counter), optional
  int draw = 0;                      // draw index inside the step (0: unknown sample, 1: known re-noise)
};
int nar_posterior(const PosteriorCall& c, cudaStream_t stream);

// Known-region re-noise q_sample(x_known, t) (diffuser.py:230-236, 386-390) and the final merge (diffuser.py:393, 467-468)
struct RenoiseCall {
  int R = 0, Q = 8, K = 1025;
  const int* x_known = nullptr;      // [R, Q]
  const uint8_t* known = nullptr;    // [R, Q] 1 where the code is known (mask m)
  int* x = nullptr;                  // [R, Q] in/out: unknown entries already hold x_{t-1} samples
  const int* x_q0 = nullptr;         // [R] clean L0 codes
  float log_cum_t = 0, log_1m_cum_t = 0;
  int t = 0;
  int q0_override = 0;               // write clean L0 when q0_override_steps < t
  const float* u = nullptr;          // optional uniforms [R, Q, K]
  uint64_t seed = 0;
  const int64_t* row_utt = nullptr;
  const int* row_pos = nullptr;
  // forward = 1: RePaint forward step x_t -> x_{t+1} (forward_diffusion, diffuser.py:336-342): every entry is re-drawn
  // from q_pred_one_timestep of ITSELF (x_known aliases x, `known` is ignored, log_cum_t / log_1m_cum_t carry
  // log_alpha[t] / log_1_min_alpha[t], no t == 0 shortcut)
  int forward = 0;
};
int nar_renoise(const RenoiseCall& c, cudaStream_t stream);

float philox_uniform_host(uint64_t seed, uint64_t stream, uint32_t step, uint32_t idx);

}  // namespace m5
