/*
 * libmars5_b200.so — C ABI of the B200-native MARS5 hot path (AR decode -> multinomial-DDPM NAR -> Vocos iSTFT).
 *
 * The reference (Camb-ai/MARS5-TTS) is pure PyTorch and has no FFI; the three call sites this library replaces are
 *   ar_generate(...)                inference.py:260-269   (def mars5/ar_generate.py:16-21)
 *   perform_simple_inference(...)   inference.py:296-298   (def mars5/diffuser.py:399-400)
 *   Mars5TTS.vocode(tokens)         inference.py:160-172   (3rd-party vocos.decode)
 * Each entry point below names the reference function it stands in for.  Conventions: extern "C", int status
 * returns (0 = ok), no exceptions, plain pointers and sizes, no torch types.  One context per GPU, not thread-safe.
 * Data buffers are HOST or DEVICE pointers as selected by `mem` (M5_MEM_*); per-utterance length arrays are always
 * HOST.  All work is enqueued on the context's own CUDA stream; calls return after the results are in the caller's
 * buffers (host) or enqueued (device; use m5_sync).  Packed layouts: utterance b's rows follow utterance b-1's.
 */
#ifndef MARS5_B200_H
#define MARS5_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M5_OK 0
#define M5_ERR_ARG 1
#define M5_ERR_CUDA 2
#define M5_ERR_STATE 3
#define M5_ERR_NOMEM 4
#define M5_ERR_MISSING_WEIGHT 5

#define M5_MEM_HOST 0
#define M5_MEM_DEVICE 1

#define M5_DT_F16 0
#define M5_DT_F32 1
#define M5_DT_U8 2 /* fp8 (e4m3) weight copies of the mixed8 numerics, one byte per element */

typedef struct m5_ctx m5_ctx;

/* Model hyper-parameters (reference: inference.py:105-110, mars5/model.py:44-67,165-242; Appendix C of SURVEY.md). */
typedef struct {
  /* AR CodecLM */
  int32_t ar_dim, ar_heads, ar_layers, ar_hidden, ar_vocab, ar_text_vocab, ar_spk_layers, ar_spk_ff;
  float ar_norm_eps;
  float ar_pos_alpha; /* CodecLM.pos_embedding.alpha (learned scalar, nn_future.py:47) */
  /* NAR ResidualTransformer */
  int32_t nar_dim, nar_heads, nar_enc_layers, nar_dec_layers, nar_spk_layers, nar_ff, nar_text_vocab;
  int32_t n_classes; /* 1025 */
  int32_t n_quant;   /* 8 */
  float ln_eps;      /* 4e-5, model.py:13 */
  float head_ln_eps; /* 1e-5, model.py:237 */
  float nar_pos_alpha, nar_cond_alpha, nar_ref_alpha; /* pos_embedding / cond_pos_embedding / ref_pos_embedding .alpha */
  /* Vocos encodec-24khz head */
  int32_t voc_feat, voc_dim, voc_inter, voc_layers, voc_nfft, voc_hop, voc_n_bw, voc_codebook;
  /* limits used to size workspaces */
  int32_t max_pos; /* rows of the sinusoidal tables */
} m5_model_cfg;

/* Named weight tensor living in DEVICE memory owned by the caller (kept alive for the context's lifetime). */
typedef struct {
  const char* name;
  const void* ptr;
  int64_t numel;
  int32_t dtype; /* M5_DT_* */
} m5_tensor;

int m5_create(int device, const m5_model_cfg* cfg, const m5_tensor* tensors, int32_t n_tensors, m5_ctx** out);
void m5_destroy(m5_ctx* ctx);
const char* m5_last_error(m5_ctx* ctx);
int m5_sync(m5_ctx* ctx);
/* Number of kernels this library launched on ctx's stream since creation (bench.py's gpu_launches). */
int64_t m5_launch_count(m5_ctx* ctx);
int m5_num_sms(m5_ctx* ctx);
/* cudaStream_t the context enqueues on (for the caller's CUDA events). */
void* m5_stream(m5_ctx* ctx);
/* Kernel-class timing with CUDA events on the context's stream (bench.py's roofline leg).  kind 0 = tcgen05 GEMM,
 * 1 = flash attention, 2 = the AR decode loop (one record per m5_ar_generate call: `launches` counts decode steps, bytes
 * = weights + KV-cache bytes those steps had to move).  m5_profile_read returns, accumulated since the
 * last enable: total device ms, algorithmic FLOPs (2*M*N*K per GEMM; 4*q*k*64 per attention head pair), algorithmic
 * bytes and number of launches. */
int m5_profile_enable(m5_ctx* ctx, int32_t on);
int m5_profile_read(m5_ctx* ctx, int32_t kind, double* ms, double* flops, double* bytes, int64_t* launches);

/* ---- AR: replaces ar_generate (mars5/ar_generate.py:15-165) for B independent utterances ------------------- */
typedef struct {
  float temperature;     /* InferenceConfig.temperature */
  int32_t top_k;         /* 0 disables */
  float top_p;           /* 1.0 disables */
  float alpha_frequency; /* freq_penalty */
  float alpha_presence;  /* presence_penalty */
  int32_t penalty_window;
  float eos_penalty_decay, eos_penalty_factor;
  int32_t max_len;      /* total length incl. prompt (ar_generate.py:62) */
  int32_t eos_id;       /* len(texttok.vocab) + speechtok '<|endofspeech|>' (ar_generate.py:47) */
  int32_t force_len;    /* benchmark only: >0 masks EOS until this many tokens were generated, then forces it */
  int32_t sync_every;   /* host polls the all-done flag every this many steps (default 16) */
  float typical_p;      /* InferenceConfig.typical_p: locally-typical mass tau (samplers.py:96-122); > 0.999 (or 0) disables */
} m5_ar_cfg;

/*
 * prompt_ids  [sum prompt_len]      text + (offset) speech-BPE ids, as built at inference.py:255
 * spk_codes   [sum spk_len][8]      reference Encodec codes (spk_ref_codec, inference.py:242)
 * n_phones_gen[B] (host)            round(factor*len(text)) for the EOS penalty (inference.py:268)
 * noise       [B][noise_steps][V]   optional Exp(1) draws replacing torch.multinomial's (parity mode), else NULL
 * out_ids     [B][max_len]          prompt followed by generated ids (EOS not appended, ar_generate.py:135)
 * out_len, hit_maxlen [B] (host)
 * logits_dump [B][dump_steps][V]    optional raw fp32 logits of the first dump_steps steps (same `mem` as data)
 */
int m5_ar_generate(m5_ctx* ctx, int32_t B, const int32_t* prompt_ids, const int32_t* prompt_len,
                   const int32_t* spk_codes, const int32_t* spk_len, const int32_t* n_phones_gen,
                   const m5_ar_cfg* cfg, int32_t mem, const float* noise, int32_t noise_steps, uint64_t seed,
                   const int64_t* utt_ids, int32_t* out_ids, int32_t* out_len, int32_t* hit_maxlen,
                   float* logits_dump, int32_t dump_steps);

/* ---- NAR: replaces perform_simple_inference (mars5/diffuser.py:398-472) ---------------------------------- */
typedef struct {
  int32_t T;          /* reverse steps (default_T = 200, inference.py:113) */
  float x0_temp;      /* DSH.x_0_temp */
  float guidance_w;   /* DSH.guidance_w */
  int32_t q0_override_steps;
  int32_t deep_clone;
  int32_t precise;    /* NAR numerics: 0 fast   = every GEMM / attention operand is one fp16 value (fp32 accumulate);
                       *               1 precise = every activation operand is an fp16 (hi, lo) pair (1e-5 on the logits);
                       *               2 mixed   = GEMM activations, keys and values are pairs, queries and probabilities
                       *                           single fp16, attention on tcgen05: the cheapest setting that keeps the
                       *                           logits within 1e-3 max-abs of the fp32 reference
                       *               3 mixed8  = mixed, with the lo half of the big decoder GEMMs' activation pairs as an
                       *                           fp8 (e5m2 x e4m3) tcgen05 pass into the same accumulator
                       *               4 mixed8k = mixed8, with the keys of the decoder self-attention as single fp16 values
                       *                           (values stay pairs): one S = Q K^T pass instead of two */
  /* optional HOST tables [4][T]: log_alpha, log_1_min_alpha, log_cumprod_alpha, log_1_min_cumprod_alpha
   * (MultinomialDiffusion.__init__, diffuser.py:76-95); NULL -> computed inside the library */
  const float* schedule;
  /* RePaint resampling (get_schedule, diffuser.py:318-333; DSH.jump_len / jump_n_sample): 0 or 1 = none (what
   * inference.py:291 uses).  With jumps the reverse loop is interleaved with forward steps x_t -> x_{t+1} ~
   * q_pred_one_timestep (diffuser.py:119-134,336-342).  scaled_forward = DSH.enable_kevin_scaled_inference: the
   * reference's q_pred_one_timestep_scaled (diffuser.py:136-159) raises a broadcasting error for every S != 8, so
   * scaled_forward != 0 together with jumps is rejected with M5_ERR_ARG. */
  int32_t jump_len, jump_n_sample, scaled_forward;
} m5_nar_cfg;

/*
 * c_text   [sum c_text_len]      text BPE ids
 * c_codes  [sum c_codes_len][8]  reference codes (prompt)
 * x_l0     [sum x_len]           AR L0 codes (the `_x[...,0]` column, inference.py:282)
 * x_init   [sum x_len][8]        optional initial randint draw (diffuser.py:409) for parity, else NULL (Philox)
 * noise    [n_steps][2][sum S][8][K]  optional uniforms for the rand_like draws (parity, tiny cases), else NULL; n_steps =
 *          T without jumps, else len(get_schedule) - 1; a reverse step uses draw 0 (unknown sample) and draw 1 (known
 *          re-noise), a forward step draw 0
 * out_codes[sum x_len][8]        result after the deep-clone crop (diffuser.py:471)
 */
int m5_nar_infer(m5_ctx* ctx, int32_t B, const int32_t* c_text, const int32_t* c_text_len, const int32_t* c_codes,
                 const int32_t* c_codes_len, const int32_t* x_l0, const int32_t* x_len, const m5_nar_cfg* cfg,
                 int32_t mem, const int32_t* x_init, const float* noise, uint64_t seed, const int64_t* utt_ids,
                 int32_t* out_codes);

/* One ResidualTransformer.forward (mars5/model.py:264-343): x [sum S][8] codes at timestep t -> logits
 * [sum S][8][n_classes] fp32 (the reference's (bs,S,K,8) permuted as reverse_diffusion does, diffuser.py:359). */
int m5_nar_forward(m5_ctx* ctx, int32_t B, const int32_t* c_text, const int32_t* c_text_len, const int32_t* c_codes,
                   const int32_t* c_codes_len, const int32_t* x, const int32_t* x_len, int32_t t, int32_t drop_cond,
                   int32_t precise, int32_t mem, float* logits_out);

/* One CodecLM.forward over full prompts (mars5/model.py:95-141, no cache): logits [sum prompt_len][V] fp32. */
int m5_ar_forward(m5_ctx* ctx, int32_t B, const int32_t* prompt_ids, const int32_t* prompt_len,
                  const int32_t* spk_codes, const int32_t* spk_len, int32_t mem, float* logits_out);

/* ---- Vocoder: replaces Mars5TTS.vocode (inference.py:160-172) -------------------------------------------- */
/* codes [sum n_frames][8] -> wav [sum 320*n_frames] fp32 */
int m5_vocode(m5_ctx* ctx, int32_t B, const int32_t* codes, const int32_t* n_frames, int32_t bandwidth_id,
              int32_t mem, float* wav_out);

/* Vocoder + silence trim without leaving the device (inference.py:304-305: trim(vocode(tokens), top_db=cfg.trim_db),
 * mars5/trim.py:110-177).  wav_out receives the UNTRIMMED waveforms exactly like m5_vocode; start / end [B] (HOST) receive
 * the sample range of utterance b inside its own waveform -- the same numbers m5_trim_bounds returns for it (frame powers
 * over frame_length-sample frames every hop_length samples of the reflect-padded signal, top_db below the loudest frame). */
int m5_vocode_trim(m5_ctx* ctx, int32_t B, const int32_t* codes, const int32_t* n_frames, int32_t bandwidth_id,
                   int32_t mem, float top_db, int32_t frame_length, int32_t hop_length, float* wav_out, int64_t* start,
                   int64_t* end);

/* ---- Encodec 24 kHz encoder + RVQ: replaces `self.codec.encode(ref_audio[None])` at
 * inference.py:233 (EncodecModel.encodec_model_24khz() at 6 kbps, inference.py:87-88) for B reference clips at once.
 * wav [sum n_samples] fp32 mono 24 kHz (HOST or DEVICE per `mem`), n_samples [B] (host); codes_out [sum ceil(n_b/320)][n_q]
 * int32, clip after clip, n_q = 8 at 6 kbps.  Needs the "enc.*" tensors in the table passed to m5_create
 * (mars5_tts_b200.weights.repack_encodec); M5_ERR_MISSING_WEIGHT otherwise. */
int m5_encodec_encode(m5_ctx* ctx, int32_t B, const float* wav, const int32_t* n_samples, int32_t mem, int32_t n_q,
                      int32_t* codes_out);

/* ---- Tokenisers either side of the path (host code, no GPU work) ----------------------------
 * Merge engine for the two "minbpe v1" tokenisers: text (base = 256 bytes, mars5/minbpe/regex.py) and speech
 * (base = 1024 Encodec L0 codes, mars5/minbpe/codebook.py).  The regex split of text into chunks and the special-token
 * handling stay with the caller (mars5_tts_b200/bpe.py mirrors RegexTokenizer / CodebookTokenizer on top of this). */
typedef struct m5_bpe m5_bpe;

/* merges[n_merges][2]: pair i creates token id base + i -- the line order of the model file
 * (Tokenizer.load, minbpe/base.py:140-168; CodebookTokenizer.load, minbpe/codebook.py:178-205). */
int m5_bpe_create(int32_t base, const int32_t* merges, int32_t n_merges, m5_bpe** out);
void m5_bpe_destroy(m5_bpe* bpe);

/* Applies the merge table to n_seq independent id sequences, sequence s = ids[offsets[s] .. offsets[s+1]), exactly like
 * `_encode_chunk` (regex.py:92-111 / codebook.py:96-115): repeatedly the adjacent pair with the lowest merge index, all of
 * its non-overlapping occurrences left to right.  Sequence s is written to out_ids + offsets[s] (out_ids may alias ids),
 * its new length to out_len[s].  n_threads <= 0: one per hardware thread (capped by n_seq). */
int m5_bpe_encode(const m5_bpe* bpe, const int32_t* ids, const int64_t* offsets, int32_t n_seq, int32_t* out_ids,
                  int32_t* out_len, int32_t n_threads);

/* Expands tokens into base symbols (what Tokenizer.decode / CodebookTokenizer.decode_int produce, regex.py:77-90,
 * codebook.py:74-94).  An id outside the merge table must be one of special_ids and is emitted as -(k+1), k its index
 * in special_ids; any other id is an error (the reference raises ValueError).  Returns the number of symbols (also
 * when out_syms is NULL: size query), or -M5_ERR_ARG.  out_offsets[n_seq+1] (optional) receives the per-sequence
 * boundaries. */
int64_t m5_bpe_expand(const m5_bpe* bpe, const int32_t* ids, const int64_t* offsets, int32_t n_seq,
                      const int32_t* special_ids, int32_t n_special, int32_t* out_syms, int64_t* out_offsets,
                      int64_t capacity);

/* ---- Silence trim after the vocoder (host code) -----------------------------------------------
 * Replaces trim() of mars5/trim.py:110-177 as called at inference.py:305: for waveform b = wav[offsets[b] ..
 * offsets[b+1]) (host, fp32 mono) the sample range [start[b], end[b]) from the first to one past the last frame whose
 * mean power is within top_db of the loudest frame (frames of frame_length every hop_length samples of the reflect-padded
 * signal); (0, 0) when no frame qualifies.  A waveform not longer than frame_length/2 is an error, as in the reference. */
int m5_trim_bounds(int32_t B, const float* wav, const int64_t* offsets, float top_db, int32_t frame_length,
                   int32_t hop_length, int64_t* start, int64_t* end, int32_t n_threads);

/* ---- kernel-level entry points (device pointers) used by tests/ and bench.py's roofline leg -------------- */
int m5_dbg_gemm(m5_ctx* ctx, const void* A_f16, const void* W_f16, int32_t M, int32_t N, int32_t K, int32_t kwrap,
                const float* bias, const float* colscale, void* out, void* out_lo, int32_t ldc, int32_t mode,
                int32_t act, int32_t accumulate, int32_t force_bn);
/* fp16 hi pass + fp8 lo pass into one accumulator (mixed8): out[M][N] fp32 = A16[M][K] W16[N][K]^T + A8[M][K] W8[N][K]^T, A8 e5m2
 * (lo halves x 2^-2), W8 e4m3 (weights x 2^+2); CTA-pair kernel (M, N large enough), K % 128 == 0. */
int m5_dbg_gemm_f8lo(m5_ctx* ctx, const void* A16, int32_t lda, const void* A8, const void* W16, const void* W8, int32_t M,
                     int32_t N, int32_t K, float* out, int32_t ldc);
int m5_dbg_skinny(m5_ctx* ctx, const void* X_f16, const void* W_f16, int32_t B, int32_t N, int32_t K, float* out_f32,
                  void* out_f16, int32_t ldc, int32_t swiglu, int32_t accumulate);
int m5_dbg_norm(m5_ctx* ctx, const float* x, int32_t M, int32_t D, const float* gamma, const float* beta, float eps,
                int32_t rms, void* out_f16, void* out_lo_f16);
int m5_dbg_attn(m5_ctx* ctx, const void* Q, const void* K, const void* V, int32_t ldq, int32_t ldk, int32_t ldv,
                void* O, int32_t ldo, int32_t n_heads, int32_t n_seqs, int32_t max_q, const int32_t* q_start,
                const int32_t* q_len, const int32_t* k_start, const int32_t* k_len, int32_t causal, int32_t impl,
                int32_t q_rows, int32_t k_rows); /* impl: 1 = mma.sync kernel, 2 = tcgen05 kernel (needs q_rows/k_rows) */
/* tcgen05 kernel with keys / values as fp16 (hi, lo) pairs and the output written as a pair ("mixed" numerics).
 * Klo == NULL: the keys are single fp16 values, only the values are pairs ("mixed8k"). */
int m5_dbg_attn_split(m5_ctx* ctx, const void* Q, const void* K, const void* V, const void* Klo, const void* Vlo,
                      int32_t ldq, int32_t ldk, int32_t ldv, void* O, void* Olo, int32_t ldo, int32_t n_heads,
                      int32_t n_seqs, int32_t max_q, const int32_t* q_start, const int32_t* q_len, const int32_t* k_start,
                      const int32_t* k_len, int32_t q_rows, int32_t k_rows);
int m5_dbg_decode_attn(m5_ctx* ctx, const void* q, const void* kc, const void* vc, int32_t B, int32_t H, int32_t W,
                       const int32_t* kv_len, void* out, int32_t n_split);
/* AR sampler chain on fp32 logits [B][V] (ar_generate.py:73-118): writes the chosen token per row. */
int m5_dbg_sample(m5_ctx* ctx, const float* logits, int32_t B, int32_t V, const m5_ar_cfg* cfg, int32_t text_vocab,
                  const int32_t* hist, int32_t hist_stride, const int32_t* n_gen, const int32_t* n_phones,
                  const float* noise, uint64_t seed, int32_t* out_tok, float* out_logprobs);
/* One reverse-diffusion posterior + sample (diffuser.py:359-393) from cond/uncond logits [R][8][K]. */
/* sched6 (HOST): log_alpha[t], log_1_min_alpha[t], log_cumprod_alpha[t-1], log_1_min_cumprod_alpha[t-1],
 *                log_cumprod_alpha[t], log_1_min_cumprod_alpha[t] */
int m5_dbg_posterior(m5_ctx* ctx, const float* cond, const float* uncond, int32_t R, int32_t t, const float* sched6,
                     float guidance_w, float x0_temp, const int32_t* x_t, const int32_t* x_known, const uint8_t* mask,
                     const float* u_unknown, const float* u_known, uint64_t seed, int32_t* x_out);
/* ISTFT head on [N][nfft+2] (mag-logits | phase) rows -> wav [hop*N] (Appendix C). */
int m5_dbg_istft(m5_ctx* ctx, const float* spec, int32_t B, const int32_t* n_frames_dev_host, float* wav);

#ifdef __cplusplus
}
#endif
#endif
