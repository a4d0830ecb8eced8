// Fused categorical sampler of the AR loop: one CTA per utterance row performs the whole logit-warping chain of
// mars5/ar_generate.py:73-118 on the fp32 logits produced by the vocabulary projection and appends the sampled id.
//
//   1. frequency/presence penalty over the last `window` GENERATED ids, only once >= 2 were generated
//      (samplers.py:20-36, gate ar_generate.py:77)
//   2. mask ids < text_vocab-1 (the off-by-one keeps id text_vocab-1 alive; ar_generate.py:82-83)
//   3. early-EOS penalty  z[eos] -= factor * max(est - n, 1)^decay  while n <= est  (samplers.py:39-56)
//   4. z /= temperature
//   5. top-k: exact k-th largest by an 8-bit radix select; everything < kth is removed, ties kept (samplers.py:70-74)
//   6. top-p over the descending-sorted survivors with the "keep the first token over the threshold" shift
//      (samplers.py:76-91)
//   6b. typical-p (samplers.py:96-122; identity for mass > 0.999): over the tokens still alive, score_i = |-log p_i - H|;
//      the tokens are ranked by score, the score at the first rank whose cumulative probability reaches `mass` is the
//      threshold, every token with a larger score is removed (ties at the threshold survive)
//   7. log_softmax, p = exp(logp), sample argmax_i p_i / e_i with e_i ~ Exp(1) -- exactly torch.multinomial's n=1 path
//      (ar_generate.py:102-118).  e_i comes from the caller's noise tensor (parity) or Philox4x32-10 keyed by
//      (seed, utterance id, step, i) (production).
//   8. EOS stops the row without being appended (ar_generate.py:121-135).
#include "sampler_body.cuh"

namespace m5 {

__global__ void __launch_bounds__(SP_THREADS) ar_sample_kernel(SampleCall p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t sp_smem[];
  ar_sample_row(p, blockIdx.x, sp_smem);
}

size_t sample_smem_bytes(int V, int cap) { return (size_t)((V + 3) & ~3) * 4 + (size_t)cap * 16; }  // z | sv | si | typical-p scores, probs

int sample_cap(int V, int top_k) {
  int want = top_k > 0 ? min(V, 2 * top_k + 64) : V;
  int cap = 64;
  while (cap < want) cap <<= 1;
  return cap;
}

int ar_sample(SampleCall& c, cudaStream_t stream) {
  if (c.B <= 0) return M5_OK;
  c.cap = sample_cap(c.V, c.cfg.top_k);
  const size_t smem = sample_smem_bytes(c.V, c.cap);
  // the opt-in shared-memory size is a per-device, per-function attribute: set on every launch (no process-wide cache)
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(ar_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return M5_ERR_CUDA;
  return launch_k(ar_sample_kernel, dim3(c.B), dim3(SP_THREADS), smem, stream, c) == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

}  // namespace m5
