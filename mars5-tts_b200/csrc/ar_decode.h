// Parameter block of the fused persistent AR decode kernel (ar_decode.cu).
#pragma once
#include <algorithm>

#include "m5_internal.h"
#include "sampler.h"

namespace m5 {

struct ArLayerDev {   // device-resident per-layer weight pointers
  const float* attn_norm; const float* ffn_norm;
  const __half* wqkv; const __half* wo; const __half* w13; const __half* w2;
};

// One weight-streaming GEMM phase: out[b, n] = sum_k X[b, k] W[n, k]; work item = (128-row tile, K slice).
struct ArGemm {
  const __half* W = nullptr;
  int N = 0, K = 0, kslice = 0, ksplit = 1, tiles = 0;
};

struct ArDecodeParams {
  int B = 0, n_layers = 0, D = 0, F = 0, H = 0, V = 0, Wc = 0;
  float eps = 1e-5f;
  const ArLayerDev* layers = nullptr;   // device [n_layers]
  const float* final_norm = nullptr; const __half* embed = nullptr;
  const float* inv_freq = nullptr;
  // row state (device)
  const int* ids = nullptr; int ids_stride = 0; const int* tok_len = nullptr; const int* kv_len = nullptr; const int* done = nullptr;
  // activations (device)
  float* x = nullptr;        // [B, D] fp32 residual stream
  float* ssq = nullptr;      // [D / 128][32] per-tile row sums of squares of x
  int ssq_tiles = 0;
  float* qkv = nullptr;      // [B, 3D] fp32
  __half* g16 = nullptr;     // [B, F]
  float* logits = nullptr;   // [B, V]
  float* attn_part = nullptr;  // [B, H, n_split, 68] split-KV partials (m, l, -, -, acc[64])
  int n_split = 1;
  int split_keys = 256;         // cached keys per attention work item (multiple of 32), ar_decode_split_keys()
  int* attn_tickets = nullptr;  // [B * H] zero between steps: the last split of a (row, head) merges
  __half* att16 = nullptr;      // [B, D] merged attention output (fp16, what the reference's SDPA returns)
  __half* kc = nullptr; __half* vc = nullptr;   // [layer][B][Wc][D] fp16
  float* scratch = nullptr;  // split-K partial tiles, ar_decode_scratch_floats()
  int* counters = nullptr;   // one ticket per row tile, zero between phases
  unsigned* gbar = nullptr;  // device-wide barrier counter
  unsigned long long* prof = nullptr;   // optional [1 + 5 * n_layers][2] globaltimer stamps of CTA 0 (work done, barrier passed)
  ArGemm g_qkv, g_wo, g_w13, g_w2, g_out;   // g_out.W = vocabulary projection
  // last phase: the categorical sampler of the step (sampler_body.cuh), CTA b = row b, after one more device-wide barrier.
  // fuse_sample = 0: the caller launches ar_sample_kernel behind this kernel instead.
  int fuse_sample = 0;
  SampleCall sample;
};

int ar_decode_plan(ArDecodeParams& p, int num_sms);          // fills the ArGemm splits and ssq_tiles
size_t ar_decode_scratch_floats(const ArDecodeParams& p);
int ar_decode_max_tiles(const ArDecodeParams& p);
int ar_decode_split_keys(int B, int H, int max_kv, int num_sms);
int ar_decode_splits_for(int max_kv, int split_keys);
size_t ar_decode_attn_floats(int B, int H, int n_split);
bool ar_decode_can_fuse_sampler(int V, int top_k);           // the sampler's shared-memory working set fits the kernel's
int ar_decode_launch(const ArDecodeParams& p, int num_sms, cudaStream_t stream);   // memset(gbar) + cooperative launch

}  // namespace m5
