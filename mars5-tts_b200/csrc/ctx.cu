// Context lifetime, workspace arena, derived tables and the kernel-level debug entry points of the C ABI.
#include "ctx.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "rowops.h"

namespace m5 {

bool g_use_pdl = false;  // set only while the AR decode step is being captured into its CUDA graph

int Arena::reserve(size_t bytes) {
  c->arena_off = 0;
  if (bytes <= c->arena_cap) return M5_OK;
  if (c->arena) {
    cudaStreamSynchronize(c->stream);
    cudaFree(c->arena);
    c->arena = nullptr;
    c->arena_cap = 0;
  }
  bytes = (bytes + (size_t(1) << 20)) & ~((size_t(1) << 20) - 1);
  if (cudaMalloc(&c->arena, bytes) != cudaSuccess) {
    cudaGetLastError();
    return c->fail(M5_ERR_NOMEM, "workspace cudaMalloc failed (" + std::to_string(bytes >> 20) + " MiB)");
  }
  c->arena_cap = bytes;
  return M5_OK;
}

const m5_tensor* find_weight(m5_ctx* c, const std::string& name) {
  auto it = c->weights.find(name);
  if (it == c->weights.end()) {
    c->last_error = "missing weight tensor: " + name;
    return nullptr;
  }
  return &it->second;
}

static cudaEvent_t prof_get(m5_ctx* c) {
  if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
cudaEvent_t prof_begin(m5_ctx* c) {
  if (!c->prof_on) return nullptr;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(c->stream, &st);
  if (st != cudaStreamCaptureStatusNone) return nullptr;
  cudaEvent_t a = prof_get(c);
  cudaEventRecord(a, c->stream);
  return a;
}
void prof_end(m5_ctx* c, cudaEvent_t a, int kind, double flops, double bytes, int64_t n) {
  if (!a) return;
  cudaEvent_t b = prof_get(c);
  cudaEventRecord(b, c->stream);
  c->prof_pending.push_back({a, b, kind, flops, bytes, n});
  if (c->prof_pending.size() >= 4096) { cudaStreamSynchronize(c->stream); prof_resolve(c); }
}
void prof_resolve(m5_ctx* c) {
  for (auto& r : c->prof_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      c->prof_ms[r.kind] += ms; c->prof_flops[r.kind] += r.flops; c->prof_bytes[r.kind] += r.bytes; c->prof_n[r.kind] += r.n;
    }
    c->prof_pool.push_back(r.a); c->prof_pool.push_back(r.b);
  }
  c->prof_pending.clear();
  cudaGetLastError();
}

}  // namespace m5

using namespace m5;

extern "C" {

int m5_create(int device, const m5_model_cfg* cfg, const m5_tensor* tensors, int32_t n_tensors, m5_ctx** out) {
  if (!cfg || !out) return M5_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return M5_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return M5_ERR_CUDA;
  if (prop.major != 10) {
    fprintf(stderr, "libmars5_b200: device %d is sm_%d%d; this library is built for sm_100a only\n", device, prop.major,
            prop.minor);
    return M5_ERR_STATE;
  }
  m5_ctx* ctx = new (std::nothrow) m5_ctx();
  if (!ctx) return M5_ERR_NOMEM;
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->cfg = *cfg;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return M5_ERR_CUDA;
  }
  for (int i = 0; i < n_tensors; ++i) ctx->weights[tensors[i].name] = tensors[i];

  // RoPE inverse frequencies: 1 / (10000 ** (arange(0,64,2).float() / 64)) in fp32 (nn_future.py:194-198)
  float inv[32];
  for (int i = 0; i < 32; ++i) inv[i] = 1.0f / powf(10000.0f, (float)(2 * i) / 64.0f);
  if (cudaMalloc(&ctx->rope_inv_freq, sizeof(inv)) != cudaSuccess ||
      cudaMemcpy(ctx->rope_inv_freq, inv, sizeof(inv), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMalloc(&ctx->skinny_scratch, gemm_skinny_scratch_bytes(ctx->num_sms)) != cudaSuccess ||
      cudaMalloc(&ctx->skinny_counters, 1024 * sizeof(int)) != cudaSuccess ||
      cudaMemset(ctx->skinny_counters, 0, 1024 * sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    m5_destroy(ctx);   // frees whatever was allocated
    return M5_ERR_NOMEM;
  }
  // The caller may override the three derived tables with torch-computed ones (bit-exact with the reference):
  //   "tab.rope_inv_freq" [32], "tab.pe_ar" [max_pos, ar_dim], "tab.pe_nar" [max_pos, nar_dim]
  *out = ctx;
  return M5_OK;
}

void m5_destroy(m5_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->rope_inv_freq) cudaFree(ctx->rope_inv_freq);
  if (ctx->twiddle) cudaFree(ctx->twiddle);
  if (ctx->skinny_scratch) cudaFree(ctx->skinny_scratch);
  if (ctx->skinny_counters) cudaFree(ctx->skinny_counters);
  prof_resolve(ctx);   // returns pending profiling events to the pool
  for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* m5_last_error(m5_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int m5_sync(m5_ctx* ctx) {
  if (!ctx) return M5_ERR_ARG;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return ctx->fail(M5_ERR_CUDA, std::string("sync: ") + cudaGetErrorString(e));
  return M5_OK;
}

int64_t m5_launch_count(m5_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* m5_stream(m5_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int m5_profile_enable(m5_ctx* ctx, int32_t on) {
  if (!ctx) return M5_ERR_ARG;
  cudaStreamSynchronize(ctx->stream);
  prof_resolve(ctx);
  ctx->prof_on = on != 0;
  for (int i = 0; i < 4; ++i) { ctx->prof_ms[i] = ctx->prof_flops[i] = ctx->prof_bytes[i] = 0; ctx->prof_n[i] = 0; }
  return M5_OK;
}
int m5_profile_read(m5_ctx* ctx, int32_t kind, double* ms, double* flops, double* bytes, int64_t* launches) {
  if (!ctx || kind < 0 || kind >= 4) return M5_ERR_ARG;
  cudaStreamSynchronize(ctx->stream);
  prof_resolve(ctx);
  if (ms) *ms = ctx->prof_ms[kind];
  if (flops) *flops = ctx->prof_flops[kind];
  if (bytes) *bytes = ctx->prof_bytes[kind];
  if (launches) *launches = ctx->prof_n[kind];
  return M5_OK;
}
int m5_num_sms(m5_ctx* ctx) { return ctx ? ctx->num_sms : 0; }

// ------------------------------------------------------------------------------------------------ debug entry points
int m5_dbg_gemm(m5_ctx* ctx, const void* A, const void* Wt, int32_t M, int32_t N, int32_t K, int32_t kwrap,
                const float* bias, const float* colscale, void* out, void* out_lo, int32_t ldc, int32_t mode,
                int32_t act, int32_t accumulate, int32_t force_bn) {
  if (!ctx) return M5_ERR_ARG;
  GemmCall g;
  g.A = (const __half*)A; g.W = (const __half*)Wt; g.M = M; g.N = N; g.K = K; g.lda = K;
  g.ldw = kwrap > 0 ? kwrap : K; g.kwrap = kwrap; g.bias = bias; g.colscale = colscale; g.out = out; g.out_lo = out_lo;
  g.ldc = ldc; g.mode = mode; g.act = act; g.accumulate = accumulate; g.force_bn = force_bn;
  int r = gemm_tc5(g, ctx->stream, ctx->num_sms);
  if (r != M5_OK) return ctx->fail(r, "gemm_tc5 launch failed");
  ctx->launches += 1;
  return M5_OK;
}

int m5_dbg_gemm_f8lo(m5_ctx* ctx, const void* A16, int32_t lda, const void* A8, const void* W16, const void* W8, int32_t M,
                     int32_t N, int32_t K, float* out, int32_t ldc) {
  if (!ctx || !A16 || !A8 || !W16 || !W8 || !out) return M5_ERR_ARG;
  GemmCall g;
  g.A = (const __half*)A16; g.W = (const __half*)W16; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldw = K;
  g.A8 = (const uint8_t*)A8; g.W8 = (const uint8_t*)W8; g.K8 = K; g.lda8 = K; g.ldw8 = K;
  g.out = out; g.ldc = ldc; g.mode = M5_OUT_F32;
  int r = gemm_tc5(g, ctx->stream, ctx->num_sms);
  if (r != M5_OK) return ctx->fail(r, "gemm_tc5 (fp8 lo pass) failed: shape not eligible for the CTA-pair kernel?");
  ctx->launches += 1;
  return M5_OK;
}

int m5_dbg_norm(m5_ctx* ctx, const float* x, int32_t M, int32_t D, const float* gamma, const float* beta, float eps,
                int32_t rms, void* out_f16, void* out_lo_f16) {
  if (!ctx) return M5_ERR_ARG;
  NormCall c;
  c.x = x; c.M = M; c.D = D; c.ldx = D; c.gamma = gamma; c.beta = beta; c.eps = eps; c.rms = rms;
  c.out = (__half*)out_f16; c.out_lo = (__half*)out_lo_f16; c.ldo = D;
  int r = norm_rows(c, ctx->stream);
  if (r != M5_OK) return ctx->fail(r, "norm_rows failed");
  ctx->launches += 1;
  return M5_OK;
}

int m5_dbg_attn(m5_ctx* ctx, const void* Q, const void* K, const void* V, int32_t ldq, int32_t ldk, int32_t ldv,
                void* O, int32_t ldo, int32_t n_heads, int32_t n_seqs, int32_t max_q, const int32_t* q_start,
                const int32_t* q_len, const int32_t* k_start, const int32_t* k_len, int32_t causal, int32_t impl,
                int32_t q_rows, int32_t k_rows) {
  if (!ctx) return M5_ERR_ARG;
  AttnCall c;
  c.Q = (const __half*)Q; c.K = (const __half*)K; c.V = (const __half*)V; c.ldq = ldq; c.ldk = ldk; c.ldv = ldv;
  c.O = (__half*)O; c.ldo = ldo; c.n_heads = n_heads; c.n_seqs = n_seqs; c.max_q = max_q; c.q_start = q_start;
  c.q_len = q_len; c.k_start = k_start; c.k_len = k_len; c.causal = causal; c.q_rows = q_rows; c.k_rows = k_rows;
  int r = impl == 2 ? flash_attn_tc5(c, ctx->stream) : flash_attn(c, ctx->stream);
  if (r != M5_OK) return ctx->fail(r, "flash_attn failed");
  ctx->launches += 1;
  return M5_OK;
}

int m5_dbg_attn_split(m5_ctx* ctx, const void* Q, const void* K, const void* V, const void* Klo, const void* Vlo,
                      int32_t ldq, int32_t ldk, int32_t ldv, void* O, void* Olo, int32_t ldo, int32_t n_heads,
                      int32_t n_seqs, int32_t max_q, const int32_t* q_start, const int32_t* q_len, const int32_t* k_start,
                      const int32_t* k_len, int32_t q_rows, int32_t k_rows) {
  if (!ctx || !Vlo || !Olo) return M5_ERR_ARG;   // Klo == NULL: keys are single fp16 values (mixed8k)
  AttnCall c;
  c.Q = (const __half*)Q; c.K = (const __half*)K; c.V = (const __half*)V; c.Klo = (const __half*)Klo; c.Vlo = (const __half*)Vlo;
  c.ldq = ldq; c.ldk = ldk; c.ldv = ldv; c.O = (__half*)O; c.Olo = (__half*)Olo; c.ldo = ldo; c.n_heads = n_heads;
  c.n_seqs = n_seqs; c.max_q = max_q; c.q_start = q_start; c.q_len = q_len; c.k_start = k_start; c.k_len = k_len;
  c.q_rows = q_rows; c.k_rows = k_rows;
  int r = flash_attn_tc5(c, ctx->stream);
  if (r != M5_OK) return ctx->fail(r, "flash_attn_tc5 (split) failed");
  ctx->launches += 1;
  return M5_OK;
}

int m5_dbg_decode_attn(m5_ctx* ctx, const void* q, const void* kc, const void* vc, int32_t B, int32_t H, int32_t W,
                       const int32_t* kv_len, void* out, int32_t n_split) {
  if (!ctx) return M5_ERR_ARG;
  n_split = std::max(n_split, decode_attn_splits_for(W));  // the kernel owns a fixed 128-key slice per split
  Arena ar(ctx);
  M5_TRY(ar.reserve(decode_attn_scratch_bytes(B, H, n_split) + 4096));
  DecodeAttnCall c;
  c.q = (const __half*)q; c.kc = (const __half*)kc; c.vc = (const __half*)vc; c.B = B; c.H = H; c.W = W;
  c.kv_len = kv_len; c.out = (__half*)out; c.n_split = n_split;
  c.scratch = ar.get<float>(decode_attn_scratch_bytes(B, H, n_split) / 4);
  int r = decode_attn(c, ctx->stream);
  if (r != M5_OK) return ctx->fail(r, "decode_attn failed");
  ctx->launches += 2;
  return M5_OK;
}

}  // extern "C"
