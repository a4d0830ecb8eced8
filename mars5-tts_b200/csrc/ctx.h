// Context object behind the C ABI: device, stream, weights (caller-owned device pointers), grow-only workspace arena.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "m5_internal.h"

struct m5_ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  m5_model_cfg cfg{};
  std::map<std::string, m5_tensor> weights;
  std::string last_error;
  int64_t launches = 0;

  // optional kernel-class profiling (bench.py roofline leg): event pairs resolved at the end of a public call
  struct ProfRec { cudaEvent_t a, b; int kind; double flops, bytes; int64_t n; };
  bool prof_on = false;
  std::vector<ProfRec> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[4] = {0, 0, 0, 0}, prof_flops[4] = {0, 0, 0, 0}, prof_bytes[4] = {0, 0, 0, 0};
  int64_t prof_n[4] = {0, 0, 0, 0};

  // grow-only device arena, reset at the start of every public call
  char* arena = nullptr;
  size_t arena_cap = 0, arena_off = 0;
  std::vector<void*> retired;  // older, smaller arenas kept until the call ends
  // pinned host staging
  char* pinned = nullptr;
  size_t pinned_cap = 0;

  // derived tables (device, fp32), built at create
  float* rope_inv_freq = nullptr;  // [32]
  float* pe_ar = nullptr;          // [max_pos, ar_dim]
  float* pe_nar = nullptr;         // [max_pos, nar_dim]
  float* twiddle = nullptr;        // iSTFT tables
  float* skinny_scratch = nullptr; // split-K partial tiles of the decode GEMMs
  int* skinny_counters = nullptr;  // [1024] tickets

  int fail(int code, const std::string& msg) {
    last_error = msg;
    return code;
  }
};

namespace m5 {

struct Arena {
  m5_ctx* c;
  explicit Arena(m5_ctx* ctx) : c(ctx) {}
  // Reserve total capacity up front (one cudaMalloc per high-water mark), then bump-allocate.
  int reserve(size_t bytes);
  template <typename T>
  T* get(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    if (c->arena_off + bytes > c->arena_cap) return nullptr;
    T* p = reinterpret_cast<T*>(c->arena + c->arena_off);
    c->arena_off += bytes;
    return p;
  }
};

const m5_tensor* find_weight(m5_ctx* c, const std::string& name);
// profiling helpers: bracket one launch with events when profiling is on
cudaEvent_t prof_begin(m5_ctx* c);
void prof_end(m5_ctx* c, cudaEvent_t a, int kind, double flops, double bytes, int64_t n = 1);
void prof_resolve(m5_ctx* c);  // call after a stream synchronize
template <typename T>
inline const T* W(m5_ctx* c, const std::string& name) {
  const m5_tensor* t = find_weight(c, name);
  return t ? reinterpret_cast<const T*>(t->ptr) : nullptr;
}

#define M5_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ctx->fail(M5_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));       \
  } while (0)
#define M5_TRY(call)                                                                           \
  do {                                                                                         \
    int _r = (call);                                                                           \
    if (_r != M5_OK) {                                                                         \
      if (ctx->last_error.empty()) ctx->last_error = std::string(#call) + " failed";           \
      return _r;                                                                               \
    }                                                                                          \
  } while (0)

}  // namespace m5
