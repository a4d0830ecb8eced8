// Pieces shared by the 1-CTA and the 2-CTA (cta_group::2) tcgen05 GEMM kernels: epilogue parameter block, the per-kind
// store routine and the TMA tensor-map encoder.
#pragma once
#include <cuda.h>
#include <cuda_fp8.h>

#include "m5_internal.h"
#include "ptx.cuh"

namespace m5 {

// epilogue kinds (compile-time)
enum { E_F32 = 0, E_F32_ACC = 1, E_F16 = 2, E_SWIGLU = 3, E_GENERIC = 4, E_F16_SPLIT = 5, E_SWIGLU_SPLIT = 6, E_SWIGLU_SPLIT8 = 7 };

// lo half of an (hi, lo) activation pair as e5m2, scaled by 2^-2 (the fp8 weights carry 2^+2): one byte
__device__ __forceinline__ uint8_t lo_to_e5m2(float lo) { return (uint8_t)__nv_cvt_float_to_fp8(lo * 0.25f, __NV_SATFINITE, __NV_E5M2); }

struct GemmEpi {
  const float* bias;      // [N] fp32 or null
  const float* colscale;  // [N] fp32 or null (Vocos layer-scale gamma)
  void* out;              // fp32 or fp16, row stride ldc (elements)
  void* out_lo;           // split modes: low halves
  int lo_from_col;        // E_F16_SPLIT: columns below it get no lo half (nobody reads it: e.g. the queries of the pair attention)
  int ldc;
  uint8_t* out_lo8;  // E_SWIGLU_SPLIT8: lo halves as e5m2
  int ldc8;
  int mode;        // M5_OUT_*
  int act;         // M5_ACT_*
  int accumulate;  // fp32 out: out += value (residual stream update)
};

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ float silu_f(float a) { return __fdividef(a, 1.0f + __expf(-a)); }  // 2 MUFU + 2 FP ops

// Stores 4 consecutive columns [col, col+4) of one row. v already holds accumulator + bias.
template <int KIND>
__device__ __forceinline__ void epi_store4(const GemmEpi& epi, float (&v)[4], const float (&s4)[4], const float4& prev, int row,
                                           int col, int N, bool full4) {
  if constexpr (KIND == E_F32 || KIND == E_F32_ACC) {
    float* o = reinterpret_cast<float*>(epi.out) + (size_t)row * epi.ldc + col;
    if constexpr (KIND == E_F32_ACC) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] *= s4[e];
    }
    if (full4) {
      float4 w = make_float4(v[0], v[1], v[2], v[3]);
      if constexpr (KIND == E_F32_ACC) { w.x += prev.x; w.y += prev.y; w.z += prev.z; w.w += prev.w; }
      *reinterpret_cast<float4*>(o) = w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (col + e < N) o[e] = (KIND == E_F32_ACC) ? o[e] + v[e] : v[e];
    }
  } else if constexpr (KIND == E_F16) {
    __half* o = reinterpret_cast<__half*>(epi.out) + (size_t)row * epi.ldc + col;
    const __half h0 = __float2half_rn(v[0]), h1 = __float2half_rn(v[1]), h2 = __float2half_rn(v[2]), h3 = __float2half_rn(v[3]);
    if (full4) {
      *reinterpret_cast<uint2*>(o) = make_uint2(pack_h2(h0, h1), pack_h2(h2, h3));
    } else {
      if (col + 0 < N) o[0] = h0;
      if (col + 1 < N) o[1] = h1;
      if (col + 2 < N) o[2] = h2;
    }
  } else if constexpr (KIND == E_SWIGLU) {
    // columns (2j, 2j+1) = (W_j x, V_j x) -> silu(W x) * V x ; N is even
    __half* o = reinterpret_cast<__half*>(epi.out) + (size_t)row * epi.ldc + (col >> 1);
    const __half h0 = __float2half_rn(silu_f(v[0]) * v[1]), h1 = __float2half_rn(silu_f(v[2]) * v[3]);
    if (full4) *reinterpret_cast<uint32_t*>(o) = pack_h2(h0, h1);
    else if (col + 1 < N) o[0] = h0;
  } else if constexpr (KIND == E_F16_SPLIT) {
    // fp16 (hi, lo) pair of every value: hi = rn(v), lo = rn(v - hi); the mixed / precise NAR modes
    __half* o = reinterpret_cast<__half*>(epi.out) + (size_t)row * epi.ldc + col;
    __half* ol = reinterpret_cast<__half*>(epi.out_lo) + (size_t)row * epi.ldc + col;
    const bool want_lo = col >= epi.lo_from_col;   // uniform over the 4 columns (lo_from_col % 4 == 0)
    __half h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { h[e] = __float2half_rn(v[e]); l[e] = __float2half_rn(v[e] - __half2float(h[e])); }
    if (full4) {
      *reinterpret_cast<uint2*>(o) = make_uint2(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]));
      if (want_lo) *reinterpret_cast<uint2*>(ol) = make_uint2(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]));
    } else {
#pragma unroll
      for (int e = 0; e < 3; ++e)
        if (col + e < N) { o[e] = h[e]; if (want_lo) ol[e] = l[e]; }
    }
  } else if constexpr (KIND == E_SWIGLU_SPLIT) {
    __half* o = reinterpret_cast<__half*>(epi.out) + (size_t)row * epi.ldc + (col >> 1);
    __half* ol = reinterpret_cast<__half*>(epi.out_lo) + (size_t)row * epi.ldc + (col >> 1);
    const float g0 = silu_f(v[0]) * v[1], g1 = silu_f(v[2]) * v[3];
    const __half h0 = __float2half_rn(g0), h1 = __float2half_rn(g1);
    const __half l0 = __float2half_rn(g0 - __half2float(h0)), l1 = __float2half_rn(g1 - __half2float(h1));
    if (full4) {
      *reinterpret_cast<uint32_t*>(o) = pack_h2(h0, h1);
      *reinterpret_cast<uint32_t*>(ol) = pack_h2(l0, l1);
    } else if (col + 1 < N) { o[0] = h0; ol[0] = l0; }
  } else if constexpr (KIND == E_SWIGLU_SPLIT8) {
    __half* o = reinterpret_cast<__half*>(epi.out) + (size_t)row * epi.ldc + (col >> 1);
    uint8_t* o8 = epi.out_lo8 + (size_t)row * epi.ldc8 + (col >> 1);
    const float g0 = silu_f(v[0]) * v[1], g1 = silu_f(v[2]) * v[3];
    const __half h0 = __float2half_rn(g0), h1 = __float2half_rn(g1);
    const uint8_t l0 = lo_to_e5m2(g0 - __half2float(h0)), l1 = lo_to_e5m2(g1 - __half2float(h1));
    if (full4) {
      *reinterpret_cast<uint32_t*>(o) = pack_h2(h0, h1);
      *reinterpret_cast<uint16_t*>(o8) = (uint16_t)l0 | ((uint16_t)l1 << 8);
    } else if (col + 1 < N) { o[0] = h0; o8[0] = l0; }
  } else {  // E_GENERIC: activations, column scale, split (hi | lo) outputs -- cold paths (vocoder, timestep MLPs, precise mode)
    if (epi.act == M5_ACT_GELU) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = 0.5f * v[e] * (1.0f + erff(v[e] * 0.70710678118654752440f));
    } else if (epi.act == M5_ACT_SILU) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = silu_f(v[e]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] *= s4[e];
    if (epi.mode == M5_OUT_F32) {
      float* o = reinterpret_cast<float*>(epi.out) + (size_t)row * epi.ldc + col;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (col + e < N) o[e] = epi.accumulate ? o[e] + v[e] : v[e];
    } else if (epi.mode == M5_OUT_F16 || epi.mode == M5_OUT_F16_SPLIT) {
      __half* o = reinterpret_cast<__half*>(epi.out) + (size_t)row * epi.ldc + col;
      __half* ol = reinterpret_cast<__half*>(epi.out_lo) + (size_t)row * epi.ldc + col;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (col + e < N) {
          const __half h = __float2half_rn(v[e]);
          o[e] = h;
          if (epi.mode == M5_OUT_F16_SPLIT) ol[e] = __float2half_rn(v[e] - __half2float(h));
        }
      }
    } else {  // SwiGLU (optionally split)
      __half* o = reinterpret_cast<__half*>(epi.out) + (size_t)row * epi.ldc + (col >> 1);
      __half* ol = reinterpret_cast<__half*>(epi.out_lo) + (size_t)row * epi.ldc + (col >> 1);
      const float g[2] = {silu_f(v[0]) * v[1], silu_f(v[2]) * v[3]};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (col + 2 * e + 1 < N) {
          const __half h = __float2half_rn(g[e]);
          o[e] = h;
          if (epi.mode == M5_OUT_SWIGLU_F16_SPLIT) ol[e] = __float2half_rn(g[e] - __half2float(h));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp16 row-major [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128B swizzle.
static inline int make_tmap_k64(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return M5_ERR_CUDA;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? M5_OK : M5_ERR_CUDA;
}


// 2-D byte matrix [rows, cols] (fp8 operands) with row stride ld BYTES; box = [box_rows, 128 bytes], 128B swizzle.
static inline int make_tmap_u8_k128(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return M5_ERR_CUDA;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? M5_OK : M5_ERR_CUDA;
}

static inline int gemm_epi_kind(const GemmCall& g) {
  const bool plain = g.act == M5_ACT_NONE;
  if (plain && g.mode == M5_OUT_SWIGLU_F16_SPLIT && !g.colscale && g.out_lo8) return E_SWIGLU_SPLIT8;
  if (plain && g.mode == M5_OUT_F32 && !g.accumulate && !g.colscale) return E_F32;
  if (plain && g.mode == M5_OUT_F32 && g.accumulate) return E_F32_ACC;
  if (plain && g.mode == M5_OUT_F16 && !g.colscale) return E_F16;
  if (plain && g.mode == M5_OUT_SWIGLU_F16 && !g.colscale) return E_SWIGLU;
  if (plain && g.mode == M5_OUT_F16_SPLIT && !g.colscale) return E_F16_SPLIT;
  if (plain && g.mode == M5_OUT_SWIGLU_F16_SPLIT && !g.colscale) return E_SWIGLU_SPLIT;
  return E_GENERIC;
}

int gemm_tc5_2cta(const GemmCall& g, cudaStream_t stream, int num_sms);  // gemm_tc5_2cta.cu

}  // namespace m5
