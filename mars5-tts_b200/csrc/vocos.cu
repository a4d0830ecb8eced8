// Vocos (encodec-24khz) vocoder kernels that are not GEMMs.  The reference calls the third-party `vocos` package
// (inference.py:119,164-171); its source is not vendored in the reference tree, so these kernels follow the published
// vocos 0.1.0 algorithm as restated in SURVEY.md Appendix C:
//   codes_to_features : features[n] = sum_q codebook[q*1024 + code[n,q]]                   (128 channels)
//   backbone.embed    : Conv1d(128 -> 384, k=7, pad=3)   -> im2col rows here, GEMM in gemm_tc5.cu
//   ConvNeXtBlock     : depthwise Conv1d(384, k=7, pad=3, groups=384) -> AdaLayerNorm -> pw GEMMs
//   ISTFTHead         : mag = min(exp(.),100), phase -> S = mag (cos + j sin) -> irfft(1280) * hann -> overlap-add
//                       (hop 320) -> crop 480 ("same") -> / window^2 envelope
// The vocoder needs fp32-class accuracy (1e-4 RMS on the waveform), so GEMM operands are written as [hi | lo] fp16
// pairs and multiplied against [W_hi | W_hi | W_lo] (three-term split), see gemm_tc5.cu `awrap`.
#include "vocos.h"

#include "ptx.cuh"

namespace m5 {

// features fp32 [N, C]
__global__ void voc_features_kernel(const int* codes, const float* codebook, float* feat, int N, int Q, int C, int cb) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < Q; ++q) s += codebook[((size_t)q * cb + codes[(size_t)n * Q + q]) * C + c];
    feat[(size_t)n * C + c] = s;
  }
}
int voc_features(const int* codes, const float* codebook, float* feat, int N, int Q, int C, int cb, cudaStream_t s) {
  if (N <= 0) return M5_OK;
  voc_features_kernel<<<N, 128, 0, s>>>(codes, codebook, feat, N, Q, C, cb);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

// im2col for the k=7 embed conv with zero padding at utterance edges; output row n = [hi(7*C) | lo(7*C)] fp16
__global__ void voc_im2col_kernel(const float* feat, const int* fpos, const int* flen, __half* out, int N, int C) {
  const int n = blockIdx.x;
  const int pos = fpos[n], len = flen[n];
  const int KC = 7 * C;
  for (int i = threadIdx.x; i < KC; i += blockDim.x) {
    const int tap = i / C, c = i - tap * C;
    const int sp = pos + tap - 3;
    const float v = (sp >= 0 && sp < len) ? feat[(size_t)(n + tap - 3) * C + c] : 0.f;
    const __half h = __float2half_rn(v);
    out[(size_t)n * 2 * KC + i] = h;
    out[(size_t)n * 2 * KC + KC + i] = __float2half_rn(v - __half2float(h));
  }
}
int voc_im2col(const float* feat, const int* fpos, const int* flen, __half* out, int N, int C, cudaStream_t s) {
  if (N <= 0) return M5_OK;
  voc_im2col_kernel<<<N, 256, 0, s>>>(feat, fpos, flen, out, N, C);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

// depthwise conv k=7 along time: y[n, c] = b[c] + sum_j w[c, j] * x[n + j - 3, c]   (zero outside the utterance)
__global__ void voc_dwconv_kernel(const float* x, const float* w, const float* b, const int* fpos, const int* flen,
                                  float* y, int N, int C) {
  const int n = blockIdx.x;
  const int pos = fpos[n], len = flen[n];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = b[c];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int sp = pos + j - 3;
      if (sp >= 0 && sp < len) s += w[c * 7 + j] * x[(size_t)(n + j - 3) * C + c];
    }
    y[(size_t)n * C + c] = s;
  }
}
int voc_dwconv(const float* x, const float* w, const float* b, const int* fpos, const int* flen, float* y, int N, int C,
               cudaStream_t s) {
  if (N <= 0) return M5_OK;
  voc_dwconv_kernel<<<N, 128, 0, s>>>(x, w, b, fpos, flen, y, N, C);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------ iSTFT
// One warp per frame.  N = 1280 real output samples come from a 640-point complex inverse DFT (even/odd packing of
// the half spectrum), evaluated as a 20 x 32 Cooley-Tukey split: lane m2 computes 20 32-point sums over shared memory
// (broadcast reads), applies the 640-th root twiddles, then a fully unrolled 20-point DFT in registers.
static constexpr int NFFT = 1280, MH = 640, N1 = 20, N2 = 32;
static constexpr int ISTFT_WARPS = 4;

__constant__ float2 c_w20[20];   // exp(+2 pi i j / 20)
__constant__ float2 c_w32[32];   // exp(+2 pi i j / 32)

__global__ void __launch_bounds__(ISTFT_WARPS * 32)
istft_frames_kernel(const float* spec, int ld, int n_frames, const float2* w1280, const float2* w640, const float* window,
                    float* frames) {
  extern __shared__ __align__(16) uint8_t is_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * ISTFT_WARPS + warp;
  float2* S = reinterpret_cast<float2*>(is_smem) + (size_t)warp * (MH + 1 + MH);
  float2* Z = S + (MH + 1);
  if (f >= n_frames) return;
  const float* row = spec + (size_t)f * ld;
  for (int k = lane; k <= MH; k += 32) {
    const float mag = fminf(expf(row[k]), 100.0f);
    float sn, cs;
    sincosf(row[MH + 1 + k], &sn, &cs);
    S[k] = make_float2(mag * cs, mag * sn);
  }
  __syncwarp();
  for (int k = lane; k < MH; k += 32) {
    float2 a = S[k], b = S[MH - k];
    if (k == 0) { a.y = 0.f; b.y = 0.f; }  // C2R ignores the imaginary parts of DC and Nyquist
    b.y = -b.y;                            // conj
    const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y + b.y));
    const float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y - b.y));
    const float2 tw = w1280[k];
    const float2 o = make_float2(d.x * tw.x - d.y * tw.y, d.x * tw.y + d.y * tw.x);
    Z[k] = make_float2(e.x - o.y, e.y + o.x);  // E + i*O
  }
  __syncwarp();
  // step A/B: Y[k1] = w640^(k1*m2) * sum_k2 Z[k1 + 20 k2] * w32^(k2*m2)      (m2 = lane)
  float2 Y[N1];
#pragma unroll
  for (int k1 = 0; k1 < N1; ++k1) {
    float ax = 0.f, ay = 0.f;
#pragma unroll 8
    for (int k2 = 0; k2 < N2; ++k2) {
      const float2 z = Z[k1 + N1 * k2];
      const float2 w = c_w32[(k2 * lane) & 31];
      ax += z.x * w.x - z.y * w.y;
      ay += z.x * w.y + z.y * w.x;
    }
    const float2 t = w640[(k1 * lane) % MH];
    Y[k1] = make_float2(ax * t.x - ay * t.y, ax * t.y + ay * t.x);
  }
  // step C: z[32 m1 + m2] = sum_k1 Y[k1] * w20^(k1*m1)
  float* out = frames + (size_t)f * NFFT;
  const float inv = 1.0f / (float)MH;
#pragma unroll
  for (int m1 = 0; m1 < N1; ++m1) {
    float zx = 0.f, zy = 0.f;
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) {
      const float2 w = c_w20[(k1 * m1) % N1];
      zx += Y[k1].x * w.x - Y[k1].y * w.y;
      zy += Y[k1].x * w.y + Y[k1].y * w.x;
    }
    const int m = N2 * m1 + lane;
    float2 r;
    r.x = zx * inv * window[2 * m];
    r.y = zy * inv * window[2 * m + 1];
    *reinterpret_cast<float2*>(out + 2 * m) = r;
  }
}

// overlap-add + "same" crop + envelope normalisation; one thread per output sample, fixed summation order.
// grid.y = utterance; utterance b owns frames [frame0[b], frame0[b] + nframes[b]) and samples [hop*frame0[b], ...).
__global__ void istft_ola_kernel(const float* frames, const float* window, const int* frame0, const int* nframes,
                                 float* wav) {
  const int hop = 320, pad = 480;
  const int b = blockIdx.y;
  const int nf = nframes[b], f0 = frame0[b];
  const int total = nf * hop;
  float* out = wav + (size_t)f0 * hop;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < total; s += gridDim.x * blockDim.x) {
    // sample s (after the crop) sits at padded position p = s + pad; frames f with f*hop <= p < f*hop + 1280
    const int p = s + pad;
    const int fhi = min(p / hop, nf - 1);
    const int flo = p >= NFFT ? (p - NFFT + hop) / hop : 0;
    float acc = 0.f, env = 0.f;
    for (int f = flo; f <= fhi; ++f) {
      const int j = p - f * hop;
      if (j >= 0 && j < NFFT) {
        acc += frames[(size_t)(f0 + f) * NFFT + j];
        const float w = window[j];
        env += w * w;
      }
    }
    out[s] = acc / env;
  }
}

int istft_setup_constants() {
  float2 h20[20], h32[32];
  const double PI = 3.14159265358979323846;
  for (int j = 0; j < 20; ++j) h20[j] = make_float2((float)cos(2 * PI * j / 20), (float)sin(2 * PI * j / 20));
  for (int j = 0; j < 32; ++j) h32[j] = make_float2((float)cos(2 * PI * j / 32), (float)sin(2 * PI * j / 32));
  if (cudaMemcpyToSymbol(c_w20, h20, sizeof(h20)) != cudaSuccess) return M5_ERR_CUDA;
  if (cudaMemcpyToSymbol(c_w32, h32, sizeof(h32)) != cudaSuccess) return M5_ERR_CUDA;
  return M5_OK;
}

int istft_frames(const float* spec, int ld, int n_frames, const float2* w1280, const float2* w640, const float* window,
                 float* frames, cudaStream_t s) {
  if (n_frames <= 0) return M5_OK;
  const size_t smem = (size_t)ISTFT_WARPS * (2 * MH + 1) * sizeof(float2);
  istft_frames_kernel<<<(n_frames + ISTFT_WARPS - 1) / ISTFT_WARPS, ISTFT_WARPS * 32, smem, s>>>(spec, ld, n_frames, w1280,
                                                                                                w640, window, frames);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

int istft_ola(const float* frames, const float* window, const int* frame0, const int* nframes, int B, int max_frames,
              float* wav, cudaStream_t s) {
  if (B <= 0 || max_frames <= 0) return M5_OK;
  dim3 grid(min(64, (max_frames * 320 + 255) / 256), B);
  istft_ola_kernel<<<grid, 256, 0, s>>>(frames, window, frame0, nframes, wav);
  return cudaGetLastError() == cudaSuccess ? M5_OK : M5_ERR_CUDA;
}

}  // namespace m5
