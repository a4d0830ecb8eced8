#pragma once
#include "m5_internal.h"

namespace m5 {

int voc_features(const int* codes, const float* codebook, float* feat, int N, int Q, int C, int cb, cudaStream_t s);
int voc_im2col(const float* feat, const int* fpos, const int* flen, __half* out, int N, int C, cudaStream_t s);
int voc_dwconv(const float* x, const float* w, const float* b, const int* fpos, const int* flen, float* y, int N, int C,
               cudaStream_t s);
int istft_setup_constants();
int istft_frames(const float* spec, int ld, int n_frames, const float2* w1280, const float2* w640, const float* window,
                 float* frames, cudaStream_t s);
int istft_ola(const float* frames, const float* window, const int* frame0, const int* nframes, int B, int max_frames,
              float* wav, cudaStream_t s);

}  // namespace m5
