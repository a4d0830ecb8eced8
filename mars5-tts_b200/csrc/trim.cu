// Leading / trailing silence trim of the synthesised waveforms -- the step right after the vocoder in the reference's tts()
// (inference.py:304-305 -> mars5/trim.py:110-177, a port of librosa.effects.trim): frame-wise mean power over 2048-sample
// frames every 512 samples of the reflect-padded signal, expressed in dB below the loudest frame; the output is the sample
// range from the first to one past the last frame above -top_db.  
//
// Host code, like the reference (which trims `final_audio.cpu()`): one pass of prefix sums of squares in double precision
// per waveform, waveforms of a batch in parallel threads.  The decision per frame is the reference's
//   10 log10(max(1e-10, p_f)) - 10 log10(max(1e-10, max_f p_f)) > -top_db
// evaluated in double instead of float32; the two can only differ for a frame within rounding distance of the threshold.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/mars5_b200.h"

namespace {
// y reflect-padded by `pad` on both sides (torch F.pad mode="reflect": the edge sample is not repeated)
inline float padded_at(const float* y, int64_t n, int64_t pad, int64_t i) {
  int64_t j = i - pad;
  if (j < 0) j = -j;
  else if (j >= n) j = 2 * (n - 1) - j;
  return y[j];
}

int trim_one(const float* y, int64_t n, double top_db, int frame_length, int hop, int64_t* start, int64_t* end) {
  const int64_t pad = frame_length / 2;
  if (n <= pad) return M5_ERR_ARG;   // reflect padding needs pad < n (torch raises here too)
  const int64_t total = n + 2 * pad;
  const int64_t n_frames = 1 + (total - frame_length) / hop;
  std::vector<double> prefix((size_t)total + 1);
  prefix[0] = 0.0;
  for (int64_t i = 0; i < total; ++i) {
    const double v = padded_at(y, n, pad, i);
    prefix[i + 1] = prefix[i] + v * v;
  }
  std::vector<double> power((size_t)n_frames);
  double peak = 0.0;
  for (int64_t f = 0; f < n_frames; ++f) {
    power[f] = (prefix[f * hop + frame_length] - prefix[f * hop]) / frame_length;
    peak = std::max(peak, power[f]);
  }
  const double amin = 1e-10;
  const double ref_db = 10.0 * std::log10(std::max(amin, peak));
  int64_t first = -1, last = -1;
  for (int64_t f = 0; f < n_frames; ++f) {
    const double db = 10.0 * std::log10(std::max(amin, power[f])) - ref_db;
    if (db > -top_db) {
      if (first < 0) first = f;
      last = f;
    }
  }
  if (first < 0) { *start = 0; *end = 0; return M5_OK; }   // "the signal only contains zeros"
  *start = first * hop;
  *end = std::min<int64_t>(n, (last + 1) * hop);
  return M5_OK;
}
}  // namespace

extern "C" int m5_trim_bounds(int32_t B, const float* wav, const int64_t* offsets, float top_db, int32_t frame_length,
                              int32_t hop_length, int64_t* start, int64_t* end, int32_t n_threads) {
  if (B < 0 || (B > 0 && (!wav || !offsets || !start || !end)) || frame_length < 2 || hop_length < 1 || top_db < 0.f)
    return M5_ERR_ARG;
  std::atomic<int> next(0), status(M5_OK);
  auto work = [&]() {
    for (int b; (b = next.fetch_add(1)) < B;) {
      const int rc = trim_one(wav + offsets[b], offsets[b + 1] - offsets[b], top_db, frame_length, hop_length, &start[b], &end[b]);
      if (rc != M5_OK) status.store(rc);
    }
  };
  int nt = n_threads > 0 ? n_threads : std::max(1, (int)std::thread::hardware_concurrency());
  nt = std::min(nt, (int)B);
  if (nt <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int i = 0; i < nt; ++i) pool.emplace_back(work);
    for (auto& th : pool) th.join();
  }
  return status.load();
}
