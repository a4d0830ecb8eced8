// Host-side byte-pair-encoding engine for the two minbpe-v1 tokenisers either side of the hot path:
// the text tokeniser (base 256 bytes, reference minbpe/regex.py) and the speech tokeniser over Encodec L0 codes (base 1024,
// reference minbpe/codebook.py).  The reference applies merges with pure-Python loops (`_encode_chunk`: recount every
// adjacent pair and rebuild the list once per merge round, regex.py:92-111 / codebook.py:96-115) on one core, one
// utterance at a time; between the AR and the NAR stage that is ~0.4 s per utterance for a 450-frame prompt.  Here a batch
// of sequences is merged in parallel threads, each with an exact O(n) round: ranks of the adjacent pairs are kept in an
// array next to the ids, a round takes the minimum rank, rewrites the sequence left to right (non-overlapping occurrences,
// like `merge()` in minbpe/base.py:26-41) and only looks up the ranks of pairs that touch a new token.
//
// No CUDA in this file (plain host C++ compiled into the same shared library); results are integers and must equal the
// reference's exactly -- tests/test_bpe_cpu.py checks them against outputs of the unmodified reference (tests/golden).
#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/mars5_b200.h"

struct m5_bpe {
  int base = 0, n_merges = 0;
  // open-addressing hash: key = left << 32 | right, value = merge index (rank); EMPTY_KEY marks a free slot
  std::vector<uint64_t> keys;
  std::vector<int32_t> vals;
  uint64_t mask = 0;
  // expansion of every token into base symbols (CSR)
  std::vector<int64_t> exp_off;
  std::vector<int32_t> exp_sym;
};

namespace {
constexpr uint64_t EMPTY_KEY = ~0ull;
constexpr int32_t NO_RANK = INT32_MAX;

inline uint64_t pair_key(int32_t a, int32_t b) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; }
inline uint64_t mix(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}
inline int32_t rank_of(const m5_bpe* t, int32_t a, int32_t b) {
  const uint64_t key = pair_key(a, b);
  for (uint64_t h = mix(key) & t->mask;; h = (h + 1) & t->mask) {
    if (t->keys[h] == key) return t->vals[h];
    if (t->keys[h] == EMPTY_KEY) return NO_RANK;
  }
}

// One sequence, in place; returns the new length.  Invariant between rounds: rk[i] = rank of (ids[i], ids[i+1]) for
// i < n-1 and rk[n-1] = NO_RANK.
int encode_one(const m5_bpe* t, int32_t* ids, int n, std::vector<int32_t>& rk) {
  if (n < 2) return n;
  rk.resize(n);
  for (int i = 0; i + 1 < n; ++i) rk[i] = rank_of(t, ids[i], ids[i + 1]);
  rk[n - 1] = NO_RANK;
  for (;;) {
    int32_t best = NO_RANK;
    for (int i = 0; i + 1 < n; ++i) best = std::min(best, rk[i]);
    if (best == NO_RANK) break;   // "nothing else can be merged anymore"
    const int32_t tok = t->base + best;
    // Rewrite left to right (write index w <= read index i).  An occurrence is decided on the OLD sequence and skips
    // both of its elements, exactly like minbpe's merge(): (a, a, a) with pair (a, a) becomes (X, a).
    int w = 0;
    bool prev_new = false;        // the element just written is a freshly merged token
    for (int i = 0; i < n;) {
      const bool hit = (i + 1 < n) && rk[i] == best;
      const int32_t sym = hit ? tok : ids[i];
      const int32_t carried = hit ? NO_RANK : rk[i];   // rank of (sym, old next element) while both stay old
      if (w > 0 && (hit || prev_new)) rk[w - 1] = rank_of(t, ids[w - 1], sym);   // a pair with a new token on either side
      ids[w] = sym;
      rk[w] = carried;
      prev_new = hit;
      ++w;
      i += hit ? 2 : 1;
    }
    n = w;
    rk[n - 1] = NO_RANK;
    if (n < 2) break;
  }
  return n;
}
}  // namespace

extern "C" {

int m5_bpe_create(int32_t base, const int32_t* merges, int32_t n_merges, m5_bpe** out) {
  if (!out || base <= 0 || n_merges < 0 || (n_merges > 0 && !merges)) return M5_ERR_ARG;
  m5_bpe* t = new (std::nothrow) m5_bpe();
  if (!t) return M5_ERR_NOMEM;
  t->base = base;
  t->n_merges = n_merges;
  uint64_t cap = 16;
  while (cap < (uint64_t)n_merges * 2 + 2) cap <<= 1;
  t->keys.assign(cap, EMPTY_KEY);
  t->vals.assign(cap, 0);
  t->mask = cap - 1;
  t->exp_off.assign((size_t)base + n_merges + 1, 0);
  for (int i = 0; i < base; ++i) t->exp_off[i + 1] = i + 1;
  t->exp_sym.resize(base);
  for (int i = 0; i < base; ++i) t->exp_sym[i] = i;
  for (int m = 0; m < n_merges; ++m) {
    const int32_t a = merges[2 * m], b = merges[2 * m + 1];
    const int32_t tok = base + m;
    // a merge may only reference tokens that already exist (minbpe's _build_vocab would raise KeyError otherwise)
    if (a < 0 || b < 0 || a >= tok || b >= tok) { delete t; return M5_ERR_ARG; }
    const uint64_t key = pair_key(a, b);
    uint64_t h = mix(key) & t->mask;
    while (t->keys[h] != EMPTY_KEY && t->keys[h] != key) h = (h + 1) & t->mask;
    // a duplicated pair keeps the LAST index, like the dict assignment in Tokenizer.load (base.py:163-165)
    t->keys[h] = key;
    t->vals[h] = m;
    const size_t la = t->exp_off[a + 1] - t->exp_off[a], lb = t->exp_off[b + 1] - t->exp_off[b];
    const size_t at = t->exp_sym.size();
    t->exp_sym.resize(at + la + lb);
    std::copy_n(t->exp_sym.begin() + t->exp_off[a], la, t->exp_sym.begin() + at);
    std::copy_n(t->exp_sym.begin() + t->exp_off[b], lb, t->exp_sym.begin() + at + la);
    t->exp_off[tok + 1] = (int64_t)(at + la + lb);
  }
  *out = t;
  return M5_OK;
}

void m5_bpe_destroy(m5_bpe* t) { delete t; }

int m5_bpe_encode(const m5_bpe* t, const int32_t* ids, const int64_t* offsets, int32_t n_seq, int32_t* out_ids,
                  int32_t* out_len, int32_t n_threads) {
  if (!t || n_seq < 0 || (n_seq > 0 && (!ids || !offsets || !out_ids || !out_len))) return M5_ERR_ARG;
  if (n_seq == 0) return M5_OK;
  for (int s = 0; s < n_seq; ++s)
    if (offsets[s + 1] < offsets[s] || offsets[s + 1] - offsets[s] > INT32_MAX) return M5_ERR_ARG;
  const int64_t total = offsets[n_seq];
  for (int64_t i = offsets[0]; i < total; ++i)
    if (ids[i] < 0 || ids[i] >= t->base + t->n_merges) return M5_ERR_ARG;
  if (out_ids != ids) std::memcpy(out_ids + offsets[0], ids + offsets[0], (size_t)(total - offsets[0]) * sizeof(int32_t));
  int hw = (int)std::thread::hardware_concurrency();
  int nt = n_threads > 0 ? n_threads : std::max(1, hw);
  nt = std::min(nt, (int)n_seq);
  std::atomic<int> next(0);
  auto work = [&]() {
    std::vector<int32_t> rk;
    for (int s; (s = next.fetch_add(1)) < n_seq;)
      out_len[s] = encode_one(t, out_ids + offsets[s], (int)(offsets[s + 1] - offsets[s]), rk);
  };
  if (nt <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (int i = 0; i < nt; ++i) pool.emplace_back(work);
    for (auto& th : pool) th.join();
  }
  return M5_OK;
}

int64_t m5_bpe_expand(const m5_bpe* t, const int32_t* ids, const int64_t* offsets, int32_t n_seq, const int32_t* special_ids,
                      int32_t n_special, int32_t* out_syms, int64_t* out_offsets, int64_t capacity) {
  if (!t || n_seq < 0 || (n_seq > 0 && (!ids || !offsets))) return -M5_ERR_ARG;
  const int32_t n_tok = t->base + t->n_merges;
  int64_t w = 0;
  for (int s = 0; s < n_seq; ++s) {
    if (out_offsets) out_offsets[s] = w;
    for (int64_t i = offsets[s]; i < offsets[s + 1]; ++i) {
      const int32_t id = ids[i];
      if (id >= 0 && id < n_tok) {
        const int64_t a = t->exp_off[id], b = t->exp_off[id + 1];
        if (out_syms) {
          if (w + (b - a) > capacity) return -M5_ERR_ARG;
          std::copy_n(t->exp_sym.begin() + a, b - a, out_syms + w);
        }
        w += b - a;
      } else {
        int k = -1;
        for (int j = 0; j < n_special; ++j)
          if (special_ids[j] == id) { k = j; break; }
        if (k < 0) return -M5_ERR_ARG;   // "invalid token id" (ValueError in the reference)
        if (out_syms) {
          if (w + 1 > capacity) return -M5_ERR_ARG;
          out_syms[w] = -(k + 1);        // special token k, in the order of special_ids
        }
        w += 1;
      }
    }
  }
  if (out_offsets) out_offsets[n_seq] = w;
  return w;
}

}  // extern "C"
