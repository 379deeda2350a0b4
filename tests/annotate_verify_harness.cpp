// Host harness of mirge3.0_b200/csrc/annotate_verify.cuh (built by tests/test_annotate_verify_host.py with g++): the
// helpers of the annotation kernels compiled for the host, driven as an exhaustive scan over every alignment start
// (what the kernel's "degenerate query" path does for very short queries).
#define ANNOTATE_VERIFY_HOST
#include "annotate_verify.cuh"

// canonical pick (minimum hit word) of key under the round policy, MIRGE_NO_HIT when nothing is valid;
// *searched = 0 when the round does not search this key (window empty, no T{3,} tail for the poly-T round)
extern "C" uint64_t hv_best_hit(const mirge_library *lib, const mirge_round_policy *pol, const uint32_t *key, int *searched,
                                uint32_t *query_words_out, int *qlen_out) {
  const KeyView kv = key_view(key);
  int tlen = -1, qs = 0, qe = 0;
  *searched = 0;
  *qlen_out = 0;
  if (!round_window(kv, *pol, tlen, qs, qe)) return MIRGE_NO_HIT;
  *searched = 1;
  static uint32_t qw[QW_MAX], qnx[QW_MAX];
  build_query(kv, qs, qe, qw, qnx);
  const int L = qe - qs, R = pol->seed_len == 0 ? L : min(pol->seed_len, L);
  *qlen_out = L;
  for (int w = 0; w < (L + 15) / 16; ++w) { query_words_out[2 * w] = qw[w]; query_words_out[2 * w + 1] = qnx[w]; }
  uint64_t best = MIRGE_NO_HIT;
  for (uint64_t a = 0; a + (uint64_t)L <= lib->n_bases; ++a) {
    const uint64_t h = verify(*lib, qw, qnx, L, *pol, R, a, (uint32_t)a);
    if (h < best) best = h;
  }
  return best;
}

// seed pieces as search_round cuts them: out[2 * pi] = first base, out[2 * pi + 1] = k-mer of the piece's first 16 bases
extern "C" int hv_pieces(const uint32_t *qw, int L, int seed_len, int seed_mm, uint32_t *out) {
  const int R = seed_len == 0 ? L : min(seed_len, L), np = seed_mm + 1, nw = (L + 15) >> 4;
  for (int pi = 0; pi <= np; ++pi) out[2 * pi] = (uint32_t)piece_bound(pi, R, np);
  for (int pi = 0; pi < np; ++pi) out[2 * pi + 1] = query_kmer16(qw, piece_bound(pi, R, np), nw);
  return np;
}
