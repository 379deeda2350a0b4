// annotate_verify.cuh -- the per-sequence arithmetic of the annotation kernels that restates bowtie's policies
// (3P): the query window of a round, 2-bit query words with always-mismatch masks, seed piece
// boundaries, k-mer extraction, and the verification of one alignment start (text under the round policy, reference
// bounds, ambiguous reference bases).  Included by annotate.cu; tests/test_annotate_verify_host.py compiles the same
// text for the host (ANNOTATE_VERIFY_HOST) and holds an exhaustive scan built from these functions against the oracle.
#pragma once
#ifdef ANNOTATE_VERIFY_HOST
#include "host_shims.h"
#else
#include "common.cuh"
#endif

#define ANN_THREADS 128
#define QW_MAX ((MIRGE_MAX_READ_LEN + 15) / 16)
#define MIN_SEED 4

__device__ __forceinline__ uint32_t lib_base(const uint32_t *packed, uint32_t p) { return (packed[p >> 4] >> (2 * (p & 15))) & 3u; }

// 16 bases starting at base position p as a 2-bit word (base p in bits 0..1).  The packed text is followed by
// MIRGE_LIB_PAD_WORDS zero words (mirge_library contract), so a candidate that runs past the end of the library
// -- rejected afterwards by its reference bounds -- reads zeros, and no bound check is needed here.
__device__ __forceinline__ uint32_t lib_word16(const uint32_t *packed, uint64_t p) {
  const uint64_t w = p >> 4;
  const uint32_t sh = 2 * (uint32_t)(p & 15);
  return __funnelshift_r(packed[w], packed[w + 1], sh);
}

// reference holding base `pos`: coarse block table, then a short forward scan (largest r with ref_off[r] <= pos)
__device__ __forceinline__ uint32_t find_ref(const mirge_library &lib, uint32_t pos) {
  uint32_t r = lib.d_ref_block[pos >> lib.ref_block_shift];
  while (r + 1 < lib.n_refs && lib.d_ref_off[r + 1] <= pos) ++r;
  return r;
}

// Hamming distance of the query (2-bit words qw, always-mismatch mask qnx, L bases) against library text
// at astart under the round policy, text only (XOR + popcount on packed words); returns the mismatch
// count or -1 when the policy is violated.
__device__ __forceinline__ int verify_text(const mirge_library &lib, const uint32_t *qw, const uint32_t *qnx, int L,
                                           const mirge_round_policy &pol, int R, uint64_t astart) {
  int mm = 0, smm = 0;
  const int nw = (L + 15) >> 4;
  for (int w = 0; w < nw; ++w) {
    const uint32_t refw = lib_word16(lib.d_packed, astart + 16 * (uint64_t)w);
    uint32_t x = qw[w] ^ refw;
    x = ((x | (x >> 1)) & 0x55555555u) | qnx[w];
    const int rem = L - 16 * w;
    if (rem < 16) x &= (1u << (2 * rem)) - 1u;
    if (x) {
      mm += __popc(x);
      const int srem = R - 16 * w;
      if (srem >= 16) smm += __popc(x);
      else if (srem > 0) smm += __popc(x & ((1u << (2 * srem)) - 1u));
      if (mm > pol.total_mm || smm > pol.seed_mm) return -1;
    }
  }
  return mm;
}

// any ambiguous reference base under [a, b)
__device__ __forceinline__ bool ref_has_n(const mirge_library &lib, uint64_t a, uint64_t b) {
  for (uint64_t w = a >> 5; w <= (b - 1) >> 5; ++w) {
    uint32_t bits = lib.d_nmask[w];
    if (w == (a >> 5)) bits &= 0xFFFFFFFFu << (a & 31);
    if (w == ((b - 1) >> 5) && (b & 31)) bits &= 0xFFFFFFFFu >> (32 - (b & 31));
    if (bits) return true;
  }
  return false;
}

// full check of one alignment start: text first (rejects almost every candidate), then the reference
// it falls in, its bounds and ambiguous bases; returns the packed hit or NO_HIT
__device__ __forceinline__ uint64_t verify(const mirge_library &lib, const uint32_t *qw, const uint32_t *qnx, int L,
                                           const mirge_round_policy &pol, int R, uint64_t astart, uint32_t pos_in_ref) {
  const int mm = verify_text(lib, qw, qnx, L, pol, R, astart);
  if (mm < 0) return MIRGE_NO_HIT;
  const uint32_t r = find_ref(lib, pos_in_ref);
  const uint32_t rlo = lib.d_ref_off[r], rhi = lib.d_ref_off[r + 1];
  if (astart < rlo || astart + (uint64_t)L > rhi) return MIRGE_NO_HIT;
  if (ref_has_n(lib, astart, astart + L)) return MIRGE_NO_HIT;
  return ((uint64_t)mm << 56) | ((uint64_t)r << 28) | (uint64_t)(astart - rlo);
}

// 16 query bases starting at base a, first base most significant (the index's k-mer order)
__device__ __forceinline__ uint32_t query_kmer16(const uint32_t *qw, int a, int nw) {
  const int wi = a >> 4, sh = 2 * (a & 15);
  const uint32_t lo = qw[wi], hi = (wi + 1 < nw) ? qw[wi + 1] : 0u;
  const uint32_t v = sh ? __funnelshift_r(lo, hi, sh) : lo;
  const uint32_t r = __brev(v);  // reverses base order and the two bits inside each base
  return ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
}

// any always-mismatch (non-ACGT) query base in [a, b)
__device__ __forceinline__ bool query_has_n(const uint32_t *qnx, int a, int b) {
  for (int p = a; p < b; ++p)
    if ((qnx[p >> 4] >> (2 * (p & 15))) & 1u) return true;
  return false;
}

// bases [a, b) of seed piece pi: a = pi * R / np without an integer division (np = seed mismatches + 1 <= 4; the
// multiply-shift for np == 3 is exact while pi * R < 2^17, and R <= MIRGE_MAX_READ_LEN = 512)
__device__ __forceinline__ int piece_bound(int pi, int R, int np) {
  const uint32_t x = (uint32_t)(pi * R);
  return np == 1 ? (int)x : np == 2 ? (int)(x >> 1) : np == 3 ? (int)((x * 43691u) >> 17) : (int)(x >> 2);
}

// A packed key as the rounds see it
struct KeyView {
  const uint32_t *pay, *exc;
  int len, nexc;
};

__device__ __forceinline__ KeyView key_view(const uint32_t *key) {
  KeyView k;
  const uint32_t hdr = key[0];
  k.len = (int)key_len(hdr);
  k.nexc = (int)key_nexc(hdr);
  k.pay = key + 1;
  k.exc = key + 1 + ((k.len + 15) >> 4);
  return k;
}

// query window [qs, qe) of the key in one round (manifoldAlign.py:118-126: round 3 searches the sequence without
// its trailing T{3,} run; -5/-3 of round 8); false = the round does not search this key.  tlen caches the length
// without the trailing run of upper-case T (-1 = not computed yet).
__device__ __forceinline__ bool round_window(const KeyView &k, const mirge_round_policy &pol, int &tlen, int &qs, int &qe) {
  qs = 0;
  qe = k.len;
  if (pol.strip_polyT) {
    if (tlen < 0) {
      int tpos = k.len;
      while (tpos > 0) {
        const int j = tpos - 1;
        if (((k.pay[j >> 4] >> (2 * (j & 15))) & 3u) != 3u) break;
        bool is_exc = false;  // a lower-case 't' (or any non-"ACGT" byte) is stored as an exception
        for (int x = 0; x < k.nexc; ++x) is_exc |= (int)(k.exc[x] >> 8) == j;
        if (is_exc) break;
        --tpos;
      }
      tlen = tpos;
    }
    if (k.len - tlen < 3) return false;
    qe = tlen;
  }
  qs += pol.trim5;
  qe -= pol.trim3;
  return qe > qs;
}

// 2-bit query words and always-mismatch mask of key[qs:qe)
__device__ __forceinline__ void build_query(const KeyView &k, int qs, int qe, uint32_t *qw, uint32_t *qnx) {
  const int L = qe - qs, nw = (L + 15) >> 4, npay = (k.len + 15) >> 4;
  for (int w = 0; w < nw; ++w) {
    const int p = qs + 16 * w, wi = p >> 4, sh = 2 * (p & 15);
    const uint32_t lo = k.pay[wi], hi = (wi + 1 < npay) ? k.pay[wi + 1] : 0u;
    uint32_t v = sh ? __funnelshift_r(lo, hi, sh) : lo;
    const int rem = L - 16 * w;
    if (rem < 16) v &= (1u << (2 * rem)) - 1u;
    qw[w] = v;
    qnx[w] = 0;
  }
  for (int x = 0; x < k.nexc; ++x) {
    const int pos = (int)(k.exc[x] >> 8) - qs;
    if (pos < 0 || pos >= L) continue;
    const uint32_t code = base_code_upper(k.exc[x] & 0xFFu);
    const int w = pos >> 4, sh = 2 * (pos & 15);
    if (code < 4u) qw[w] = (qw[w] & ~(3u << sh)) | (code << sh);
    else qnx[w] |= 1u << sh;
  }
}
