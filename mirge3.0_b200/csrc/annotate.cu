// annotate.cu -- the ordered bowtie rounds of bwtAlign (mirge/libs/manifoldAlign.py:68-146) as an
// end-to-end, ungapped, forward-strand (--norc) mismatch search over 2-bit packed libraries.
//
// bowtie 1.x semantics restated (3P binary; oracle/pyoracle.py::hits): -v N = Hamming distance <= N
// over the whole read; -n N (-l 28 -e 70, FASTA input => all qualities 'I' => Maq-rounded 30) =
// <= N mismatches in the first min(28, len) bases and <= 2 in total; a read character outside ACGT
// always mismatches; a reference position outside ACGT may not be overlapped.  Because bowtie's
// choice among equally valid alignments is an FM-index artefact, the reported hit is the canonical
// minimum of (n_mismatch, reference index, offset) over the full valid hit set.
//
// Search = pigeonhole seeds: the seed region is cut into seed_mm + 1 pieces, at least one of which
// must match exactly; each piece is looked up in a sorted 16-mer index of the library (bucket table
// + binary search, prefix ranges for pieces shorter than 16) and every candidate is verified with
// XOR/popcount on the packed text.  One thread per unique sequence; libraries and indexes are
// L2-resident for all but the mRNA library.
#include "common.cuh"

#define ANN_THREADS 128
#define QW_MAX ((MIRGE_MAX_READ_LEN + 15) / 16)
#define MIN_SEED 4

__device__ __forceinline__ uint32_t lib_base(const uint32_t *packed, uint32_t p) { return (packed[p >> 4] >> (2 * (p & 15))) & 3u; }

// 16 bases starting at base position p as a 2-bit word (base p in bits 0..1)
__device__ __forceinline__ uint32_t lib_word16(const uint32_t *packed, uint64_t p, uint64_t n_words) {
  const uint64_t w = p >> 4;
  const uint32_t sh = 2 * (uint32_t)(p & 15);
  const uint32_t lo = packed[w];
  const uint32_t hi = (w + 1 < n_words) ? packed[w + 1] : 0u;
  return sh ? __funnelshift_r(lo, hi, sh) : lo;
}

__device__ __forceinline__ uint32_t find_ref(const uint32_t *ref_off, uint32_t n_refs, uint32_t pos) {
  uint32_t lo = 0, hi = n_refs;  // largest r with ref_off[r] <= pos
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (ref_off[mid] <= pos) lo = mid; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------ index construction -------

__global__ void __launch_bounds__(256)
lib_kmers_kernel(mirge_library lib, uint32_t *__restrict__ kmer, uint8_t *__restrict__ valid) {
  const uint32_t p = blockIdx.x * 256u + threadIdx.x;
  if (p >= lib.n_bases) return;
  const uint32_t r = find_ref(lib.d_ref_off, lib.n_refs, p);
  const uint32_t end = lib.d_ref_off[r + 1];
  uint32_t k = 0, v = 0;
  for (uint32_t i = 0; i < 16; ++i) {
    const uint32_t q = p + i;
    if (q >= end) break;
    if ((lib.d_nmask[q >> 5] >> (q & 31)) & 1u) break;
    k |= lib_base(lib.d_packed, q) << (2 * (15 - i));
    ++v;
  }
  kmer[p] = k;
  valid[p] = (uint8_t)v;
}

extern "C" int mirge_lib_kmers(mirge_ctx *ctx, const mirge_library *lib, uint32_t *d_kmer, uint8_t *d_valid, void *stream_) {
  if (!ctx) return MIRGE_ERR_ARG;
  if (!lib || !lib->d_packed || !lib->d_nmask || !lib->d_ref_off || !d_kmer || !d_valid) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "lib_kmers: null buffer");
  if (lib->n_bases == 0 || lib->n_refs == 0) return MIRGE_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  MIRGE_CUDA(ctx, cudaSetDevice(ctx->device));
  lib_kmers_kernel<<<(lib->n_bases + 255) / 256, 256, 0, stream>>>(*lib, d_kmer, d_valid);
  MIRGE_LAUNCH_CHECK(ctx, "lib_kmers_kernel");
  return MIRGE_OK;
}

// ------------------------------------------------------------------ search -------------------

struct Query {
  uint32_t w[QW_MAX];   // 2-bit codes, 16 per word
  uint32_t nx[QW_MAX];  // bit 2q set <=> base 16w+q of the query is not ACGT (always mismatches)
  int len;
};

// Hamming distance of the query against library text at astart under the round policy, text only
// (cheap: XOR + popcount on packed words); returns the mismatch count or -1 when the policy is violated.
__device__ __forceinline__ int verify_text(const mirge_library &lib, const Query &q, const mirge_round_policy &pol, int R,
                                           uint64_t astart) {
  const int L = q.len;
  const uint64_t n_words = ((uint64_t)lib.n_bases + 15) >> 4;
  int mm = 0, smm = 0;
  const int nw = (L + 15) >> 4;
  for (int w = 0; w < nw; ++w) {
    const uint32_t refw = lib_word16(lib.d_packed, astart + 16 * (uint64_t)w, n_words);
    uint32_t x = q.w[w] ^ refw;
    x = ((x | (x >> 1)) & 0x55555555u) | q.nx[w];
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
__device__ __forceinline__ uint64_t verify(const mirge_library &lib, const Query &q, const mirge_round_policy &pol, int R,
                                           uint64_t astart, uint32_t pos_in_ref) {
  const int mm = verify_text(lib, q, pol, R, astart);
  if (mm < 0) return MIRGE_NO_HIT;
  const uint32_t r = find_ref(lib.d_ref_off, lib.n_refs, pos_in_ref);
  const uint32_t rlo = lib.d_ref_off[r], rhi = lib.d_ref_off[r + 1];
  if (astart < rlo || astart + (uint64_t)q.len > rhi) return MIRGE_NO_HIT;
  if (ref_has_n(lib, astart, astart + q.len)) return MIRGE_NO_HIT;
  return ((uint64_t)mm << 56) | ((uint64_t)r << 28) | (uint64_t)(astart - rlo);
}

__global__ void __launch_bounds__(ANN_THREADS)
annotate_kernel(mirge_library lib, mirge_round_policy pol, mirge_table t, uint64_t n_keys, uint8_t *__restrict__ annot_round,
                uint64_t *__restrict__ hit) {
  const uint64_t id = (uint64_t)blockIdx.x * ANN_THREADS + threadIdx.x;
  if (id >= n_keys) return;
  const uint32_t *key = t.d_arena + t.d_key_ref[id];
  const uint32_t hdr = key[0];
  const int len = (int)key_len(hdr), nexc = (int)key_nexc(hdr);
  if (pol.select == MIRGE_SELECT_LEN_LT26) { if (!(len < 26)) return; }
  else if (pol.select == MIRGE_SELECT_LEN_GT25) { if (!(len > 25)) return; }
  else if (annot_round[id] != 0xFF) return;
  const uint32_t *pay = key + 1;
  const uint32_t *exc = key + 1 + ((len + 15) >> 4);
  // query window [qs, qe) of the key (manifoldAlign.py:118-126; -5/-3 of round 8)
  int qs = 0, qe = len;
  if (pol.strip_polyT) {
    int tpos = len;
    while (tpos > 0) {
      const int j = tpos - 1;
      if (((pay[j >> 4] >> (2 * (j & 15))) & 3u) != 3u) break;
      bool is_exc = false;  // a lower-case 't' (or any non-"ACGT" byte) is stored as an exception
      for (int x = 0; x < nexc; ++x) is_exc |= (int)(exc[x] >> 8) == j;
      if (is_exc) break;
      --tpos;
    }
    if (len - tpos < 3) return;
    qe = tpos;
  }
  qs += pol.trim5;
  qe -= pol.trim3;
  if (qe <= qs) return;
  Query q;
  q.len = qe - qs;
  const int L = q.len, nw = (L + 15) >> 4;
  {
    const int npay = (len + 15) >> 4;
    for (int w = 0; w < nw; ++w) {
      const int p = qs + 16 * w, wi = p >> 4, sh = 2 * (p & 15);
      const uint32_t lo = pay[wi], hi = (wi + 1 < npay) ? pay[wi + 1] : 0u;
      uint32_t v = sh ? __funnelshift_r(lo, hi, sh) : lo;
      const int rem = L - 16 * w;
      if (rem < 16) v &= (1u << (2 * rem)) - 1u;
      q.w[w] = v;
      q.nx[w] = 0;
    }
    for (int x = 0; x < nexc; ++x) {
      const int pos = (int)(exc[x] >> 8) - qs;
      if (pos < 0 || pos >= L) continue;
      const uint32_t code = base_code_upper(exc[x] & 0xFFu);
      const int w = pos >> 4, sh = 2 * (pos & 15);
      if (code < 4u) q.w[w] = (q.w[w] & ~(3u << sh)) | (code << sh);
      else q.nx[w] |= 1u << sh;
    }
  }
  const int R = pol.seed_len == 0 ? L : min(pol.seed_len, L);
  const int np = pol.seed_mm + 1;
  uint64_t best = MIRGE_NO_HIT;
  if (R / np < MIN_SEED) {
    // degenerate (very short query): exhaustive scan keeps the result exact
    for (uint32_t r = 0; r < lib.n_refs; ++r) {
      const uint32_t lo = lib.d_ref_off[r], hi = lib.d_ref_off[r + 1];
      for (uint64_t a = lo; a + L <= hi; ++a) {
        const uint64_t h = verify(lib, q, pol, R, a, (uint32_t)a);
        if (h < best) best = h;
      }
    }
  } else {
    for (int pi = 0; pi < np; ++pi) {
      const int a = (int)((long long)pi * R / np), b = (int)((long long)(pi + 1) * R / np);
      const int s = min(16, b - a);
      // piece k-mer, first base most significant; a piece containing a non-ACGT read base cannot be exact
      uint32_t kmer = 0;
      bool has_n = false;
      for (int i = 0; i < b - a; ++i) {
        const int p = a + i;
        has_n |= (q.nx[p >> 4] >> (2 * (p & 15))) & 1u;
        if (i < s) kmer |= ((q.w[p >> 4] >> (2 * (p & 15))) & 3u) << (2 * (15 - i));
      }
      if (has_n) continue;
      const uint32_t span = (s == 16) ? 0u : ((1u << (2 * (16 - s))) - 1u);
      const uint32_t k_lo = kmer, k_hi = kmer | span;
      const uint32_t bsh = 32 - lib.bucket_bits;
      uint32_t lo = lib.d_idx_bucket[k_lo >> bsh], hi = lib.d_idx_bucket[(k_hi >> bsh) + 1];
      // lower_bound(k_lo) and upper_bound(k_hi) inside [lo, hi)
      {
        uint32_t l = lo, h = hi;
        while (l < h) { const uint32_t m = (l + h) >> 1; if (lib.d_idx_kmer[m] < k_lo) l = m + 1; else h = m; }
        lo = l;
        h = hi;
        while (l < h) { const uint32_t m = (l + h) >> 1; if (lib.d_idx_kmer[m] <= k_hi) l = m + 1; else h = m; }
        hi = l;
      }
      for (uint32_t e = lo; e < hi; ++e) {
        const uint32_t pos = lib.d_idx_pos[e];
        if (pos < (uint32_t)a) continue;
        const uint32_t astart = pos - (uint32_t)a;
        const uint64_t h = verify(lib, q, pol, R, astart, pos);
        if (h < best) best = h;
      }
    }
  }
  if (best != MIRGE_NO_HIT) {
    annot_round[id] = (uint8_t)pol.round;
    hit[id] = best;
  }
}

extern "C" int mirge_annotate_round(mirge_ctx *ctx, const mirge_library *lib, const mirge_round_policy *policy, const mirge_table *t,
                                    uint64_t n_keys, uint8_t *d_annot_round, uint64_t *d_hit, void *stream_) {
  if (!ctx) return MIRGE_ERR_ARG;
  if (!lib || !policy || !t || !d_annot_round || !d_hit) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: null argument");
  if (n_keys == 0 || lib->n_refs == 0 || lib->n_bases == 0) return MIRGE_OK;
  if (!lib->d_packed || !lib->d_nmask || !lib->d_ref_off || !lib->d_idx_kmer || !lib->d_idx_pos || !lib->d_idx_bucket)
    MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: library has no index");
  if (lib->bucket_bits < 1 || lib->bucket_bits > 28) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: bucket_bits out of range");
  if (policy->seed_mm < 0 || policy->seed_mm > 3 || policy->total_mm < policy->seed_mm || policy->trim5 < 0 || policy->trim3 < 0)
    MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: unsupported policy");
  if (lib->n_refs >= (1u << 28)) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: too many references");
  cudaStream_t stream = (cudaStream_t)stream_;
  MIRGE_CUDA(ctx, cudaSetDevice(ctx->device));
  annotate_kernel<<<(unsigned)((n_keys + ANN_THREADS - 1) / ANN_THREADS), ANN_THREADS, 0, stream>>>(*lib, *policy, *t, n_keys,
                                                                                                   d_annot_round, d_hit);
  MIRGE_LAUNCH_CHECK(ctx, "annotate_kernel");
  return MIRGE_OK;
}
