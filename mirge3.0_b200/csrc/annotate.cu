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
// XOR/popcount on the packed text.  One thread prepares one unique sequence; long candidate lists are
// verified warp-cooperatively.  Libraries and indexes are L2-resident for all but the mRNA library.
#include "common.cuh"

// lib_base .. build_query: the pure per-sequence helpers (also compiled for the host by the tests)
#include "annotate_verify.cuh"

// ------------------------------------------------------------------ index construction -------

__global__ void __launch_bounds__(256)
lib_kmers_kernel(mirge_library lib, uint32_t *__restrict__ kmer, uint8_t *__restrict__ valid) {
  const uint32_t p = blockIdx.x * 256u + threadIdx.x;
  if (p >= lib.n_bases) return;
  const uint32_t r = find_ref(lib, p);
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
  if (!lib || !lib->d_packed || !lib->d_nmask || !lib->d_ref_off || !lib->d_ref_block || !d_kmer || !d_valid)
    MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "lib_kmers: null buffer");
  if (lib->n_bases == 0 || lib->n_refs == 0) return MIRGE_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  MIRGE_CUDA(ctx, cudaSetDevice(ctx->device));
  lib_kmers_kernel<<<(lib->n_bases + 255) / 256, 256, 0, stream>>>(*lib, d_kmer, d_valid);
  MIRGE_LAUNCH_CHECK(ctx, "lib_kmers_kernel");
  return MIRGE_OK;
}

// Presence bitmap over the first `bases` bases of every indexed position: bit (k-mer >> (32 - 2 * bases)) is set
// when some position has at least `bases` usable bases starting with that prefix.  A seed piece of >= `bases`
// bases whose prefix bit is clear cannot occur in the library (no false negatives), which ends most look-ups
// of sequences the library does not contain with one load from a table small enough to stay in L2.
__global__ void __launch_bounds__(256)
lib_filter_kernel(const uint32_t *__restrict__ kmer, const uint8_t *__restrict__ valid, uint32_t n, uint32_t bases,
                  uint32_t *__restrict__ filter) {
  const uint32_t p = blockIdx.x * 256u + threadIdx.x;
  if (p >= n || valid[p] < bases) return;
  const uint32_t fi = bases >= 16 ? kmer[p] : (kmer[p] >> (32 - 2 * bases));
  atomicOr(filter + (fi >> 5), 1u << (fi & 31));
}

extern "C" int mirge_lib_filter(mirge_ctx *ctx, const uint32_t *d_kmer, const uint8_t *d_valid, uint32_t n_bases, uint32_t prefix_bases,
                                uint32_t *d_filter, void *stream_) {
  if (!ctx) return MIRGE_ERR_ARG;
  if (!d_kmer || !d_valid || !d_filter) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "lib_filter: null buffer");
  if (prefix_bases < 4 || prefix_bases > 16) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "lib_filter: prefix length must be 4..16");
  if (n_bases == 0) return MIRGE_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  MIRGE_CUDA(ctx, cudaSetDevice(ctx->device));
  MIRGE_CUDA(ctx, cudaMemsetAsync(d_filter, 0, (size_t)1 << (2 * prefix_bases - 3), stream));
  lib_filter_kernel<<<(n_bases + 255) / 256, 256, 0, stream>>>(d_kmer, d_valid, n_bases, prefix_bases, d_filter);
  MIRGE_LAUNCH_CHECK(ctx, "lib_filter_kernel");
  return MIRGE_OK;
}

// Presence bitmap of complete 16-mers through a 32-bit mix (no false negatives; see mirge_lib_filter16 in the header)
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ bool filter16_pass(const mirge_library &lib, uint32_t kmer) {
  const uint32_t fi = mix32(kmer) >> (32 - lib.filter16_bits);
  return (lib.d_filter16[fi >> 5] >> (fi & 31)) & 1u;
}

__global__ void __launch_bounds__(256)
lib_filter16_kernel(const uint32_t *__restrict__ kmer, const uint8_t *__restrict__ valid, uint32_t n, uint32_t bits,
                    uint32_t *__restrict__ filter) {
  const uint32_t p = blockIdx.x * 256u + threadIdx.x;
  if (p >= n || valid[p] < 16) return;
  const uint32_t fi = mix32(kmer[p]) >> (32 - bits);
  atomicOr(filter + (fi >> 5), 1u << (fi & 31));
}

extern "C" int mirge_lib_filter16(mirge_ctx *ctx, const uint32_t *d_kmer, const uint8_t *d_valid, uint32_t n_bases, uint32_t bits,
                                  uint32_t *d_filter16, void *stream_) {
  if (!ctx) return MIRGE_ERR_ARG;
  if (!d_kmer || !d_valid || !d_filter16) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "lib_filter16: null buffer");
  if (bits < 16 || bits > 30) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "lib_filter16: 16..30 bits");
  cudaStream_t stream = (cudaStream_t)stream_;
  MIRGE_CUDA(ctx, cudaSetDevice(ctx->device));
  MIRGE_CUDA(ctx, cudaMemsetAsync(d_filter16, 0, (size_t)1 << (bits - 3), stream));
  if (n_bases == 0) return MIRGE_OK;
  lib_filter16_kernel<<<(n_bases + 255) / 256, 256, 0, stream>>>(d_kmer, d_valid, n_bases, bits, d_filter16);
  MIRGE_LAUNCH_CHECK(ctx, "lib_filter16_kernel");
  return MIRGE_OK;
}

// ------------------------------------------------------------------ search -------------------

#define MAX_PIECES 4
#define MAX_ROUNDS 10

struct RoundSet {
  int n;
  mirge_library lib[MAX_ROUNDS];
  mirge_round_policy pol[MAX_ROUNDS];
};

// Per-warp scratch of the candidate redistribution: every lane publishes its query (up to COOP_QW words) and the
// index ranges of its seed pieces; the candidates of all 32 lanes then form one list that the lanes share out
// evenly, so a sequence with many candidates does not hold up 31 idle lanes and candidate-free lanes help.
#define COOP_QW 6  // queries of up to 96 bases take part; longer ones verify their own candidates
struct WarpScratch {
  uint32_t qw[32][COOP_QW], qnx[32][COOP_QW];
  uint32_t lo[32][MAX_PIECES], hi[32][MAX_PIECES], off[32][MAX_PIECES];
  uint32_t len[32];
  uint32_t pref[33];  // exclusive prefix of the candidate counts
  unsigned long long best[32];
};

// One bowtie round for the query of this lane (active lanes only); all 32 lanes must call it together.
// republish: this lane's query words changed since the last round (window change) and must be copied again.
__device__ __forceinline__ uint64_t search_round(const mirge_library &lib, const mirge_round_policy &pol, bool active,
                                                 const uint32_t *qw, const uint32_t *qnx, int L, bool has_exc, WarpScratch &ws,
                                                 bool &republish, int lane) {
  uint32_t p_lo[MAX_PIECES], p_hi[MAX_PIECES], p_off[MAX_PIECES];
#pragma unroll
  for (int i = 0; i < MAX_PIECES; ++i) p_lo[i] = p_hi[i] = p_off[i] = 0;
  // an end-to-end alignment lies inside one reference: a query longer than the longest reference has none
  if (lib.max_ref_len && (uint32_t)L > lib.max_ref_len) active = false;
  const int R = pol.seed_len == 0 ? L : min(pol.seed_len, L);
  const int np = pol.seed_mm + 1;
  const bool degenerate = active && (R < MIN_SEED * np);
  const int nw = (L + 15) >> 4;
  if (active && !degenerate) {
    for (int pi = 0; pi < np; ++pi) {
      const int a = piece_bound(pi, R, np), b = piece_bound(pi + 1, R, np);
      const int s = min(16, b - a);
      // a piece with a non-ACGT read base cannot be exact
      if (has_exc && query_has_n(qnx, a, b)) continue;
      const uint32_t span = (s == 16) ? 0u : ((1u << (2 * (16 - s))) - 1u);
      const uint32_t k_lo = query_kmer16(qw, a, nw) & ~span, k_hi = k_lo | span;
      if (s == 16 && lib.filter16_bits && lib.filter_bases < 16) {  // complete 16-mer absent from the library
        if (!filter16_pass(lib, k_lo)) continue;
      } else if (lib.filter_bases && (uint32_t)s >= lib.filter_bases) {  // prefix absent from the library: no candidates
        const uint32_t fi = lib.filter_bases >= 16 ? k_lo : (k_lo >> (32 - 2 * lib.filter_bases));
        if (!((lib.d_filter[fi >> 5] >> (fi & 31)) & 1u)) continue;
      }
      const uint32_t bsh = 32 - lib.bucket_bits;
      uint32_t l = lib.d_idx_bucket[k_lo >> bsh], h = lib.d_idx_bucket[(k_hi >> bsh) + 1];
      const uint32_t hi0 = h;  // lower_bound(k_lo), then upper_bound(k_hi), inside the bucket range
      while (l < h) { const uint32_t mid = (l + h) >> 1; if (lib.d_idx_kmer[mid] < k_lo) l = mid + 1; else h = mid; }
      p_lo[pi] = l;
      h = hi0;
      while (l < h) { const uint32_t mid = (l + h) >> 1; if (lib.d_idx_kmer[mid] <= k_hi) l = mid + 1; else h = mid; }
      p_hi[pi] = l;
      p_off[pi] = (uint32_t)a;
    }
  }
  uint64_t best = MIRGE_NO_HIT;
  uint32_t total = 0;
#pragma unroll
  for (int i = 0; i < MAX_PIECES; ++i) total += p_hi[i] - p_lo[i];
  const bool shared = active && !degenerate && total > 0 && nw <= COOP_QW;
  if (degenerate) {
    // very short query: exhaustive scan keeps the result exact
    for (uint32_t r = 0; r < lib.n_refs; ++r) {
      const uint32_t lo = lib.d_ref_off[r], hi = lib.d_ref_off[r + 1];
      for (uint64_t a = lo; a + L <= hi; ++a) {
        const uint64_t h = verify(lib, qw, qnx, L, pol, R, a, (uint32_t)a);
        if (h < best) best = h;
      }
    }
  } else if (active && total > 0 && !shared) {  // long query: its own candidates, serially
#pragma unroll
    for (int pi = 0; pi < MAX_PIECES; ++pi)
      for (uint32_t e = p_lo[pi]; e < p_hi[pi]; ++e) {
        const uint32_t pos = lib.d_idx_pos[e];
        if (pos < p_off[pi]) continue;
        const uint64_t h = verify(lib, qw, qnx, L, pol, R, pos - p_off[pi], pos);
        if (h < best) best = h;
      }
  }
  // ---- the candidates of all lanes as one list
  const unsigned sm = __ballot_sync(0xffffffffu, shared);
  if (sm == 0) return best;
  uint32_t incl = shared ? total : 0u;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t x = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += x;
  }
  const uint32_t T = __shfl_sync(0xffffffffu, incl, 31);
  ws.pref[lane + 1] = incl;
  if (lane == 0) ws.pref[0] = 0;
  if (shared) {
    if (republish) {
      for (int w = 0; w < nw; ++w) { ws.qw[lane][w] = qw[w]; ws.qnx[lane][w] = qnx[w]; }
      republish = false;
    }
#pragma unroll
    for (int pi = 0; pi < MAX_PIECES; ++pi) { ws.lo[lane][pi] = p_lo[pi]; ws.hi[lane][pi] = p_hi[pi]; ws.off[lane][pi] = p_off[pi]; }
    ws.len[lane] = (uint32_t)L;
    ws.best[lane] = MIRGE_NO_HIT;
  }
  __syncwarp();
  for (uint32_t w = lane; w < T; w += 32) {
    // owner = last lane whose prefix is <= w (binary search over the 32 prefixes)
    int o = 0;
#pragma unroll
    for (int st = 16; st > 0; st >>= 1)
      if (ws.pref[o + st] <= w) o += st;
    uint32_t j = w - ws.pref[o];
    int pi = 0;
#pragma unroll
    for (int q = 0; q < MAX_PIECES - 1; ++q) {
      const uint32_t c = ws.hi[o][pi] - ws.lo[o][pi];
      if (j >= c && pi == q) { j -= c; ++pi; }
    }
    const uint32_t pos = lib.d_idx_pos[ws.lo[o][pi] + j], off = ws.off[o][pi];
    if (pos < off) continue;
    const int oL = (int)ws.len[o];
    const int oR = pol.seed_len == 0 ? oL : min(pol.seed_len, oL);
    const uint64_t h = verify(lib, ws.qw[o], ws.qnx[o], oL, pol, oR, pos - off, pos);
    if (h != MIRGE_NO_HIT) atomicMin(&ws.best[o], (unsigned long long)h);
  }
  __syncwarp();
  if (shared) {
    const uint64_t b = ws.best[lane];
    if (b < best) best = b;
  }
  return best;
}

// All rounds of bwtAlign for one unique sequence per thread, in order; a sequence leaves at the first round
// that hits it (manifoldAlign.py:120,129).  The key is read once and the query words are rebuilt only when
// a round's window differs (poly-T stripping of round 3, -5/-3 trimming of round 8).
// 10 CTAs per SM (48 registers): 17.5 ms per 39 M sequences; 85 registers / 6 CTAs measured 28.3 ms, 40 registers / 12
// CTAs 17.8 ms -- the kernel lives on warps in flight, up to the point where spills eat the gain.
__global__ void __launch_bounds__(ANN_THREADS, 10)
annotate_kernel(const __grid_constant__ RoundSet rs, mirge_table t, uint64_t n_keys, uint8_t *__restrict__ annot_round,
                uint64_t *__restrict__ hit) {
  __shared__ WarpScratch s_ws[ANN_THREADS / 32];
  const uint64_t slot = (uint64_t)blockIdx.x * ANN_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool in_range = slot < n_keys;
  const uint64_t id = slot;
  uint32_t qw[QW_MAX], qnx[QW_MAX];
  KeyView kv;
  kv.pay = kv.exc = nullptr; kv.len = kv.nexc = 0;
  int cur_qs = -1, cur_qe = -1, tlen = -1;
  uint32_t state = 0xFF;  // round that annotated this sequence, 0xFF = none yet
  if (in_range) {
    kv = key_view(t.d_arena + t.d_key_ref[id]);
    state = annot_round[id];
  }
  const int len = kv.len, nexc = kv.nexc;
  uint64_t my_hit = MIRGE_NO_HIT;
  int my_round = -1;
  bool republish = true;  // the warp scratch does not hold this lane's current query words yet
  for (int ri = 0; ri < rs.n; ++ri) {
    const mirge_round_policy &pol = rs.pol[ri];
    bool active = in_range;
    if (active) {
      if (pol.select == MIRGE_SELECT_LEN_LT26) active = len < 26;
      else if (pol.select == MIRGE_SELECT_LEN_GT25) active = len > 25;
      else active = state == 0xFF;
    }
    int L = 0;
    if (active) {
      int qs, qe;
      active = round_window(kv, pol, tlen, qs, qe);
      if (active) {
        L = qe - qs;
        if (qs != cur_qs || qe != cur_qe) {  // the query words are rebuilt only when the window changes
          build_query(kv, qs, qe, qw, qnx);
          cur_qs = qs;
          cur_qe = qe;
          republish = true;
        }
      }
    }
    const uint64_t best = search_round(rs.lib[ri], pol, active, qw, qnx, L, nexc > 0, s_ws[warp], republish, lane);
    if (active && best != MIRGE_NO_HIT) {
      state = (uint32_t)pol.round;
      my_round = pol.round;
      my_hit = best;
    }
  }
  if (my_round >= 0) {
    annot_round[id] = (uint8_t)my_round;
    hit[id] = my_hit;
  }
}

// ---- every hit of the best stratum (rounds run with -a --best --strata; SAM emission) ------------------------
// For the listed sequences (annotated by this round, best stratum = MIRGE_HIT_MM(hit[id])): all valid alignments
// with that many mismatches.  fill == 0: counts[i] = number of hits found (an alignment reachable through
// several seed pieces is found once per piece; the host removes duplicates); fill != 0: the hits are written to
// out[offs[i] ...].  Thread per sequence: these lists are short (tRNA libraries).
__global__ void __launch_bounds__(ANN_THREADS)
allhits_kernel(mirge_library lib, mirge_round_policy pol, mirge_table t, const uint32_t *__restrict__ ids, uint64_t n,
               const uint64_t *__restrict__ hit, int fill, uint32_t *__restrict__ counts, const uint64_t *__restrict__ offs,
               uint64_t *__restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * ANN_THREADS + threadIdx.x;
  if (i >= n) return;
  const uint32_t id = ids[i];
  const uint64_t want_mm = hit[id] >> 56;
  const KeyView kv = key_view(t.d_arena + t.d_key_ref[id]);
  uint32_t qw[QW_MAX], qnx[QW_MAX];
  int tlen = -1, qs, qe;
  uint32_t k = 0;
  const uint64_t o0 = fill ? offs[i] : 0;
  if (round_window(kv, pol, tlen, qs, qe)) {
    build_query(kv, qs, qe, qw, qnx);
    const int L = qe - qs, nw = (L + 15) >> 4;
    const int R = pol.seed_len == 0 ? L : min(pol.seed_len, L);
    const int np = pol.seed_mm + 1;
#define ALLHITS_TAKE(h_)                                    \
    if ((h_) != MIRGE_NO_HIT && ((h_) >> 56) == want_mm) {  \
      if (fill) out[o0 + k] = (h_);                         \
      ++k;                                                  \
    }
    if (R < MIN_SEED * np) {
      for (uint32_t r = 0; r < lib.n_refs; ++r) {
        const uint32_t lo = lib.d_ref_off[r], hi = lib.d_ref_off[r + 1];
        for (uint64_t a = lo; a + L <= hi; ++a) {
          const uint64_t h = verify(lib, qw, qnx, L, pol, R, a, (uint32_t)a);
          ALLHITS_TAKE(h)
        }
      }
    } else {
      for (int pi = 0; pi < np; ++pi) {
        const int a = piece_bound(pi, R, np), b = piece_bound(pi + 1, R, np);
        const int s = min(16, b - a);
        if (kv.nexc > 0 && query_has_n(qnx, a, b)) continue;
        const uint32_t span = (s == 16) ? 0u : ((1u << (2 * (16 - s))) - 1u);
        const uint32_t k_lo = query_kmer16(qw, a, nw) & ~span, k_hi = k_lo | span;
        const uint32_t bsh = 32 - lib.bucket_bits;
        uint32_t l = lib.d_idx_bucket[k_lo >> bsh], h = lib.d_idx_bucket[(k_hi >> bsh) + 1];
        const uint32_t hi0 = h;
        while (l < h) { const uint32_t mid = (l + h) >> 1; if (lib.d_idx_kmer[mid] < k_lo) l = mid + 1; else h = mid; }
        const uint32_t e0 = l;
        h = hi0;
        while (l < h) { const uint32_t mid = (l + h) >> 1; if (lib.d_idx_kmer[mid] <= k_hi) l = mid + 1; else h = mid; }
        for (uint32_t e = e0; e < l; ++e) {
          const uint32_t pos = lib.d_idx_pos[e];
          if (pos < (uint32_t)a) continue;
          const uint64_t hh = verify(lib, qw, qnx, L, pol, R, pos - (uint32_t)a, pos);
          ALLHITS_TAKE(hh)
        }
      }
    }
#undef ALLHITS_TAKE
  }
  if (!fill) counts[i] = k;
}

// ---- two-pass rounds: filter masks for all rounds -> per-warp compaction and search ---------------------------
// In the fused kernel above a warp runs nine rounds at the pace of its busiest lane, and most lanes hold sequences no
// library contains (HEAD counting keeps the untrimmed reads): per round they evaluate their seed pieces' filters,
// find nothing and wait.  Pass 1 (annot_mask_kernel) does that part for ALL rounds of every sequence in one
// streaming kernel -- no search, no dependent index look-ups -- and leaves a round mask per sequence.  Pass 2
// (annot_search_kernel) hands a warp 512 consecutive sequences: round by round the warp compacts the ones that round
// still has to search (mask bit set, and not annotated by an earlier round) into a shared-memory list and searches
// them 32 at a time: every lane holds a sequence with at least one piece the library may contain, all lanes use
// the same library and policy, and the order of the rounds is the warp's own program order.  The masks are a
// superset (pass 1 ignores what earlier rounds will find; the search re-checks everything), so results cannot
// differ from the fused form.

// rounds (bit ri of the result) in which the key may have a candidate; state = annotation before this call
__device__ __forceinline__ uint32_t round_mask(const RoundSet &rs, const KeyView &kv, uint32_t state) {
  uint32_t qw[QW_MAX], qnx[QW_MAX];
  int cur_qs = -1, cur_qe = -1, tlen = -1;
  const int len = kv.len;
  const bool has_exc = kv.nexc > 0;
  uint32_t mask = 0;
  for (int ri = 0; ri < rs.n; ++ri) {
    const mirge_round_policy &pol = rs.pol[ri];
    const mirge_library &lib = rs.lib[ri];
    bool active;
    if (pol.select == MIRGE_SELECT_LEN_LT26) active = len < 26;
    else if (pol.select == MIRGE_SELECT_LEN_GT25) active = len > 25;
    else active = state == 0xFF;  // annotated before this call: stays annotated
    if (!active) continue;
    int qs, qe;
    if (!round_window(kv, pol, tlen, qs, qe)) continue;
    const int L = qe - qs;
    if (lib.max_ref_len && (uint32_t)L > lib.max_ref_len) continue;
    if (qs != cur_qs || qe != cur_qe) {
      build_query(kv, qs, qe, qw, qnx);
      cur_qs = qs;
      cur_qe = qe;
    }
    const int R = pol.seed_len == 0 ? L : min(pol.seed_len, L);
    const int np = pol.seed_mm + 1;
    const int nw = (L + 15) >> 4;
    bool pass = false;
    if (R < MIN_SEED * np) {
      pass = true;  // exhaustive scan in the search
    } else {
      for (int pi = 0; pi < np; ++pi) {
        const int a = piece_bound(pi, R, np), b = piece_bound(pi + 1, R, np);
        const int s = min(16, b - a);
        if (has_exc && query_has_n(qnx, a, b)) continue;
        if (s == 16 && lib.filter16_bits && lib.filter_bases < 16) {
          pass |= filter16_pass(lib, query_kmer16(qw, a, nw));
        } else if (lib.filter_bases && (uint32_t)s >= lib.filter_bases) {
          const uint32_t span = (s == 16) ? 0u : ((1u << (2 * (16 - s))) - 1u);
          const uint32_t k_lo = query_kmer16(qw, a, nw) & ~span;
          const uint32_t fi = lib.filter_bases >= 16 ? k_lo : (k_lo >> (32 - 2 * lib.filter_bases));
          pass |= ((lib.d_filter[fi >> 5] >> (fi & 31)) & 1u) != 0u;
        } else {
          pass = true;  // piece shorter than the filter prefix: look it up
        }
      }
    }
    if (pass) mask |= 1u << ri;
  }
  return mask;
}

// Lean form of round_mask for the common key (no exception words, <= LEAN_WORDS payload words): the payload sits in a
// shared-memory column of the thread (dynamic word index without local memory), a piece's 16 bases are one funnel
// shift of two of those words, and nothing is copied per window.  Bases of the 16-base window beyond the piece are
// masked by the prefix span exactly as in search_round, so the look-ups are the same ones.
#define LEAN_WORDS 8
#define MASK_THREADS 256
template <int S>
__device__ __forceinline__ uint32_t lean_word16(const uint32_t *col, int p) {  // 16 bases from base p, base p in bits 0..1
  const int wi = p >> 4, sh = 2 * (p & 15);
  return __funnelshift_r(col[wi * S], col[(wi + 1) * S], sh);
}
template <int S>
__device__ __forceinline__ uint32_t lean_kmer16(const uint32_t *col, int p) {  // the same, first base most significant
  const uint32_t v = lean_word16<S>(col, p);
  const uint32_t r = __brev(v);
  return ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
}

// query window of a round for a key without exception words whose payload is the column `col` (round_window)
template <int S>
__device__ __forceinline__ bool lean_window(const uint32_t *col, int len, const mirge_round_policy &pol, int &tlen, int &qs, int &qe) {
  qs = 0;
  qe = len;
  if (pol.strip_polyT) {
    if (tlen < 0) {
      int tpos = len;
      while (tpos > 0 && ((col[((tpos - 1) >> 4) * S] >> (2 * ((tpos - 1) & 15))) & 3u) == 3u) --tpos;
      tlen = tpos;
    }
    if (len - tlen < 3) return false;
    qe = tlen;
  }
  qs += pol.trim5;
  qe -= pol.trim3;
  return qe > qs;
}

__device__ __forceinline__ uint32_t round_mask_lean(const RoundSet &rs, const uint32_t *col, int len, uint32_t state) {
  int tlen = -1;
  uint32_t mask = 0;
  for (int ri = 0; ri < rs.n; ++ri) {
    const mirge_round_policy &pol = rs.pol[ri];
    const mirge_library &lib = rs.lib[ri];
    bool active;
    if (pol.select == MIRGE_SELECT_LEN_LT26) active = len < 26;
    else if (pol.select == MIRGE_SELECT_LEN_GT25) active = len > 25;
    else active = state == 0xFF;
    if (!active) continue;
    int qs, qe;
    if (!lean_window<MASK_THREADS>(col, len, pol, tlen, qs, qe)) continue;
    const int L = qe - qs;
    if (lib.max_ref_len && (uint32_t)L > lib.max_ref_len) continue;
    const int R = pol.seed_len == 0 ? L : min(pol.seed_len, L);
    const int np = pol.seed_mm + 1;
    bool pass = false;
    if (R < MIN_SEED * np) {
      pass = true;
    } else {
      int a = 0;
      for (int pi = 0; pi < np; ++pi) {
        const int b = piece_bound(pi + 1, R, np);
        const int s = min(16, b - a);
        const uint32_t k16 = lean_kmer16<MASK_THREADS>(col, qs + a);
        if (s == 16 && lib.filter16_bits && lib.filter_bases < 16) {
          pass |= filter16_pass(lib, k16);
        } else if (lib.filter_bases && (uint32_t)s >= lib.filter_bases) {
          const uint32_t fi = lib.filter_bases >= 16 ? k16 : (k16 >> (32 - 2 * lib.filter_bases));
          pass |= ((lib.d_filter[fi >> 5] >> (fi & 31)) & 1u) != 0u;
        } else {
          pass = true;
        }
        a = b;
      }
    }
    if (pass) mask |= 1u << ri;
  }
  return mask;
}

// pass 1: the round mask of every sequence (streaming: one thread per sequence, nothing but filter look-ups)
__global__ void __launch_bounds__(MASK_THREADS)
annot_mask_kernel(const __grid_constant__ RoundSet rs, mirge_table t, uint64_t n_keys, const uint8_t *__restrict__ annot_round,
                  uint16_t *__restrict__ masks) {
  __shared__ uint32_t s_pay[LEAN_WORDS + 2][MASK_THREADS];
  const uint64_t id = (uint64_t)blockIdx.x * MASK_THREADS + threadIdx.x;
  if (id >= n_keys) return;
  const uint32_t *key = t.d_arena + t.d_key_ref[id];
  const uint32_t state = annot_round[id];
  const KeyView kv = key_view(key);
  const int npay = (kv.len + 15) >> 4;
  uint32_t mask;
  if (kv.nexc == 0 && npay <= LEAN_WORDS) {
    uint32_t *col = &s_pay[0][threadIdx.x];
#pragma unroll
    for (int w = 0; w < LEAN_WORDS + 2; ++w) col[w * MASK_THREADS] = w < npay ? kv.pay[w] : 0u;
    mask = round_mask_lean(rs, col, kv.len, state);
  } else {
    mask = round_mask(rs, kv, state);
  }
  masks[id] = (uint16_t)mask;
}

// One round for one sequence from its shared-memory column: the same pieces, index ranges, candidates and checks as
// search_round + verify, without the query copies in local memory and without the warp-level candidate sharing --
// for the common sequence (no exception words, <= LEAN_WORDS payload words, a handful of candidates).  Returns
// false when the sequence needs the general path (degenerate seed, many candidates); `best` is then untouched.
#ifndef LEAN_MAX_CAND
#define LEAN_MAX_CAND 12
#endif
template <int S>
__device__ __forceinline__ bool search_lean(const mirge_library &lib, const mirge_round_policy &pol, const uint32_t *col, int qs,
                                            int L, uint64_t &best) {
  if (lib.max_ref_len && (uint32_t)L > lib.max_ref_len) return true;  // longer than every reference: no alignment
  const int R = pol.seed_len == 0 ? L : min(pol.seed_len, L);
  const int np = pol.seed_mm + 1;
  if (R < MIN_SEED * np) return false;
  uint32_t p_lo[MAX_PIECES], p_hi[MAX_PIECES];
  uint32_t total = 0;
  int a = 0;
#pragma unroll
  for (int pi = 0; pi < MAX_PIECES; ++pi) {
    p_lo[pi] = p_hi[pi] = 0;
    if (pi < np) {
      const int b = piece_bound(pi + 1, R, np);
      const int s = min(16, b - a);
      const uint32_t k16 = lean_kmer16<S>(col, qs + a);
      const uint32_t span = (s == 16) ? 0u : ((1u << (2 * (16 - s))) - 1u);
      const uint32_t k_lo = k16 & ~span, k_hi = k_lo | span;
      bool look = true;
      if (s == 16 && lib.filter16_bits && lib.filter_bases < 16) {
        look = filter16_pass(lib, k_lo);
      } else if (lib.filter_bases && (uint32_t)s >= lib.filter_bases) {
        const uint32_t fi = lib.filter_bases >= 16 ? k_lo : (k_lo >> (32 - 2 * lib.filter_bases));
        look = (lib.d_filter[fi >> 5] >> (fi & 31)) & 1u;
      }
      if (look) {
        const uint32_t bsh = 32 - lib.bucket_bits;
        uint32_t l = lib.d_idx_bucket[k_lo >> bsh], h = lib.d_idx_bucket[(k_hi >> bsh) + 1];
        const uint32_t hi0 = h;
        while (l < h) { const uint32_t mid = (l + h) >> 1; if (lib.d_idx_kmer[mid] < k_lo) l = mid + 1; else h = mid; }
        p_lo[pi] = l;
        h = hi0;
        while (l < h) { const uint32_t mid = (l + h) >> 1; if (lib.d_idx_kmer[mid] <= k_hi) l = mid + 1; else h = mid; }
        p_hi[pi] = l;
        total += p_hi[pi] - p_lo[pi];
      }
      a = b;
    }
  }
  if (total > LEAN_MAX_CAND) return false;
  const int nw = (L + 15) >> 4;
  a = 0;
#pragma unroll
  for (int pi = 0; pi < MAX_PIECES; ++pi) {
    if (pi < np) {
      for (uint32_t e = p_lo[pi]; e < p_hi[pi]; ++e) {
        const uint32_t pos = lib.d_idx_pos[e];
        if (pos < (uint32_t)a) continue;
        const uint64_t astart = pos - (uint32_t)a;
        // text under the round policy (verify_text on the column)
        int mm = 0, smm = 0;
        for (int w = 0; w < nw && mm >= 0; ++w) {
          uint32_t x = lean_word16<S>(col, qs + 16 * w) ^ lib_word16(lib.d_packed, astart + 16 * (uint64_t)w);
          x = (x | (x >> 1)) & 0x55555555u;
          const int rem = L - 16 * w;
          if (rem < 16) x &= (1u << (2 * rem)) - 1u;
          if (x) {
            mm += __popc(x);
            const int srem = R - 16 * w;
            if (srem >= 16) smm += __popc(x);
            else if (srem > 0) smm += __popc(x & ((1u << (2 * srem)) - 1u));
            if (mm > pol.total_mm || smm > pol.seed_mm) mm = -1;
          }
        }
        if (mm < 0) continue;
        const uint32_t r = find_ref(lib, pos);
        const uint32_t rlo = lib.d_ref_off[r], rhi = lib.d_ref_off[r + 1];
        if (astart < rlo || astart + (uint64_t)L > rhi) continue;
        if (ref_has_n(lib, astart, astart + L)) continue;
        const uint64_t hword = ((uint64_t)mm << 56) | ((uint64_t)r << 28) | (uint64_t)(astart - rlo);
        if (hword < best) best = hword;
      }
      a = piece_bound(pi + 1, R, np);
    }
  }
  return true;
}

// pass 2: a WARP owns SUB_KEYS consecutive sequences and runs the rounds on them in order, alone: per round it
// compacts the sequences to search (mask bit set, not annotated by an earlier round) into its shared-memory list and
// searches them 32 at a time.  No CTA barrier: warps do not wait for each other's searches.
#define SUB_KEYS 256
struct WarpTile {
  uint16_t mask[SUB_KEYS];
  uint16_t list[SUB_KEYS];
  uint8_t state[SUB_KEYS];
  uint32_t pay[LEAN_WORDS + 2][32];  // payload column of the sequence a lane is searching
};

// Sequences the lean search does not take (exception words, > 128 bases, a degenerate seed, many candidates) leave
// pass 2 at the round that meets them: (id, round index) goes to a list and pass 3 (annot_general_kernel) runs that
// round and every later one for them with the general code, thread per sequence, candidates shared across the warp.
struct GeneralList {
  unsigned long long *count;
  uint32_t *id;
  uint8_t *round;
};

__global__ void __launch_bounds__(ANN_THREADS, 10)
annot_search_kernel(const __grid_constant__ RoundSet rs, mirge_table t, uint64_t n_keys, const uint16_t *__restrict__ masks,
                    uint8_t *__restrict__ annot_round, uint64_t *__restrict__ hit, const GeneralList gl) {
  __shared__ WarpTile s_wt[ANN_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpTile &W = s_wt[warp];
  const uint64_t tile0 = ((uint64_t)blockIdx.x * (ANN_THREADS / 32) + warp) * SUB_KEYS;
  if (tile0 >= n_keys) return;
  const uint32_t n_tile = (uint32_t)min((uint64_t)SUB_KEYS, n_keys - tile0);
  uint32_t any = 0;
  for (uint32_t k = lane; k < SUB_KEYS; k += 32) {
    const uint32_t m = k < n_tile ? masks[tile0 + k] : 0u;
    W.mask[k] = (uint16_t)m;
    W.state[k] = k < n_tile ? annot_round[tile0 + k] : (uint8_t)0;
    any |= m;
  }
  any = __reduce_or_sync(0xffffffffu, any);
  __syncwarp();
  for (int ri = 0; ri < rs.n; ++ri) {
    if (!((any >> ri) & 1u)) continue;
    const mirge_round_policy &pol = rs.pol[ri];
    uint32_t cnt = 0;
    for (uint32_t k0 = 0; k0 < SUB_KEYS; k0 += 32) {
      const uint32_t k = k0 + lane;
      bool p = (W.mask[k] >> ri) & 1u;
      if (p && pol.select == MIRGE_SELECT_UNANNOTATED) p = W.state[k] == 0xFF;
      const unsigned bal = __ballot_sync(0xffffffffu, p);
      if (p) W.list[cnt + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)k;
      cnt += __popc(bal);
    }
    __syncwarp();
    for (uint32_t base = 0; base < cnt; base += 32) {
      const uint32_t idx = base + lane;
      const bool active = idx < cnt;
      const uint32_t k = active ? W.list[idx] : 0u;
      const uint64_t id = tile0 + k;
      uint64_t best = MIRGE_NO_HIT;
      bool general = false;
      if (active) {
        const KeyView kv = key_view(t.d_arena + t.d_key_ref[id]);
        const int npay = (kv.len + 15) >> 4;
        general = true;
        if (kv.nexc == 0 && npay <= LEAN_WORDS) {
          uint32_t *col = &W.pay[0][lane];
#pragma unroll
          for (int w = 0; w < LEAN_WORDS + 2; ++w) col[w * 32] = w < npay ? kv.pay[w] : 0u;
          int qs, qe, tlen = -1;
          if (!lean_window<32>(col, kv.len, pol, tlen, qs, qe)) general = false;  // (cannot happen: the mask bit says the round searches it)
          else general = !search_lean<32>(rs.lib[ri], pol, col, qs, qe - qs, best);
        }
      }
      const unsigned gm = __ballot_sync(0xffffffffu, general);
      if (gm) {  // hand these sequences to pass 3 from this round on; no later round of this pass looks at them
        const int leader = __ffs(gm) - 1;
        unsigned long long gb = 0;
        if (lane == leader) gb = atomicAdd(gl.count, (unsigned long long)__popc(gm));
        gb = __shfl_sync(0xffffffffu, gb, leader);
        if (general) {
          const unsigned long long e = gb + __popc(gm & ((1u << lane) - 1u));
          gl.id[e] = (uint32_t)id;
          gl.round[e] = (uint8_t)ri;
          W.mask[k] = 0;
        }
      }
      if (active && best != MIRGE_NO_HIT) {
        W.state[k] = (uint8_t)pol.round;
        annot_round[id] = (uint8_t)pol.round;
        hit[id] = best;
      }
    }
    __syncwarp();
  }
}

// pass 3: the listed sequences, rounds from the listed one on (entries beyond *count: CTAs exit at once)
__global__ void __launch_bounds__(ANN_THREADS, 10)
annot_general_kernel(const __grid_constant__ RoundSet rs, mirge_table t, const GeneralList gl, uint8_t *__restrict__ annot_round,
                     uint64_t *__restrict__ hit) {
  __shared__ WarpScratch s_ws[ANN_THREADS / 32];
  const unsigned long long n = *gl.count;
  if ((uint64_t)blockIdx.x * ANN_THREADS >= n) return;  // uniform per CTA
  const uint64_t i = (uint64_t)blockIdx.x * ANN_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool in_range = i < n;
  const uint64_t id = in_range ? gl.id[i] : 0;
  const int first = in_range ? gl.round[i] : rs.n;
  uint32_t qw[QW_MAX], qnx[QW_MAX];
  KeyView kv;
  kv.pay = kv.exc = nullptr; kv.len = kv.nexc = 0;
  int cur_qs = -1, cur_qe = -1, tlen = -1;
  uint32_t state = 0xFF;
  if (in_range) {
    kv = key_view(t.d_arena + t.d_key_ref[id]);
    state = annot_round[id];
  }
  const int len = kv.len, nexc = kv.nexc;
  uint64_t my_hit = MIRGE_NO_HIT;
  int my_round = -1;
  bool republish = true;
  const int lo = __reduce_min_sync(0xffffffffu, first);
  for (int ri = lo; ri < rs.n; ++ri) {
    const mirge_round_policy &pol = rs.pol[ri];
    bool active = in_range && ri >= first;
    if (active) {
      if (pol.select == MIRGE_SELECT_LEN_LT26) active = len < 26;
      else if (pol.select == MIRGE_SELECT_LEN_GT25) active = len > 25;
      else active = state == 0xFF;
    }
    int L = 0;
    if (active) {
      int qs, qe;
      active = round_window(kv, pol, tlen, qs, qe);
      if (active) {
        L = qe - qs;
        if (qs != cur_qs || qe != cur_qe) {
          build_query(kv, qs, qe, qw, qnx);
          cur_qs = qs;
          cur_qe = qe;
          republish = true;
        }
      }
    }
    const uint64_t best = search_round(rs.lib[ri], pol, active, qw, qnx, L, nexc > 0, s_ws[warp], republish, lane);
    if (active && best != MIRGE_NO_HIT) {
      state = (uint32_t)pol.round;
      my_round = pol.round;
      my_hit = best;
    }
  }
  if (my_round >= 0) {
    annot_round[id] = (uint8_t)my_round;
    hit[id] = my_hit;
  }
}

extern "C" uint64_t mirge_annotate_scratch_bytes(uint64_t n_keys) {
  const uint64_t n = (n_keys + 7) & ~7ull;
  return 16 + 2 * n + 4 * n + n;  // counter, round masks, list of (id, round) for the general pass
}

static int check_round(mirge_ctx *ctx, const mirge_library *lib, const mirge_round_policy *policy) {
  if (!lib->d_packed || !lib->d_nmask || !lib->d_ref_off || !lib->d_idx_kmer || !lib->d_idx_pos || !lib->d_idx_bucket ||
      !lib->d_ref_block)
    MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: library has no index");
  if (lib->ref_block_shift > 20) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: ref_block_shift out of range");
  if (lib->filter_bases && (!lib->d_filter || lib->filter_bases < 4 || lib->filter_bases > 16))
    MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: bad prefix filter");
  if (lib->filter16_bits && (!lib->d_filter16 || lib->filter16_bits < 16 || lib->filter16_bits > 30))
    MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: bad 16-mer filter");
  if (lib->bucket_bits < 1 || lib->bucket_bits > 28) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: bucket_bits out of range");
  if (policy->seed_mm < 0 || policy->seed_mm > 3 || policy->total_mm < policy->seed_mm || policy->trim5 < 0 || policy->trim3 < 0)
    MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: unsupported policy");
  if (lib->n_refs >= (1u << 28)) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: too many references");
  return MIRGE_OK;
}

extern "C" int mirge_annotate_rounds(mirge_ctx *ctx, const mirge_library *libs, const mirge_round_policy *policies, int n_rounds,
                                     const mirge_table *t, uint64_t n_keys, uint8_t *d_annot_round, uint64_t *d_hit,
                                     void *d_scratch, void *stream_) {
  if (!ctx) return MIRGE_ERR_ARG;
  if (!libs || !policies || !t || !d_annot_round || !d_hit) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: null argument");
  if (n_keys > 0xFFFFFFFFull) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: more than 2^32 sequences in one call");
  if (n_rounds < 0 || n_rounds > MAX_ROUNDS) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: at most %d rounds per call", MAX_ROUNDS);
  if (n_keys == 0 || n_rounds == 0) return MIRGE_OK;
  RoundSet rs;
  rs.n = 0;
  for (int i = 0; i < n_rounds; ++i) {
    if (libs[i].n_refs == 0 || libs[i].n_bases == 0) continue;  // empty library: nothing can hit
    int rc = check_round(ctx, &libs[i], &policies[i]);
    if (rc) return rc;
    rs.lib[rs.n] = libs[i];
    rs.pol[rs.n] = policies[i];
    ++rs.n;
  }
  if (rs.n == 0) return MIRGE_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  MIRGE_CUDA(ctx, cudaSetDevice(ctx->device));
  if (d_scratch) {
    if ((uintptr_t)d_scratch & 7) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: misaligned scratch");
    const uint64_t n8 = (n_keys + 7) & ~7ull;
    GeneralList gl;
    gl.count = (unsigned long long *)d_scratch;
    uint16_t *masks = (uint16_t *)((uint8_t *)d_scratch + 16);
    gl.id = (uint32_t *)((uint8_t *)d_scratch + 16 + 2 * n8);
    gl.round = (uint8_t *)d_scratch + 16 + 6 * n8;
    MIRGE_CUDA(ctx, cudaMemsetAsync(gl.count, 0, 16, stream));
    annot_mask_kernel<<<(unsigned)((n_keys + MASK_THREADS - 1) / MASK_THREADS), MASK_THREADS, 0, stream>>>(rs, *t, n_keys, d_annot_round, masks);
    MIRGE_LAUNCH_CHECK(ctx, "annot_mask_kernel");
    const uint64_t per_cta = (uint64_t)SUB_KEYS * (ANN_THREADS / 32);
    annot_search_kernel<<<(unsigned)((n_keys + per_cta - 1) / per_cta), ANN_THREADS, 0, stream>>>(rs, *t, n_keys, masks, d_annot_round, d_hit, gl);
    MIRGE_LAUNCH_CHECK(ctx, "annot_search_kernel");
    annot_general_kernel<<<(unsigned)((n_keys + ANN_THREADS - 1) / ANN_THREADS), ANN_THREADS, 0, stream>>>(rs, *t, gl, d_annot_round, d_hit);
    MIRGE_LAUNCH_CHECK(ctx, "annot_general_kernel");
    return MIRGE_OK;
  }
  annotate_kernel<<<(unsigned)((n_keys + ANN_THREADS - 1) / ANN_THREADS), ANN_THREADS, 0, stream>>>(rs, *t, n_keys, d_annot_round, d_hit);
  MIRGE_LAUNCH_CHECK(ctx, "annotate_kernel");
  return MIRGE_OK;
}

extern "C" int mirge_annotate_round(mirge_ctx *ctx, const mirge_library *lib, const mirge_round_policy *policy, const mirge_table *t,
                                    uint64_t n_keys, uint8_t *d_annot_round, uint64_t *d_hit, void *stream_) {
  if (!ctx) return MIRGE_ERR_ARG;
  if (!lib || !policy) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: null argument");
  return mirge_annotate_rounds(ctx, lib, policy, 1, t, n_keys, d_annot_round, d_hit, nullptr, stream_);
}

extern "C" int mirge_annotate_allhits(mirge_ctx *ctx, const mirge_library *lib, const mirge_round_policy *policy, const mirge_table *t,
                                      const uint32_t *d_ids, uint64_t n_ids, const uint64_t *d_hit, int fill, uint32_t *d_counts,
                                      const uint64_t *d_offs, uint64_t *d_out, void *stream_) {
  if (!ctx) return MIRGE_ERR_ARG;
  if (!lib || !policy || !t || !d_ids || !d_hit) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "allhits: null argument");
  if (fill ? (!d_offs || !d_out) : !d_counts) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "allhits: null output buffer");
  if (n_ids == 0) return MIRGE_OK;
  int rc = check_round(ctx, lib, policy);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  MIRGE_CUDA(ctx, cudaSetDevice(ctx->device));
  allhits_kernel<<<(unsigned)((n_ids + ANN_THREADS - 1) / ANN_THREADS), ANN_THREADS, 0, stream>>>(*lib, *policy, *t, d_ids, n_ids, d_hit,
                                                                                                   fill, d_counts, d_offs, d_out);
  MIRGE_LAUNCH_CHECK(ctx, "allhits_kernel");
  return MIRGE_OK;
}
