// trim.cu -- per-read digest kernel: the body of the reference worker ``cutadapt(n)``
// (mirge/libs/digest.py:320-375) for one read per thread:
//   NextSeq trim -> quality trim -> adapter removal -> N-end trim -> unconditional cuts
//   (modifier order of stipulate(), digest.py:87-99), key emission + length filter
//   (digest.py:332-373, HEAD per-stage counting or release post-pipeline counting), and 2-bit
//   packing of every emitted key for the collapse table.
// cutadapt semantics (3P, restated in oracle/pyoracle.py): qualtrim.pyx nextseq_trim_index /
// quality_trim_index; _align.pyx Aligner.locate with (cost, origin, matches) cells and the
// "most matches, then lowest cost, then first found" objective; NEndTrimmer; UnconditionalCutter.
//
// Data movement: a CTA owns TRIM_THREADS consecutive records, whose bytes are one contiguous span
// of the FASTQ stream; the span is staged in shared memory with coalesced 128-bit streaming loads
// and every thread then works on its own record out of shared memory.  All arithmetic is integer.
#include "common.cuh"

#define TRIM_THREADS 128
#define OB 64  // origin bias inside the packed (origin, matches) word

struct DevAdapter {
  int where, m, min_overlap, indel_cost, wildcard_ref, k, effective_length, pad;
  uint64_t peq[5];  // bit i-1 set <=> adapter row i matches read class c (A,C,G,T,other)
  int n_counts[MIRGE_MAX_ADAPTER_LEN + 1];
  int max_err[MIRGE_MAX_ADAPTER_LEN + 1];
  int acc[MIRGE_MAX_ADAPTER_LEN + 1];  // 3' adapters: max errors of a candidate ending in adapter row i, -1 = never
  uint64_t a2;                         // 2-bit text of the adapter (row t at bits 2(t-1)); plain ACGT adapters <= 32 nt
};
struct DevParams {
  int n_mods, kind[MIRGE_MAX_MODS], a[MIRGE_MAX_MODS], b[MIRGE_MAX_MODS], c[MIRGE_MAX_MODS];
  int n_adapters, times, min_len, umi_mode, umi5, umi3, qia_len, slots;
  DevAdapter ad[MIRGE_MAX_ADAPTERS];
};
__constant__ DevParams c_p;

static int trim_slots_of(const mirge_trim_params *p) {
  return (p->umi_mode != MIRGE_UMI_QIAGEN && p->count_mode == MIRGE_COUNT_HEAD) ? p->n_mods : 1;
}

extern "C" int mirge_trim_slots(const mirge_ctx *ctx) {
  if (!ctx || !ctx->params_set) return MIRGE_ERR_ARG;
  return trim_slots_of(&ctx->params);
}

extern "C" int mirge_set_trim_params(mirge_ctx *ctx, const mirge_trim_params *p) {
  if (!ctx || !p) return MIRGE_ERR_ARG;
  if (p->n_mods < 0 || p->n_mods > MIRGE_MAX_MODS) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "n_mods out of range");
  if (p->n_adapters < 0 || p->n_adapters > MIRGE_MAX_ADAPTERS) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "n_adapters out of range");
  DevParams d;
  memset(&d, 0, sizeof(d));
  d.n_mods = p->n_mods;
  for (int i = 0; i < p->n_mods; ++i) {
    if (p->mod_kind[i] < MIRGE_MOD_NEXTSEQ || p->mod_kind[i] > MIRGE_MOD_CUT) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "unknown modifier kind %d", p->mod_kind[i]);
    d.kind[i] = p->mod_kind[i];
    d.a[i] = p->mod_a[i];
    d.b[i] = p->mod_b[i];
    d.c[i] = p->mod_c[i];
  }
  d.n_adapters = p->n_adapters;
  d.times = p->times;
  d.min_len = p->min_len;
  d.umi_mode = p->umi_mode;
  d.umi5 = p->umi5;
  d.umi3 = p->umi3;
  d.qia_len = p->qia_adapter_len;
  d.slots = trim_slots_of(p);
  if (p->umi_mode < MIRGE_UMI_NONE || p->umi_mode > MIRGE_UMI_QIAGEN) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "bad umi_mode");
  if (p->umi_mode != MIRGE_UMI_NONE && (p->umi5 < 0 || p->umi3 < 0)) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "negative UMI length");
  if (p->umi_mode == MIRGE_UMI_QIAGEN && p->n_adapters < 1) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "qiagen UMI mode needs an adapter");
  int maxm = 0, fast_ok = 1;
  for (int a = 0; a < p->n_adapters; ++a) {
    const mirge_adapter *s = &p->adapters[a];
    if (s->m < 1 || s->m > MIRGE_MAX_ADAPTER_LEN) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "adapter %d length %d unsupported", a, s->m);
    if (s->where != 0 && s->where != 1) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "adapter %d: unknown type", a);
    DevAdapter *o = &d.ad[a];
    o->where = s->where; o->m = s->m; o->min_overlap = s->min_overlap; o->indel_cost = s->indel_cost;
    o->wildcard_ref = s->wildcard_ref; o->k = s->k; o->effective_length = s->effective_length;
    for (int c = 0; c < 4; ++c) {
      uint64_t bits = 0;
      for (int i = 0; i < s->m; ++i)
        if (s->mask[i] & (1 << c)) bits |= 1ull << i;
      o->peq[c] = bits;
    }
    o->peq[4] = 0;  // a read character outside ACGT never matches (match_read_wildcards=False)
    memcpy(o->n_counts, s->n_counts, sizeof(o->n_counts));
    memcpy(o->max_err, s->max_err, sizeof(o->max_err));
    for (int i = 0; i <= s->m; ++i) {
      const int eff = s->wildcard_ref ? i - s->n_counts[i] : i;
      o->acc[i] = (i >= s->min_overlap && i >= 1 && eff >= 0) ? s->max_err[eff] : -1;
    }
    o->a2 = 0;
    if (!s->wildcard_ref && s->m <= 32)
      for (int i = 0; i < s->m; ++i) {
        const uint64_t code = s->ascii[i] == 'A' ? 0 : s->ascii[i] == 'C' ? 1 : s->ascii[i] == 'G' ? 2 : 3;
        o->a2 |= code << (2 * i);
      }
    if (s->where != 0 || s->indel_cost != 1 || s->m > 32 || s->min_overlap < 1 || s->m + 2 * s->k + 3 > 48) fast_ok = 0;
    if (s->m > maxm) maxm = s->m;
  }
  MIRGE_CUDA(ctx, cudaSetDevice(ctx->device));
  MIRGE_CUDA(ctx, cudaMemcpyToSymbol(c_p, &d, sizeof(d)));
  ctx->params = *p;
  ctx->params_set = 1;
  ctx->max_adapter_len = maxm;
  ctx->fast_ok = fast_ok;
  return MIRGE_OK;
}

// ------------------------------------------------------------------ quality trimming --------

__device__ __forceinline__ int nextseq_trim_index(const uint8_t *seq, const uint8_t *qual, int len, int cutoff, int base) {
  int s = 0, max_qual = 0, max_i = len;
  for (int i = len - 1; i >= 0; --i) {
    int q = (int)qual[i] - base;
    if (seq[i] == 'G') q = cutoff - 1;
    s += cutoff - q;
    if (s < 0) break;
    if (s > max_qual) { max_qual = s; max_i = i; }
  }
  return max_i;
}

__device__ __forceinline__ void quality_trim_index(const uint8_t *qual, int len, int q5, int q3, int base, int &start, int &stop) {
  int s = 0, max_qual = 0;
  start = 0; stop = len;
  for (int i = 0; i < len; ++i) {
    s += q5 - ((int)qual[i] - base);
    if (s < 0) break;
    if (s > max_qual) { max_qual = s; start = i + 1; }
  }
  max_qual = 0; s = 0;
  for (int i = len - 1; i >= 0; --i) {
    s += q3 - ((int)qual[i] - base);
    if (s < 0) break;
    if (s > max_qual) { max_qual = s; stop = i; }
  }
  if (start >= stop) { start = 0; stop = 0; }
}

// ------------------------------------------------------------------ adapter alignment -------

struct Match { int rstart, rstop, matches, errors; };

// cutadapt Aligner.locate: full-column DP in registers, one column per read base.
template <int MAXM>
__device__ __noinline__ bool locate(const int a, const uint8_t *read, const int n, Match &out) {
  const DevAdapter &ad = c_p.ad[a];
  const int m = ad.m, ic = ad.indel_cost, k = ad.k;
  const bool back = ad.where == 0;
  const uint64_t p0 = ad.peq[0], p1 = ad.peq[1], p2 = ad.peq[2], p3 = ad.peq[3];
  int cost[MAXM + 1], om[MAXM + 1];
#pragma unroll
  for (int i = 0; i <= MAXM; ++i) {
    cost[i] = back ? i * ic : 0;
    om[i] = ((back ? 0 : -i) + OB) << 8;
  }
  int best_cost = m + n, best_om = OB << 8, best_ref_stop = m, best_query_stop = n;
  bool stopped = false;
  for (int j = 1; j <= n; ++j) {
    const uint32_t rc = base_code_upper(read[j - 1]);
    const uint64_t eq = rc == 0 ? p0 : rc == 1 ? p1 : rc == 2 ? p2 : rc == 3 ? p3 : 0ull;
    int dc = cost[0], dom = om[0];
    om[0] = (j + OB) << 8;
    int cm = 0, omm = 0;
#pragma unroll
    for (int i = 1; i <= MAXM; ++i) {
      if (i <= m) {
        int c, o;
        if ((eq >> (i - 1)) & 1ull) {
          c = dc; o = dom + 1;
        } else {
          const int cd = dc + 1, cdel = cost[i] + ic, cins = cost[i - 1] + ic;
          if (cd <= cdel && cd <= cins) { c = cd; o = dom; }
          else if (cins <= cdel) { c = cins; o = om[i - 1]; }
          else { c = cdel; o = om[i]; }
        }
        dc = cost[i]; dom = om[i];
        cost[i] = c; om[i] = o;
        if (i == m) { cm = c; omm = o; }
      }
    }
    if (cm <= k) {
      const int origin = (omm >> 8) - OB, mt = omm & 0xFF;
      const int length = m + min(origin, 0);
      int eff = length;
      if (ad.wildcard_ref) eff = (length < m) ? length - (ad.n_counts[m] - ad.n_counts[m - length]) : ad.effective_length;
      const int bm = best_om & 0xFF;
      if (length >= ad.min_overlap && cm <= ad.max_err[eff] && (mt > bm || (mt == bm && cm < best_cost))) {
        best_cost = cm; best_om = omm; best_ref_stop = m; best_query_stop = j;
        if (cm == 0 && mt == m) { stopped = true; break; }
      }
    }
  }
  if (!stopped) {
    const int first_i = back ? 0 : m;
#pragma unroll
    for (int i = 0; i <= MAXM; ++i) {
      if (i >= first_i && i <= m) {
        const int origin = (om[i] >> 8) - OB, mt = om[i] & 0xFF, c = cost[i];
        const int length = i + min(origin, 0);
        int eff = length;
        if (ad.wildcard_ref) {
          if (length < m) { const int ref_start = origin < 0 ? -origin : 0; eff = length - (ad.n_counts[i] - ad.n_counts[ref_start]); }
          else eff = ad.effective_length;
        }
        const int bm = best_om & 0xFF;
        if (length >= ad.min_overlap && eff >= 0 && c <= ad.max_err[eff] && (mt > bm || (mt == bm && c < best_cost))) {
          best_cost = c; best_om = om[i]; best_ref_stop = i; best_query_stop = n;
        }
      }
    }
  }
  (void)best_ref_stop;
  if (best_cost == m + n) return false;
  const int origin = (best_om >> 8) - OB;
  out.rstart = origin >= 0 ? origin : 0;
  out.rstop = best_query_stop;
  out.matches = best_om & 0xFF;
  out.errors = best_cost;
  return true;
}

// ------------------------------------------------------------------ bit-parallel locate ------
// Same result as locate() for 3' adapters with unit indel cost and m <= 32, at ~15 instructions per
// read base instead of ~15 per DP cell:
//   1. Myers/Hyyro bit-vector recurrence gives the exact DP cost column (vertical deltas VP/VN) for
//      every read position; the cost of adapter row m is tracked incrementally.  Nothing is stored.
//   2. cutadapt's candidates are the row-m cell of every column and all rows of the last column; their
//      costs come from (1).  A candidate's (origin, matches) are those of the path cutadapt's
//      tie-breaking (mismatch, then insertion, then deletion) propagates into that cell; the path is
//      recovered by a traceback that needs only DP *costs* of neighbouring cells.  Those are recomputed
//      on demand for the last m + 2k + 2 columns before the candidate with a fresh-start Myers pass into
//      a thread-local buffer: a path with <= k errors into row i spans <= i + k columns, and every
//      neighbour whose cost can tie has an optimal path starting inside that window, so the costs the
//      rule compares are exact (larger values only lose).  cost == 0 cells are pure diagonals.
//   3. The winner is the maximum of (matches, -cost, -scan order) as in Aligner.locate, so candidates
//      may be evaluated in any order.  Cheap exact pruning keeps tracebacks rare: a row-m candidate
//      entered by a deletion is dominated by its left neighbour; one entered by an insertion is dominated
//      when the next column's cell is a character match; a candidate whose first step is not a match has
//      at most row-1 matches; candidates are tried best-first and skipped when they cannot win.
//   All tracebacks run after the column loop, so the lanes of a warp execute them together.
#define RB 48  // recompute window capacity in columns (m + 2k + 3 <= RB is checked on the host)

struct FastCtx {
  const uint32_t *s_eq;  // [n_adapters][256]: bit i-1 set <=> adapter row i matches this read byte
  const uint32_t *ps;    // this thread's 2-bit packed read (16 bases per word), stride TRIM_THREADS words
  bool jump_ok;          // ps is valid and the read is pure "ACGT": match runs can be skipped with bit tricks
  int rbase;             // offset of the window being searched inside the read
};

struct ColBuf {
  uint32_t vp[RB], vn[RB];
  int j0;  // entry t holds column j0 + t
};

__device__ __forceinline__ int cell_cost(uint32_t vp, uint32_t vn, int r) {
  const uint32_t mask = r >= 32 ? 0xFFFFFFFFu : ((1u << r) - 1u);
  return __popc(vp & mask) - __popc(vn & mask);
}

#define MYERS_STEP(eq, vp, vn, hp_out, hn_out)                    \
  {                                                               \
    const uint32_t xv_ = (eq) | (vn);                             \
    const uint32_t xh_ = ((((eq) & (vp)) + (vp)) ^ (vp)) | (eq);  \
    uint32_t hp_ = (vn) | ~(xh_ | (vp));                          \
    uint32_t hn_ = (vp) & xh_;                                    \
    hp_out = hp_;                                                 \
    hn_out = hn_;                                                 \
    hp_ <<= 1;                                                    \
    hn_ <<= 1;                                                    \
    (vp) = hn_ | ~(xv_ | hp_);                                    \
    (vn) = hp_ & xv_;                                             \
  }

// cost columns jc - span .. jc recomputed with a fresh start (exact where it matters, see above)
__device__ __noinline__ void recompute(const uint32_t *eqt, const uint8_t *read, int jc, int span, ColBuf &cb) {
  const int j0 = max(0, jc - span);
  uint32_t vp = 0xFFFFFFFFu, vn = 0u;
  cb.j0 = j0;
  cb.vp[0] = vp;
  cb.vn[0] = vn;
  int t = 1;
  for (int j = j0 + 1; j <= jc; ++j, ++t) {
    const uint32_t eq = eqt[read[j - 1]];
    uint32_t hp, hn;
    MYERS_STEP(eq, vp, vn, hp, hn)
    cb.vp[t] = vp;
    cb.vn[t] = vn;
  }
}

// 64 bits = 32 bases of the packed read starting at base `a0` (zero beyond the packed words)
__device__ __forceinline__ uint64_t read_window(const uint32_t *ps, int a0) {
  const int wi = a0 >> 4, sh = 2 * (a0 & 15);
  const uint32_t w0 = ps[wi * TRIM_THREADS], w1 = ps[(wi + 1) * TRIM_THREADS], w2 = ps[(wi + 2) * TRIM_THREADS];
  const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
  return ((uint64_t)hi << 32) | lo;
}

// (matches, origin) cutadapt's DP holds in cell (i, j) whose cost is c > 0, from the cost columns in cb.
// jump: adapter without wildcards and a pure-ACGT packed read -- the run of matches up a diagonal is skipped
// with one XOR + count-leading-zeros instead of one step per cell.
__device__ __noinline__ void traceback(const int a, const uint32_t *eqt, const uint8_t *read, const FastCtx &fc, const ColBuf &cb,
                                          int i, int j, int c, int &matches, int &origin) {
  const bool jump = fc.jump_ok && !c_p.ad[a].wildcard_ref;
  const uint64_t a2 = c_p.ad[a].a2;
  int r = i, col = j, cost = c, nonmatch = 0;
  while (r > 0 && cost > 0) {
    if (col == 0) {  // initial column: cost r, origin 0, no further matches
      nonmatch += r;
      r = 0;
      break;
    }
    const int delta = col - r;
    if (jump && delta >= 0) {
      // rows t = 1..r of this diagonal face read bases delta + t - 1 (all in columns >= 1)
      uint64_t x = a2 ^ read_window(fc.ps, fc.rbase + delta);
      x = (x | (x >> 1)) & 0x5555555555555555ull;
      if (r < 32) x &= (1ull << (2 * r)) - 1ull;
      if (x == 0) {  // cannot happen while cost > 0 (an all-match diagonal has cost 0); kept for safety
        col -= r;
        r = 0;
        break;
      }
      const int t = (64 - __clzll((long long)x) + 1) >> 1;  // highest mismatching row <= r
      col -= r - t;
      r = t;
    } else if ((eqt[read[col - 1]] >> (r - 1)) & 1u) {  // equal characters: diagonal, cost unchanged
      --r; --col;
      continue;
    }
    const int t0 = col - cb.j0;  // >= 1 by the window bound
    const uint32_t vp0 = cb.vp[t0], vn0 = cb.vn[t0], vp1 = cb.vp[t0 - 1], vn1 = cb.vn[t0 - 1];
    const int cd = cell_cost(vp1, vn1, r - 1) + 1, cdel = cell_cost(vp1, vn1, r) + 1, cins = cell_cost(vp0, vn0, r - 1) + 1;
    if (cd <= cdel && cd <= cins) { --r; --col; ++nonmatch; cost = cd - 1; }
    else if (cins <= cdel) { --r; ++nonmatch; cost = cins - 1; }
    else { --col; cost = cdel - 1; }
  }
  origin = col - r;  // r > 0 here means cost == 0: r more diagonal matches
  matches = i - nonmatch;
}

// Cost-1 cells need no cost columns: slide up the diagonal to the only error and decide its kind by asking
// which neighbour has cost 0 -- a cost-0 cell (a, b) simply means adapter[0:a] == read[b-a:b], one packed
// compare.  cutadapt's rule picks mismatch if (t-1, col-1) is such a cell, else insertion if (t-1, col) is,
// else deletion.  Needs the packed pure-ACGT read and a plain adapter; returns false when it does not apply.
__device__ __forceinline__ bool traceback_cost1(const int a, const FastCtx &fc, int i, int j, int &matches, int &origin) {
  const int delta = j - i;
  if (!fc.jump_ok || c_p.ad[a].wildcard_ref || delta < 0) return false;
  const uint64_t a2 = c_p.ad[a].a2;
  const uint64_t even = 0x5555555555555555ull;
  uint64_t x = a2 ^ read_window(fc.ps, fc.rbase + delta);
  x = (x | (x >> 1)) & even;
  if (i < 32) x &= (1ull << (2 * i)) - 1ull;
  if (x == 0) return false;  // not a cost-1 cell of this diagonal (cannot happen); use the general path
  const int t = (64 - __clzll((long long)x) + 1) >> 1;  // row of the error
  const uint64_t below = t - 1 >= 32 ? ~0ull : ((1ull << (2 * (t - 1))) - 1ull);  // rows 1..t-1
  if ((x & below) == 0) {  // mismatch: rows above continue cleanly on the same diagonal
    matches = i - 1;
    origin = delta;
    return true;
  }
  uint64_t y = a2 ^ read_window(fc.ps, fc.rbase + delta + 1);
  y = (y | (y >> 1)) & even & below;
  if (y == 0) {  // insertion: adapter[0:t-1] ends at the same read base
    matches = i - 1;
    origin = delta + 1;
    return true;
  }
  if (delta < 1) return false;
  const uint64_t upto = t >= 32 ? ~0ull : ((1ull << (2 * t)) - 1ull);  // rows 1..t
  uint64_t z = a2 ^ read_window(fc.ps, fc.rbase + delta - 1);
  z = (z | (z >> 1)) & even & upto;
  if (z != 0) return false;  // inconsistent with cost 1: let the general path decide
  matches = i;  // deletion: a read base is skipped, every adapter row matches
  origin = delta - 1;
  return true;
}

// A cell (i, j) of cost c reaches i matches only if every error is a deletion; that path leaves row 0 at
// column j - i - c with a match, so adapter[0] must equal that read base (cheap necessary condition).
__device__ __forceinline__ bool all_deletions_possible(const uint32_t *eqt, const uint8_t *read, int i, int j, int c) {
  const int o = j - i - c;
  return o >= 0 && (eqt[read[o]] & 1u);
}

// candidate with at most `u` matches, cost c, scan index idx can still beat the best so far
#define MAY_WIN(u, c, idx) (!have || (u) > b_m || ((u) == b_m && ((c) < b_c || ((c) == b_c && (idx) < b_idx))))
#define TAKE_IF_BETTER(mt, c, org, idx)                                                      \
  if (!have || (mt) > b_m || ((mt) == b_m && ((c) < b_c || ((c) == b_c && (idx) < b_idx)))) { \
    have = true; b_m = (mt); b_c = (c); b_o = (org); b_idx = (idx);                          \
  }
// queued row-m candidate: column | cost << 16 | (first step is a match) << 24
#define Q_COL(v) ((int)((v)&0xFFFFu))
#define Q_COST(v) ((int)(((v) >> 16) & 0xFFu))
#define Q_UB(v, m) ((((v) >> 24) & 1u) ? (m) : (m)-1)

// DEFER (first pass of the two-pass scheme): as soon as a candidate would need cost columns (recompute +
// general traceback) the search gives up with 2 and the read is queued for the second pass, which runs
// this function without DEFER on a compacted list, so that the rare expensive path executes with full warps.
// Returns 0 = no match, 1 = match in `out`, 2 = deferred.
template <bool DEFER>
__device__ __forceinline__ int locate_fast(const int a, const uint8_t *read, const int n, const FastCtx &fc, Match &out) {
  const unsigned lanes = __activemask();  // lanes searching together; re-converged after the divergent loops
  const DevAdapter &ad = c_p.ad[a];
  const int m = ad.m;
  const int span = m + 2 * ad.k + 2;
  const uint32_t *eqt = fc.s_eq + a * 256;
  const uint32_t top = 1u << (m - 1);
  const int acc_m = ad.acc[m];
  uint32_t vp = 0xFFFFFFFFu, vn = 0u, pvp = vp, pvn = vn, eq = 0;
  int score = m;
  bool have = false;
  int b_m = 0, b_c = 0, b_o = 0, b_idx = 0;
  bool stopped = false;
  ColBuf cb;
  // row-m candidates wait here until the column loop is over (0 = empty slot)
  uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0;
  // a new candidate waits one column: the next column may prove it dominated
  bool pend = false, pend_ins = false, need_slow = false;
  uint32_t pend_v = 0;

// trace the queued candidates best-first (lowest cost, then leftmost), skipping those that cannot win
#define FLUSH_QUEUE()                                                                            \
  for (int it_ = 0; it_ < 4; ++it_) {                                                            \
    uint32_t v_ = 0;                                                                             \
    int which_ = -1;                                                                             \
    if (q0 && (!v_ || (q0 >> 16 & 0xFF) < (v_ >> 16 & 0xFF))) { v_ = q0; which_ = 0; }           \
    if (q1 && (!v_ || (q1 >> 16 & 0xFF) < (v_ >> 16 & 0xFF))) { v_ = q1; which_ = 1; }           \
    if (q2 && (!v_ || (q2 >> 16 & 0xFF) < (v_ >> 16 & 0xFF))) { v_ = q2; which_ = 2; }           \
    if (q3 && (!v_ || (q3 >> 16 & 0xFF) < (v_ >> 16 & 0xFF))) { v_ = q3; which_ = 3; }           \
    if (which_ < 0) break;                                                                       \
    if (which_ == 0) q0 = 0; else if (which_ == 1) q1 = 0; else if (which_ == 2) q2 = 0; else q3 = 0; \
    const int jc_ = Q_COL(v_), cc_ = Q_COST(v_);                                                 \
    int ub_ = Q_UB(v_, m);                                                                       \
    if (ub_ == m && !all_deletions_possible(eqt, read, m, jc_, cc_)) ub_ = m - 1;               \
    if (MAY_WIN(ub_, cc_, jc_)) {                                                                \
      int mt_, org_;                                                                             \
      if (!(cc_ == 1 && traceback_cost1(a, fc, m, jc_, mt_, org_))) {                            \
        if (DEFER) { need_slow = true; break; }                                                  \
        recompute(eqt, read, jc_, span, cb);                                                     \
        traceback(a, eqt, read, fc, cb, m, jc_, cc_, mt_, org_);                                 \
      }                                                                                          \
      TAKE_IF_BETTER(mt_, cc_, org_, jc_)                                                        \
    }                                                                                            \
  }

  for (int j = 1; j <= n; ++j) {
    eq = eqt[read[j - 1]];
    const int score_prev = score;
    pvp = vp;
    pvn = vn;
    uint32_t hp, hn;
    MYERS_STEP(eq, vp, vn, hp, hn)
    score += (hp & top) ? 1 : 0;
    score -= (hn & top) ? 1 : 0;
    if (pend) {
      // Rule 2: the candidate (m, j-1) was entered by an insertion from (m-1, j-1) and cell (m, j) is a
      // character match from that same cell: (m, j) has one more match and one error less, is itself a
      // candidate (row m of column j, or of the last column) and therefore beats (m, j-1).
      if (!(pend_ins && (eq & top))) {
        if (!q0) q0 = pend_v;
        else if (!q1) q1 = pend_v;
        else if (!q2) q2 = pend_v;
        else if (!q3) q3 = pend_v;
        else {  // queue full (low-complexity read): resolve what is queued, then go on
          FLUSH_QUEUE()
          if (DEFER && need_slow) break;
          q0 = pend_v;
        }
      }
      pend = false;
    }
    if (score <= acc_m && j < n) {  // row-m candidate (column n is handled with the last column)
      if (score == 0) {              // exact full adapter: cutadapt stops here
        have = true; b_m = m; b_c = 0; b_o = j - m; b_idx = j;
        stopped = true;
        break;
      }
      // first traceback step of cell (m, j) from the two cost columns at hand
      bool is_del = false, is_ins = false, is_match = true;
      if (!(eq & top)) {
        is_match = false;
        const int dm1_prev = score_prev - (int)((pvp >> (m - 1)) & 1u) + (int)((pvn >> (m - 1)) & 1u);  // D[m-1][j-1]
        const int dm1_cur = score - (int)((vp >> (m - 1)) & 1u) + (int)((vn >> (m - 1)) & 1u);          // D[m-1][j]
        const int cd = dm1_prev + 1, cdel = score_prev + 1, cins = dm1_cur + 1;
        if (!(cd <= cdel && cd <= cins)) {
          if (cins <= cdel) is_ins = true;
          else is_del = true;
        }
      }
      // Rule 1: entered by a deletion from (m, j-1): same matches and origin as that cell, one error
      // more; (m, j-1) is an accepted earlier candidate, so (m, j) can never win.
      if (!is_del) {
        pend = true;
        pend_ins = is_ins;
        pend_v = (uint32_t)j | ((uint32_t)score << 16) | (is_match ? (1u << 24) : 0u);
      }
    }
  }
  __syncwarp(lanes);
  const unsigned scan_lanes = __ballot_sync(lanes, !stopped && !need_slow);
  if (!stopped && !need_slow) {
    // last column: rows with cost 0 are pure diagonals; the others wait in rowmask / rowub
    uint32_t rowmask = 0, rowub = 0;  // rowub bit: first step of that cell is a match (up to i matches), else i - 1
    int c = 0, cprev = 0;             // D[i][n], D[i][n-1]
    for (int i = 1; i <= m; ++i) {
      const int c_up = c, cprev_up = cprev;  // row i - 1
      c += (int)((vp >> (i - 1)) & 1u) - (int)((vn >> (i - 1)) & 1u);
      cprev += (int)((pvp >> (i - 1)) & 1u) - (int)((pvn >> (i - 1)) & 1u);
      if (c > ad.acc[i]) continue;
      if (c == 0) {
        const int idx = (i == m) ? n : n + 1 + i;  // row m of column n precedes the last-column scan
        TAKE_IF_BETTER(i, 0, n - i, idx)
      } else {
        bool full = true;  // may reach i matches only if its first step is a match or a deletion
        if (!((eq >> (i - 1)) & 1u) && n >= 1) {
          const int cd = cprev_up + 1, cdel = cprev + 1, cins = c_up + 1;
          if (cd <= cdel && cd <= cins) full = false;  // mismatch
          else if (cins <= cdel) {
            // Rule 3: entered by an insertion from (i-1, n): same matches and origin as that cell, one
            // error more; when (i-1, n) is an accepted candidate of the same column (scanned earlier),
            // (i, n) can never win.
            if (ad.acc[i - 1] >= c - 1) continue;
            full = false;
          }
        }
        rowmask |= 1u << (i - 1);
        if (full) rowub |= 1u << (i - 1);
      }
    }
    __syncwarp(scan_lanes);
    FLUSH_QUEUE()
    bool have_cols = false;
    while (rowmask) {  // descending rows: the first traceback usually prunes the rest
      const int i = 32 - __clz(rowmask);
      rowmask &= ~(1u << (i - 1));
      const int ci = cell_cost(vp, vn, i);
      const int idx = (i == m) ? n : n + 1 + i;
      int ub = ((rowub >> (i - 1)) & 1u) ? i : i - 1;
      if (ub == i && !all_deletions_possible(eqt, read, i, n, ci)) ub = i - 1;
      if (MAY_WIN(ub, ci, idx)) {
        int mt, org;
        if (!(ci == 1 && traceback_cost1(a, fc, i, n, mt, org))) {
          if (DEFER) { need_slow = true; break; }
          if (!have_cols) {
            recompute(eqt, read, n, span, cb);
            have_cols = true;
          }
          traceback(a, eqt, read, fc, cb, i, n, ci, mt, org);
        }
        TAKE_IF_BETTER(mt, ci, org, idx)
      }
    }
  }
  if (DEFER && need_slow) return 2;
  if (!have) return 0;
  out.rstart = b_o;
  out.rstop = b_idx <= n ? b_idx : n;
  out.matches = b_m;
  out.errors = b_c;
  return 1;
}

// AdapterCutter._best_match: most matches, then fewer errors, first adapter wins ties.
// returns the index of the winning adapter, -1 for none, -2 when the search was deferred (DEFER only)
template <int MAXM, bool FAST, bool DEFER>
__device__ __noinline__ int best_match(const uint8_t *read, int n, Match &best, const FastCtx &fc) {
  int which = -1;
  for (int a = 0; a < c_p.n_adapters; ++a) {
    Match mt;
    if (FAST) {
      const int rc = locate_fast<DEFER>(a, read, n, fc, mt);
      if (rc == 2) return -2;
      if (rc == 0) continue;
    } else if (!locate<MAXM>(a, read, n, mt)) continue;
    if (which < 0 || mt.matches > best.matches || (mt.matches == best.matches && mt.errors < best.errors)) {
      best = mt;
      which = a;
    }
  }
  return which;
}

// returns true when the adapter search was deferred to the second pass (DEFER only)
template <int MAXM, bool FAST, bool DEFER>
__device__ __noinline__ bool apply_mod(int mi, const uint8_t *seq, const uint8_t *qual, int &start, int &stop, FastCtx &fc) {
  const int len = stop - start;
  switch (c_p.kind[mi]) {
    case MIRGE_MOD_NEXTSEQ:
      stop = start + nextseq_trim_index(seq + start, qual + start, len, c_p.a[mi], c_p.b[mi]);
      break;
    case MIRGE_MOD_QUALITY: {
      int s, e;
      quality_trim_index(qual + start, len, c_p.a[mi], c_p.b[mi], c_p.c[mi], s, e);
      stop = start + e;
      start = start + s;
      break;
    }
    case MIRGE_MOD_ADAPTER:
      for (int t = 0; t < c_p.times; ++t) {
        Match mt;
        fc.rbase = start;
        const int a = best_match<MAXM, FAST, DEFER>(seq + start, stop - start, mt, fc);
        if (a == -2) return true;
        if (a < 0) break;
        if (c_p.ad[a].where == 0) stop = start + mt.rstart;
        else start = start + mt.rstop;
      }
      break;
    case MIRGE_MOD_NEND:
      while (start < stop && seq[start] == 'N') ++start;
      while (stop > start && seq[stop - 1] == 'N') --stop;
      break;
    case MIRGE_MOD_CUT: {
      const int c = c_p.a[mi];
      if (c > 0) start += min(c, len);
      else stop = start + max(len + c, 0);
      break;
    }
    default: break;
  }
  return false;
}

__device__ __forceinline__ int find_sub(const uint8_t *s, int n, const uint8_t *t, int tl, int from) {
  for (int p = from; p + tl <= n; ++p) {
    int j = 0;
    while (j < tl && s[p + j] == t[j]) ++j;
    if (j == tl) return p;
  }
  return -1;
}

// byte p of the emitted text read[start:stop] + read[us:ue]
__device__ __forceinline__ uint32_t key_byte(const uint8_t *seq, int start, int l1, int us, int p) {
  return p < l1 ? seq[start + p] : seq[us + p - l1];
}

#define PACK_WORDS 8  // reads up to 128 bases keep their 2-bit text in registers for key emission

// Everything a thread does for record r once its bytes are addressable through B (shared-memory staging or
// the global stream).  PASS 0: single pass; PASS 1: first pass (reads whose adapter search needs cost
// columns are appended to d_slow and emit nothing); PASS 2: second pass over those reads.
template <int MAXM, bool FAST, int PASS>
__device__ __forceinline__ void process_record(const bool valid, const uint64_t r, const uint8_t *B, uint64_t nbytes,
                                               const uint32_t *__restrict__ line_start, ushort4 *__restrict__ win,
                                               uint32_t *__restrict__ key_off, uint32_t *__restrict__ keys, uint64_t keys_cap,
                                               unsigned long long *__restrict__ ctrl, uint32_t *__restrict__ d_slow, FastCtx &fc,
                                               uint32_t *ps_mine) {
  const int tid = threadIdx.x;
  bool is_slow = false;
  const int E = c_p.slots;
  // per-slot windows live in local memory (dynamic slot index keeps the code small)
  int w_start[MIRGE_MAX_MODS], w_stop[MIRGE_MAX_MODS], w_us[MIRGE_MAX_MODS], w_ue[MIRGE_MAX_MODS];
  uint32_t w_words[MIRGE_MAX_MODS];  // 0 = not kept
#pragma unroll 1
  for (int s = 0; s < E; ++s) { w_start[s] = w_stop[s] = w_us[s] = w_ue[s] = 0; w_words[s] = 0; }
  const uint8_t *seq = nullptr;
  uint32_t my_words = 0, my_kept = 0;
  bool fast_emit = FAST;
  if (valid) {
    const uint4 ls = *(const uint4 *)(line_start + 4 * r);
    const uint32_t nxt = line_start[4 * r + 4];
    int sl = (int)(ls.z - 1 - ls.y), ql = (int)(nxt - 1 - ls.w);
    if (sl > 0 && B[ls.y + sl - 1] == '\r') --sl;
    if (ql > 0 && B[ls.w + ql - 1] == '\r') --ql;
    const bool bad = B[ls.x] != '@' || B[ls.z] != '+' || sl != ql || sl > MIRGE_MAX_READ_LEN;
    // lanes that run the modifier pipeline together; used to re-converge them after every modifier,
    // whose data-dependent loops (quality scans, adapter search) otherwise leave the warp split
    const unsigned good_lanes = __ballot_sync(__activemask(), !bad);
    if (bad) {
      atomicOr(ctrl + 2, (sl > MIRGE_MAX_READ_LEN && sl == ql) ? 4ull : 1ull);
      atomicMax(ctrl + 3, ~(unsigned long long)r);
    } else {
      seq = B + ls.y;
      const uint8_t *qual = B + ls.w;
      int start = 0, stop = sl;
      if (FAST) {
        // 2-bit text of the whole read once, four bytes per step (SIMD-in-register), into shared memory:
        // used by the traceback jumps and by key emission
        fast_emit = sl <= 16 * PACK_WORDS;
        if (PASS == 2 && (uint64_t)ls.y + 16 * ((sl + 15) >> 4) + 8 > nbytes) fast_emit = false;  // no over-read past the stream
        if (fast_emit) {
          const uint32_t *ap = (const uint32_t *)((uintptr_t)seq & ~(uintptr_t)3);
          const uint32_t bs = ((uint32_t)(uintptr_t)seq & 3u) * 8u;
          const int nwords = (sl + 15) >> 4;
          uint32_t anyexc = 0, prev = ap[0];
          int k = 1;
#pragma unroll 1
          for (int w = 0; w < nwords; ++w) {
            uint32_t word = 0;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const uint32_t nx = ap[k++];
              const uint32_t quad = __funnelshift_r(prev, nx, bs);
              prev = nx;
              const int nb = sl - (16 * w + 4 * q4);  // bytes of this quad that belong to the read
              uint32_t ok = __vcmpeq4(quad, 0x41414141u) | __vcmpeq4(quad, 0x43434343u) | __vcmpeq4(quad, 0x47474747u) |
                            __vcmpeq4(quad, 0x54545454u);
              uint32_t codes = ((quad >> 1) ^ (quad >> 2)) & 0x03030303u;
              if (nb < 4) {
                const uint32_t tm = nb <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - nb)));
                codes &= tm;
                ok |= ~tm;
              }
              anyexc |= ~ok;
              word |= ((codes * 0x01041040u) >> 24) << (8 * q4);
            }
            ps_mine[w * TRIM_THREADS] = word;
          }
#pragma unroll 1
          for (int w = nwords; w < PACK_WORDS + 2; ++w) ps_mine[w * TRIM_THREADS] = 0;
          if (anyexc) fast_emit = false;
        }
        fc.jump_ok = fast_emit;
      }
      if (c_p.umi_mode == MIRGE_UMI_QIAGEN) {
#pragma unroll 1
        for (int mi = 0; mi < c_p.n_mods; ++mi) {
          if (!is_slow) is_slow = apply_mod<MAXM, FAST, PASS == 1>(mi, seq, qual, start, stop, fc);
          __syncwarp(good_lanes);
        }
        const int tl = stop - start, U = c_p.umi3;
        int us = 0, ue = 0;
        if (tl > 0) {
          const int first = find_sub(seq, sl, seq + start, tl, 0);
          const int after = first + tl;
          const int nx = find_sub(seq, sl, seq + start, tl, after);
          int seg_end = nx < 0 ? sl : nx;
          seg_end = min(seg_end, after + c_p.qia_len + U);
          ue = seg_end;
          us = (U != 0) ? max(seg_end - U, after) : after;
        }
        w_start[0] = start; w_stop[0] = stop; w_us[0] = us; w_ue[0] = ue;
        w_words[0] = (tl >= c_p.min_len) ? 1u : 0u;
        if (ue > us) fast_emit = false;  // concatenated key: generic packing
      } else {
#pragma unroll 1
        for (int mi = 0; mi < c_p.n_mods; ++mi) {
          if (!is_slow) is_slow = apply_mod<MAXM, FAST, PASS == 1>(mi, seq, qual, start, stop, fc);
          __syncwarp(good_lanes);
          if (E != 1 || mi == c_p.n_mods - 1) {
            const int slot = (E == 1) ? 0 : mi;
            int ln = stop - start;
            if (c_p.umi_mode == MIRGE_UMI_FLANKS) ln = max(ln - c_p.umi5 - c_p.umi3, 0);
            w_start[slot] = start; w_stop[slot] = stop; w_words[slot] = ln >= c_p.min_len;
          }
        }
      }
      if (PASS == 1 && is_slow) {
        // second pass will redo this read from scratch; emit nothing now
#pragma unroll 1
        for (int s = 0; s < E; ++s) { w_start[s] = w_stop[s] = w_us[s] = w_ue[s] = 0; w_words[s] = 0; }
      }
      // size of every kept key: header + payload + exceptions
#pragma unroll 1
      for (int s = 0; s < E; ++s) {
        if (!w_words[s]) continue;
        if (FAST && fast_emit) {
          w_words[s] = 1u + ((uint32_t)(w_stop[s] - w_start[s] + 15) >> 4);
        } else {
          const int l1 = w_stop[s] - w_start[s], len = l1 + (w_ue[s] - w_us[s]);
          uint32_t nexc = 0;
          for (int p = 0; p < len; ++p) nexc += base_code_exact(key_byte(seq, w_start[s], l1, w_us[s], p)) == 4u;
          w_words[s] = 1u + ((uint32_t)(len + 15) >> 4) + nexc;
        }
        my_words += w_words[s];
        ++my_kept;
      }
    }
  }
  // key space: one atomic per warp (warps finish independently, no CTA barrier)
  const int lane = tid & 31;
  uint32_t inc = my_words;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  const uint32_t warp_total = __shfl_sync(0xffffffffu, inc, 31);
  const uint32_t kept_warp = __reduce_add_sync(0xffffffffu, my_kept);
  unsigned long long warp_base = 0;
  if (lane == 0) {
    if (warp_total) warp_base = atomicAdd(ctrl + 0, (unsigned long long)warp_total);
    if (kept_warp) atomicAdd(ctrl + 1, (unsigned long long)kept_warp);
  }
  warp_base = __shfl_sync(0xffffffffu, warp_base, 0);
  const bool overflow = warp_base + warp_total > keys_cap || warp_base + warp_total > 0xFFFFFFF0ull;
  if (overflow && lane == 0 && warp_total) atomicOr(ctrl + 2, 2ull);
  if (PASS == 1) {  // queue the deferred reads for the second pass (one atomic per warp)
    const unsigned sm_ = __ballot_sync(0xffffffffu, is_slow);
    if (sm_) {
      unsigned long long base = 0;
      if (lane == __ffs(sm_) - 1) base = atomicAdd(ctrl + 5, (unsigned long long)__popc(sm_));
      base = __shfl_sync(0xffffffffu, base, __ffs(sm_) - 1);
      if (is_slow) d_slow[base + __popc(sm_ & ((1u << lane) - 1u))] = (uint32_t)r;
    }
  }
  if (!valid) return;
  uint32_t off = (uint32_t)warp_base + inc - my_words;
#pragma unroll 1
  for (int s = 0; s < E; ++s) {
    const uint64_t e = r * (uint64_t)E + s;
    win[e] = make_ushort4((unsigned short)w_start[s], (unsigned short)w_stop[s], (unsigned short)w_us[s], (unsigned short)w_ue[s]);
    if (!w_words[s] || overflow) {
      key_off[e] = 0xFFFFFFFFu;
      continue;
    }
    key_off[e] = off;
    const int l1 = w_stop[s] - w_start[s], len = l1 + (w_ue[s] - w_us[s]);
    const uint32_t npay = (uint32_t)(len + 15) >> 4;
    uint32_t *k = keys + off;
    if (FAST && fast_emit) {
      // key = 2-bit text of read[w_start : w_stop): a slice of the packed read in shared memory
      k[0] = (uint32_t)len;
      const int wi = w_start[s] >> 4;
      const uint32_t bs = 2u * (uint32_t)(w_start[s] & 15);
      uint32_t lo = ps_mine[wi * TRIM_THREADS];
#pragma unroll 1
      for (uint32_t w = 0; w < npay; ++w) {
        const uint32_t hi = ps_mine[(wi + w + 1) * TRIM_THREADS];
        uint32_t v = __funnelshift_r(lo, hi, bs);
        lo = hi;
        const int rem = len - 16 * (int)w;
        if (rem < 16) v &= (1u << (2 * rem)) - 1u;
        k[1 + w] = v;
      }
    } else {
      const uint32_t nexc = w_words[s] - 1u - npay;
      k[0] = (uint32_t)len | (nexc << 16);
      uint32_t word = 0, xi = 0;
      for (int p = 0; p < len; ++p) {
        const uint32_t ch = key_byte(seq, w_start[s], l1, w_us[s], p);
        const uint32_t code = base_code_exact(ch);
        if (code == 4u) k[1 + npay + xi++] = ((uint32_t)p << 8) | ch;
        else word |= code << (2 * (p & 15));
        if ((p & 15) == 15) { k[1 + (p >> 4)] = word; word = 0; }
      }
      if (len & 15) k[1 + (len >> 4)] = word;
    }
    off += w_words[s];
  }
}

// FAST = bit-parallel adapter search (locate_fast) + packed-read key emission; requires the CTA's span to be
// staged in shared memory, otherwise the batch is flagged (ctrl[2] bit 3) for the generic kernel.
template <int MAXM, bool FAST, int PASS>
__global__ void __launch_bounds__(TRIM_THREADS, (FAST && PASS == 1) ? 7 : 1)
trim_kernel(const uint8_t *__restrict__ fq, uint64_t nbytes, const uint32_t *__restrict__ line_start, uint64_t n_records,
            ushort4 *__restrict__ win, uint32_t *__restrict__ key_off, uint32_t *__restrict__ keys, uint64_t keys_cap,
            unsigned long long *__restrict__ ctrl, uint32_t smem_bytes, uint32_t *__restrict__ d_slow) {
  extern __shared__ uint4 smem4[];
  uint8_t *sbuf = (uint8_t *)smem4;
  const int tid = threadIdx.x;
  FastCtx fc;
  fc.s_eq = nullptr; fc.ps = nullptr; fc.jump_ok = false; fc.rbase = 0;
  uint32_t *ps_mine = nullptr;
  if (FAST) {
    // smem: [staging smem_bytes][eq tables n_adapters * 256 words][packed reads (PACK_WORDS + 2) * T words]
    uint32_t *eq = (uint32_t *)(sbuf + smem_bytes);
    for (int e = tid; e < c_p.n_adapters * 256; e += TRIM_THREADS) {
      const uint32_t code = base_code_upper((uint32_t)(e & 255));
      eq[e] = code < 4u ? (uint32_t)c_p.ad[e >> 8].peq[code] : 0u;
    }
    fc.s_eq = eq;
    ps_mine = eq + c_p.n_adapters * 256 + tid;
    fc.ps = ps_mine;
  }
  if (PASS == 2) {
    // second pass: the deferred reads, addressed in the global stream; whole warps iterate together
    __syncthreads();
    const unsigned long long n_slow = ctrl[5];
    const uint64_t first = (uint64_t)blockIdx.x * TRIM_THREADS + (tid & ~31);
    for (uint64_t base = first; base < n_slow; base += (uint64_t)gridDim.x * TRIM_THREADS) {
      const uint64_t idx = base + (tid & 31);
      const bool valid = idx < n_slow;
      const uint64_t r = valid ? d_slow[idx] : 0;
      fc.jump_ok = false;
      process_record<MAXM, FAST, 2>(valid, r, fq, nbytes, line_start, win, key_off, keys, keys_cap, ctrl, d_slow, fc, ps_mine);
      __syncwarp();
    }
    return;
  }
  const uint64_t r0 = (uint64_t)blockIdx.x * TRIM_THREADS;
  const uint64_t r = r0 + tid;
  const bool valid = r < n_records;
  const uint64_t r_end = min(r0 + (uint64_t)TRIM_THREADS, n_records);
  const uint32_t span_lo = line_start[4 * r0];
  const uint64_t span_hi = min((uint64_t)line_start[4 * r_end], nbytes);
  const uint32_t alo = span_lo & ~15u;
  const bool staged = (span_hi - alo) <= smem_bytes;
  if (FAST && !staged) {  // uniform per CTA
    if (PASS == 1) {
      // this group's bytes do not fit the staging buffer: hand all of its reads to the second pass
      const unsigned vm = __ballot_sync(0xffffffffu, valid);
      if (vm) {
        const int lane = tid & 31;
        unsigned long long base = 0;
        if (lane == __ffs(vm) - 1) base = atomicAdd(ctrl + 5, (unsigned long long)__popc(vm));
        base = __shfl_sync(0xffffffffu, base, __ffs(vm) - 1);
        if (valid) d_slow[base + __popc(vm & ((1u << lane) - 1u))] = (uint32_t)r;
      }
      return;
    }
    if (tid == 0) atomicOr(ctrl + 2, 8ull);
    return;
  }
  if (staged) {
    for (uint64_t o = (uint64_t)tid * 16; alo + o < span_hi; o += TRIM_THREADS * 16) {
      const uint64_t g = alo + o;
      if (g + 16 <= nbytes) {
        *(uint4 *)(sbuf + o) = ld_stream_u4(fq + g);
      } else {
        for (int b = 0; b < 16 && g + b < nbytes; ++b) sbuf[o + b] = fq[g + b];
      }
    }
  }
  __syncthreads();
  const uint8_t *B = staged ? (const uint8_t *)(sbuf - alo) : fq;  // B[absolute stream offset]
  process_record<MAXM, FAST, PASS>(valid, r, B, nbytes, line_start, win, key_off, keys, keys_cap, ctrl, d_slow, fc, ps_mine);
}

extern "C" int mirge_trim(mirge_ctx *ctx, const uint8_t *d_fastq, uint64_t nbytes, const uint32_t *d_line_start, uint64_t n_records,
                          uint16_t *d_win, uint32_t *d_key_off, uint32_t *d_keys, uint64_t keys_capacity_words,
                          uint64_t *d_trim_ctrl, uint32_t *d_slow, void *stream_) {
  if (!ctx) return MIRGE_ERR_ARG;
  if (!ctx->params_set) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "trim: mirge_set_trim_params has not been called");
  if (n_records == 0) return MIRGE_OK;
  if (!d_fastq || !d_line_start || !d_win || !d_key_off || !d_keys || !d_trim_ctrl) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "trim: null buffer");
  if (((uintptr_t)d_line_start & 15) || ((uintptr_t)d_win & 7)) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "trim: misaligned buffer");
  // same convention as the tokeniser: line_start offsets are relative to the 16-byte aligned stream
  {
    const uint32_t skew = (uint32_t)((uintptr_t)d_fastq & 15);
    d_fastq -= skew;
    nbytes += skew;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  MIRGE_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t avg = (nbytes + n_records - 1) / n_records;
  const bool fast = ctx->fast_ok && ctx->trim_mode == 0;
  // staging buffer: the bit-parallel kernel keeps it tight (7 CTAs per SM); a record group that does not fit
  // is handed to its second pass.  The generic kernel reads such groups straight from global memory.
  uint64_t want = fast ? avg * TRIM_THREADS * 33 / 32 + 512 : avg * TRIM_THREADS * 5 / 4 + 64;
  want = (want + 255) & ~255ull;
  if (want < 4096) want = 4096;
  if (want > 96 * 1024) want = 96 * 1024;
  const uint32_t smem = (uint32_t)want;
  const unsigned grid = (unsigned)((n_records + TRIM_THREADS - 1) / TRIM_THREADS);
  unsigned long long *ctrl = (unsigned long long *)d_trim_ctrl;
  ushort4 *win = (ushort4 *)d_win;
  if (fast) {
    if (!d_slow) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "trim: d_slow scratch (u32 per record) is required");
    const size_t extra = (size_t)ctx->params.n_adapters * 1024 + (size_t)(PACK_WORDS + 2) * TRIM_THREADS * 4;
    MIRGE_CUDA(ctx, cudaFuncSetAttribute(trim_kernel<32, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    // pass 1: every read; adapter searches that need cost columns are deferred to the list d_slow
    trim_kernel<32, true, 1><<<grid, TRIM_THREADS, smem + extra, stream>>>(d_fastq, nbytes, d_line_start, n_records, win, d_key_off,
                                                                          d_keys, keys_capacity_words, ctrl, smem, d_slow);
    MIRGE_LAUNCH_CHECK(ctx, "trim_kernel(pass 1)");
    // pass 2: the deferred reads with full warps (count read on the device; grid-stride over the list)
    unsigned grid2 = grid / 16 + 1;
    if (grid2 > (unsigned)ctx->sm_count * 8) grid2 = (unsigned)ctx->sm_count * 8;
    trim_kernel<32, true, 2><<<grid2, TRIM_THREADS, extra, stream>>>(d_fastq, nbytes, d_line_start, n_records, win, d_key_off, d_keys,
                                                                   keys_capacity_words, ctrl, 0u, d_slow);
  } else if (ctx->max_adapter_len <= 32) {
    MIRGE_CUDA(ctx, cudaFuncSetAttribute(trim_kernel<32, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    trim_kernel<32, false, 0><<<grid, TRIM_THREADS, smem, stream>>>(d_fastq, nbytes, d_line_start, n_records, win, d_key_off, d_keys,
                                                                    keys_capacity_words, ctrl, smem, nullptr);
  } else {
    MIRGE_CUDA(ctx, cudaFuncSetAttribute(trim_kernel<64, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    trim_kernel<64, false, 0><<<grid, TRIM_THREADS, smem, stream>>>(d_fastq, nbytes, d_line_start, n_records, win, d_key_off, d_keys,
                                                                    keys_capacity_words, ctrl, smem, nullptr);
  }
  MIRGE_LAUNCH_CHECK(ctx, "trim_kernel");
  return MIRGE_OK;
}

extern "C" int mirge_trim_mode(mirge_ctx *ctx, int mode) {
  if (!ctx || mode < 0 || mode > 1) return MIRGE_ERR_ARG;
  ctx->trim_mode = mode;
  return MIRGE_OK;
}
