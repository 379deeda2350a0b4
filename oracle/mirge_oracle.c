/*
 * mirge_oracle.c -- fast CPU restatement (plain C + OpenMP) of miRge3.0's digest -> collapse ->
 * annotate hot path.  TEST INFRASTRUCTURE ONLY: built into oracle/_build/libmirge_oracle.so and
 * loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as
 * the checker / CPU baseline; the product (mirge3.0_b200/) never links or calls it.
 *
 * PARITY UNPINNED: cutadapt/dnaio/bowtie are third-party, not vendored in /root/reference and not
 * installable here; this file follows oracle/pyoracle.py (the readable restatement, pinned on the
 * reference's documented known-answer reads) function by function and is checked against it in
 * tests/test_oracle_c.py.  Reference call sites (relative to /root/reference):
 *   mirge/libs/digest.py:59-101 (pipeline), :320-375 (worker + key emission), :141-208 (collapse,
 *   UMI level), mirge/libs/manifoldAlign.py:84-135 (round policies and selection).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/mirge_b200.h"

#define MAXM MIRGE_MAX_ADAPTER_LEN

/* ------------------------------------------------------------------ FASTQ lines (dnaio) ---- */

/* Fills line_start (if non-NULL, capacity 4*n_rec+1) like mirge_line_index; returns the number of
 * complete records or a negative MIRGE_ERR_FORMAT.  EOF rule: a last line without '\n' counts. */
int64_t oracle_line_index(const uint8_t *fq, uint64_t n, uint32_t *line_start, uint64_t cap_records) {
  uint64_t nl = 0, pos = 0;
  if (line_start && cap_records * 4 + 1 > 0) line_start[0] = 0;
  while (pos < n) {
    const uint8_t *q = memchr(fq + pos, '\n', n - pos);
    uint64_t next;
    if (!q) next = n + 1; /* virtual newline at EOF */
    else next = (uint64_t)(q - fq) + 1;
    nl++;
    if (line_start && nl <= cap_records * 4) line_start[nl] = (uint32_t)next;
    pos = next;
  }
  if (nl % 4 != 0) return MIRGE_ERR_FORMAT;
  return (int64_t)(nl / 4);
}

/* ------------------------------------------------------------------ quality trimming -------- */

static int nextseq_trim_index(const uint8_t *seq, const uint8_t *qual, int len, int cutoff, int base) {
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

static void quality_trim_index(const uint8_t *qual, int len, int q5, int q3, int base, int *pstart, int *pstop) {
  int s = 0, max_qual = 0, start = 0, stop = len;
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
  *pstart = start; *pstop = stop;
}

/* ------------------------------------------------------------------ adapter alignment ------- */

static inline int upper(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }
static inline int acgt_mask(int c) {
  switch (upper(c)) { case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': case 'U': return 8; default: return 0; }
}
/* the read's characters as IUPAC sets (--match-read-wildcards; _align.pyx IUPAC_TABLE) */
static inline int iupac_mask(int c) {
  switch (upper(c)) {
    case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': case 'U': return 8;
    case 'R': return 5; case 'Y': return 10; case 'S': return 6; case 'W': return 9; case 'K': return 12; case 'M': return 3;
    case 'B': return 14; case 'D': return 13; case 'H': return 11; case 'V': return 7; case 'N': return 15; default: return 0;
  }
}

typedef struct { int astart, astop, rstart, rstop, matches, errors; } match_t;

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* cutadapt Aligner.locate (pyoracle.locate). Returns 1 if a match was found. */
static int locate(const mirge_adapter *ad, const uint8_t *read, int n, match_t *out, int compat) {
  /* merit carried by a cell: matches (cutadapt 2.x-3.x) or the score of cutadapt >= 4 (match +1, mismatch -1, indel -2) */
  const int w_mis = compat == MIRGE_COMPAT_CUTADAPT4 ? -1 : 0, w_indel = compat == MIRGE_COMPAT_CUTADAPT4 ? -2 : 0;
  int m = ad->m;
  int cost[MAXM + 1], origin[MAXM + 1], matches[MAXM + 1];
  /* cutadapt's Where flags (pyoracle.WHERE_FLAGS): which ends of the alignment are free */
  const int w = ad->where;
  const int start_in_ref = w == MIRGE_WHERE_FRONT || w == MIRGE_WHERE_FRONT_NOT_INTERNAL;
  const int stop_in_ref = w == MIRGE_WHERE_BACK || w == MIRGE_WHERE_BACK_NOT_INTERNAL;
  const int start_in_query = w == MIRGE_WHERE_BACK || w == MIRGE_WHERE_FRONT || w == MIRGE_WHERE_SUFFIX || w == MIRGE_WHERE_BACK_NOT_INTERNAL;
  const int stop_in_query = w == MIRGE_WHERE_BACK || w == MIRGE_WHERE_FRONT || w == MIRGE_WHERE_PREFIX || w == MIRGE_WHERE_FRONT_NOT_INTERNAL;
  int ic = ad->indel_cost;
  int k = ad->k;
  /* an anchored start cannot use more than m + k read bases, an anchored end only the last m + k */
  const int max_n = start_in_query ? n : imin(n, m + k);
  const int min_n = stop_in_query ? 0 : imax(0, n - m - k);
  for (int i = 0; i <= m; ++i) {
    if (!start_in_ref && !start_in_query) { cost[i] = imax(i, min_n) * ic; origin[i] = 0; }
    else if (start_in_ref && !start_in_query) { cost[i] = min_n * ic; origin[i] = imin(0, min_n - i); }
    else if (!start_in_ref && start_in_query) { cost[i] = i * ic; origin[i] = imax(0, min_n - i); }
    else { cost[i] = imin(i, min_n) * ic; origin[i] = min_n - i; }
    matches[i] = start_in_ref ? 0 : i * w_indel;
  }
  int best_cost = m + n, best_origin = 0, best_matches = compat == MIRGE_COMPAT_CUTADAPT4 ? -(1 << 30) : 0, best_ref_stop = m,
      best_query_stop = n;
  int stopped = 0;
  for (int j = min_n + 1; j <= max_n; ++j) {
    int dc = cost[0], dor = origin[0], dm = matches[0];
    if (start_in_query) origin[0] = j;
    else { cost[0] = j * ic; matches[0] = j * w_indel; }
    /* without any wildcards _align.pyx compares the characters themselves: a U in the read is a T only through its
     * translation tables (adapter or read wildcards active) */
    int rc = ad->wildcard_read ? iupac_mask(read[j - 1]) : acgt_mask(read[j - 1]);
    if (!ad->wildcard_read && !ad->wildcard_ref && upper(read[j - 1]) == 'U') rc = 0;
    for (int i = 1; i <= m; ++i) {
      int c, o, mt;
      if (ad->mask[i - 1] & rc) { c = dc; o = dor; mt = dm + 1; }
      else {
        int cd = dc + 1, cdel = cost[i] + ic, cins = cost[i - 1] + ic;
        if (cd <= cdel && cd <= cins) { c = cd; o = dor; mt = dm + w_mis; }
        else if (cins <= cdel) { c = cins; o = origin[i - 1]; mt = matches[i - 1] + w_indel; }
        else { c = cdel; o = origin[i]; mt = matches[i] + w_indel; }
      }
      dc = cost[i]; dor = origin[i]; dm = matches[i];
      cost[i] = c; origin[i] = o; matches[i] = mt;
    }
    if (cost[m] <= k && stop_in_query) {
      int length = m + (origin[m] < 0 ? origin[m] : 0);
      int eff = length;
      if (ad->wildcard_ref) eff = (length < m) ? length - (ad->n_counts[m] - ad->n_counts[m - length]) : ad->effective_length;
      int c = cost[m], mt = matches[m];
      if (length >= ad->min_overlap && c <= ad->max_err[eff] && (mt > best_matches || (mt == best_matches && c < best_cost))) {
        best_matches = mt; best_cost = c; best_origin = origin[m]; best_ref_stop = m; best_query_stop = j;
        if (c == 0 && mt == m) { stopped = 1; break; }
      }
    }
  }
  if (!stopped && max_n == n) {
    int first_i = stop_in_ref ? 0 : m;
    for (int i = first_i; i <= m; ++i) {
      int length = i + (origin[i] < 0 ? origin[i] : 0);
      int c = cost[i], mt = matches[i], eff = length;
      if (ad->wildcard_ref) {
        if (length < m) { int ref_start = origin[i] < 0 ? -origin[i] : 0; eff = length - (ad->n_counts[i] - ad->n_counts[ref_start]); }
        else eff = ad->effective_length;
      }
      if (length >= ad->min_overlap && eff >= 0 && c <= ad->max_err[eff] && (mt > best_matches || (mt == best_matches && c < best_cost))) {
        best_matches = mt; best_cost = c; best_origin = origin[i]; best_ref_stop = i; best_query_stop = n;
      }
    }
  }
  if (best_cost == m + n) return 0;
  out->astart = best_origin >= 0 ? 0 : -best_origin;
  out->rstart = best_origin >= 0 ? best_origin : 0;
  out->astop = best_ref_stop; out->rstop = best_query_stop; out->matches = best_matches; out->errors = best_cost;
  return 1;
}

/* Adapter.match_to: when the adapter has no wildcards, an exact comparison on upper(read) first -- str.find for the
 * plain forms, startswith / endswith for the anchored ones, none for the non-internal ones. */
static int match_to(const mirge_adapter *ad, const uint8_t *read, int n, match_t *out, int compat) {
  int m = ad->m;
  if (!ad->wildcard_ref && ad->where <= MIRGE_WHERE_PREFIX) {
    int lo = 0, hi = n - m;
    if (ad->where == MIRGE_WHERE_PREFIX) hi = imin(hi, 0);
    if (ad->where == MIRGE_WHERE_SUFFIX) lo = imax(n - m, 0);
    for (int p = lo; p <= hi; ++p) {
      int j = 0;
      while (j < m && upper(read[p + j]) == ad->ascii[j]) ++j;
      if (j == m) { out->astart = 0; out->astop = m; out->rstart = p; out->rstop = p + m; out->matches = m; out->errors = 0; return 1; }
    }
  }
  return locate(ad, read, n, out, compat);
}

static int best_match(const mirge_trim_params *p, const uint8_t *read, int n, match_t *best) {
  int have = 0, which = -1;
  for (int a = 0; a < p->n_adapters; ++a) {
    match_t mt;
    int link = p->adapters[a].link;
    if (link == MIRGE_LINK_BACK_HALF) continue;
    if (link != 0) {
      /* cutadapt LinkedAdapter.match_to: the 3' half is searched in read[front.rstop:]; a missing half ends the search
       * unless it is optional (-g "A...B": both required; -a "A...B": the 3' half optional), and a pair needs its 5' half
       * or, failing that, its 3' half.
       * The match of a pair: rstart = first base kept, rstop = end of what is kept, matches / errors = sums. */
      match_t f, b;
      const int have_f = match_to(&p->adapters[a], read, n, &f, p->compat);
      if (!have_f && !(link & MIRGE_LINK_FRONT_OPTIONAL)) continue;
      const int rest = have_f ? f.rstop : 0;
      const int have_b = match_to(&p->adapters[(link & MIRGE_LINK_INDEX_MASK) - 1], read + rest, n - rest, &b, p->compat);
      if (!have_b && (!(link & MIRGE_LINK_BACK_OPTIONAL) || !have_f)) continue;
      memset(&mt, 0, sizeof(mt));
      mt.rstart = rest; mt.rstop = have_b ? rest + b.rstart : n;
      mt.matches = (have_f ? f.matches : 0) + (have_b ? b.matches : 0); mt.errors = (have_f ? f.errors : 0) + (have_b ? b.errors : 0);
    } else if (!match_to(&p->adapters[a], read, n, &mt, p->compat)) continue;
    if (!have || mt.matches > best->matches || (mt.matches == best->matches && mt.errors < best->errors)) { *best = mt; have = 1; which = a; }
  }
  return which;
}

/* ------------------------------------------------------------------ per-read pipeline ------- */

static void apply_mod(const mirge_trim_params *p, int mi, const uint8_t *seq, const uint8_t *qual, int *pstart, int *pstop) {
  int start = *pstart, stop = *pstop, len = stop - start;
  switch (p->mod_kind[mi]) {
    case MIRGE_MOD_NEXTSEQ:
      stop = start + nextseq_trim_index(seq + start, qual + start, len, p->mod_a[mi], p->mod_b[mi]);
      break;
    case MIRGE_MOD_QUALITY: {
      int s, e; quality_trim_index(qual + start, len, p->mod_a[mi], p->mod_b[mi], p->mod_c[mi], &s, &e);
      stop = start + e; start = start + s; break; }
    case MIRGE_MOD_ADAPTER:
      for (int t = 0; t < p->times; ++t) {
        match_t mt; int a = best_match(p, seq + start, stop - start, &mt);
        if (a < 0) break;
        if (p->adapters[a].link != 0) { stop = start + mt.rstop; start = start + mt.rstart; }
        else if (!MIRGE_WHERE_IS_FRONT(p->adapters[a].where)) stop = start + mt.rstart; /* 3' forms */
        else start = start + mt.rstop;                                 /* 5' forms */
      }
      break;
    case MIRGE_MOD_NEND:
      while (start < stop && seq[start] == 'N') ++start;
      while (stop > start && seq[stop - 1] == 'N') --stop;
      break;
    case MIRGE_MOD_CUT: {
      int c = p->mod_a[mi];
      if (c > 0) start += (c < len ? c : len);
      else { int nl = len + c; stop = start + (nl > 0 ? nl : 0); }
      break; }
    default: break;
  }
  *pstart = start; *pstop = stop;
}

static int find_sub(const uint8_t *s, int n, const uint8_t *t, int tl, int from) {
  for (int p = from; p + tl <= n; ++p) if (memcmp(s + p, t, (size_t)tl) == 0) return p;
  return -1;
}

int oracle_trim_slots(const mirge_trim_params *p) {
  return (p->umi_mode != MIRGE_UMI_QIAGEN && p->count_mode == MIRGE_COUNT_HEAD) ? p->n_mods : 1;
}

/* digest.py:325-373 for one read: win[4*s..] and kept[s] for each emission slot. */
static void digest_read(const mirge_trim_params *p, const uint8_t *seq, const uint8_t *qual, int len, uint16_t *win, uint8_t *kept) {
  int start = 0, stop = len;
  int E = oracle_trim_slots(p);
  if (p->umi_mode == MIRGE_UMI_QIAGEN) {
    for (int mi = 0; mi < p->n_mods; ++mi) apply_mod(p, mi, seq, qual, &start, &stop);
    int tl = stop - start, U = p->umi3, us = 0, ue = 0;
    if (tl > 0) {
      int first = find_sub(seq, len, seq + start, tl, 0);
      int after = first + tl;
      int nxt = find_sub(seq, len, seq + start, tl, after);
      int seg_end = nxt < 0 ? len : nxt;
      if (seg_end > after + p->qia_adapter_len + U) seg_end = after + p->qia_adapter_len + U;
      ue = seg_end;
      us = (U != 0) ? (seg_end - U > after ? seg_end - U : after) : after;
    }
    win[0] = (uint16_t)start; win[1] = (uint16_t)stop; win[2] = (uint16_t)us; win[3] = (uint16_t)ue;
    kept[0] = tl >= p->min_len;
    return;
  }
  for (int mi = 0; mi < p->n_mods; ++mi) {
    apply_mod(p, mi, seq, qual, &start, &stop);
    int slot = (E == 1) ? 0 : mi;
    if (E == 1 && mi != p->n_mods - 1) continue;
    int ln = stop - start;
    if (p->umi_mode == MIRGE_UMI_FLANKS) { ln = ln - p->umi5 - p->umi3; if (ln < 0) ln = 0; }
    win[4 * slot] = (uint16_t)start; win[4 * slot + 1] = (uint16_t)stop; win[4 * slot + 2] = 0; win[4 * slot + 3] = 0;
    kept[slot] = ln >= p->min_len;
  }
  if (p->n_mods == 0) { win[0] = 0; win[1] = (uint16_t)len; win[2] = win[3] = 0; kept[0] = 0; }
}

static inline int line_len(const uint8_t *fq, const uint32_t *ls, uint64_t li) {
  int l = (int)(ls[li + 1] - 1 - ls[li]);
  if (l > 0 && fq[ls[li] + l - 1] == '\r') --l;
  return l;
}

/* Validate + trim every record. Returns n_records or MIRGE_ERR_FORMAT. win: u16[n*E*4], kept: u8[n*E]. */
int64_t oracle_trim(const uint8_t *fq, uint64_t n, const mirge_trim_params *p, const uint32_t *line_start,
                    uint64_t n_records, uint16_t *win, uint8_t *kept, int nthreads) {
  int E = oracle_trim_slots(p);
  int bad = 0;
  (void)n;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1) reduction(| : bad)
  for (int64_t r = 0; r < (int64_t)n_records; ++r) {
    const uint32_t *ls = line_start + 4 * r;
    int sl = line_len(fq, line_start, 4 * (uint64_t)r + 1), ql = line_len(fq, line_start, 4 * (uint64_t)r + 3);
    if (fq[ls[0]] != '@' || fq[ls[2]] != '+' || sl != ql) { bad |= 1; continue; }
    digest_read(p, fq + ls[1], fq + ls[3], sl, win + (size_t)r * E * 4, kept + (size_t)r * E);
  }
  if (bad) return MIRGE_ERR_FORMAT;
  return (int64_t)n_records;
}

/* ------------------------------------------------------------------ collapse (string map) --- */

typedef struct { uint64_t hash; uint64_t off; uint32_t len; uint64_t count; } ent_t;
typedef struct {
  ent_t *e; uint64_t cap, n;
  uint8_t *arena; uint64_t arena_cap, arena_n;
  uint64_t total; /* sum of counts added */
} tab_t;

static uint64_t hash_bytes(const uint8_t *s, uint32_t len) {
  uint64_t h = 1469598103934665603ull;
  for (uint32_t i = 0; i < len; ++i) { h ^= s[i]; h *= 1099511628211ull; }
  h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
  return h | 1;
}

static tab_t *tab_new(uint64_t cap_hint) {
  tab_t *t = calloc(1, sizeof(tab_t));
  uint64_t cap = 1024; while (cap < cap_hint * 2) cap <<= 1;
  t->cap = cap; t->e = calloc(cap, sizeof(ent_t));
  t->arena_cap = 1 << 16; t->arena = malloc(t->arena_cap);
  return t;
}
static void tab_grow(tab_t *t) {
  uint64_t ncap = t->cap * 2; ent_t *ne = calloc(ncap, sizeof(ent_t));
  for (uint64_t i = 0; i < t->cap; ++i) if (t->e[i].hash) {
    uint64_t j = t->e[i].hash & (ncap - 1);
    while (ne[j].hash) j = (j + 1) & (ncap - 1);
    ne[j] = t->e[i];
  }
  free(t->e); t->e = ne; t->cap = ncap;
}
static void tab_add(tab_t *t, const uint8_t *a, uint32_t la, const uint8_t *b, uint32_t lb, uint64_t c) {
  uint8_t tmp[2 * MIRGE_MAX_READ_LEN + 64];
  const uint8_t *s = a; uint32_t len = la;
  if (lb) { memcpy(tmp, a, la); memcpy(tmp + la, b, lb); s = tmp; len = la + lb; }
  uint64_t h = hash_bytes(s, len);
  if ((t->n + 1) * 2 > t->cap) tab_grow(t);
  uint64_t j = h & (t->cap - 1);
  while (t->e[j].hash) {
    if (t->e[j].hash == h && t->e[j].len == len && memcmp(t->arena + t->e[j].off, s, len) == 0) { t->e[j].count += c; t->total += c; return; }
    j = (j + 1) & (t->cap - 1);
  }
  if (t->arena_n + len > t->arena_cap) { while (t->arena_n + len > t->arena_cap) t->arena_cap *= 2; t->arena = realloc(t->arena, t->arena_cap); }
  memcpy(t->arena + t->arena_n, s, len);
  t->e[j].hash = h; t->e[j].off = t->arena_n; t->e[j].len = len; t->e[j].count = c;
  t->arena_n += len; t->n++; t->total += c;
}
void oracle_table_free(void *h) { tab_t *t = h; if (!t) return; free(t->e); free(t->arena); free(t); }
uint64_t oracle_table_size(void *h) { return ((tab_t *)h)->n; }
uint64_t oracle_table_bytes(void *h) { return ((tab_t *)h)->arena_n; }
uint64_t oracle_table_total(void *h) { return ((tab_t *)h)->total; }
/* keys concatenated in table order; key_off[n+1]; counts[n] */
void oracle_table_export(void *h, uint8_t *keys, uint64_t *key_off, uint64_t *counts) {
  tab_t *t = h; uint64_t k = 0, o = 0;
  for (uint64_t i = 0; i < t->cap; ++i) if (t->e[i].hash) {
    memcpy(keys + o, t->arena + t->e[i].off, t->e[i].len);
    key_off[k] = o; counts[k] = t->e[i].count; o += t->e[i].len; ++k;
  }
  key_off[k] = o;
}

/* completeDict of one sample (digest.py:141-163): trim every record, collapse emitted keys.
 * Per-thread maps (the worker dicts) merged by the parent. Returns a table handle or NULL. */
void *oracle_digest_collapse(const uint8_t *fq, uint64_t n, const mirge_trim_params *p, const uint32_t *line_start,
                             uint64_t n_records, int nthreads, int64_t *status) {
  int E = oracle_trim_slots(p);
  if (nthreads < 1) nthreads = 1;
  tab_t **parts = calloc((size_t)nthreads, sizeof(tab_t *));
  int bad = 0;
  (void)n;
#pragma omp parallel num_threads(nthreads) reduction(| : bad)
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    int tid = 0, nt = 1;
#endif
    tab_t *t = tab_new(n_records / (uint64_t)nt / 4 + 16);
    parts[tid] = t;
    uint64_t r0 = n_records * (uint64_t)tid / (uint64_t)nt, r1 = n_records * (uint64_t)(tid + 1) / (uint64_t)nt;
    uint16_t win[4 * MIRGE_MAX_MODS]; uint8_t kept[MIRGE_MAX_MODS];
    for (uint64_t r = r0; r < r1; ++r) {
      const uint32_t *ls = line_start + 4 * r;
      int sl = line_len(fq, line_start, 4 * r + 1), ql = line_len(fq, line_start, 4 * r + 3);
      if (fq[ls[0]] != '@' || fq[ls[2]] != '+' || sl != ql) { bad |= 1; continue; }
      const uint8_t *seq = fq + ls[1];
      digest_read(p, seq, fq + ls[3], sl, win, kept);
      for (int s = 0; s < E; ++s) if (kept[s])
        tab_add(t, seq + win[4 * s], (uint32_t)(win[4 * s + 1] - win[4 * s]), seq + win[4 * s + 2], (uint32_t)(win[4 * s + 3] - win[4 * s + 2]), 1);
    }
  }
  tab_t *out = parts[0];
  for (int i = 1; i < nthreads; ++i) if (parts[i]) {
    tab_t *t = parts[i];
    for (uint64_t j = 0; j < t->cap; ++j) if (t->e[j].hash) tab_add(out, t->arena + t->e[j].off, t->e[j].len, NULL, 0, t->e[j].count);
    oracle_table_free(t);
  }
  free(parts);
  *status = bad ? MIRGE_ERR_FORMAT : 0;
  return out;
}

/* UMI second level (digest.py:164-205) */
void *oracle_umi_collapse(void *first, int f, int b, int min_len, int dedup) {
  tab_t *t = first; tab_t *o = tab_new(t->n + 16);
  for (uint64_t j = 0; j < t->cap; ++j) if (t->e[j].hash) {
    int len = (int)t->e[j].len, cl = len - f - b; if (cl < 0) cl = 0;
    if (cl >= min_len) tab_add(o, t->arena + t->e[j].off + (cl > 0 ? f : 0), (uint32_t)cl, NULL, 0, dedup ? 1 : t->e[j].count);
  }
  return o;
}

/* ------------------------------------------------------------------ annotation (bowtie) ----- */

static inline int base_code(int c) { switch (upper(c)) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; } }

/* Valid-hit search by exhaustive scan (the trustworthy definition; pyoracle.hits + canonical_pick).
 * refs: concatenated upper-case text, ref_off[n_refs+1]. Returns packed hit or MIRGE_NO_HIT. */
static uint64_t scan_query(const uint8_t *q, int L, const uint8_t *refs, const uint32_t *ref_off, uint32_t n_refs, const mirge_round_policy *pol) {
  uint64_t best = MIRGE_NO_HIT;
  if (L <= 0) return best;
  int seed = pol->seed_len == 0 ? L : (pol->seed_len < L ? pol->seed_len : L);
  uint8_t qc[MIRGE_MAX_READ_LEN];
  for (int j = 0; j < L; ++j) qc[j] = (uint8_t)base_code(q[j]);
  for (uint32_t r = 0; r < n_refs; ++r) {
    int rl = (int)(ref_off[r + 1] - ref_off[r]);
    const uint8_t *ref = refs + ref_off[r];
    for (int off = 0; off + L <= rl; ++off) {
      int mm = 0, smm = 0, ok = 1;
      for (int j = 0; j < L; ++j) {
        int rc = base_code(ref[off + j]);
        if (rc == 4) { ok = 0; break; }
        if (qc[j] != rc) { ++mm; if (j < seed) ++smm; if (mm > pol->total_mm || smm > pol->seed_mm) { ok = 0; break; } }
      }
      if (ok) { uint64_t h = ((uint64_t)mm << 56) | ((uint64_t)r << 28) | (uint64_t)off; if (h < best) best = h; }
    }
  }
  return best;
}

/* Query rewrite of a round (manifoldAlign.py:118-126 and the -5/-3 trimming of round 8).
 * Returns the query length or -1 when the sequence is not submitted. *qstart receives the offset. */
static int round_query(const uint8_t *s, int len, const mirge_round_policy *pol, int *qstart) {
  int a = 0, b = len;
  if (pol->strip_polyT) {
    int t = len; while (t > 0 && s[t - 1] == 'T') --t;
    if (len - t < 3) return -1;
    b = t;
  }
  a += pol->trim5; b -= pol->trim3;
  if (b < a) b = a;
  *qstart = a; return b - a;
}

/* One round over n sequences (keys concatenated, key_off[n+1]). annot_round[i] (0xFF = none) and
 * hit[i] are updated in place following manifoldAlign.py:90-135. */
void oracle_annotate_round(const uint8_t *keys, const uint64_t *key_off, uint64_t n, const uint8_t *refs,
                           const uint32_t *ref_off, uint32_t n_refs, const mirge_round_policy *pol,
                           uint8_t *annot_round, uint64_t *hit, int nthreads) {
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    const uint8_t *s = keys + key_off[i]; int len = (int)(key_off[i + 1] - key_off[i]);
    if (pol->select == MIRGE_SELECT_LEN_LT26) { if (!(len < 26)) continue; }
    else if (pol->select == MIRGE_SELECT_LEN_GT25) { if (!(len > 25)) continue; }
    else if (annot_round[i] != 0xFF) continue;
    int qs = 0, ql = round_query(s, len, pol, &qs);
    if (ql < 0) continue;
    uint64_t h = scan_query(s + qs, ql, refs, ref_off, n_refs, pol);
    if (h != MIRGE_NO_HIT) { annot_round[i] = (uint8_t)pol->round; hit[i] = h; }
  }
}

/* ------------------------------------------------------------------ indexed search (CPU port) -
 * Same valid-hit definition as scan_query, but candidates come from a sorted 16-mer index of the
 * library and pigeonhole seeds, the way an FM-index aligner avoids scanning the text.  Used as the
 * CPU baseline on libraries where the exhaustive scan is impractical (mRNA); checked against
 * scan_query in tests/test_oracle_c.py. */
typedef struct {
  const uint8_t *text; const uint32_t *ref_off; uint32_t n_refs; uint64_t n_bases;
  uint32_t *kmer, *pos; uint64_t n; uint32_t *bucket; int bucket_bits;
} oidx_t;

void oracle_index_free(void *h) { oidx_t *x = h; if (!x) return; free(x->kmer); free(x->pos); free(x->bucket); free(x); }

void *oracle_index_build(const uint8_t *text, const uint32_t *ref_off, uint32_t n_refs, int nthreads) {
  oidx_t *x = calloc(1, sizeof(oidx_t));
  x->text = text; x->ref_off = ref_off; x->n_refs = n_refs; x->n_bases = ref_off[n_refs];
  uint64_t nb = x->n_bases;
  uint32_t *km = malloc((nb + 1) * 4), *ps = malloc((nb + 1) * 4);
  uint64_t n = 0;
  (void)nthreads;
  for (uint32_t r = 0; r < n_refs; ++r) {
    uint32_t lo = ref_off[r], hi = ref_off[r + 1];
    for (uint32_t p = lo; p < hi; ++p) {
      uint32_t k = 0; int v = 0;
      for (int i = 0; i < 16 && p + i < hi; ++i) { int c = base_code(text[p + i]); if (c == 4) break; k |= (uint32_t)c << (2 * (15 - i)); ++v; }
      if (v >= 4) { km[n] = k; ps[n] = p; ++n; }
    }
  }
  /* stable LSD radix sort on the k-mer (positions are ascending already) */
  uint32_t *km2 = malloc((n + 1) * 4), *ps2 = malloc((n + 1) * 4);
  for (int pass = 0; pass < 2; ++pass) {
    uint64_t *cnt = calloc(65537, sizeof(uint64_t));
    int sh = 16 * pass;
    for (uint64_t i = 0; i < n; ++i) cnt[((km[i] >> sh) & 0xFFFF) + 1]++;
    for (int b = 0; b < 65536; ++b) cnt[b + 1] += cnt[b];
    for (uint64_t i = 0; i < n; ++i) { uint64_t d = cnt[(km[i] >> sh) & 0xFFFF]++; km2[d] = km[i]; ps2[d] = ps[i]; }
    free(cnt);
    uint32_t *t = km; km = km2; km2 = t; t = ps; ps = ps2; ps2 = t;
  }
  free(km2); free(ps2);
  x->kmer = km; x->pos = ps; x->n = n;
  int bb = 4; while ((1ull << (bb + 2)) < n && bb < 24) ++bb;
  x->bucket_bits = bb;
  x->bucket = malloc(((1ull << bb) + 1) * 4);
  uint64_t j = 0;
  for (uint64_t b = 0; b <= (1ull << bb); ++b) {
    uint64_t bound = b << (32 - bb);
    while (j < n && (uint64_t)km[j] < bound) ++j;
    x->bucket[b] = (uint32_t)j;
  }
  return x;
}

static uint64_t verify_at(const oidx_t *x, const uint8_t *qc, int L, int seed, const mirge_round_policy *pol, uint32_t r, uint32_t astart) {
  const uint8_t *ref = x->text + astart;
  int mm = 0, smm = 0;
  for (int j = 0; j < L; ++j) {
    int rc = base_code(ref[j]);
    if (rc == 4) return MIRGE_NO_HIT;
    if (qc[j] != rc) { ++mm; if (j < seed) ++smm; if (mm > pol->total_mm || smm > pol->seed_mm) return MIRGE_NO_HIT; }
  }
  return ((uint64_t)mm << 56) | ((uint64_t)r << 28) | (uint64_t)(astart - x->ref_off[r]);
}

static uint64_t search_query(const oidx_t *x, const uint8_t *q, int L, const mirge_round_policy *pol) {
  uint64_t best = MIRGE_NO_HIT;
  if (L <= 0) return best;
  int seed = pol->seed_len == 0 ? L : (pol->seed_len < L ? pol->seed_len : L);
  int np = pol->seed_mm + 1;
  if (seed / np < 4) return scan_query(q, L, x->text, x->ref_off, x->n_refs, pol);
  uint8_t qc[MIRGE_MAX_READ_LEN];
  for (int j = 0; j < L; ++j) qc[j] = (uint8_t)base_code(q[j]);
  for (int pi = 0; pi < np; ++pi) {
    int a = (int)((long long)pi * seed / np), b = (int)((long long)(pi + 1) * seed / np);
    int s = b - a < 16 ? b - a : 16, has_n = 0;
    uint32_t k = 0;
    for (int i = 0; i < b - a; ++i) { if (qc[a + i] == 4) has_n = 1; else if (i < s) k |= (uint32_t)qc[a + i] << (2 * (15 - i)); }
    if (has_n) continue;
    uint32_t span = s == 16 ? 0u : ((1u << (2 * (16 - s))) - 1u), k_hi = k | span;
    int bsh = 32 - x->bucket_bits;
    uint64_t lo = x->bucket[k >> bsh], hi = x->bucket[(k_hi >> bsh) + 1], l, h;
    l = lo; h = hi; while (l < h) { uint64_t m = (l + h) >> 1; if (x->kmer[m] < k) l = m + 1; else h = m; } lo = l;
    h = hi; while (l < h) { uint64_t m = (l + h) >> 1; if (x->kmer[m] <= k_hi) l = m + 1; else h = m; } hi = l;
    for (uint64_t e = lo; e < hi; ++e) {
      uint32_t pos = x->pos[e];
      if (pos < (uint32_t)a) continue;
      uint32_t astart = pos - (uint32_t)a;
      uint32_t rl = 0, rh = x->n_refs;
      while (rh - rl > 1) { uint32_t m = (rl + rh) >> 1; if (x->ref_off[m] <= pos) rl = m; else rh = m; }
      if (astart < x->ref_off[rl] || (uint64_t)astart + (uint64_t)L > x->ref_off[rl + 1]) continue;
      uint64_t hh = verify_at(x, qc, L, seed, pol, rl, astart);
      if (hh < best) best = hh;
    }
  }
  return best;
}

void oracle_annotate_round_indexed(const uint8_t *keys, const uint64_t *key_off, uint64_t n, void *index,
                                   const mirge_round_policy *pol, uint8_t *annot_round, uint64_t *hit, int nthreads) {
  const oidx_t *x = index;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    const uint8_t *s = keys + key_off[i]; int len = (int)(key_off[i + 1] - key_off[i]);
    if (pol->select == MIRGE_SELECT_LEN_LT26) { if (!(len < 26)) continue; }
    else if (pol->select == MIRGE_SELECT_LEN_GT25) { if (!(len > 25)) continue; }
    else if (annot_round[i] != 0xFF) continue;
    int qs = 0, ql = round_query(s, len, pol, &qs);
    if (ql < 0) continue;
    uint64_t h = search_query(x, s + qs, ql, pol);
    if (h != MIRGE_NO_HIT) { annot_round[i] = (uint8_t)pol->round; hit[i] = h; }
  }
}
