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
  int maxm = 0;
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
    if (s->m > maxm) maxm = s->m;
  }
  MIRGE_CUDA(ctx, cudaSetDevice(ctx->device));
  MIRGE_CUDA(ctx, cudaMemcpyToSymbol(c_p, &d, sizeof(d)));
  ctx->params = *p;
  ctx->params_set = 1;
  ctx->max_adapter_len = maxm;
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

// AdapterCutter._best_match: most matches, then fewer errors, first adapter wins ties.
template <int MAXM>
__device__ __forceinline__ int best_match(const uint8_t *read, int n, Match &best) {
  int which = -1;
  for (int a = 0; a < c_p.n_adapters; ++a) {
    Match mt;
    if (!locate<MAXM>(a, read, n, mt)) continue;
    if (which < 0 || mt.matches > best.matches || (mt.matches == best.matches && mt.errors < best.errors)) {
      best = mt;
      which = a;
    }
  }
  return which;
}

template <int MAXM>
__device__ __forceinline__ void apply_mod(int mi, const uint8_t *seq, const uint8_t *qual, int &start, int &stop) {
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
        const int a = best_match<MAXM>(seq + start, stop - start, mt);
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

template <int MAXM>
__global__ void __launch_bounds__(TRIM_THREADS)
trim_kernel(const uint8_t *__restrict__ fq, uint64_t nbytes, const uint32_t *__restrict__ line_start, uint64_t n_records,
            ushort4 *__restrict__ win, uint32_t *__restrict__ key_off, uint32_t *__restrict__ keys, uint64_t keys_cap,
            unsigned long long *__restrict__ ctrl, uint32_t smem_bytes) {
  extern __shared__ uint4 smem4[];
  __shared__ uint32_t s_scan[TRIM_THREADS / 32];
  __shared__ unsigned long long s_base;
  uint8_t *sbuf = (uint8_t *)smem4;
  const int tid = threadIdx.x;
  const uint64_t r0 = (uint64_t)blockIdx.x * TRIM_THREADS;
  const uint64_t r = r0 + tid;
  const bool valid = r < n_records;
  const uint64_t r_end = min(r0 + (uint64_t)TRIM_THREADS, n_records);
  const uint32_t span_lo = line_start[4 * r0];
  const uint64_t span_hi = min((uint64_t)line_start[4 * r_end], nbytes);
  const uint32_t alo = span_lo & ~15u;
  const bool staged = (span_hi - alo) <= smem_bytes;
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

  const int E = c_p.slots;
  int w_start[MIRGE_MAX_MODS], w_stop[MIRGE_MAX_MODS], w_us[MIRGE_MAX_MODS], w_ue[MIRGE_MAX_MODS];
  uint32_t w_words[MIRGE_MAX_MODS];  // 0 = not kept
#pragma unroll
  for (int s = 0; s < MIRGE_MAX_MODS; ++s) { w_start[s] = w_stop[s] = w_us[s] = w_ue[s] = 0; w_words[s] = 0; }
  const uint8_t *seq = nullptr;
  uint32_t my_words = 0, my_kept = 0;
  if (valid) {
    const uint4 ls = *(const uint4 *)(line_start + 4 * r);
    const uint32_t nxt = line_start[4 * r + 4];
    int sl = (int)(ls.z - 1 - ls.y), ql = (int)(nxt - 1 - ls.w);
    if (sl > 0 && B[ls.y + sl - 1] == '\r') --sl;
    if (ql > 0 && B[ls.w + ql - 1] == '\r') --ql;
    const bool bad = B[ls.x] != '@' || B[ls.z] != '+' || sl != ql || sl > MIRGE_MAX_READ_LEN;
    if (bad) {
      atomicOr(ctrl + 2, (sl > MIRGE_MAX_READ_LEN && sl == ql) ? 4ull : 1ull);
      atomicMax(ctrl + 3, ~(unsigned long long)r);
    } else {
      seq = B + ls.y;
      const uint8_t *qual = B + ls.w;
      int start = 0, stop = sl;
      if (c_p.umi_mode == MIRGE_UMI_QIAGEN) {
        for (int mi = 0; mi < c_p.n_mods; ++mi) apply_mod<MAXM>(mi, seq, qual, start, stop);
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
      } else {
#pragma unroll
        for (int mi = 0; mi < MIRGE_MAX_MODS; ++mi) {
          if (mi < c_p.n_mods) {
            apply_mod<MAXM>(mi, seq, qual, start, stop);
            if (E != 1 || mi == c_p.n_mods - 1) {
              const int slot = (E == 1) ? 0 : mi;
              int ln = stop - start;
              if (c_p.umi_mode == MIRGE_UMI_FLANKS) ln = max(ln - c_p.umi5 - c_p.umi3, 0);
              // slot is compile-time mi in HEAD mode; in release mode only slot 0 is used
              if (slot == 0) { w_start[0] = start; w_stop[0] = stop; w_words[0] = ln >= c_p.min_len; }
              else { w_start[mi] = start; w_stop[mi] = stop; w_words[mi] = ln >= c_p.min_len; }
            }
          }
        }
      }
      // size of every kept key: header + payload + exceptions
#pragma unroll
      for (int s = 0; s < MIRGE_MAX_MODS; ++s) {
        if (s < E && w_words[s]) {
          const int l1 = w_stop[s] - w_start[s], len = l1 + (w_ue[s] - w_us[s]);
          uint32_t nexc = 0;
          for (int p = 0; p < len; ++p) nexc += base_code_exact(key_byte(seq, w_start[s], l1, w_us[s], p)) == 4u;
          w_words[s] = 1u + ((uint32_t)(len + 15) >> 4) + nexc;
          my_words += w_words[s];
          ++my_kept;
        }
      }
    }
  }
  // CTA-level allocation of key space: one atomic per CTA
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t inc = my_words;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  const uint32_t kept_warp = __reduce_add_sync(0xffffffffu, my_kept);
  if (lane == 31) s_scan[warp] = inc;
  __syncthreads();
  uint32_t warp_base = 0, cta_total = 0;
#pragma unroll
  for (int w = 0; w < TRIM_THREADS / 32; ++w) {
    if (w < warp) warp_base += s_scan[w];
    cta_total += s_scan[w];
  }
  if (tid == 0) s_base = cta_total ? atomicAdd(ctrl + 0, (unsigned long long)cta_total) : 0ull;
  if (lane == 0 && kept_warp) atomicAdd(ctrl + 1, (unsigned long long)kept_warp);
  __syncthreads();
  const unsigned long long cta_base = s_base;
  const bool overflow = cta_base + cta_total > keys_cap || cta_base + cta_total > 0xFFFFFFF0ull;
  if (overflow && tid == 0 && cta_total) atomicOr(ctrl + 2, 2ull);
  if (!valid) return;
  uint32_t off = (uint32_t)cta_base + warp_base + inc - my_words;
#pragma unroll
  for (int s = 0; s < MIRGE_MAX_MODS; ++s) {
    if (s < E) {
      const uint64_t e = r * (uint64_t)E + s;
      win[e] = make_ushort4((unsigned short)w_start[s], (unsigned short)w_stop[s], (unsigned short)w_us[s], (unsigned short)w_ue[s]);
      if (w_words[s] && !overflow) {
        key_off[e] = off;
        const int l1 = w_stop[s] - w_start[s], len = l1 + (w_ue[s] - w_us[s]);
        const uint32_t npay = (uint32_t)(len + 15) >> 4;
        const uint32_t nexc = w_words[s] - 1u - npay;
        uint32_t *k = keys + off;
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
        off += w_words[s];
      } else {
        key_off[e] = 0xFFFFFFFFu;
      }
    }
  }
}

extern "C" int mirge_trim(mirge_ctx *ctx, const uint8_t *d_fastq, uint64_t nbytes, const uint32_t *d_line_start, uint64_t n_records,
                          uint16_t *d_win, uint32_t *d_key_off, uint32_t *d_keys, uint64_t keys_capacity_words,
                          uint64_t *d_trim_ctrl, void *stream_) {
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
  uint64_t want = avg * TRIM_THREADS * 5 / 4 + 64;
  want = (want + 1023) & ~1023ull;
  if (want < 8192) want = 8192;
  if (want > 96 * 1024) want = 96 * 1024;
  const uint32_t smem = (uint32_t)want;
  const unsigned grid = (unsigned)((n_records + TRIM_THREADS - 1) / TRIM_THREADS);
  unsigned long long *ctrl = (unsigned long long *)d_trim_ctrl;
  ushort4 *win = (ushort4 *)d_win;
  if (ctx->max_adapter_len <= 32) {
    MIRGE_CUDA(ctx, cudaFuncSetAttribute(trim_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    trim_kernel<32><<<grid, TRIM_THREADS, smem, stream>>>(d_fastq, nbytes, d_line_start, n_records, win, d_key_off, d_keys,
                                                          keys_capacity_words, ctrl, smem);
  } else {
    MIRGE_CUDA(ctx, cudaFuncSetAttribute(trim_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    trim_kernel<64><<<grid, TRIM_THREADS, smem, stream>>>(d_fastq, nbytes, d_line_start, n_records, win, d_key_off, d_keys,
                                                          keys_capacity_words, ctrl, smem);
  }
  MIRGE_LAUNCH_CHECK(ctx, "trim_kernel");
  return MIRGE_OK;
}
