/* pinflate.c -- chunk-parallel inflate of ONE gzip stream on the host cores ("next" row f-3:
 * "parallel gz inflate").  What the reference does here is `xopen(FQfile, "rb")` (mirge/libs/digest.py:136), i.e. one
 * gzip stream inflated by one core (or by an external pigz / igzip process, which also inflate serially); a single
 * sample's .fastq.gz therefore reaches the device path at ~0.1-0.2 GB/s.  DEFLATE is serial by construction -- block
 * boundaries are only known after decoding what precedes them and every block may refer to the 32 KB before it -- so
 * this decoder does what pugz / rapidgzip published:
 *
 *   1. the compressed stream is cut every `chunk` bytes; near every cut a thread looks for a bit position at which a
 *      dynamic-Huffman block header parses (complete code-length code, complete literal/length code, end-of-block
 *      symbol present);
 *   2. from there it decodes into 16-bit symbols with an UNKNOWN window: a back-reference that reaches before the
 *      chunk yields markers (256 + position in the 32 KB window) that are copied like any other symbol;
 *   3. a chunk stops exactly when its bit position equals the next chunk's start at a block boundary -- which proves
 *      that start to be a real boundary; a start the chain never lands on is discarded and the previous chunk decodes
 *      on, so a false positive of step 1 costs time, never correctness;
 *   4. the windows are chained (only the last 32 KB of each chunk are resolved serially), then all chunks replace
 *      their markers in parallel; CRC-32 is computed per chunk and combined, and checked with ISIZE against the gzip
 *      trailer of every member.
 *
 * Multi-member files and zero padding between members are handled as Python's gzip module does.  Host-only code (no
 * CUDA): built by __graft_entry__.build() with gcc -fopenmp into libmirge_inflate.so and bound by ingest.py; when the
 * library is missing ingest.py keeps its serial zlib reader (same bytes, slower).  */
#ifdef _OPENMP
#include <omp.h>
#else /* built without OpenMP: the same decoder on one core (the pragmas are ignored) */
#include <time.h>
static double omp_get_wtime(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
#endif
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define WIN 32768
#define MAXBITS 15
#define LIT_PB 11 /* primary table bits, literal/length code */
#define DST_PB 9  /* primary table bits, distance code */
#define POOL_MAX 512

/* ------------------------------------------------------------------ bit reader ------- */
typedef struct {
  const uint8_t *p;
  uint64_t n;   /* bytes */
  uint64_t pos; /* next byte to load */
  uint64_t buf;
  int cnt; /* valid bits in buf */
} bitr;

static inline void br_init(bitr *b, const uint8_t *p, uint64_t n, uint64_t bitpos) {
  b->p = p; b->n = n; b->pos = bitpos >> 3; b->buf = 0; b->cnt = 0;
  int skip = (int)(bitpos & 7);
  if (skip) {
    if (b->pos < n) b->buf = (uint64_t)p[b->pos] >> skip;
    b->pos++;
    b->cnt = 8 - skip;
  }
}
static inline void br_refill(bitr *b) {
  if (b->pos + 8 <= b->n) { /* one unaligned load tops the buffer up to >= 56 bits */
    uint64_t v;
    memcpy(&v, b->p + b->pos, 8);
    b->buf |= v << b->cnt;
    const int adv = (63 - b->cnt) >> 3;
    b->pos += (uint64_t)adv;
    b->cnt += adv << 3;
    return;
  }
  while (b->cnt <= 56) {
    uint64_t byte = b->pos < b->n ? b->p[b->pos] : 0; /* zeros beyond the end; br_overrun() tells */
    b->buf |= byte << b->cnt;
    b->pos++;
    b->cnt += 8;
  }
}
static inline uint64_t br_bitpos(const bitr *b) { return b->pos * 8 - (uint64_t)b->cnt; }
static inline int br_overrun(const bitr *b) { return br_bitpos(b) > b->n * 8; }
static inline uint32_t br_peek(const bitr *b, int k) { return (uint32_t)(b->buf & ((1ull << k) - 1ull)); }
static inline void br_drop(bitr *b, int k) { b->buf >>= k; b->cnt -= k; }
static inline uint32_t br_bits(bitr *b, int k) { uint32_t v = br_peek(b, k); br_drop(b, k); return v; }

/* ------------------------------------------------------------------ Huffman tables ------- */
typedef struct {
  uint16_t count[MAXBITS + 1];
  uint16_t symbol[288];
  uint16_t fast[1 << LIT_PB]; /* (symbol << 4) | length, 0 = longer than the primary table */
  int pb;
  /* the primary table once more, with what the symbol means folded in (huff_pack): base << 16 | extra bits << 8 | kind |
   * code length; 0 = not in the primary table (longer code, unused pattern, symbol out of range): huff_decode decides */
  uint32_t pk[1 << LIT_PB];
} huff;
#define PK_LIT 0x10u /* base = the literal */
#define PK_LIT2 0x80u /* (with PK_LIT) two literals whose codes fit the primary table together: bits 16-23 the first,
                       * bits 24-31 the second, code length = both codes */
#define PK_EOB 0x20u
#define PK_LEN 0x40u /* base = length base; distance tables: always a distance */

/* returns 0 complete, > 0 incomplete (left over), < 0 over-subscribed */
static int huff_build(huff *h, const uint8_t *len, int n, int pb) {
  uint16_t offs[MAXBITS + 1];
  h->pb = pb;
  memset(h->count, 0, sizeof(h->count));
  for (int s = 0; s < n; ++s) h->count[len[s]]++;
  int left = 1;
  for (int l = 1; l <= MAXBITS; ++l) {
    left <<= 1;
    left -= h->count[l];
    if (left < 0) return left;
  }
  offs[1] = 0;
  for (int l = 1; l < MAXBITS; ++l) offs[l + 1] = offs[l] + h->count[l];
  for (int s = 0; s < n; ++s)
    if (len[s]) h->symbol[offs[len[s]]++] = (uint16_t)s;
  /* primary table: canonical codes, bit-reversed (DEFLATE packs codes MSB first into an LSB-first stream) */
  memset(h->fast, 0, sizeof(uint16_t) << pb);
  uint32_t code = 0;
  int idx = 0;
  for (int l = 1; l <= pb; ++l) {
    for (int k = 0; k < h->count[l]; ++k, ++idx, ++code) {
      uint32_t rev = 0;
      for (int b = 0; b < l; ++b) rev |= ((code >> b) & 1u) << (l - 1 - b);
      const uint16_t e = (uint16_t)((h->symbol[idx] << 4) | l);
      for (uint32_t j = rev; j < (1u << pb); j += 1u << l) h->fast[j] = e;
    }
    code <<= 1;
  }
  return left;
}

/* decode one symbol; -1 = invalid code.  The caller has >= 15 bits in the buffer. */
static inline int huff_decode(const huff *h, bitr *b) {
  const uint16_t e = h->fast[br_peek(b, h->pb)];
  if (e) { br_drop(b, e & 15); return e >> 4; }
  /* canonical, bit by bit (codes longer than the primary table, or an incomplete code's unused pattern) */
  int code = 0, first = 0, index = 0;
  uint64_t bits = b->buf;
  for (int l = 1; l <= MAXBITS; ++l) {
    code |= (int)(bits & 1u);
    bits >>= 1;
    const int c = h->count[l];
    if (code - c < first) { br_drop(b, l); return h->symbol[index + (code - first)]; }
    index += c;
    first += c;
    first <<= 1;
    code <<= 1;
  }
  return -1;
}

static const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t CLORD[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

/* fold base values and extra-bit counts into the primary table of a literal/length (is_dist = 0) or distance code */
static void huff_pack(huff *h, int is_dist) {
  const uint32_t n = 1u << h->pb;
  for (uint32_t j = 0; j < n; ++j) {
    const uint32_t e = h->fast[j], l = e & 15u, sym = e >> 4;
    uint32_t v = 0;
    if (e) {
      if (is_dist) {
        if (sym < 30) v = ((uint32_t)DBASE[sym] << 16) | ((uint32_t)DEXT[sym] << 8) | PK_LEN | l;
      } else if (sym < 256) {
        v = (sym << 16) | PK_LIT | l;
        /* what follows the code inside this index, zero-extended, decodes by itself iff its code needs no more than
         * the bits that are left */
        const uint32_t e2 = h->fast[j >> l], l2 = e2 & 15u, sym2 = e2 >> 4;
        if (e2 && sym2 < 256 && l + l2 <= (uint32_t)h->pb) v = (sym2 << 24) | (sym << 16) | PK_LIT | PK_LIT2 | (l + l2);
      }
      else if (sym == 256) v = PK_EOB | l;
      else if (sym < 286) v = ((uint32_t)LBASE[sym - 257] << 16) | ((uint32_t)LEXT[sym - 257] << 8) | PK_LEN | l;
    }
    h->pk[j] = v;
  }
}

/* Header of a dynamic block (the 3 block bits already consumed).  strict: what the block finder demands of a
 * candidate (complete codes; zlib never writes anything else); otherwise what inflate accepts.  0 = ok. */
static int dynamic_header(bitr *b, huff *lit, huff *dst, int strict) {
  br_refill(b);
  const int nlen = (int)br_bits(b, 5) + 257, ndist = (int)br_bits(b, 5) + 1, ncode = (int)br_bits(b, 4) + 4;
  if (nlen > 286 || ndist > 30) return -1;
  uint8_t lens[320];
  uint8_t cl[19];
  memset(cl, 0, sizeof(cl));
  for (int i = 0; i < ncode; ++i) {
    if ((i & 7) == 0) br_refill(b);
    cl[CLORD[i]] = (uint8_t)br_bits(b, 3);
  }
  huff clh;
  int err = huff_build(&clh, cl, 19, 7);
  if (err < 0 || (err > 0 && strict)) return -2;
  int idx = 0;
  while (idx < nlen + ndist) {
    br_refill(b);
    int sym = huff_decode(&clh, b);
    if (sym < 0) return -3;
    if (sym < 16) { lens[idx++] = (uint8_t)sym; continue; }
    int rep, val = 0;
    if (sym == 16) {
      if (idx == 0) return -4;
      val = lens[idx - 1];
      rep = 3 + (int)br_bits(b, 2);
    } else if (sym == 17) rep = 3 + (int)br_bits(b, 3);
    else rep = 11 + (int)br_bits(b, 7);
    if (idx + rep > nlen + ndist) return -5;
    while (rep--) lens[idx++] = (uint8_t)val;
  }
  if (br_overrun(b)) return -6;
  if (lens[256] == 0) return -7; /* no end-of-block code */
  err = huff_build(lit, lens, nlen, LIT_PB);
  if (err < 0) return -8;
  if (err > 0 && (strict || nlen != lit->count[0] + lit->count[1])) return -8; /* incomplete: only a single 1-bit code */
  err = huff_build(dst, lens + nlen, ndist, DST_PB);
  if (err < 0) return -9;
  if (err > 0 && ndist != dst->count[0] + dst->count[1]) return -9;
  return 0;
}

static huff g_fixed_lit, g_fixed_dst;
static int g_fixed_ready = 0;
static void fixed_tables(void) {
  uint8_t l[288];
  int s = 0;
  for (; s < 144; ++s) l[s] = 8;
  for (; s < 256; ++s) l[s] = 9;
  for (; s < 280; ++s) l[s] = 7;
  for (; s < 288; ++s) l[s] = 8;
  huff_build(&g_fixed_lit, l, 288, LIT_PB);
  huff_pack(&g_fixed_lit, 0);
  for (s = 0; s < 30; ++s) l[s] = 5;
  huff_build(&g_fixed_dst, l, 30, DST_PB);
  huff_pack(&g_fixed_dst, 1);
  g_fixed_ready = 1;
}

/* ------------------------------------------------------------------ chunk decoder ------- */
enum { END_HIT = 1, END_FINAL = 2, END_LIMIT = 3, END_ERROR = 4 };

typedef struct {
  /* A speculative chunk is decoded in symbol mode: 16-bit symbols behind a prefix of WIN markers, resolved into bytes
   * once its window is known.  A chunk whose window is known (the first of every wave) is decoded straight into bytes.
   * (Changing a speculative chunk to byte mode once WIN marker-free symbols lie behind it was built and measured: on
   * FASTQ it almost never happens -- the markers of header and separator bytes are copied forward from record to record
   * for the whole chunk -- and tracking it cost more than it saved.) */
  uint16_t *sym;  /* [WIN prefix][symbols] */
  uint64_t cap;   /* symbols allocated */
  uint64_t n_sym; /* output symbols still to be resolved (0: byte mode) */
  uint8_t *bytes; /* known window: [WIN window bytes][output]; speculative: filled when the symbols are resolved */
  uint64_t bytes_cap;
  uint64_t boff;  /* where the output starts in bytes (WIN or 0) */
  uint64_t n;     /* output bytes */
  uint64_t start_bit, end_bit;
  int end_kind;
  int hit;        /* END_HIT: index of the candidate the chunk stopped at */
  int err;
  int known;      /* 1: the real window is known (no markers can occur), 2: empty window (start of a member) */
  uint64_t max_out; /* speculative chunks: give up beyond this much output (0 = no limit) */
  uint32_t crc; /* of the whole output (computed when the head is resolved; inside the decode loop it cost 3x as much) */
} chunk;

static int chunk_reserve(chunk *c, uint64_t need_total) {
  if (need_total <= c->cap) return 0;
  uint64_t nc = c->cap + c->cap / 2;
  if (nc < need_total) nc = need_total;
  uint16_t *p = (uint16_t *)realloc(c->sym, nc * sizeof(uint16_t));
  if (!p) return -1;
  c->sym = p;
  c->cap = nc;
  return 0;
}
static int bytes_reserve(chunk *c, uint64_t need_total) {
  if (need_total <= c->bytes_cap) return 0;
  uint64_t nc = c->bytes_cap + c->bytes_cap / 2;
  if (nc < need_total) nc = need_total;
  uint8_t *p = (uint8_t *)realloc(c->bytes, nc);
  if (!p) return -1;
  c->bytes = p;
  c->bytes_cap = nc;
  return 0;
}

/* One length / distance pair after the literal fast path has declined: `e` is the packed entry of the next
 * literal/length code (0: not in the primary table).  Sets *plen / *pdist; returns 0 = a match to copy, 1 = a literal in
 * *plen, 2 = end of block, < 0 error.  The caller refilled (>= 56 bits) before looking `e` up and has consumed nothing
 * since: code (<= 15) + length extra (<= 5) + distance code (<= 15) + distance extra (<= 13) = 48 bits. */
static inline int lenpair(bitr *b, const huff *L, const huff *D, uint32_t e, uint32_t *plen, uint32_t *pdist) {
  uint32_t len;
  if (e) {
    br_drop(b, e & 15);
    if (e & PK_EOB) return 2;
    len = (e >> 16) + br_bits(b, (e >> 8) & 15);
  } else {
    int s = huff_decode(L, b);
    if (s < 256) {
      if (s < 0) return -14;
      *plen = (uint32_t)s;
      return 1;
    }
    if (s == 256) return 2;
    s -= 257;
    if (s >= 29) return -15;
    len = LBASE[s] + br_bits(b, LEXT[s]);
  }
  const uint32_t d = D->pk[br_peek(b, DST_PB)];
  if (d) {
    br_drop(b, d & 15);
    *pdist = (d >> 16) + br_bits(b, (d >> 8) & 15);
  } else {
    const int ds = huff_decode(D, b);
    if (ds < 0 || ds >= 30) return -16;
    *pdist = DBASE[ds] + br_bits(b, DEXT[ds]);
  }
  *plen = len;
  return 0;
}

/* The symbols of one Huffman block, symbol mode.  *po: write index into c->sym.  0 = end of block reached, < 0 error. */
static int block_symbols(chunk *c, bitr *b, const huff *L, const huff *D, uint64_t nbytes, uint64_t *po) {
  uint64_t o = *po;
  int bad = 0;
  for (;;) {
    if (o + 300 > c->cap) {
      if (c->max_out && o - WIN > c->max_out) { bad = -20; break; } /* a speculative chunk that runs away */
      if (chunk_reserve(c, o + 300 + c->cap / 2)) { bad = -12; break; }
    }
    br_refill(b);
    if (b->pos > nbytes + 16) { bad = -18; break; } /* reading zeros beyond the end of the file */
    uint32_t e = L->pk[br_peek(b, LIT_PB)];
    if (e & PK_LIT) { /* runs of literals as in block_bytes, two 16-bit symbols stored per entry */
#define PUT_LITS(e_)                                          \
  do {                                                        \
    br_drop(b, (e_) & 15);                                    \
    c->sym[o] = (uint16_t)(((e_) >> 16) & 0xFFu);             \
    c->sym[o + 1] = (uint16_t)((e_) >> 24);                   \
    o += 1 + (((e_) >> 7) & 1u);                              \
  } while (0)
      PUT_LITS(e);
      e = L->pk[br_peek(b, LIT_PB)];
      if (e & PK_LIT) {
        PUT_LITS(e);
        e = L->pk[br_peek(b, LIT_PB)];
        if (e & PK_LIT) {
          PUT_LITS(e);
          e = L->pk[br_peek(b, LIT_PB)];
          if (e & PK_LIT) PUT_LITS(e);
        }
      }
#undef PUT_LITS
      continue;
    }
    uint32_t len, dist;
    const int k = lenpair(b, L, D, e, &len, &dist);
    if (k) {
      if (k == 1) { c->sym[o++] = (uint16_t)len; continue; }
      if (k < 0) bad = k;
      break;
    }
    uint16_t *w = c->sym + o;
    const uint16_t *r = w - dist;
    o += len;
    if (dist >= 4) { /* four symbols (8 bytes) at a time; the 300 symbols of slack take the overshoot */
      uint16_t *const end = w + len;
      do { memcpy(w, r, 8); w += 4; r += 4; } while (w < end);
    } else {
      for (uint32_t i = 0; i < len; ++i) w[i] = r[i];
    }
  }
  *po = o;
  return bad;
}

/* The same in byte mode: *po is the write index into c->bytes (at least WIN bytes of history lie in front of it). */
static int block_bytes(chunk *c, bitr *b, const huff *L, const huff *D, uint64_t nbytes, uint64_t *po) {
  uint64_t o = *po;
  int bad = 0;
  for (;;) {
    if (o + 300 > c->bytes_cap) {
      if (c->max_out && o - c->boff > c->max_out) { bad = -20; break; }
      if (bytes_reserve(c, o + 300 + c->bytes_cap / 2)) { bad = -12; break; }
    }
    br_refill(b);
    if (b->pos > nbytes + 16) { bad = -18; break; }
    uint32_t e = L->pk[br_peek(b, LIT_PB)];
    if (e & PK_LIT) {
      /* runs of literals: up to four primary-table entries (<= LIT_PB = 11 bits each, one or two literals) per refill
       * of >= 56 bits; both bytes of an entry are stored, the write index moves by one or two */
#define PUT_LITS(e_)                                          \
  do {                                                        \
    const uint16_t two_ = (uint16_t)((e_) >> 16);             \
    br_drop(b, (e_) & 15);                                    \
    memcpy(c->bytes + o, &two_, 2);                           \
    o += 1 + (((e_) >> 7) & 1u);                              \
  } while (0)
      PUT_LITS(e);
      e = L->pk[br_peek(b, LIT_PB)];
      if (e & PK_LIT) {
        PUT_LITS(e);
        e = L->pk[br_peek(b, LIT_PB)];
        if (e & PK_LIT) {
          PUT_LITS(e);
          e = L->pk[br_peek(b, LIT_PB)];
          if (e & PK_LIT) PUT_LITS(e);
        }
      }
#undef PUT_LITS
      continue;
    }
    uint32_t len, dist;
    const int k = lenpair(b, L, D, e, &len, &dist);
    if (k) {
      if (k == 1) { c->bytes[o++] = (uint8_t)len; continue; }
      if (k < 0) bad = k;
      break;
    }
    if (c->known == 2 && dist > o - c->boff) { bad = -17; break; } /* before the start of the member */
    uint8_t *w = c->bytes + o;
    const uint8_t *r = w - dist;
    o += len;
    if (dist >= 8) { /* eight bytes at a time; the 300 bytes of slack take the overshoot */
      uint8_t *const end = w + len;
      do { memcpy(w, r, 8); w += 8; r += 8; } while (w < end);
    } else if (dist == 1) {
      memset(w, r[0], len);
    } else {
      for (uint32_t i = 0; i < len; ++i) w[i] = r[i];
    }
  }
  *po = o;
  return bad;
}

/* Decode from c->start_bit.  Stops at a block boundary whose bit position equals cand[j] for some j >= first_cand
 * (END_HIT), or is >= limit_bit once no candidate is left (END_LIMIT), or after the final block (END_FINAL).
 * c->known != 0: c->bytes[0 .. WIN) holds the window and the output follows it; else c->sym[0 .. WIN) holds markers. */
static void chunk_decode(chunk *c, const uint8_t *data, uint64_t nbytes, const uint64_t *cand, int n_cand, int first_cand,
                         uint64_t limit_bit) {
  bitr b;
  br_init(&b, data, nbytes, c->start_bit);
  int byte_mode = c->known != 0;
  uint64_t o = WIN;  /* write index: into c->sym (symbol mode) or c->bytes (byte mode) */
  c->boff = byte_mode ? WIN : 0;
  c->n_sym = 0;
  int j = first_cand;
  huff lit, dst;
  c->end_kind = END_ERROR;
  c->err = 0;
  c->hit = -1;
  int first_block = 1;
  for (;;) {
    const uint64_t bp = br_bitpos(&b);
    if (!first_block) {
      while (j < n_cand && cand[j] < bp) ++j; /* starts the chain did not land on: not real boundaries */
      if (j < n_cand && cand[j] == bp) { c->end_kind = END_HIT; c->hit = j; break; }
      if (j >= n_cand && bp >= limit_bit) { c->end_kind = END_LIMIT; break; }
    }
    first_block = 0;
    br_refill(&b);
    const int final = (int)br_bits(&b, 1);
    const int type = (int)br_bits(&b, 2);
    if (type == 3) { c->err = -10; break; }
    if (type == 0) {
      br_drop(&b, b.cnt & 7); /* to the byte boundary */
      br_refill(&b);
      const uint32_t len = br_bits(&b, 16), nlen = br_bits(&b, 16);
      if ((len ^ 0xFFFFu) != nlen) { c->err = -11; break; }
      /* the bit buffer holds whole bytes now: hand them back and copy from the stream */
      uint64_t at = br_bitpos(&b) >> 3;
      if (at + len > nbytes) { c->err = -13; break; }
      if (byte_mode) {
        if (bytes_reserve(c, o + len + 16)) { c->err = -12; break; }
        memcpy(c->bytes + o, data + at, len);
      } else {
        if (chunk_reserve(c, o + len + 16)) { c->err = -12; break; }
        for (uint32_t i = 0; i < len; ++i) c->sym[o + i] = data[at + i];
      }
      o += len;
      br_init(&b, data, nbytes, (at + len) * 8);
    } else {
      const huff *L, *D;
      if (type == 1) {
        L = &g_fixed_lit; D = &g_fixed_dst;
      } else {
        if ((c->err = dynamic_header(&b, &lit, &dst, 0)) != 0) break;
        huff_pack(&lit, 0);
        huff_pack(&dst, 1);
        L = &lit; D = &dst;
      }
      const int bad = byte_mode ? block_bytes(c, &b, L, D, nbytes, &o) : block_symbols(c, &b, L, D, nbytes, &o);
      if (bad) { c->err = bad; break; }
      if (br_overrun(&b)) { c->err = -18; break; }
    }
    if (final) { c->end_kind = END_FINAL; break; }
  }
  if (byte_mode) {
    c->n = o - c->boff;
  } else {
    c->n = o - WIN;
    c->n_sym = c->n; /* never left symbol mode: everything is resolved later */
  }
  c->end_bit = br_bitpos(&b);
}

/* first bit position in [from, to) at which a non-final dynamic block header parses strictly; UINT64_MAX: none */
static uint64_t find_block(const uint8_t *data, uint64_t nbytes, uint64_t from, uint64_t to) {
  huff lit, dst;
  for (uint64_t bp = from; bp < to; ++bp) {
    const uint64_t byte = bp >> 3;
    if (byte + 8 >= nbytes) break;
    uint64_t v;
    memcpy(&v, data + byte, 8);
    v >>= (bp & 7);
    if ((v & 7) != 4) continue;                    /* BFINAL = 0, BTYPE = 2 */
    if (((v >> 3) & 31) > 29 || ((v >> 8) & 31) > 29) continue; /* HLIT, HDIST */
    bitr b;
    br_init(&b, data, nbytes, bp + 3);
    if (dynamic_header(&b, &lit, &dst, 1) == 0) return bp;
  }
  return UINT64_MAX;
}

/* ------------------------------------------------------------------ CRC-32 (slice by 8) + combine ------- */
static uint32_t g_crc[8][256];
static void crc_tables(void) {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    g_crc[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc[t][i] = (g_crc[t - 1][i] >> 8) ^ g_crc[0][g_crc[t - 1][i] & 0xFF];
}
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
/* CRC-32 by carry-less multiplication (folding four 128-bit lanes over 64-byte blocks, then Barrett reduction; the
 * constants are x^n mod P for the reflected gzip polynomial, as in Intel's "Fast CRC computation using PCLMULQDQ").
 * crc in / out in the register form (already inverted); n >= 64, a multiple of 16.  Used when the CPU has PCLMULQDQ;
 * tests/test_ingest_host.py holds every path against zlib. */
__attribute__((target("pclmul,sse4.1"))) static uint32_t crc32_clmul(uint32_t crc, const uint8_t *buf, uint64_t len) {
  const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596ll, 0x0154442bd4ll);
  const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009ell, 0x01751997d0ll);
  const __m128i k5k0 = _mm_set_epi64x(0, 0x0163cd6124ll);
  const __m128i poly = _mm_set_epi64x(0x01f7011641ll, 0x01db710641ll);
  __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
  x1 = _mm_loadu_si128((const __m128i *)(buf + 0x00));
  x2 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
  x3 = _mm_loadu_si128((const __m128i *)(buf + 0x20));
  x4 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
  x0 = k1k2;
  buf += 64;
  len -= 64;
  while (len >= 64) {
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
    x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
    x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
    x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
    y5 = _mm_loadu_si128((const __m128i *)(buf + 0x00));
    y6 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
    y7 = _mm_loadu_si128((const __m128i *)(buf + 0x20));
    y8 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
    x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
    x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
    x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
    buf += 64;
    len -= 64;
  }
  x0 = k3k4; /* four lanes -> one */
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
  while (len >= 16) { /* single 16-byte blocks */
    x2 = _mm_loadu_si128((const __m128i *)buf);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    buf += 16;
    len -= 16;
  }
  x2 = _mm_clmulepi64_si128(x1, x0, 0x10); /* 128 -> 64 bits */
  x3 = _mm_setr_epi32(~0, 0, ~0, 0);
  x1 = _mm_srli_si128(x1, 8);
  x1 = _mm_xor_si128(x1, x2);
  x0 = k5k0;
  x2 = _mm_srli_si128(x1, 4);
  x1 = _mm_and_si128(x1, x3);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  x0 = poly; /* Barrett reduction to 32 bits */
  x2 = _mm_and_si128(x1, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
  x2 = _mm_and_si128(x2, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}
static int g_have_clmul = -1;
static int have_clmul(void) {
  if (g_have_clmul < 0) {
    const char *off = getenv("MIRGE_B200_PGZ_NO_CLMUL");
    __builtin_cpu_init();
    g_have_clmul = (__builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1") && !(off && off[0] == '1')) ? 1 : 0;
  }
  return g_have_clmul;
}
#else
static int have_clmul(void) { return 0; }
static uint32_t crc32_clmul(uint32_t crc, const uint8_t *buf, uint64_t len) { (void)buf; (void)len; return crc; }
#endif

static uint32_t crc32_buf(uint32_t crc, const uint8_t *p, uint64_t n) {
  crc = ~crc;
  if (n >= 64 && have_clmul()) {
    const uint64_t k = n & ~15ull;
    crc = crc32_clmul(crc, p, k);
    p += k;
    n -= k;
  }
  while (n && ((uintptr_t)p & 7)) { crc = g_crc[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8); --n; }
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    v ^= crc;
    crc = g_crc[7][v & 0xFF] ^ g_crc[6][(v >> 8) & 0xFF] ^ g_crc[5][(v >> 16) & 0xFF] ^ g_crc[4][(v >> 24) & 0xFF] ^
          g_crc[3][(v >> 32) & 0xFF] ^ g_crc[2][(v >> 40) & 0xFF] ^ g_crc[1][(v >> 48) & 0xFF] ^ g_crc[0][v >> 56];
    p += 8;
    n -= 8;
  }
  while (n--) crc = g_crc[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
  return ~crc;
}
/* crc of A ++ B from crc(A), crc(B), len(B): multiplication by x^(8 len) in GF(2)[x] mod the CRC polynomial */
static uint32_t gf2_times(const uint32_t *mat, uint32_t vec) {
  uint32_t s = 0;
  for (int i = 0; vec; vec >>= 1, ++i)
    if (vec & 1) s ^= mat[i];
  return s;
}
static void gf2_square(uint32_t *sq, const uint32_t *mat) {
  for (int i = 0; i < 32; ++i) sq[i] = gf2_times(mat, mat[i]);
}
static uint32_t crc32_comb(uint32_t crc1, uint32_t crc2, uint64_t len2) {
  if (len2 == 0) return crc1;
  uint32_t even[32], odd[32];
  odd[0] = 0xEDB88320u;
  uint32_t row = 1;
  for (int i = 1; i < 32; ++i) { odd[i] = row; row <<= 1; }
  gf2_square(even, odd);
  gf2_square(odd, even);
  do {
    gf2_square(even, odd);
    if (len2 & 1) crc1 = gf2_times(even, crc1);
    len2 >>= 1;
    if (!len2) break;
    gf2_square(odd, even);
    if (len2 & 1) crc1 = gf2_times(odd, crc1);
    len2 >>= 1;
  } while (len2);
  return crc1 ^ crc2;
}

/* ------------------------------------------------------------------ the stream ------- */
typedef struct {
  const uint8_t *data;
  uint64_t n;
  int threads;
  int wave_factor;
  uint64_t chunk_bytes;
  /* position: next wave starts at bit `pos` inside a member (in_member) or at byte `pos / 8` between members */
  uint64_t pos;
  int in_member;
  int eof;
  uint8_t window[WIN]; /* last WIN bytes of the output so far (zero-padded at the front while shorter) */
  uint64_t member_len; /* bytes of the current member so far */
  uint32_t member_crc;
  /* output of the last wave */
  chunk *out;
  int n_out, cur_out;
  uint64_t cur_off;
  uint64_t total_out;
  /* buffers kept between waves (fresh allocations of this size are page faults, several per output page) */
  void *pool[2][POOL_MAX];
  uint64_t pool_cap[2][POOL_MAX];
  int pool_n[2];
  /* statistics */
  uint64_t waves, chunks_used, chunks_wasted, serial_repairs;
  double t_find, t_decode, t_chain, t_resolve;
  char err[256];
  /* The reader's side.  With more than one thread the waves are decoded by a producer thread one wave ahead of the
   * reader: while pgz_read copies wave k out (a serial memcpy during which the decoding threads would idle), wave k + 1
   * is being decoded; `ready` is the hand-over slot between the two.  Everything above belongs to whoever runs
   * run_wave (the producer thread, or the reader itself when there is none); the pools are shared under `mu`. */
  chunk *c_out;
  int c_n, c_cur, c_eof;
  uint64_t c_off;
  int async, started, stop, failed;
  pthread_t producer;
  pthread_mutex_t mu;
  pthread_cond_t cv_data, cv_space;
  chunk *ready_out;
  int ready_n, ready_rc, ready_eof, ready_full;
} pgz;

static int fail(pgz *z, const char *msg, uint64_t at) {
  snprintf(z->err, sizeof(z->err), "%s (compressed byte %llu)", msg, (unsigned long long)at);
  return -1;
}

/* which: 0 = symbol buffers (cap in symbols), 1 = byte buffers.  Returns a buffer of at least `need` units or NULL. */
static void *pool_take(pgz *z, int which, uint64_t need, uint64_t *cap) {
  const uint64_t unit = which == 0 ? sizeof(uint16_t) : 1;
  pthread_mutex_lock(&z->mu);
  if (z->pool_n[which] > 0) {
    const int i = --z->pool_n[which];
    void *p = z->pool[which][i];
    uint64_t c = z->pool_cap[which][i];
    pthread_mutex_unlock(&z->mu);
    if (c < need) {
      void *q = realloc(p, need * unit);
      if (!q) { free(p); return NULL; }
      p = q;
      c = need;
    }
    *cap = c;
    return p;
  }
  pthread_mutex_unlock(&z->mu);
  *cap = need;
  return malloc(need ? need * unit : 1);
}
static void pool_give(pgz *z, int which, void *p, uint64_t cap) {
  if (!p) return;
  pthread_mutex_lock(&z->mu);
  if (z->pool_n[which] < POOL_MAX) {
    z->pool[which][z->pool_n[which]] = p;
    z->pool_cap[which][z->pool_n[which]] = cap;
    z->pool_n[which]++;
    p = NULL;
  }
  pthread_mutex_unlock(&z->mu);
  free(p);
}

/* gzip member header at byte `at` (RFC 1952); returns the byte after it, 0 on error */
static uint64_t gzip_header(pgz *z, uint64_t at) {
  const uint8_t *d = z->data;
  const uint64_t n = z->n;
  if (at + 10 > n) { fail(z, "truncated gzip header", at); return 0; }
  if (d[at] != 0x1f || d[at + 1] != 0x8b) { fail(z, "not a gzip member (bad magic number)", at); return 0; }
  if (d[at + 2] != 8) { fail(z, "unknown compression method", at); return 0; }
  const int flg = d[at + 3];
  uint64_t p = at + 10;
  if (flg & 4) {
    if (p + 2 > n) { fail(z, "truncated gzip header", at); return 0; }
    p += 2 + (uint64_t)(d[p] | (d[p + 1] << 8));
  }
  if (flg & 8) { while (p < n && d[p]) ++p; ++p; }
  if (flg & 16) { while (p < n && d[p]) ++p; ++p; }
  if (flg & 2) p += 2;
  if (p > n) { fail(z, "truncated gzip header", at); return 0; }
  return p;
}

static void chunk_release(pgz *z, chunk *c) { /* buffers back to the pools */
  pool_give(z, 0, c->sym, c->cap);
  pool_give(z, 1, c->bytes, c->bytes_cap);
  c->sym = NULL;
  c->bytes = NULL;
}

static void prefix_unknown(chunk *c) {
  for (int i = 0; i < WIN; ++i) c->sym[i] = (uint16_t)(256 + i);
  c->known = 0;
}
static void prefix_known(chunk *c, const uint8_t *win, int empty) {
  memcpy(c->bytes, win, WIN);
  c->known = empty ? 2 : 1;
}

/* speculative: a symbol buffer for the head (grown on demand: a chunk that never gets marker-free stays in it) and the
 * byte buffer; known window: the byte buffer only */
static int chunk_alloc(pgz *z, chunk *c, uint64_t est_out, int speculative) {
  memset(c, 0, sizeof(*c));
  if (speculative) {
    c->sym = (uint16_t *)pool_take(z, 0, WIN + (est_out < (1u << 20) ? est_out : (1u << 20)) + 1024, &c->cap);
    if (!c->sym) return -1;
  }
  c->bytes = (uint8_t *)pool_take(z, 1, WIN + est_out + 1024, &c->bytes_cap);
  return c->bytes ? 0 : -1;
}

/* byte i of a chunk's output before / after its head has been resolved */
static inline uint8_t chunk_byte(const chunk *a, const uint8_t *win, uint64_t i) {
  if (i >= a->n_sym) return a->bytes[a->boff + i];
  const uint16_t t = a->sym[WIN + i];
  return t < 256 ? (uint8_t)t : win[t - 256];
}

/* One wave: up to `threads` chunks decoded in parallel from z->pos on.  Fills z->out. */
static int run_wave(pgz *z) {
  const uint8_t *data = z->data;
  const uint64_t n = z->n;
  for (int i = 0; i < z->n_out; ++i) chunk_release(z, &z->out[i]);
  free(z->out);
  z->out = NULL;
  z->n_out = z->cur_out = 0;
  z->cur_off = 0;
  if (!z->in_member) {
    uint64_t at = z->pos >> 3;
    while (at < n && data[at] == 0) ++at; /* zero padding between / after members */
    if (at >= n) { z->eof = 1; return 0; }
    const uint64_t body = gzip_header(z, at);
    if (!body) return -1;
    z->pos = body * 8;
    z->in_member = 1;
    z->member_len = 0;
    z->member_crc = 0;
    memset(z->window, 0, WIN);
    /* (the window of a new member is empty: marked by known = 2 on its first chunk) */
  }
  const int new_member = z->member_len == 0;
  const int T = z->threads;
  const int K = T == 1 ? 1 : T * z->wave_factor; /* chunks per wave: more than threads, so that unequal chunks even out */
  const uint64_t S = z->chunk_bytes;
  const uint64_t w0 = z->pos >> 3;
  uint64_t w1 = w0 + (uint64_t)K * S;
  if (w1 > n) w1 = n;
  const uint64_t est = S * 8;
  /* 1. candidate starts near every cut */
  uint64_t *cand = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(K + 1));
  chunk *ch = (chunk *)calloc((size_t)K + 1, sizeof(chunk));
  if (!cand || !ch) { free(cand); free(ch); return fail(z, "out of memory", w0); }
  int n_cut = 0;
  for (int k = 1; k < K; ++k)
    if (w0 + (uint64_t)k * S + 64 < w1) n_cut = k;
  uint64_t *found = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n_cut + 1));
  if (!found) { free(cand); free(ch); return fail(z, "out of memory", w0); }
  double t0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
  for (int k = 1; k <= n_cut; ++k) {
    const uint64_t from = (w0 + (uint64_t)k * S) * 8;
    uint64_t to = from + S * 8;
    if (to > w1 * 8) to = w1 * 8;
    found[k] = find_block(data, n, from, to);
  }
  int n_cand = 0;
  for (int k = 1; k <= n_cut; ++k)
    if (found[k] != UINT64_MAX && (n_cand == 0 || found[k] > cand[n_cand - 1])) cand[n_cand++] = found[k];
  free(found);
  z->t_find += omp_get_wtime() - t0; t0 = omp_get_wtime();
  /* 2. decode: chunk 0 from the known position with the known window, chunk j + 1 from candidate j, unknown window */
  const uint64_t limit_bit = w1 < n ? w1 * 8 : UINT64_MAX; /* the last wave runs to the final block */
  int alloc_fail = 0;
  for (int k = 0; k <= n_cand; ++k)
    if (chunk_alloc(z, &ch[k], est, k != 0)) alloc_fail = 1;
  if (alloc_fail) {
    for (int k = 0; k <= n_cand; ++k) chunk_release(z, &ch[k]);
    free(cand); free(ch);
    return fail(z, "out of memory", w0);
  }
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
  for (int k = 0; k <= n_cand; ++k) {
    chunk *c = &ch[k];
    if (k == 0) {
      c->start_bit = z->pos;
      prefix_known(c, z->window, new_member);
      chunk_decode(c, data, n, cand, n_cand, 0, limit_bit);
    } else {
      c->start_bit = cand[k - 1];
      c->max_out = 64 * S + (64u << 20);
      prefix_unknown(c);
      chunk_decode(c, data, n, cand, n_cand, k, limit_bit);
    }
  }
  z->t_decode += omp_get_wtime() - t0; t0 = omp_get_wtime();
  /* 3. the chain: which chunks are real; a speculative chunk that failed is re-decoded here from where the chain
   *    stands (its candidate was a real boundary, the chain landed on it, so the failure is the stream's) */
  chunk *acc = (chunk *)calloc((size_t)n_cand + 2, sizeof(chunk));
  if (!acc) { for (int k = 0; k <= n_cand; ++k) chunk_release(z, &ch[k]); free(cand); free(ch); return fail(z, "out of memory", w0); }
  int n_acc = 0, k = 0, rc = 0, ended_final = 0;
  uint64_t next_pos = z->pos;
  for (;;) {
    chunk *c = &ch[k];
    if (c->end_kind == END_ERROR && c->err == -20) {
      /* the chain landed on this start, so it is real: decode it again without the speculative output limit */
      c->max_out = 0;
      prefix_unknown(c);
      chunk_decode(c, data, n, cand, n_cand, k, limit_bit);
      z->serial_repairs++;
    }
    if (c->end_kind == END_ERROR) {
      /* everything before this chunk is accepted and its start is a real block boundary: the error is the stream's
       * (decoding with an unknown window fails exactly where decoding with the real one would) */
      rc = fail(z, c->err == -18 || c->err == -13 ? "compressed file ended before the end-of-stream marker was reached"
                                                  : "invalid deflate data", c->end_bit >> 3);
      break;
    }
    acc[n_acc++] = *c;
    memset(c, 0, sizeof(*c)); /* ownership moved */
    chunk *a = &acc[n_acc - 1];
    next_pos = a->end_bit;
    if (a->end_kind == END_HIT) { z->chunks_wasted += (uint64_t)(a->hit - k); k = a->hit + 1; continue; } /* chunks k+1..hit: bogus starts */
    if (a->end_kind == END_FINAL) ended_final = 1;
    break;
  }
  for (int s = 0; s <= n_cand; ++s) chunk_release(z, &ch[s]); /* whatever the chain did not take (a failed first chunk has no symbol buffer) */
  free(ch);
  free(cand);
  if (rc) { for (int s = 0; s < n_acc; ++s) chunk_release(z, &acc[s]); free(acc); return rc; }
  /* 4. windows (serial over the tails), then markers -> bytes and CRC in parallel */
  uint8_t *wins = (uint8_t *)malloc((size_t)WIN * (size_t)(n_acc + 1));
  if (!wins) { for (int s = 0; s < n_acc; ++s) chunk_release(z, &acc[s]); free(acc); return fail(z, "out of memory", w0); }
  memcpy(wins, z->window, WIN);
  for (int s = 0; s < n_acc; ++s) {
    const uint8_t *win = wins + (size_t)WIN * s;
    uint8_t *nw = wins + (size_t)WIN * (s + 1);
    const chunk *a = &acc[s];
    if (a->n >= WIN) {
      if (a->n - a->n_sym >= WIN) memcpy(nw, a->bytes + a->boff + (a->n - WIN), WIN); /* the tail is final already */
      else for (uint64_t i = 0; i < WIN; ++i) nw[i] = chunk_byte(a, win, a->n - WIN + i);
    } else {
      const uint64_t keep = WIN - a->n;
      memcpy(nw, win + a->n, keep);
      for (uint64_t i = 0; i < a->n; ++i) nw[keep + i] = chunk_byte(a, win, i);
    }
  }
  z->t_chain += omp_get_wtime() - t0; t0 = omp_get_wtime();
  int oom = 0;
  for (int s = 0; s < n_acc; ++s) /* (a chunk that stayed in symbol mode needs its whole output resolved into bytes) */
    if (acc[s].n_sym && bytes_reserve(&acc[s], acc[s].boff + acc[s].n + 8)) oom = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
  for (int s = 0; s < n_acc; ++s) {
    chunk *a = &acc[s];
    if (oom) continue;
    const uint8_t *win = wins + (size_t)WIN * s;
    const uint16_t *t = a->sym ? a->sym + WIN : NULL;
    uint8_t *o = a->bytes + a->boff;
    uint32_t crc = 0;
    /* symbol -> byte as one table (literals map to themselves, markers to the window): groups that hold markers are
     * translated without a branch per symbol -- on FASTQ the markers of header and separator bytes never die out */
    uint8_t lut[256 + WIN];
    if (a->n_sym) {
      for (int q = 0; q < 256; ++q) lut[q] = (uint8_t)q;
      memcpy(lut + 256, win, WIN);
    }
    for (uint64_t blk = 0; blk < a->n; blk += 32768) { /* block-wise: the CRC reads the bytes while they are in cache */
      const uint64_t blk_end = blk + 32768 < a->n ? blk + 32768 : a->n;
      const uint64_t head_end = blk_end < a->n_sym ? blk_end : a->n_sym; /* symbols of this block still to resolve */
      uint64_t i = blk;
      for (; i + 8 <= head_end; i += 8) { /* eight symbols at a time; markers are rare beyond a chunk's first 32 KB */
        uint64_t lo, hi;
        memcpy(&lo, t + i, 8);
        memcpy(&hi, t + i + 4, 8);
        if (((lo | hi) & 0xFF00FF00FF00FF00ull) == 0) {
          lo = (lo | (lo >> 8)) & 0x0000FFFF0000FFFFull;
          lo = (lo | (lo >> 16)) & 0x00000000FFFFFFFFull;
          hi = (hi | (hi >> 8)) & 0x0000FFFF0000FFFFull;
          hi = (hi | (hi >> 16)) & 0x00000000FFFFFFFFull;
          const uint64_t v = lo | (hi << 32);
          memcpy(o + i, &v, 8);
        } else {
          for (int q = 0; q < 8; ++q) o[i + q] = lut[t[i + q]];
        }
      }
      for (; i < head_end; ++i) o[i] = lut[t[i]];
      crc = crc32_buf(crc, o + blk, blk_end - blk);
    }
    a->crc = crc;
  }
  for (int s = 0; s < n_acc; ++s) { pool_give(z, 0, acc[s].sym, acc[s].cap); acc[s].sym = NULL; }
  memcpy(z->window, wins + (size_t)WIN * n_acc, WIN);
  free(wins);
  z->t_resolve += omp_get_wtime() - t0;
  if (oom) { for (int s = 0; s < n_acc; ++s) chunk_release(z, &acc[s]); free(acc); return fail(z, "out of memory", w0); }
  for (int s = 0; s < n_acc; ++s) {
    z->member_crc = crc32_comb(z->member_crc, acc[s].crc, acc[s].n);
    z->member_len += acc[s].n;
  }
  z->chunks_used += (uint64_t)n_acc;
  z->waves++;
  z->out = acc;
  z->n_out = n_acc;
  z->pos = next_pos;
  if (ended_final) {
    /* gzip trailer: CRC-32 and ISIZE, little endian, at the next byte boundary */
    const uint64_t at = (next_pos + 7) >> 3;
    if (at + 8 > n) return fail(z, "compressed file ended before the end-of-stream marker was reached", at);
    uint32_t crc, isz;
    memcpy(&crc, data + at, 4);
    memcpy(&isz, data + at + 4, 4);
    if (crc != z->member_crc) return fail(z, "CRC check failed", at);
    if (isz != (uint32_t)(z->member_len & 0xFFFFFFFFu)) return fail(z, "incorrect length of data produced", at);
    z->pos = (at + 8) * 8;
    z->in_member = 0;
  } else if (w1 >= n) {
    return fail(z, "compressed file ended before the end-of-stream marker was reached", n);
  }
  return 0;
}

/* ------------------------------------------------------------------ C ABI (ctypes: ingest.py) ------- */
__attribute__((constructor)) static void pgz_init_tables(void) {
  crc_tables();
  fixed_tables();
}

void *pgz_open(const uint8_t *data, uint64_t n, int threads, uint64_t chunk_bytes) {
  if (!g_fixed_ready) return NULL;
  pgz *z = (pgz *)calloc(1, sizeof(pgz));
  if (!z) return NULL;
  z->data = data;
  z->n = n;
  z->threads = threads < 1 ? 1 : threads;
  {
    const char *e = getenv("MIRGE_B200_PGZ_WAVE_FACTOR");
    z->wave_factor = e ? atoi(e) : 1; /* (2 and 4 measured: no gain on 8 threads, twice / four times the buffers) */
    if (z->wave_factor < 1 || z->wave_factor > 16) z->wave_factor = 1;
  }
  z->chunk_bytes = chunk_bytes < 65536 ? 65536 : chunk_bytes;
  pthread_mutex_init(&z->mu, NULL);
  pthread_cond_init(&z->cv_data, NULL);
  pthread_cond_init(&z->cv_space, NULL);
  {
    const char *e = getenv("MIRGE_B200_PGZ_PREFETCH"); /* 0: decode a wave only when the reader asks for it */
    z->async = z->threads > 1 && !(e && e[0] == '0');
  }
  return z;
}

/* the producer thread: one wave ahead of the reader, never more (two waves of buffers exist at a time) */
static void *producer_main(void *arg) {
  pgz *z = (pgz *)arg;
  for (;;) {
    pthread_mutex_lock(&z->mu);
    while (z->ready_full && !z->stop) pthread_cond_wait(&z->cv_space, &z->mu);
    const int stop = z->stop;
    pthread_mutex_unlock(&z->mu);
    if (stop) break;
    const int rc = run_wave(z);
    pthread_mutex_lock(&z->mu);
    z->ready_out = z->out;
    z->ready_n = z->n_out;
    z->ready_rc = rc;
    z->ready_eof = z->eof;
    z->out = NULL;
    z->n_out = 0;
    z->ready_full = 1;
    pthread_cond_signal(&z->cv_data);
    pthread_mutex_unlock(&z->mu);
    if (rc || z->eof) break;
  }
  return NULL;
}

/* the next wave into the reader's hands; < 0: the wave failed (z->err) */
static int next_wave(pgz *z) {
  int rc;
  if (!z->async) {
    rc = run_wave(z);
    z->c_out = z->out;
    z->c_n = z->n_out;
    z->c_eof = z->eof;
    z->out = NULL;
    z->n_out = 0;
    return rc;
  }
  if (!z->started) {
    if (pthread_create(&z->producer, NULL, producer_main, z) != 0) { /* no thread to be had: decode on demand */
      z->async = 0;
      return next_wave(z);
    }
    z->started = 1;
  }
  pthread_mutex_lock(&z->mu);
  while (!z->ready_full) pthread_cond_wait(&z->cv_data, &z->mu);
  z->c_out = z->ready_out;
  z->c_n = z->ready_n;
  z->c_eof = z->ready_eof;
  rc = z->ready_rc;
  z->ready_out = NULL;
  z->ready_n = 0;
  z->ready_full = 0;
  pthread_cond_signal(&z->cv_space);
  pthread_mutex_unlock(&z->mu);
  return rc;
}

/* next decompressed bytes into dst (at most cap): > 0 bytes written, 0 end of file, < 0 error (pgz_error) */
int64_t pgz_read(void *h, uint8_t *dst, uint64_t cap) {
  pgz *z = (pgz *)h;
  uint64_t got = 0;
  if (z->failed) return -1;
  while (got < cap) {
    if (z->c_cur < z->c_n) {
      chunk *a = &z->c_out[z->c_cur];
      uint64_t k = a->n - z->c_off;
      if (k > cap - got) k = cap - got;
      memcpy(dst + got, a->bytes + a->boff + z->c_off, k);
      got += k;
      z->c_off += k;
      if (z->c_off == a->n) { chunk_release(z, a); z->c_cur++; z->c_off = 0; }
      continue;
    }
    if (z->c_out) { /* the wave is drained */
      for (int i = 0; i < z->c_n; ++i) chunk_release(z, &z->c_out[i]);
      free(z->c_out);
      z->c_out = NULL;
      z->c_n = z->c_cur = 0;
      z->c_off = 0;
    }
    if (z->c_eof) break;
    if (got) break; /* hand over what is there before the next wave blocks */
    if (next_wave(z)) { /* the wave failed: nothing of it is handed out, and every later call fails as well */
      for (int i = 0; i < z->c_n; ++i) chunk_release(z, &z->c_out[i]);
      free(z->c_out);
      z->c_out = NULL;
      z->c_n = z->c_cur = 0;
      z->failed = 1;
      return -1;
    }
  }
  z->total_out += got;
  return (int64_t)got;
}

const char *pgz_error(void *h) { return ((pgz *)h)->err; }

void pgz_stats(void *h, uint64_t *out4) {
  pgz *z = (pgz *)h;
  out4[0] = z->waves; out4[1] = z->chunks_used; out4[2] = z->chunks_wasted; out4[3] = z->total_out;
}

/* seconds spent looking for block starts, decoding, chaining windows, resolving markers + CRC */
void pgz_times(void *h, double *out4) {
  pgz *z = (pgz *)h;
  out4[0] = z->t_find; out4[1] = z->t_decode; out4[2] = z->t_chain; out4[3] = z->t_resolve;
}

uint32_t pgz_crc(uint32_t crc, const uint8_t *p, uint64_t n) {
  return crc32_buf(crc, p, n); /* (the tables are filled when the library is loaded) */
}

void pgz_close(void *h) {
  pgz *z = (pgz *)h;
  if (!z) return;
  if (z->started) {
    pthread_mutex_lock(&z->mu);
    z->stop = 1;
    pthread_cond_signal(&z->cv_space);
    pthread_mutex_unlock(&z->mu);
    pthread_join(z->producer, NULL);
  }
  for (int i = 0; i < z->n_out; ++i) chunk_release(z, &z->out[i]);
  free(z->out);
  for (int i = 0; i < z->ready_n; ++i) chunk_release(z, &z->ready_out[i]);
  free(z->ready_out);
  for (int i = 0; i < z->c_n; ++i) chunk_release(z, &z->c_out[i]);
  free(z->c_out);
  for (int w = 0; w < 2; ++w)
    for (int i = 0; i < z->pool_n[w]; ++i) free(z->pool[w][i]);
  pthread_mutex_destroy(&z->mu);
  pthread_cond_destroy(&z->cv_data);
  pthread_cond_destroy(&z->cv_space);
  free(z);
}
