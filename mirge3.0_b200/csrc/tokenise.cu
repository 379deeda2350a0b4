// tokenise.cu -- FASTQ tokeniser: line-break census + record/line index.
// Replaces dnaio.read_chunks' record-boundary search and the line splitting of dnaio's FastqIter
// (reference call sites: mirge/libs/digest.py:140 and :324).
//
// Layout: the byte stream is cut into tiles of TOK_TILE bytes (one CTA each, 32 contiguous bytes
// per thread, two coalesced 128-bit loads).  Pass 1 counts '\n' per tile and per group of
// TOK_GROUP tiles and keeps every thread's 32-bit newline mask (1/8 of the data); a one-CTA scan turns group
// totals into prefixes; pass 2 reads the masks instead of the bytes and scatters the start offset
// of every line: line_start[g + 1] = position after the g-th '\n'.  Four consecutive entries are
// the header / sequence / '+' / quality line of one record, so the trim kernel fetches one
// aligned uint4 per read.
#include "common.cuh"

#define TOK_TILE 8192
#define TOK_THREADS 256
#define TOK_GROUP 1024

struct TokScratch {
  uint32_t *masks;  // newline mask of every 32-byte piece (one per thread of pass 1)
  uint32_t *tile_excl;  // '\n' before the tile inside its group
  uint32_t *tile_counts;
  uint32_t *group_totals;
  uint64_t *group_prefix;
  uint64_t n_tiles, n_groups;
};

static inline uint64_t align16(uint64_t x) { return (x + 15) & ~15ull; }

static TokScratch tok_layout(void *scratch, uint64_t nbytes) {
  TokScratch s;
  s.n_tiles = (nbytes + TOK_TILE - 1) / TOK_TILE;
  s.n_groups = (s.n_tiles + TOK_GROUP - 1) / TOK_GROUP;
  uint8_t *p = (uint8_t *)scratch;
  s.masks = (uint32_t *)p;
  p += align16(s.n_tiles * TOK_THREADS * 4);
  s.tile_excl = (uint32_t *)p;
  p += align16(s.n_tiles * 4 + 4);
  s.tile_counts = (uint32_t *)p;
  p += align16(s.n_tiles * 4 + 4);
  s.group_totals = (uint32_t *)p;
  p += align16(s.n_groups * 4 + 4);
  s.group_prefix = (uint64_t *)p;
  return s;
}

extern "C" uint64_t mirge_tokenise_scratch_bytes(uint64_t nbytes) {
  nbytes += 16;  // alignment skew
  uint64_t n_tiles = (nbytes + TOK_TILE - 1) / TOK_TILE;
  uint64_t n_groups = (n_tiles + TOK_GROUP - 1) / TOK_GROUP;
  return align16(n_tiles * TOK_THREADS * 4) + 2 * align16(n_tiles * 4 + 4) + align16(n_groups * 4 + 4) + align16((n_groups + 1) * 8) + 64;
}

// 32 bytes of the stream starting at pos -> 8 words (zero filled past n)
__device__ __forceinline__ void load32(const uint8_t *fq, uint64_t n, uint64_t pos, uint32_t w[8]) {
  if (pos + 32 <= n) {
    uint4 a = ld_stream_u4(fq + pos), b = ld_stream_u4(fq + pos + 16);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint32_t v = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        uint64_t q = pos + 4 * k + b;
        if (q < n) v |= (uint32_t)fq[q] << (8 * b);
      }
      w[k] = v;
    }
  }
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *smem, uint32_t *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t s = lane < (TOK_THREADS / 32) ? smem[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += t;
    }
    if (lane < (TOK_THREADS / 32)) smem[lane] = s;
  }
  __syncthreads();
  uint32_t base = warp ? smem[warp - 1] : 0;
  *total = smem[TOK_THREADS / 32 - 1];
  return base + inc - v;
}

// Pass 1: one CTA per TOK_PAIR consecutive tiles; the loads of both tiles (4 x 128 bits per thread) are issued before
// any of them is used.
#define TOK_PAIR 2
static_assert(TOK_GROUP % TOK_PAIR == 0, "a tile pair must not straddle two groups");
__global__ void __launch_bounds__(TOK_THREADS) tok_count_kernel(const uint8_t *__restrict__ fq, uint64_t n, uint32_t skew,
                                                                 uint32_t *__restrict__ masks, uint32_t *__restrict__ tile_counts,
                                                                 uint32_t *__restrict__ group_totals, uint64_t n_tiles) {
  __shared__ uint32_t sm[TOK_PAIR][TOK_THREADS / 32];
  const uint64_t tile0 = (uint64_t)blockIdx.x * TOK_PAIR;
  uint32_t w[TOK_PAIR][8], mk[TOK_PAIR], c[TOK_PAIR];
  const uint64_t pos0 = tile0 * TOK_TILE + (uint64_t)threadIdx.x * 32;
  if (pos0 + (TOK_PAIR - 1) * TOK_TILE + 32 <= n) {  // all pieces complete: straight-line loads
#pragma unroll
    for (int k = 0; k < TOK_PAIR; ++k) {
      const uint4 a = ld_stream_u4(fq + pos0 + k * TOK_TILE), b = ld_stream_u4(fq + pos0 + k * TOK_TILE + 16);
      w[k][0] = a.x; w[k][1] = a.y; w[k][2] = a.z; w[k][3] = a.w;
      w[k][4] = b.x; w[k][5] = b.y; w[k][6] = b.z; w[k][7] = b.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < TOK_PAIR; ++k) {
      const uint64_t pos = pos0 + (uint64_t)k * TOK_TILE;
      if (pos < n) load32(fq, n, pos, w[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < TOK_PAIR; ++k) {
    const uint64_t pos = (tile0 + k) * TOK_TILE + (uint64_t)threadIdx.x * 32;
    mk[k] = 0;
    if (pos < n) {
      mk[k] = newline_mask(w[k]);
      if (pos == 0) mk[k] &= ~((1u << skew) - 1u);  // bytes before the stream start are not ours
    }
    if (tile0 + k < n_tiles) masks[(tile0 + k) * TOK_THREADS + threadIdx.x] = mk[k];
    c[k] = __reduce_add_sync(0xffffffffu, __popc(mk[k]));
    if ((threadIdx.x & 31) == 0) sm[k][threadIdx.x >> 5] = c[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t both = 0;
#pragma unroll
    for (int k = 0; k < TOK_PAIR; ++k) {
      uint32_t t = 0;
#pragma unroll
      for (int i = 0; i < TOK_THREADS / 32; ++i) t += sm[k][i];
      if (tile0 + k < n_tiles) tile_counts[tile0 + k] = t;
      both += t;
    }
    atomicAdd(&group_totals[tile0 / TOK_GROUP], both);
  }
}

__global__ void tok_scan_groups_kernel(const uint32_t *__restrict__ group_totals, uint64_t n_groups,
                                       uint64_t *__restrict__ group_prefix, uint64_t *__restrict__ out_total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint64_t s = 0;
    for (uint64_t g = 0; g < n_groups; ++g) {
      group_prefix[g] = s;
      s += group_totals[g];
    }
    group_prefix[n_groups] = s;
    *out_total = s;
  }
}

// exclusive prefix of the tile counts inside every group of TOK_GROUP tiles (one CTA per group), so that pass 2
// finds the number of '\n' before its tile with two loads
__global__ void __launch_bounds__(TOK_THREADS) tok_scan_tiles_kernel(const uint32_t *__restrict__ tile_counts, uint64_t n_tiles,
                                                                      uint32_t *__restrict__ tile_excl) {
  __shared__ uint32_t sm[TOK_THREADS / 32];
  const uint64_t first = (uint64_t)blockIdx.x * TOK_GROUP + (uint64_t)threadIdx.x * (TOK_GROUP / TOK_THREADS);
  uint32_t c[TOK_GROUP / TOK_THREADS], mine = 0;
#pragma unroll
  for (int k = 0; k < TOK_GROUP / TOK_THREADS; ++k) {
    c[k] = first + k < n_tiles ? tile_counts[first + k] : 0u;
    mine += c[k];
  }
  uint32_t total;
  uint32_t run = block_exclusive_scan(mine, sm, &total);
#pragma unroll
  for (int k = 0; k < TOK_GROUP / TOK_THREADS; ++k) {
    if (first + k < n_tiles) tile_excl[first + k] = run;
    run += c[k];
  }
}

// Pass 2: one CTA per TOK_SUPER consecutive tiles (they share a group: TOK_SUPER divides TOK_GROUP), one 128-bit load of
// four masks (128 bytes of the stream) per thread, one block scan, then the line starts of those bytes.
#define TOK_SUPER 4
static_assert(TOK_GROUP % TOK_SUPER == 0, "a super-tile must not straddle two groups");
__global__ void __launch_bounds__(TOK_THREADS) tok_index_kernel(const uint32_t *__restrict__ masks, uint64_t n,
                                                                 const uint32_t *__restrict__ tile_excl,
                                                                 const uint64_t *__restrict__ group_prefix,
                                                                 uint32_t *__restrict__ line_start, uint64_t max_lines,
                                                                 uint32_t skew) {
  __shared__ uint32_t sm[TOK_THREADS / 32];
  const uint64_t n_tiles = (n + TOK_TILE - 1) / TOK_TILE;
  const uint64_t tile0 = (uint64_t)blockIdx.x * TOK_SUPER;
  const uint64_t before = group_prefix[tile0 / TOK_GROUP] + tile_excl[tile0];
  const uint64_t w0 = tile0 * TOK_THREADS + 4ull * threadIdx.x;  // first of this thread's four mask words
  uint4 m4 = make_uint4(0, 0, 0, 0);
  if (w0 < n_tiles * TOK_THREADS) m4 = *(const uint4 *)(masks + w0);  // written by pass 1 (zero past the end of the stream)
  const uint32_t mk[4] = {m4.x, m4.y, m4.z, m4.w};
  uint32_t total;
  const uint32_t ex = block_exclusive_scan(__popc(m4.x) + __popc(m4.y) + __popc(m4.z) + __popc(m4.w), sm, &total);
  uint64_t g = before + ex;  // ordinal (0-based) of this thread's first '\n'
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t pos = (w0 + k) * 32;
    uint32_t mask = mk[k];
    while (mask) {
      const int b = __ffs(mask) - 1;
      mask &= mask - 1;
      if (g + 1 <= max_lines) line_start[g + 1] = (uint32_t)(pos + b + 1);
      ++g;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = skew;
}

__global__ void tok_fixup_kernel(uint32_t *last, uint32_t virtual_end) {
  if (*last == 0) *last = virtual_end;
}

// position just after the target-th '\n' (1-based ordinal); 0 when target == 0
__global__ void __launch_bounds__(TOK_THREADS) tok_find_kernel(const uint32_t *__restrict__ masks, uint64_t n,
                                                                const uint32_t *__restrict__ tile_counts,
                                                                const uint64_t *__restrict__ group_prefix,
                                                                uint64_t n_groups, uint64_t n_tiles, uint64_t target,
                                                                uint32_t skew, uint64_t *__restrict__ out) {
  __shared__ uint32_t sm[TOK_THREADS / 32];
  __shared__ uint64_t s_tile, s_before;
  if (target == 0) {
    if (threadIdx.x == 0) *out = skew;
    return;
  }
  if (threadIdx.x == 0) {
    uint64_t g = 0;
    while (g + 1 < n_groups && group_prefix[g + 1] < target) ++g;
    uint64_t before = group_prefix[g], t = g * TOK_GROUP;
    while (t + 1 < n_tiles && before + tile_counts[t] < target) {
      before += tile_counts[t];
      ++t;
    }
    s_tile = t;
    s_before = before;
  }
  __syncthreads();
  const uint64_t pos = s_tile * TOK_TILE + (uint64_t)threadIdx.x * 32;
  uint32_t mask = masks[s_tile * TOK_THREADS + threadIdx.x];
  uint32_t total;
  uint32_t ex = block_exclusive_scan(__popc(mask), sm, &total);
  uint64_t g = s_before + ex;
  while (mask) {
    int b = __ffs(mask) - 1;
    mask &= mask - 1;
    ++g;
    if (g == target) *out = pos + b + 1;
  }
}

extern "C" int mirge_tokenise_sync(mirge_ctx *ctx, const uint8_t *d_fastq, uint64_t nbytes, int is_final,
                                   void *d_scratch, uint64_t *n_records, uint64_t *consumed, void *stream_) {
  if (!ctx || !n_records || !consumed) return MIRGE_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  *n_records = 0;
  *consumed = 0;
  if (nbytes == 0) return MIRGE_OK;
  if (!d_fastq || !d_scratch) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "tokenise: null buffer");
  if (nbytes > 0xFFFFFF00ull) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "tokenise: batch larger than 4 GiB - 256");
  MIRGE_CUDA(ctx, cudaSetDevice(ctx->device));
  // kernels work on the 16-byte aligned stream that contains the batch; offsets are relative to it
  const uint32_t skew = (uint32_t)((uintptr_t)d_fastq & 15);
  const uint8_t *fq_al = d_fastq - skew;
  const uint64_t n_al = nbytes + skew;
  TokScratch s = tok_layout(d_scratch, n_al);
  MIRGE_CUDA(ctx, cudaMemsetAsync(s.group_totals, 0, s.n_groups * 4, stream));
  tok_count_kernel<<<(unsigned)((s.n_tiles + TOK_PAIR - 1) / TOK_PAIR), TOK_THREADS, 0, stream>>>(fq_al, n_al, skew, s.masks, s.tile_counts,
                                                                                              s.group_totals, s.n_tiles);
  MIRGE_LAUNCH_CHECK(ctx, "tok_count_kernel");
  tok_scan_groups_kernel<<<1, 32, 0, stream>>>(s.group_totals, s.n_groups, s.group_prefix, ctx->d_small);
  MIRGE_LAUNCH_CHECK(ctx, "tok_scan_groups_kernel");
  MIRGE_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_small, 8, cudaMemcpyDeviceToHost, stream));
  MIRGE_CUDA(ctx, cudaMemcpyAsync((uint8_t *)(ctx->h_pinned + 1), d_fastq + nbytes - 1, 1, cudaMemcpyDeviceToHost, stream));
  MIRGE_CUDA(ctx, cudaStreamSynchronize(stream));
  uint64_t n_nl = ctx->h_pinned[0];
  uint8_t last = *(uint8_t *)(ctx->h_pinned + 1);
  if (is_final) {
    uint64_t n_lines = n_nl + (last != '\n' ? 1 : 0);
    if (n_lines % 4 != 0)
      MIRGE_FAIL(ctx, MIRGE_ERR_FORMAT, "FASTQ format error: %llu lines is not a multiple of 4 (premature end of file)",
                 (unsigned long long)n_lines);
    *n_records = n_lines / 4;
    *consumed = nbytes;
    return MIRGE_OK;
  }
  *n_records = n_nl / 4;
  tok_find_kernel<<<1, TOK_THREADS, 0, stream>>>(s.masks, n_al, s.tile_counts, s.group_prefix, s.n_groups, s.n_tiles,
                                                 4 * (*n_records), skew, ctx->d_small + 1);
  MIRGE_LAUNCH_CHECK(ctx, "tok_find_kernel");
  MIRGE_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_small + 1, 8, cudaMemcpyDeviceToHost, stream));
  MIRGE_CUDA(ctx, cudaStreamSynchronize(stream));
  *consumed = ctx->h_pinned[0] - skew;
  return MIRGE_OK;
}

extern "C" int mirge_line_index(mirge_ctx *ctx, const uint8_t *d_fastq, uint64_t nbytes, const void *d_scratch,
                                uint32_t *d_line_start, uint64_t n_records, void *stream_) {
  if (!ctx) return MIRGE_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!d_line_start) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "line_index: null output");
  MIRGE_CUDA(ctx, cudaSetDevice(ctx->device));
  if (nbytes == 0 || n_records == 0) {
    MIRGE_CUDA(ctx, cudaMemsetAsync(d_line_start, 0, 4, stream));
    return MIRGE_OK;
  }
  const uint32_t skew = (uint32_t)((uintptr_t)d_fastq & 15);
  const uint64_t n_al = nbytes + skew;
  TokScratch s = tok_layout((void *)d_scratch, n_al);
  // EOF rule: when the last line has no '\n' the 4n-th line break does not exist; the entry is
  // zeroed first and a fix-up kernel turns a still-zero entry into the virtual break nbytes + 1.
  MIRGE_CUDA(ctx, cudaMemsetAsync(d_line_start + 4 * n_records, 0, 4, stream));
  tok_scan_tiles_kernel<<<(unsigned)s.n_groups, TOK_THREADS, 0, stream>>>(s.tile_counts, s.n_tiles, s.tile_excl);
  MIRGE_LAUNCH_CHECK(ctx, "tok_scan_tiles_kernel");
  tok_index_kernel<<<(unsigned)((s.n_tiles + TOK_SUPER - 1) / TOK_SUPER), TOK_THREADS, 0, stream>>>(s.masks, n_al, s.tile_excl, s.group_prefix,
                                                                    d_line_start, 4 * n_records, skew);
  MIRGE_LAUNCH_CHECK(ctx, "tok_index_kernel");
  tok_fixup_kernel<<<1, 1, 0, stream>>>(d_line_start + 4 * n_records, (uint32_t)(n_al + 1));
  MIRGE_LAUNCH_CHECK(ctx, "tok_fixup_kernel");
  return MIRGE_OK;
}
