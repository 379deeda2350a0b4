// report.cu -- the per-sample reductions behind annotation.report.csv / miR.Counts.csv
// (mirge/libs/summary.py:692-698: per-library read sums; :716-770: exact-miRNA and isomiR reads grouped by
// miRNA, the inputs of the canonical-ratio filter mirge_can(), summary.py:25-45).
// One pass over the (key id, count) pairs a sample's drain produced; the annotation of a key is looked up in
// the arrays the annotation rounds filled, so the unique-sequence table never goes back to the host for this.
#include "common.cuh"

#define REP_THREADS 256
#define REP_ROUNDS 10

__global__ void __launch_bounds__(REP_THREADS)
report_reduce_kernel(const uint8_t *__restrict__ annot_round, const uint64_t *__restrict__ hit, const uint32_t *__restrict__ ids,
                     const uint32_t *__restrict__ counts, uint64_t n_pairs, uint32_t n_mirna, unsigned long long *__restrict__ round_sum,
                     unsigned long long *__restrict__ can, unsigned long long *__restrict__ iso,
                     unsigned long long *__restrict__ n_bad) {
  __shared__ unsigned long long s_sum[REP_ROUNDS];
  if (threadIdx.x < REP_ROUNDS) s_sum[threadIdx.x] = 0;
  __syncthreads();
  unsigned long long mine[REP_ROUNDS];
#pragma unroll
  for (int r = 0; r < REP_ROUNDS; ++r) mine[r] = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * REP_THREADS + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * REP_THREADS) {
    const uint32_t id = ids[i], c = counts[i];
    const uint32_t r = annot_round[id];
    if (r >= REP_ROUNDS) continue;  // 0xFF: unannotated (pdUnmapped)
#pragma unroll
    for (int q = 0; q < REP_ROUNDS; ++q)
      if (r == (uint32_t)q) mine[q] += c;
    if (r == 0 || r == 8) {
      const uint32_t ref = MIRGE_HIT_REF(hit[id]);
      if (ref < n_mirna) atomicAdd((r == 0 ? can : iso) + ref, (unsigned long long)c);
      else atomicAdd(n_bad, 1ull);
    }
  }
#pragma unroll
  for (int q = 0; q < REP_ROUNDS; ++q) {
    unsigned long long v = mine[q];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_sum[q], v);
  }
  __syncthreads();
  if (threadIdx.x < REP_ROUNDS && s_sum[threadIdx.x]) atomicAdd(round_sum + threadIdx.x, s_sum[threadIdx.x]);
}

extern "C" int mirge_report_reduce(mirge_ctx *ctx, const uint8_t *d_annot_round, const uint64_t *d_hit, const uint32_t *d_ids,
                                   const uint32_t *d_counts, uint64_t n_pairs, uint32_t n_mirna, uint64_t *d_round_sum,
                                   uint64_t *d_can, uint64_t *d_iso, void *stream_) {
  if (!ctx) return MIRGE_ERR_ARG;
  if (n_pairs == 0) return MIRGE_OK;
  if (!d_annot_round || !d_hit || !d_ids || !d_counts || !d_round_sum || !d_can || !d_iso)
    MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "report_reduce: null buffer");
  cudaStream_t stream = (cudaStream_t)stream_;
  MIRGE_CUDA(ctx, cudaSetDevice(ctx->device));
  MIRGE_CUDA(ctx, cudaMemsetAsync(ctx->d_small, 0, 8, stream));
  uint64_t want = (n_pairs + REP_THREADS - 1) / REP_THREADS;
  const uint64_t cap = (uint64_t)ctx->sm_count * 8;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  report_reduce_kernel<<<grid, REP_THREADS, 0, stream>>>(d_annot_round, d_hit, d_ids, d_counts, n_pairs, n_mirna,
                                                         (unsigned long long *)d_round_sum, (unsigned long long *)d_can,
                                                         (unsigned long long *)d_iso, (unsigned long long *)ctx->d_small);
  MIRGE_LAUNCH_CHECK(ctx, "report_reduce_kernel");
  MIRGE_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_small, 8, cudaMemcpyDeviceToHost, stream));
  MIRGE_CUDA(ctx, cudaStreamSynchronize(stream));
  if (ctx->h_pinned[0]) MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "report_reduce: %llu miRNA hits point outside the miRNA library (n_mirna too small)",
                                   (unsigned long long)ctx->h_pinned[0]);
  return MIRGE_OK;
}
