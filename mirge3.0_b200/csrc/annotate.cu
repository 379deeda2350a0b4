// annotate.cu -- placeholder while the annotation kernels are being written (replaced next commit)
#include "common.cuh"
extern "C" int mirge_lib_kmers(mirge_ctx *ctx, const mirge_library *, uint32_t *, uint8_t *, void *) {
  if (!ctx) return MIRGE_ERR_ARG;
  MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: not built yet");
}
extern "C" int mirge_annotate_round(mirge_ctx *ctx, const mirge_library *, const mirge_round_policy *, const mirge_table *,
                                    uint64_t, uint8_t *, uint64_t *, void *) {
  if (!ctx) return MIRGE_ERR_ARG;
  MIRGE_FAIL(ctx, MIRGE_ERR_ARG, "annotate: not built yet");
}
