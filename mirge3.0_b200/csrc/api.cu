// api.cu -- context lifecycle of the C ABI (include/mirge_b200.h).
#include "common.cuh"

static char g_create_err[256] = "";

extern "C" int mirge_abi_version(void) { return MIRGE_ABI_VERSION; }

extern "C" int mirge_ctx_create(int device, mirge_ctx **out) {
  if (!out) return MIRGE_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    snprintf(g_create_err, sizeof(g_create_err), "no usable CUDA device %d (%s)", device,
             e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
    return MIRGE_ERR_NODEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MIRGE_ERR_NODEVICE;
  if (prop.major < 10) {
    snprintf(g_create_err, sizeof(g_create_err), "device %d is sm_%d%d; this library is built for sm_100a only",
             device, prop.major, prop.minor);
    return MIRGE_ERR_NODEVICE;
  }
  if (cudaSetDevice(device) != cudaSuccess) return MIRGE_ERR_NODEVICE;
  mirge_ctx *ctx = new mirge_ctx();
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaMallocHost((void **)&ctx->h_pinned, 8 * sizeof(uint64_t)) != cudaSuccess ||
      cudaMalloc((void **)&ctx->d_small, 8 * sizeof(uint64_t)) != cudaSuccess) {
    snprintf(g_create_err, sizeof(g_create_err), "control block allocation failed");
    delete ctx;
    return MIRGE_ERR_CUDA;
  }
  *out = ctx;
  return MIRGE_OK;
}

extern "C" void mirge_ctx_destroy(mirge_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->d_small) cudaFree(ctx->d_small);
  delete ctx;
}

extern "C" const char *mirge_last_error(const mirge_ctx *ctx) { return ctx ? ctx->err : g_create_err; }
