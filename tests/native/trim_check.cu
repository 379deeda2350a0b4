// trim_check.cu -- a native (no Python, no torch) parity check of the trim path through the C ABI of libmirge_b200.so:
// tokenise -> line index -> trim on the device, windows and kept flags compared with what the C oracle computed for the
// same bytes and parameters.  The cases (parameters, FASTQ bytes, expected windows) are written on the build host by
// tests/native/make_trim_cases.py; this binary only reads them, so it needs neither the oracle nor Python on the GPU
// box and finishes in seconds.  Build: tests/native/build.sh.  It is test infrastructure, not part of the product.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mirge_b200.h"

#define CK(x)                                                                           \
  do {                                                                                  \
    cudaError_t e_ = (x);                                                               \
    if (e_ != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);   \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)

template <class F>
static bool sym(void *h, const char *name, F &f) {
  f = (F)dlsym(h, name);
  if (!f) printf("missing symbol %s\n", name);
  return f != nullptr;
}

int main(int argc, char **argv) {
  if (argc < 3) {
    printf("usage: trim_check libmirge_b200.so cases.bin [--dry]\n");
    return 2;
  }
  const bool dry = argc > 3 && strcmp(argv[3], "--dry") == 0;  // read and size-check the cases only (no device needed)
  void *h = dlopen(argv[1], RTLD_NOW);
  if (!h) {
    printf("dlopen: %s\n", dlerror());
    return 2;
  }
  decltype(&mirge_abi_version) p_abi;
  decltype(&mirge_ctx_create) p_create;
  decltype(&mirge_last_error) p_err;
  decltype(&mirge_set_trim_params) p_set;
  decltype(&mirge_trim_slots) p_slots;
  decltype(&mirge_tokenise_scratch_bytes) p_tsb;
  decltype(&mirge_tokenise_sync) p_tok;
  decltype(&mirge_line_index) p_li;
  decltype(&mirge_trim_scratch_bytes) p_trsb;
  decltype(&mirge_trim) p_trim;
  decltype(&mirge_trim_mode) p_mode;
  if (!(sym(h, "mirge_abi_version", p_abi) && sym(h, "mirge_ctx_create", p_create) && sym(h, "mirge_last_error", p_err) &&
        sym(h, "mirge_set_trim_params", p_set) && sym(h, "mirge_trim_slots", p_slots) &&
        sym(h, "mirge_tokenise_scratch_bytes", p_tsb) && sym(h, "mirge_tokenise_sync", p_tok) && sym(h, "mirge_line_index", p_li) &&
        sym(h, "mirge_trim_scratch_bytes", p_trsb) && sym(h, "mirge_trim", p_trim) && sym(h, "mirge_trim_mode", p_mode)))
    return 2;
  if (p_abi() != MIRGE_ABI_VERSION) {
    printf("ABI version %d, header %d\n", p_abi(), MIRGE_ABI_VERSION);
    return 2;
  }
  FILE *f = fopen(argv[2], "rb");
  if (!f) {
    printf("cannot open %s\n", argv[2]);
    return 2;
  }
  char magic[8];
  uint32_t ncases = 0;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "MRGCASE1", 8) != 0 || fread(&ncases, 4, 1, f) != 1) {
    printf("bad case file\n");
    return 2;
  }
  mirge_ctx *ctx = nullptr;
  cudaStream_t st = nullptr;
  if (!dry) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s (sm_%d%d, %d SMs); %u cases\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, ncases);
    if (p_create(0, &ctx) != MIRGE_OK) {
      printf("mirge_ctx_create failed\n");
      return 2;
    }
    CK(cudaStreamCreate(&st));
  }
  int failed = 0;
  for (uint32_t c = 0; c < ncases; ++c) {
    char name[64];
    uint32_t mode = 0, E = 0;
    uint64_t psize = 0, nbytes = 0, n = 0;
    mirge_trim_params prm;
    if (fread(name, 1, 64, f) != 64 || fread(&mode, 4, 1, f) != 1 || fread(&E, 4, 1, f) != 1 || fread(&psize, 8, 1, f) != 1 ||
        psize != sizeof(prm) || fread(&prm, 1, sizeof(prm), f) != sizeof(prm) || fread(&nbytes, 8, 1, f) != 1) {
      printf("case %u: truncated case file (params %llu vs %zu)\n", c, (unsigned long long)psize, sizeof(prm));
      return 2;
    }
    name[63] = 0;
    std::vector<uint8_t> fq(nbytes);
    if (fread(fq.data(), 1, nbytes, f) != nbytes || fread(&n, 8, 1, f) != 1) return 2;
    std::vector<uint16_t> win_o(n * E * 4);
    std::vector<uint8_t> kept_o(n * E);
    if (fread(win_o.data(), 2, win_o.size(), f) != win_o.size() || fread(kept_o.data(), 1, kept_o.size(), f) != kept_o.size()) return 2;

    if (dry) {
      printf("case %-28s mode %u: %llu reads x %u slots, %llu bytes, %d adapters, %d modifiers\n", name, mode, (unsigned long long)n, E,
             (unsigned long long)nbytes, prm.n_adapters, prm.n_mods);
      continue;
    }
    int rc = p_set(ctx, &prm);
    if (rc != MIRGE_OK) {
      printf("FAIL %-28s set_trim_params: %s\n", name, p_err(ctx));
      ++failed;
      continue;
    }
    if ((uint32_t)p_slots(ctx) != E) {
      printf("FAIL %-28s slots %d, expected %u\n", name, p_slots(ctx), E);
      ++failed;
      continue;
    }
    rc = p_mode(ctx, (int)mode);
    uint8_t *d_fq = nullptr;
    void *d_scr = nullptr, *d_slow = nullptr;
    uint32_t *d_ls = nullptr, *d_koff = nullptr, *d_keys = nullptr;
    uint16_t *d_win = nullptr;
    uint64_t *d_ctrl = nullptr, *d_ins = nullptr;
    CK(cudaMalloc(&d_fq, nbytes + 256));
    CK(cudaMemset(d_fq, '\n', nbytes + 256));
    CK(cudaMemcpy(d_fq, fq.data(), nbytes, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_scr, p_tsb(nbytes) + 256));
    uint64_t n_rec = 0, consumed = 0;
    rc = p_tok(ctx, d_fq, nbytes, 1, d_scr, &n_rec, &consumed, st);
    if (rc != MIRGE_OK || n_rec != n) {
      printf("FAIL %-28s tokenise rc=%d records=%llu expected=%llu: %s\n", name, rc, (unsigned long long)n_rec, (unsigned long long)n, p_err(ctx));
      ++failed;
      continue;
    }
    CK(cudaMalloc(&d_ls, (4 * n + 4) * 4 + 256));
    rc = p_li(ctx, d_fq, nbytes, d_scr, d_ls, n, st);
    if (rc != MIRGE_OK) {
      printf("FAIL %-28s line_index: %s\n", name, p_err(ctx));
      ++failed;
      continue;
    }
    const uint64_t cap = (uint64_t)E * (n + consumed / 2 + consumed / 32 + 64) + 4096;  // worst case of device.py
    CK(cudaMalloc(&d_win, n * E * 4 * 2 + 256));
    CK(cudaMalloc(&d_koff, n * E * 4 + 256));
    CK(cudaMalloc(&d_ins, n * E * 8 + 256));
    CK(cudaMalloc(&d_keys, cap * 4));
    CK(cudaMalloc(&d_slow, p_trsb(n) + 256));
    CK(cudaMalloc(&d_ctrl, 16 * 8));
    CK(cudaMemset(d_ctrl, 0, 16 * 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms = 0;
    uint64_t ctrl[16];
    bool dead = false;
    for (int attempt = 0; attempt < 2; ++attempt) {
      CK(cudaMemset(d_ctrl, 0, 16 * 8));
      CK(cudaEventRecord(e0, st));
      rc = p_trim(ctx, d_fq, consumed, d_ls, n, d_win, d_koff, d_keys, cap, d_ctrl, d_slow, d_ins, n * E, st);
      CK(cudaEventRecord(e1, st));
      cudaError_t se = cudaStreamSynchronize(st);
      if (rc != MIRGE_OK || se != cudaSuccess) {
        printf("FAIL %-28s trim rc=%d (%s) cuda=%s\n", name, rc, p_err(ctx), cudaGetErrorString(se));
        if (se != cudaSuccess) return 1;
        dead = true;
        break;
      }
      CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(ctrl, d_ctrl, sizeof(ctrl), cudaMemcpyDeviceToHost));
      if (!(ctrl[2] & 8) || mode == 1) break;
      // a record group did not fit the bit-parallel kernel's staging: the batch is repeated with the generic kernel (device.py)
      ctrl[2] &= ~8ull;
      mode = 1;
      p_mode(ctx, 1);
    }
    if (dead) {
      ++failed;
      continue;
    }
    std::vector<uint16_t> win_g(n * E * 4);
    std::vector<uint32_t> koff(n * E);
    CK(cudaMemcpy(win_g.data(), d_win, win_g.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(koff.data(), d_koff, koff.size() * 4, cudaMemcpyDeviceToHost));
    uint64_t bad = 0, first = ~0ull;
    for (uint64_t e = 0; e < n * E; ++e) {
      const uint8_t kept_g = koff[e] != 0xFFFFFFFFu;
      const bool same = kept_g == kept_o[e] && memcmp(&win_g[4 * e], &win_o[4 * e], 8) == 0;
      if (!same) {
        ++bad;
        if (first == ~0ull) first = e;
      }
    }
    if (ctrl[2] != 0) {
      printf("FAIL %-28s error flags %llu\n", name, (unsigned long long)ctrl[2]);
      ++failed;
    } else if (bad) {
      printf("FAIL %-28s %llu of %llu slots differ; first: record %llu slot %llu gpu=(%u,%u,%u,%u kept %d) oracle=(%u,%u,%u,%u kept %d)\n", name,
             (unsigned long long)bad, (unsigned long long)(n * E), (unsigned long long)(first / E), (unsigned long long)(first % E),
             win_g[4 * first], win_g[4 * first + 1], win_g[4 * first + 2], win_g[4 * first + 3], (int)(koff[first] != 0xFFFFFFFFu),
             win_o[4 * first], win_o[4 * first + 1], win_o[4 * first + 2], win_o[4 * first + 3], (int)kept_o[first]);
      ++failed;
    } else {
      printf("PASS %-28s mode %u: %llu reads x %u slots bit-exact (trim %.3f ms, %llu emitted keys, %llu reads via DP)\n", name, mode,
             (unsigned long long)n, E, ms, (unsigned long long)ctrl[1], (unsigned long long)ctrl[6]);
    }
    cudaFree(d_fq); cudaFree(d_scr); cudaFree(d_ls); cudaFree(d_win); cudaFree(d_koff); cudaFree(d_ins); cudaFree(d_keys);
    cudaFree(d_slow); cudaFree(d_ctrl);
  }
  printf("%s: %u cases, %d failed\n", failed ? "FAILED" : "ALL PASS", ncases, failed);
  return failed ? 1 : 0;
}
