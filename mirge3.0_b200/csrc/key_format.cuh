// key_format.cuh -- slicing of packed keys ([len | n_exc << 16][2-bit payload][exceptions pos << 8 | byte], DESIGN.md
// section 3): the UMI flanks are cut off a first-level key to give the second-level (centre) key of digest.py:164-205,
// exceptions re-based.  Included by collapse.cu; tests/test_key_format_host.py compiles it for the host.
#pragma once
#ifdef KEY_FORMAT_HOST
#include "host_shims.h"
#else
#include "common.cuh"
#endif

__device__ __forceinline__ uint32_t key_base(const uint32_t *key, uint32_t j) { return (key[1 + (j >> 4)] >> (2 * (j & 15))) & 3u; }

// out = key[f : len - b] as a packed key (exceptions re-based); returns its word count
__device__ __forceinline__ uint32_t slice_key(const uint32_t *key, int f, int b, uint32_t *out) {
  const uint32_t hdr = key[0];
  const int len = (int)key_len(hdr), nexc = (int)key_nexc(hdr);
  int cl = len - f - b;
  if (cl < 0) cl = 0;
  const int lo = (cl > 0) ? f : 0;
  const uint32_t npay_in = (uint32_t)(len + 15) >> 4, npay = (uint32_t)(cl + 15) >> 4;
  for (uint32_t w = 0; w < npay; ++w) {
    uint32_t word = 0;
    for (int q = 0; q < 16; ++q) {
      const int j = (int)w * 16 + q;
      if (j < cl) word |= key_base(key, (uint32_t)(lo + j)) << (2 * q);
    }
    out[1 + w] = word;
  }
  uint32_t xo = 0;
  for (int x = 0; x < nexc; ++x) {
    const uint32_t e = key[1 + npay_in + x];
    const int pos = (int)(e >> 8);
    if (pos >= lo && pos < lo + cl) out[1 + npay + xo++] = ((uint32_t)(pos - lo) << 8) | (e & 0xFFu);
  }
  out[0] = (uint32_t)cl | (xo << 16);
  return 1u + npay + xo;
}
