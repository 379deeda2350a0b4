// Host harness of mirge3.0_b200/csrc/key_format.cuh (built by tests/test_key_format_host.py with g++).
#define KEY_FORMAT_HOST
#include "key_format.cuh"

extern "C" uint32_t hk_slice(const uint32_t *key, int f, int b, uint32_t *out) { return slice_key(key, f, b, out); }
