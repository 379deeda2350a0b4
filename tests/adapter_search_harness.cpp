// Host harness of mirge3.0_b200/csrc/adapter_search.cuh (built by tests/test_adapter_search_host.py with g++):
// the device search code compiled for one host thread, driven through a small C interface.
#define ADAPTER_SEARCH_HOST
#include "adapter_search.cuh"

#include <vector>

static std::vector<uint32_t> g_eq_bytes, g_eq_codes, g_ps;
static int g_fast_ok = 0, g_maxm = 0;

// returns MIRGE_OK / MIRGE_ERR_ARG; *fast_ok = every adapter qualifies for the bit-parallel search
extern "C" int hs_set_params(const mirge_trim_params *p, int *fast_ok, char *err, int err_len) {
  const int rc = fill_dev_params(p, c_p, g_maxm, g_fast_ok, err, (size_t)err_len);
  if (rc != MIRGE_OK) return rc;
  // the two match-mask tables exactly as trim_kernel / trim_dp_kernel fill them
  g_eq_bytes.assign((size_t)c_p.n_adapters * 256, 0u);
  g_eq_codes.assign((size_t)c_p.n_adapters * 256, 0u);
  for (int e = 0; e < c_p.n_adapters * 256; ++e) {
    const uint32_t code = base_code_upper((uint32_t)(e & 255));
    g_eq_bytes[e] = code < 4u ? (uint32_t)c_p.ad[e >> 8].peq[code] : 0u;
  }
  for (int e = 0; e < c_p.n_adapters * 4; ++e) g_eq_codes[(e >> 2) * 256 + (e & 3)] = (uint32_t)c_p.ad[e >> 2].peq[e & 3];
  *fast_ok = g_fast_ok;
  return MIRGE_OK;
}

// mode 0: locate<MAXM> (literal DP); 1 / 2: locate_fast on the bytes without / with deferral; 3 / 4: locate_fast on the
// packed read without / with deferral.  The search runs on read[start:stop).  Returns 0 = no match, 1 = match in
// out[4] = {rstart, rstop, matches, errors}, 2 = deferred, -1 = mode not applicable to this input.
extern "C" int hs_locate(int mode, int a, const uint8_t *read, int len, int start, int stop, int32_t *out) {
  Match mt{0, 0, 0, 0};
  const int n = stop - start;
  int rc;
  if (mode == 0) {
    rc = g_maxm <= 32 ? (int)locate_any<32>(a, read + start, n, mt) : (int)locate_any<64>(a, read + start, n, mt);
  } else {
    if (!g_fast_ok) return -1;
    bool pure = true;
    for (int i = 0; i < len; ++i) pure = pure && (read[i] == 'A' || read[i] == 'C' || read[i] == 'G' || read[i] == 'T');
    const int nwords = (len + 15) >> 4, rows = nwords + 3;
    g_ps.assign((size_t)rows * TRIM_THREADS, 0u);
    if (pure)
      for (int i = 0; i < len; ++i) {
        const uint32_t code = read[i] == 'A' ? 0u : read[i] == 'C' ? 1u : read[i] == 'G' ? 2u : 3u;
        g_ps[(size_t)(i >> 4) * TRIM_THREADS] |= code << (2 * (i & 15));
      }
    FastCtx fc;
    fc.ps = g_ps.data();
    fc.jump_ok = pure && len <= 160;
    fc.rbase = start;
    if (mode == 1 || mode == 2) {
      fc.s_eq = g_eq_bytes.data();
      const ByteRead rv{read + start};
      rc = mode == 1 ? locate_fast<false>(a, rv, n, fc, mt) : locate_fast<true>(a, rv, n, fc, mt);
    } else {
      if (!fc.jump_ok) return -1;
      fc.s_eq = g_eq_codes.data();
      const PackedRead rv{g_ps.data(), start};
      rc = mode == 3 ? locate_fast<false>(a, rv, n, fc, mt) : locate_fast<true>(a, rv, n, fc, mt);
    }
  }
  out[0] = mt.rstart; out[1] = mt.rstop; out[2] = mt.matches; out[3] = mt.errors;
  return rc;
}

// the whole modifier pipeline of one read on the full-DP path (apply_mod<MAXM, false, false>: what trim_kernel<MAXM, false, 0, false>
// runs per read, linked pairs and every adapter placement included): windows[2 * mi] / [2 * mi + 1] = the read's window after modifier mi
extern "C" int hs_pipeline(const uint8_t *seq, const uint8_t *qual, int len, int32_t *windows) {
  FastCtx fc;
  memset(&fc, 0, sizeof(fc));
  int start = 0, stop = len;
  for (int mi = 0; mi < c_p.n_mods; ++mi) {
    if (g_maxm <= 32) apply_mod<32, false, false>(mi, seq, qual, start, stop, fc);
    else apply_mod<64, false, false>(mi, seq, qual, start, stop, fc);
    windows[2 * mi] = start;
    windows[2 * mi + 1] = stop;
  }
  return c_p.n_mods;
}

// the two quality scans (cutadapt qualtrim.pyx), as the trim kernels call them
extern "C" int hs_nextseq(const uint8_t *seq, const uint8_t *qual, int len, int cutoff, int base) {
  return nextseq_trim_index(seq, qual, len, cutoff, base);
}
extern "C" void hs_quality(const uint8_t *qual, int len, int q5, int q3, int base, int *start, int *stop) {
  quality_trim_index(qual, len, q5, q3, base, *start, *stop);
}
