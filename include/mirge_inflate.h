/* mirge_inflate.h -- C ABI of libmirge_inflate.so (mirge3.0_b200/csrc/pinflate.c): chunk-parallel inflate of one gzip
 * stream on the host cores.  Host-only code, no CUDA; bound with ctypes by mirge3.0_b200/ingest.py.
 *
 * Replaces, for large single-stream .fastq.gz input, what the reference gets from `xopen(FQfile, "rb")`
 * (mirge/libs/digest.py:136: Python's gzip module or an external pigz / igzip process, one inflating core either way)
 * in front of `dnaio.read_chunks` (digest.py:140).  Output bytes are exactly those of the gzip module: members are
 * concatenated, zero padding between and after members is skipped, CRC-32 and ISIZE of every member are checked.
 *
 * The compressed file is handed over as one contiguous buffer (ingest.py maps the file); the decoder keeps no
 * reference to it beyond pgz_close.  One pgz handle serves one reader thread; different handles are independent. */
#ifndef MIRGE_INFLATE_H
#define MIRGE_INFLATE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* data[0 .. n): the whole .gz file.  threads: cores to use (>= 1).  chunk_bytes: compressed bytes per speculative
 * chunk (>= 64 KB; a wave decodes threads x chunk_bytes at once).  NULL on allocation failure. */
void *pgz_open(const uint8_t *data, uint64_t n, int threads, uint64_t chunk_bytes);

/* The next decompressed bytes into dst (at most cap): > 0 = bytes written, 0 = end of file, < 0 = error, text in
 * pgz_error (truncated file: "... ended before the end-of-stream marker ..."; bad data: "invalid deflate data";
 * trailer: "CRC check failed" / "incorrect length of data produced"; header problems).  With more than one thread the
 * waves are decoded by a producer thread one wave ahead of the reader (MIRGE_B200_PGZ_PREFETCH=0: on demand), so the
 * call blocks only when the reader has caught up.  After an error every later call fails as well. */
int64_t pgz_read(void *h, uint8_t *dst, uint64_t cap);
const char *pgz_error(void *h);

/* out4 = { waves run, chunks accepted, speculative chunks discarded, bytes delivered } */
void pgz_stats(void *h, uint64_t *out4);
/* out4 = seconds spent { searching block starts, decoding, chaining windows, replacing markers + CRC } */
void pgz_times(void *h, double *out4);

void pgz_close(void *h);

/* The decoder's CRC-32 (zlib's polynomial and conventions: crc of the bytes so far in, new crc out; 0 to start).
 * Carry-less multiplication where the CPU has PCLMULQDQ (MIRGE_B200_PGZ_NO_CLMUL=1 turns it off), slice-by-8 otherwise. */
uint32_t pgz_crc(uint32_t crc, const uint8_t *p, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
