/*
 * aocl_llc_native.h -- the codecs' own C entry points, exported by the B200 library next to the
 * unified API so that callers of the reference's native interfaces (test/codec_native_api_bench.c:83-176,
 * and anything linked against the liblz4.so / libsnappy.so symlinks the reference installs,
 * CMakeLists.txt:616-635) relink without source changes.  SURVEY.md section 8(f) row 2.
 *
 * Reference interface each item replaces (paths under /root/reference):
 *   LZ4_compressBound                 algos/lz4/lz4.h:287    (lz4.c:735)
 *   LZ4_compress_default              algos/lz4/lz4.h:203    (lz4.c:2967 -> LZ4_compress_fast -> AOCL_LZ4_compress_fast_mt, lz4.c:2655-2909)
 *   LZ4_decompress_safe               algos/lz4/lz4.h:232    (lz4.c:4898 -> AOCL_LZ4_decompress_safe_mt, lz4.c:4785-4890)
 *   snappy_compress                   algos/snappy/snappy-c.h (snappy-c.cc:34-43  -> snappy::RawCompress, snappy.cc:2494-2666)
 *   snappy_uncompress                 algos/snappy/snappy-c.h (snappy-c.cc:45-63  -> snappy::RawUncompress, snappy.cc:2271-2390)
 *   snappy_max_compressed_length      algos/snappy/snappy-c.h (snappy-c.cc:65-67)
 *   snappy_uncompressed_length        algos/snappy/snappy-c.h (snappy-c.cc:69-79)
 *
 * Same conventions as the reference: the LZ4 functions return a byte count (0 / negative on failure),
 * the Snappy functions a snappy_status.  Streams are the same RAP-framed streams the unified API
 * produces (the reference's threaded build frames them in exactly these entry points).  One
 * deliberate difference: snappy_uncompress / snappy_uncompressed_length read the length varint
 * BEHIND a RAP frame when there is one; the reference's C wrappers parse it from the first bytes even
 * for framed streams (snappy.cc:570-580 vs 596-615) and so report 65 ('A' of the magic).
 * Buffers may be host or device memory.  No CPU fallback.
 */
#ifndef AOCL_LLC_NATIVE_H
#define AOCL_LLC_NATIVE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

int LZ4_compressBound(int inputSize);
int LZ4_compress_default(const char *src, char *dst, int srcSize, int dstCapacity);
int LZ4_decompress_safe(const char *src, char *dst, int compressedSize, int dstCapacity);

typedef enum { SNAPPY_OK = 0, SNAPPY_INVALID_INPUT = 1, SNAPPY_BUFFER_TOO_SMALL = 2 } snappy_status;

snappy_status snappy_compress(const char *input, size_t input_length, char *compressed, size_t *compressed_length);
snappy_status snappy_uncompress(const char *compressed, size_t compressed_length, char *uncompressed,
                                size_t *uncompressed_length);
size_t snappy_max_compressed_length(size_t source_length);
snappy_status snappy_uncompressed_length(const char *compressed, size_t compressed_length, size_t *result);

#ifdef __cplusplus
}
#endif
#endif /* AOCL_LLC_NATIVE_H */
