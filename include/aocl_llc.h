/*
 * aocl_llc.h -- the unified AOCL-Compression C API, as exported by the B200 library.
 *
 * This is the drop-in boundary: the seven entry points below are exactly the symbols a
 * caller of the reference binds for the LZ4 / Snappy path.  The struct layout, enum
 * values and return conventions are ABI-identical to the reference (LP64, 128-byte
 * descriptor) so existing callers relink without source changes; the reference's own
 * api/aocl_compression.h and api/aocl_threads.h can be used in place of this file.
 *
 * Reference interface each item replaces (paths under /root/reference):
 *   aocl_error_type          api/aocl_compression.h:95-102
 *   aocl_compression_type    api/aocl_compression.h:109-119
 *   aocl_compression_desc    api/aocl_compression.h:125-152
 *   aocl_llc_compress        api/aocl_compression.h:170   (api/api.cpp:45-83)
 *   aocl_llc_decompress      api/aocl_compression.h:189   (api/api.cpp:86-124)
 *   aocl_llc_setup           api/aocl_compression.h:207   (api/api.cpp:127-166)
 *   aocl_llc_destroy         api/aocl_compression.h:221   (api/api.cpp:169-183)
 *   aocl_llc_version         api/aocl_compression.h:229   (api/api.cpp:186-189)
 *   aocl_get_rap_frame_bound_mt  api/aocl_threads.h:109   (threads/threads.c:315-318)
 *   aocl_skip_rap_frame_mt       api/aocl_threads.h:133   (threads/threads.c:320-336)
 *
 * inBuf / outBuf may be ordinary host memory, pinned host memory or device memory; the
 * library inspects the pointer (cudaPointerGetAttributes) and stages as needed.  Only LZ4
 * and SNAPPY are served: setup for the other five methods reports ERR_EXCLUDED_METHOD,
 * exactly like a reference build configured with AOCL_EXCLUDE_<method>.
 * There is no CPU fallback: without a usable CUDA device every call fails.
 */
#ifndef AOCL_LLC_B200_H
#define AOCL_LLC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    ERR_INVALID_INPUT = -5,
    ERR_UNSUPPORTED_METHOD = -4,
    ERR_EXCLUDED_METHOD = -3,
    ERR_COMPRESSION_FAILED = -2,
    ERR_COMPRESSION_INVALID_OUTPUT = -1
} aocl_error_type;

typedef enum {
    LZ4 = 0,
    LZ4HC = 1,
    LZMA = 2,
    BZIP2 = 3,
    SNAPPY = 4,
    ZLIB = 5,
    ZSTD = 6,
    AOCL_COMPRESSOR_ALGOS_NUM = 7
} aocl_compression_type;

typedef struct {
    char *inBuf;        /* input bytes (host or device)                                  */
    char *outBuf;       /* output bytes (host or device)                                 */
    char *workBuf;      /* always NULL for LZ4 / SNAPPY, as in the reference             */
    size_t inSize;      /* bytes at inBuf                                                */
    size_t outSize;     /* capacity of outBuf                                            */
    size_t level;       /* ignored by LZ4 / SNAPPY (api/codec.cpp:129-130, 259-260)      */
    size_t optVar;      /* ignored                                                       */
    int numThreads;     /* ignored (dead field in the reference as well)                 */
    int numMPIranks;    /* ignored                                                       */
    size_t memLimit;    /* ignored                                                       */
    int measureStats;   /* 1: fill the six statistics fields below                       */
    uint64_t cSize;     /* compressed size of the last compress call                     */
    uint64_t dSize;     /* decompressed size of the last decompress call                 */
    uint64_t cTime;     /* wall-clock ns of the last compress call (incl. PCIe copies)   */
    uint64_t dTime;     /* wall-clock ns of the last decompress call                     */
    float cSpeed;       /* inSize * 1000 / cTime  (MB/s)                                 */
    float dSpeed;       /* dSize * 1000 / dTime   (MB/s)                                 */
    int optOff;         /* accepted; 1 selects the frame-less single-partition LZ4 layout */
    int optLevel;       /* overwritten by setup (reports 4: widest level of the reference) */
} aocl_compression_desc;

int64_t aocl_llc_compress(aocl_compression_desc *handle, aocl_compression_type codec_type);
int64_t aocl_llc_decompress(aocl_compression_desc *handle, aocl_compression_type codec_type);
int32_t aocl_llc_setup(aocl_compression_desc *handle, aocl_compression_type codec_type);
void aocl_llc_destroy(aocl_compression_desc *handle, aocl_compression_type codec_type);
const char *aocl_llc_version(void);

/* Upper bound of the RAP frame this library can emit for any int32-sized input. */
int32_t aocl_get_rap_frame_bound_mt(void);
/* Bytes to skip to reach the codec payload: frame length when the RAP magic is present,
 * 0 otherwise, ERR_INVALID_INPUT for NULL.  `src` must be host-readable. */
int32_t aocl_skip_rap_frame_mt(char *src, int32_t src_size);

#ifdef __cplusplus
}
#endif
#endif /* AOCL_LLC_B200_H */
