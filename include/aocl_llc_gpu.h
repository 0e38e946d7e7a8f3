/*
 * aocl_llc_gpu.h -- device-level entry points of the B200 LZ4 / Snappy RAP library.
 *
 * aocl_llc.h is the reference-facing API.  The functions here sit directly below it and
 * are what aocl_llc_compress / aocl_llc_decompress call once their buffers are in HBM;
 * they are exported so that callers that already hold device memory (a columnar reader,
 * bench.py's device-resident leg, a multi-GPU driver that shards RAP partitions) can skip
 * the PCIe staging.  Plain C ABI: pointers, sizes and a cudaStream_t passed as void*.
 *
 * Reference code each entry point replaces (paths under /root/reference):
 *   aocl_gpu_compress      LZ4_compress_default -> AOCL_LZ4_compress_fast_mt   algos/lz4/lz4.c:2655-2909, 2967
 *                          snappy::RawCompress                                algos/snappy/snappy.cc:2494-2666
 *   aocl_gpu_decompress    LZ4_decompress_safe -> AOCL_LZ4_decompress_safe_mt  algos/lz4/lz4.c:4785-4890, 4898
 *                          snappy::RawUncompress                              algos/snappy/snappy.cc:2271-2390
 *   aocl_gpu_*_batch       a loop over LZ4_compress_default / LZ4_decompress_safe /
 *                          snappy::RawCompress / RawUncompress on independent frame-less pages
 *   aocl_gpu_partition_*   aocl_setup_parallel_compress_mt / aocl_do_partition_compress_mt
 *                          threads/threads.c:46-153 (partition arithmetic only)
 *
 * Every function returns 0 / a byte count on success and a negative aocl_error_type-style
 * code on failure.  No function falls back to the CPU.
 */
#ifndef AOCL_LLC_GPU_H
#define AOCL_LLC_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AOCL_GPU_LZ4 0     /* == aocl_compression_type LZ4    */
#define AOCL_GPU_SNAPPY 4  /* == aocl_compression_type SNAPPY */

typedef struct aocl_gpu_ctx_s *aocl_gpu_ctx_t;

/* One context = one CUDA stream + a growable HBM workspace + a pinned result block.
 * `device` < 0 selects the current device.  `stream` NULL creates a private stream;
 * otherwise work is enqueued on the caller's cudaStream_t. */
int32_t aocl_gpu_ctx_create(aocl_gpu_ctx_t *ctx, int device, void *stream);
void aocl_gpu_ctx_destroy(aocl_gpu_ctx_t ctx);
void *aocl_gpu_ctx_stream(aocl_gpu_ctx_t ctx);

/* RAP partition arithmetic for an n-byte input (saturated layout, T = P(n)). */
int32_t aocl_gpu_partition_count(int32_t codec, size_t n);
/* Worst-case compressed size incl. RAP frame, for sizing d_out. */
size_t aocl_gpu_compress_bound(int32_t codec, size_t n);

/* Whole-buffer operations on device-resident data.  d_in and d_out are device pointers.
 * The *_async forms only enqueue; aocl_gpu_finish() synchronises the context's stream and
 * returns the result of the most recent enqueue (bytes produced, or < 0). */
int32_t aocl_gpu_compress_async(aocl_gpu_ctx_t ctx, int32_t codec, const void *d_in, size_t n,
                                void *d_out, size_t out_cap);
int32_t aocl_gpu_decompress_async(aocl_gpu_ctx_t ctx, int32_t codec, const void *d_in, size_t n,
                                  void *d_out, size_t out_cap);
int64_t aocl_gpu_finish(aocl_gpu_ctx_t ctx);
int64_t aocl_gpu_compress(aocl_gpu_ctx_t ctx, int32_t codec, const void *d_in, size_t n,
                          void *d_out, size_t out_cap);
int64_t aocl_gpu_decompress(aocl_gpu_ctx_t ctx, int32_t codec, const void *d_in, size_t n,
                            void *d_out, size_t out_cap);

/* Frame-less single-partition LZ4 layout for every size (what the reference emits with
 * optOff=1 or one OpenMP thread): 1 = on, 0 = off (default). */
void aocl_gpu_set_lz4_frameless(aocl_gpu_ctx_t ctx, int32_t on);

/* Compress mode of LZ4 RAP frames (env AOCL_GPU_MODE at context creation, or this call):
 *   "exact"     (default) byte-identical to the reference's LZ4_compress_fast at acceleration 1 (lz4.c:1853-2350);
 *   "fastparse" the separately named position-parallel parse (csrc/lz4_fastparse.cuh): a valid RAP stream that every
 *               LZ4 decoder reads, NOT byte-identical to the reference's and with a somewhat larger output; bench.py
 *               reports its speed and its ratio delta.  Never selected implicitly.  Frame-less blocks (T == 1), Snappy
 *               and the page batches always use the exact encoders.
 * Returns 0, -4 for an unknown mode. */
int32_t aocl_gpu_set_mode(aocl_gpu_ctx_t ctx, const char *mode);

/* Partition layout of the frames this context writes (env AOCL_GPU_PARTITIONS at context creation, or this call).
 * The reference cuts a frame into min(omp_get_max_threads(), P(n)) partitions (threads/threads.c:55-88), so what a host
 * emits depends on its thread count.  The default (0) is the saturated layout T = P(n) -- what a host with at least P(n)
 * threads writes, and the only layout with GPU-sized parallelism.  max_threads = K makes this context write, byte for
 * byte, the frame of a K-thread host: K partitions of n/K bytes.  LZ4 partitions are serial chains, one warp each: a
 * K-partition frame compresses at K x ~7 MB/s -- this is for byte comparison and interop, not for speed.  Snappy is
 * not affected in speed (its unit of work is the 64 KiB fragment whatever the layout).  fastparse falls back to the
 * exact encoder for partitions of 512 KiB and more.  The sharded calls always use the saturated layout.
 * Decompression reads every layout regardless.  Returns 0, -5 for a bad argument. */
int32_t aocl_gpu_set_partitions(aocl_gpu_ctx_t ctx, int32_t max_threads);
/* Partitions a frame of n bytes written by THIS context will have (aocl_gpu_partition_count: the saturated P(n)). */
int32_t aocl_gpu_ctx_partition_count(aocl_gpu_ctx_t ctx, int32_t codec, size_t n);

/* Streaming input for the NEXT aocl_gpu_compress_async on this context (one-shot).  d_flag points to a
 * 32-bit watermark in device memory that the caller raises (e.g. with 4-byte H2D copies ordered after
 * the payload copies on its own copy stream) while the encoder is already running:
 *   LZ4 RAP frames:  bytes present in EVERY partition (send the input in stripes across partitions:
 *                    partition i occupies [i*(n/T), (i+1)*(n/T)), T = aocl_gpu_ctx_partition_count);
 *   Snappy:          bytes present from the start of the input.
 * Raise it to 0xffffffff once everything (including the last partition's n % T extra bytes) is there.
 * Encoder warps wait on the watermark before they touch input beyond it (with a 64-byte margin, so a
 * cached sector never straddles it).  This is how aocl_llc_compress overlaps the PCIe transfer of a
 * host buffer with the encode (reference equivalent: none -- its input is already in host memory). */
void aocl_gpu_set_input_watermark(aocl_gpu_ctx_t ctx, const uint32_t *d_flag);

/* Decode only RAP partitions [first, first+count) of the stream at d_in into
 * d_out + (sum of decomp_len of partitions < first) - out_origin.  Used to shard one frame
 * across GPUs: each rank passes the same frame header and its own partition range.
 * Result (bytes produced by the range) via aocl_gpu_finish(). */
int32_t aocl_gpu_decompress_range_async(aocl_gpu_ctx_t ctx, int32_t codec, const void *d_in, size_t n,
                                        void *d_out, size_t out_cap, uint32_t first, uint32_t count,
                                        uint64_t out_origin);

/* ---- ONE frame over several GPUs ------------------------------------------------------------
 * One process (or host thread) per GPU, each with its own context.  Rank r of R owns the contiguous partition
 * range [floor(r*T/R), floor((r+1)*T/R)): it holds only that slice of the input, works only on those partitions
 * and writes only its own byte range of the result.  The library exchanges, over NCCL on the context's stream:
 * the per-partition records (compress: every rank runs the same stitch plan, rank 0 writes the RAP frame), the
 * trailing literals a rank's first partitions inherit from its predecessor (LZ4 compress, lz4.c:2808-2877), and
 * {bytes produced, error} (decompress).  Replaces the OpenMP fork/join + serial stitch of lz4.c:2684-2905,
 * 4785-4890 and snappy.cc:2506-2655, 2271-2390.  All calls are collective and blocking.
 *   unique_id : rank 0 obtains 128 bytes (an ncclUniqueId) and hands them to the other ranks by any means;
 *   init      : joins the communicator on this context's device;
 *   range     : the partitions and input bytes rank `rank` owns of an n-byte input (-2: fewer partitions than ranks);
 *   compress  : d_in_slice = input bytes [byte_off, byte_off + byte_len).  On return d_out_slice holds bytes
 *               [*out_off, *out_off + *out_len) of the final stream (rank 0: from 0, RAP frame first); the return
 *               value is the length of the whole stream, the same on every rank, or < 0;
 *   decompress: d_stream addresses the stream by its own offsets; the frame header, the entry table (and Snappy's
 *               varint) and this rank's partitions must be present.  On return d_out_slice holds output bytes
 *               [*out_off, *out_off + *out_len); the return value is the total, the same on every rank, or < 0. */
int32_t aocl_gpu_shard_unique_id(void *id_out_128_bytes);
int32_t aocl_gpu_shard_init(aocl_gpu_ctx_t ctx, const void *id_128_bytes, int32_t rank, int32_t nranks);
void aocl_gpu_shard_destroy(aocl_gpu_ctx_t ctx);
int32_t aocl_gpu_shard_range(int32_t codec, size_t n, int32_t rank, int32_t nranks, uint32_t *first_partition,
                             uint32_t *partition_count, uint64_t *byte_off, uint64_t *byte_len);
int64_t aocl_gpu_compress_sharded(aocl_gpu_ctx_t ctx, int32_t codec, const void *d_in_slice, size_t n_total,
                                  void *d_out_slice, size_t out_cap, uint64_t *out_off, uint64_t *out_len);
int64_t aocl_gpu_decompress_sharded(aocl_gpu_ctx_t ctx, int32_t codec, const void *d_stream, size_t n,
                                    void *d_out_slice, size_t out_cap, uint64_t *out_off, uint64_t *out_len);

/* Slab-wise decode of one RAP stream, for callers that overlap the decode with their own transfers
 * (aocl_llc_decompress does, for host buffers: H2D of slab k+1 | decode of slab k | D2H of slab k-1).
 *   open : parse + validate the frame at d_in (the header, the entry table and -- Snappy -- the varint
 *          behind it must already be in HBM) and lay the partitions out in the output;
 *   slab : decode partitions [first, first+count) to their final offsets in d_out, enqueued on the
 *          context's stream (order it after the slab's H2D copy with cudaStreamWaitEvent);
 *   close: fetch the result; aocl_gpu_finish() then returns the stream's total or < 0.
 * Errors are sticky across slabs. */
int32_t aocl_gpu_decompress_open_async(aocl_gpu_ctx_t ctx, int32_t codec, const void *d_in, size_t n, size_t out_cap);
int32_t aocl_gpu_decompress_slab_async(aocl_gpu_ctx_t ctx, int32_t codec, const void *d_in, void *d_out,
                                       uint32_t first, uint32_t count);
int32_t aocl_gpu_decompress_close_async(aocl_gpu_ctx_t ctx);

/* Independent frame-less pages (one LZ4 block / one Snappy stream each), all device
 * resident.  Arrays are device arrays of `count` entries.  status[i] receives bytes
 * produced or < 0.  The return value of aocl_gpu_finish() is the number of failed pages
 * negated (0 = all good). */
int32_t aocl_gpu_decompress_batch_async(aocl_gpu_ctx_t ctx, int32_t codec, const void *const *d_in_ptrs,
                                        const uint32_t *d_in_sizes, void *const *d_out_ptrs,
                                        const uint32_t *d_out_caps, int64_t *d_status, size_t count);
int32_t aocl_gpu_compress_batch_async(aocl_gpu_ctx_t ctx, int32_t codec, const void *const *d_in_ptrs,
                                      const uint32_t *d_in_sizes, void *const *d_out_ptrs,
                                      const uint32_t *d_out_caps, int64_t *d_status, size_t count);

/* aocl_llc_compress / aocl_llc_decompress calls that were split over several GPUs so far (AOCL_GPU_SHARD=1 with two
 * or more devices in AOCL_GPU_DEVICES; host buffers of at least 32 MiB). */
uint64_t aocl_gpu_sharded_host_calls(void);

/* Number of kernels this library has launched in the calling process (bench.py reports it). */
uint64_t aocl_gpu_launch_count(void);

/* Per-kernel device timing of the most recent enqueue: when enabled, every kernel launch is
 * bracketed by CUDA events on the context's stream.  Query after aocl_gpu_finish(). */
void aocl_gpu_set_profiling(aocl_gpu_ctx_t ctx, int32_t on);
int32_t aocl_gpu_profile_count(aocl_gpu_ctx_t ctx);
float aocl_gpu_profile_get(aocl_gpu_ctx_t ctx, int32_t index, char *name, int32_t name_cap);  /* ms, <0 if n/a */

/* Diagnostics of the tile decoder: copies 32 device counters (per-phase cycles when the library
 * is built with -DLLC_TILE_PROF, watchdog hits always) and optionally resets them. */
int32_t aocl_gpu_debug_counters(uint64_t *out32, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* AOCL_LLC_GPU_H */
