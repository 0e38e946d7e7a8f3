// llc_device.cu -- device-level C ABI (include/aocl_llc_gpu.h): context, workspace and the
// kernel launch sequences for RAP compress / decompress and batched pages.
#include "llc_kernels.cuh"
#include "../../include/aocl_llc_gpu.h"

#include <algorithm>
#include <atomic>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace llc;

static std::atomic<uint64_t> g_launches{0};
// Every launch goes through this macro: it counts the launch and, when profiling is enabled on
// the context, brackets the kernel with CUDA events on the launching stream.
#define LLC_LAUNCH(kernel, grid, block, smem, stream, ...)                         \
    do {                                                                           \
        const int prof_slot_ = prof_begin(c, #kernel);                             \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                \
        prof_end(c, prof_slot_);                                                   \
        g_launches.fetch_add(1, std::memory_order_relaxed);                        \
    } while (0)

constexpr int kProfSlots = 16;
constexpr int kStabMax = 11;        // shared-table LZ4 encoder CTAs per SM (16 KiB table + 4 KiB owner bytes each)

struct aocl_gpu_shard_s;
struct aocl_gpu_ctx_s {
    int device = 0;
    aocl_gpu_shard_s* shard = nullptr;   // one frame over several GPUs (llc_shard.cuh)
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side = nullptr;    // second stream for kernels that run concurrently with the main one
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // LZ4 encoder placement.  Every partition is one serial chain, so wall time is (waves) x (chain
    // latency): what matters is having ALL partitions resident at once.  Up to 14 x SMs partitions the
    // shared-memory-table flavour does that with the fastest table; beyond, only the global-table
    // flavour (32 warps per SM, tables L2 resident) keeps a 1 GiB frame (4094 partitions) in one wave.
    // Measured on B200, 1 GiB text: 14 smem CTAs/SM = 100 ms, 14 smem + 16 gtab = 75 ms, 32 gtab = 54 ms.
    int gtab_ctas_per_sm = -1;      // AOCL_GPU_GTAB_CTAS: global-table CTAs per SM (-1 auto, 0 never)
    int stab_ctas_per_sm = -1;      // AOCL_GPU_STAB_CTAS: shared-table CTAs per SM (-1 auto, 0 never; max 14)
    size_t l2_persist_bytes = 0, l2_window_bytes = 0;
    int snappy_gtab_ctas_per_sm = -1;   // AOCL_GPU_SNAPPY_GTAB_CTAS: Snappy global-table CTAs per SM (-1 auto, 0 never)
    int snappy_stab_ctas_per_sm = -1;   // AOCL_GPU_SNAPPY_STAB_CTAS: Snappy shared-memory-table CTAs per SM (-1 auto)
    int sm_count = 0;
    uint8_t* ws = nullptr;          // growable HBM workspace (scratch slots, tables, plans)
    size_t ws_bytes = 0;
    CallResult* d_res = nullptr;    // result block on the device
    CallResult* h_res = nullptr;    // pinned mirror
    int decode_blocks = 0;          // persistent grid size of decode_parts_kernel
    // Decoder organisation (AOCL_GPU_DECODER = auto | rowq | tile | warp).  Measured on B200, 1 GiB frames
    // (LZ4 text / Snappy log), ms per GiB:
    //   rowq   lane parsers + thread-per-byte row copiers, 28 partitions in flight per SM  
    //   tile   one 512-thread CTA per partition, data-parallel parse, pointer jumping:       6.45 / 5.05
    //   warp   one warp per partition: LZ4 TMA-ring pipelined decoder                        14.2 / 11.9
    // auto (default): the row decoder when a launch has at least rowq_min_units units (it needs several units
    // per SM in flight), else the tile decoder.  Round 1 also measured a parser warp + lane-per-sequence copier
    // (16-19 ms) and lane parsers + lane-per-sequence copiers against global memory (22 ms); both are gone.
    int decoder_mode = 0;           // 0 auto, 1 warp, 4 tile, 5 rowq
    uint32_t rowq_min_units = 0;    // AOCL_GPU_ROWQ_MIN_UNITS: auto mode takes the row decoder from this many units on
    bool lz4_frameless = false;
    bool fastparse = false;         // AOCL_GPU_MODE=fastparse / aocl_gpu_set_mode(): the named non-exact LZ4 RAP encoder
    uint32_t max_parts = 0;         // AOCL_GPU_PARTITIONS / aocl_gpu_set_partitions(): the host's omp_get_max_threads() to imitate (0: none)
    const uint32_t* in_flag = nullptr;   // one-shot input watermark for the next compress (aocl_gpu_set_input_watermark)
    bool batch_mode = false;        // last enqueue was a batch call (finish() returns -failures)
    int last_rc = 0;                // enqueue-time failure to report from finish()
    // optional per-kernel timing (aocl_gpu_set_profiling)
    bool prof = false;
    int prof_n = 0;
    cudaEvent_t prof_ev[kProfSlots][2] = {};
    const char* prof_name[kProfSlots] = {};
};

static int prof_begin(aocl_gpu_ctx_t c, const char* name) {
    if (!c->prof || c->prof_n >= kProfSlots) return -1;
    const int s = c->prof_n++;
    if (!c->prof_ev[s][0]) { cudaEventCreate(&c->prof_ev[s][0]); cudaEventCreate(&c->prof_ev[s][1]); }
    c->prof_name[s] = name;
    cudaEventRecord(c->prof_ev[s][0], c->stream);
    return s;
}
static void prof_end(aocl_gpu_ctx_t c, int s) {
    if (s >= 0) cudaEventRecord(c->prof_ev[s][1], c->stream);
}

static bool cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    if (getenv("AOCL_GPU_VERBOSE")) fprintf(stderr, "[aocl-llc-b200] %s: %s\n", what, cudaGetErrorString(e));
    return false;
}

static bool ensure_ws(aocl_gpu_ctx_t c, size_t bytes) {
    if (bytes <= c->ws_bytes) return true;
    if (c->ws) { cudaStreamSynchronize(c->stream); cudaFree(c->ws); c->ws = nullptr; c->ws_bytes = 0; }
    bytes = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    if (!cuda_ok(cudaMalloc(&c->ws, bytes), "cudaMalloc(workspace)")) return false;
    c->ws_bytes = bytes;
    return true;
}

extern "C" int32_t aocl_gpu_ctx_create(aocl_gpu_ctx_t* out, int device, void* stream) {
    if (!out) return -5;
    *out = nullptr;
    int count = 0;
    if (!cuda_ok(cudaGetDeviceCount(&count), "cudaGetDeviceCount") || count == 0) return -2;
    if (device < 0 && !cuda_ok(cudaGetDevice(&device), "cudaGetDevice")) return -2;
    if (device >= count || !cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return -2;
    aocl_gpu_ctx_t c = new aocl_gpu_ctx_s();
    c->device = device;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (stream) c->stream = (cudaStream_t)stream;
    else { if (!cuda_ok(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate")) { delete c; return -2; } c->own_stream = true; }
    if (!cuda_ok(cudaMalloc(&c->d_res, sizeof(CallResult)), "cudaMalloc(result)") ||
        !cuda_ok(cudaMallocHost(&c->h_res, sizeof(CallResult)), "cudaMallocHost(result)")) { aocl_gpu_ctx_destroy(c); return -2; }
    memset(c->h_res, 0, sizeof(CallResult));
    cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
    // L2 persistence for the encoders' global hash tables (they are hit once per probe round)
    {
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
        c->l2_persist_bytes = (size_t)max_persist;
        c->l2_window_bytes = (size_t)max_window;
        if (getenv("AOCL_GPU_NO_L2_PERSIST")) c->l2_persist_bytes = 0;
        if (c->l2_persist_bytes) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c->l2_persist_bytes);
        if (getenv("AOCL_GPU_VERBOSE")) fprintf(stderr, "[aocl-llc-b200] L2 persisting max %d B, window max %d B\n", max_persist, max_window);
    }
    if (const char* e = getenv("AOCL_GPU_MODE")) c->fastparse = strcmp(e, "fastparse") == 0;
    if (const char* e = getenv("AOCL_GPU_PARTITIONS")) c->max_parts = (uint32_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("AOCL_GPU_GTAB_CTAS")) c->gtab_ctas_per_sm = atoi(e);
    if (const char* e = getenv("AOCL_GPU_SNAPPY_GTAB_CTAS")) c->snappy_gtab_ctas_per_sm = atoi(e);
    if (const char* e = getenv("AOCL_GPU_SNAPPY_STAB_CTAS")) c->snappy_stab_ctas_per_sm = atoi(e);
    if (const char* e = getenv("AOCL_GPU_STAB_CTAS")) c->stab_ctas_per_sm = atoi(e) > kStabMax ? kStabMax : atoi(e);

    // opt in to the shared-memory sizes the encoders need (16 KiB LZ4 table, 32 KiB Snappy table)
    cudaFuncSetAttribute(lz4_encode_parts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    cudaFuncSetAttribute(lz4_encode_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    cudaFuncSetAttribute(snappy_encode_frags_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaFuncSetAttribute(encode_pages_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_parts_kernel, 128, 0);
    if (per_sm < 1) per_sm = 1;
    c->decode_blocks = per_sm * c->sm_count;
    c->rowq_min_units = 0xffffffffu;                           // measured slower than the tile decoder so far: opt-in
    if (const char* e = getenv("AOCL_GPU_ROWQ_MIN_UNITS")) c->rowq_min_units = (uint32_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("AOCL_GPU_DECODER"))
        c->decoder_mode = strcmp(e, "rowq") == 0 ? 5 : strcmp(e, "tile") == 0 ? 4 : strcmp(e, "warp") == 0 ? 1 : 0;
    cudaFuncSetAttribute(decode_parts_tile_kernel<TileLz4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileShared<TileLz4>));
    cudaFuncSetAttribute(decode_parts_tile_kernel<TileSnappy, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileShared<TileSnappy>));
    cudaFuncSetAttribute(decode_pages_tile_kernel<TileLz4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileShared<TileLz4>));
    cudaFuncSetAttribute(decode_pages_tile_kernel<TileSnappy, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileShared<TileSnappy>));
    cudaFuncSetAttribute(decode_parts_rowq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QShared));
    cudaFuncSetAttribute(decode_parts_rowq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QShared));
    cudaFuncSetAttribute(decode_pages_rowq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QShared));
    cudaFuncSetAttribute(decode_pages_rowq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QShared));
    *out = c;
    return 0;
}

extern "C" void aocl_gpu_shard_destroy(aocl_gpu_ctx_t c);
extern "C" void aocl_gpu_ctx_destroy(aocl_gpu_ctx_t c) {
    if (!c) return;
    cudaSetDevice(c->device);
    aocl_gpu_shard_destroy(c);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->ws) cudaFree(c->ws);
    if (c->d_res) cudaFree(c->d_res);
    if (c->h_res) cudaFreeHost(c->h_res);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (auto& e : c->prof_ev) { if (e[0]) cudaEventDestroy(e[0]); if (e[1]) cudaEventDestroy(e[1]); }
    delete c;
}

extern "C" void* aocl_gpu_ctx_stream(aocl_gpu_ctx_t c) { return c ? (void*)c->stream : nullptr; }
extern "C" void aocl_gpu_set_lz4_frameless(aocl_gpu_ctx_t c, int32_t on) { if (c) c->lz4_frameless = on != 0; }
extern "C" int32_t aocl_gpu_set_mode(aocl_gpu_ctx_t c, const char* mode) {
    if (!c || !mode) return -5;
    if (strcmp(mode, "exact") == 0) { c->fastparse = false; return 0; }
    if (strcmp(mode, "fastparse") == 0) { c->fastparse = true; return 0; }
    return -4;
}
// threads/threads.c:55-88: the frame has min(omp_get_max_threads(), P(n)) partitions
static uint32_t frame_parts(aocl_gpu_ctx_t c, size_t n, uint32_t window) {
    const uint32_t T = partition_count(n, window);
    return (c->max_parts && T > c->max_parts) ? c->max_parts : T;
}
extern "C" int32_t aocl_gpu_set_partitions(aocl_gpu_ctx_t c, int32_t max_threads) {
    if (!c || max_threads < 0) return -5;
    c->max_parts = (uint32_t)max_threads;
    return 0;
}
extern "C" int32_t aocl_gpu_ctx_partition_count(aocl_gpu_ctx_t c, int32_t codec, size_t n) {
    if (!c) return -5;
    if (codec == AOCL_GPU_LZ4 && c->lz4_frameless) return 1;
    return (int32_t)frame_parts(c, n, codec == AOCL_GPU_LZ4 ? kLz4Window : kSnappyBlock);
}
extern "C" void aocl_gpu_set_input_watermark(aocl_gpu_ctx_t c, const uint32_t* d_flag) { if (c) c->in_flag = d_flag; }
extern "C" uint64_t aocl_gpu_launch_count(void) { return g_launches.load(); }
extern "C" void aocl_gpu_set_profiling(aocl_gpu_ctx_t c, int32_t on) { if (c) c->prof = on != 0; }
extern "C" int32_t aocl_gpu_profile_count(aocl_gpu_ctx_t c) { return c ? c->prof_n : 0; }
extern "C" float aocl_gpu_profile_get(aocl_gpu_ctx_t c, int32_t i, char* name, int32_t name_cap) {
    if (!c || i < 0 || i >= c->prof_n) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, c->prof_ev[i][0], c->prof_ev[i][1]) != cudaSuccess) { cudaGetLastError(); ms = -1.f; }
    if (name && name_cap > 0) { strncpy(name, c->prof_name[i], name_cap - 1); name[name_cap - 1] = 0; }
    return ms;
}

// Diagnostics of the tile decoder: 32 counters (phase cycles with -DLLC_TILE_PROF, watchdog hits always).
extern "C" int32_t aocl_gpu_debug_counters(uint64_t* out32, int32_t reset) {
    if (out32 && cudaMemcpyFromSymbol(out32, g_tile_prof, sizeof(uint64_t) * 32) != cudaSuccess) return -2;
    if (reset) { static const uint64_t zeros[32] = {}; if (cudaMemcpyToSymbol(g_tile_prof, zeros, sizeof(zeros)) != cudaSuccess) return -2; }
    return 0;
}

extern "C" int32_t aocl_gpu_partition_count(int32_t codec, size_t n) {
    return (int32_t)partition_count(n, codec == AOCL_GPU_LZ4 ? kLz4Window : kSnappyBlock);
}
extern "C" size_t aocl_gpu_compress_bound(int32_t codec, size_t n) {
    const size_t T = (size_t)aocl_gpu_partition_count(codec, n);
    const size_t frame = T > 1 ? 16 + 12 * T : 0;
    if (codec == AOCL_GPU_LZ4) return frame + n + n / 255 + 16 + 8 * T;   // stitching can lengthen one header per partition
    return frame + 32 + n + n / 6;
}

static void begin_call(aocl_gpu_ctx_t c) {
    cudaSetDevice(c->device);
    cudaMemsetAsync(c->d_res, 0, sizeof(CallResult), c->stream);
    c->batch_mode = false;
    c->last_rc = 0;
    c->prof_n = 0;
}
static void end_call(aocl_gpu_ctx_t c) {
    cudaMemcpyAsync(c->h_res, c->d_res, sizeof(CallResult), cudaMemcpyDeviceToHost, c->stream);
}

extern "C" int64_t aocl_gpu_finish(aocl_gpu_ctx_t c) {
    if (!c) return -5;
    if (c->last_rc) return c->last_rc;
    if (!cuda_ok(cudaStreamSynchronize(c->stream), "cudaStreamSynchronize")) return -2;
    if (!cuda_ok(cudaGetLastError(), "kernel")) return -2;
    if (c->batch_mode) return -(int64_t)c->h_res->error;
    if (c->h_res->error && getenv("AOCL_GPU_VERBOSE")) {
        uint64_t cnt[32] = {};
        cudaMemcpyFromSymbol(cnt, g_tile_prof, sizeof(cnt));
        fprintf(stderr, "[aocl-llc-b200] call failed: error=%d value=%lld parts=%d next=%u watchdog=[%llu %llu %llu %llu]\n",
                c->h_res->error, c->h_res->value, c->h_res->parts, c->h_res->next, (unsigned long long)cnt[24],
                (unsigned long long)cnt[25], (unsigned long long)cnt[26], (unsigned long long)cnt[27]);
    }
    if (c->h_res->error) return -2;
    return c->h_res->value;
}

// ---------------------------------------------------------------------------------- decompress
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void launch_decode_range(aocl_gpu_ctx_t c, int32_t codec, const void* d_in, void* d_out, PartDesc* parts,
                                uint32_t first, uint32_t count, uint64_t out_origin);

extern "C" int32_t aocl_gpu_decompress_range_async(aocl_gpu_ctx_t c, int32_t codec, const void* d_in, size_t n,
                                                   void* d_out, size_t out_cap, uint32_t first, uint32_t count,
                                                   uint64_t out_origin) {
    if (!c) return -5;
    begin_call(c);
    if ((codec != AOCL_GPU_LZ4 && codec != AOCL_GPU_SNAPPY) || !d_in || n == 0 || n > 0xffffffffull || (!d_out && out_cap)) {
        c->last_rc = -2; return -2;
    }
    if (!ensure_ws(c, sizeof(PartDesc) * kMaxPartitions)) { c->last_rc = -2; return -2; }
    PartDesc* parts = reinterpret_cast<PartDesc*>(c->ws);
    const bool ranged = !(first == 0 && count == 0xffffffffu) || out_origin != 0;
    LLC_LAUNCH(rap_parse_kernel, 1, 1024, 0, c->stream, codec, (const uint8_t*)d_in, (uint64_t)n, (uint64_t)out_cap,
               ranged ? 0 : 1, parts, c->d_res);
    // a range is validated against the caller's buffer before anything is decoded (the entries are untrusted)
    if (ranged) LLC_LAUNCH(range_check_kernel, 1, 256, 0, c->stream, parts, c->d_res, first, count, out_origin, (uint64_t)out_cap);
    launch_decode_range(c, codec, d_in, d_out, parts, first, count, out_origin);
    end_call(c);
    return 0;
}

// Slab-wise decode of one stream: open (frame parse) / slab (a partition range, fresh ticket, sticky
// error) / close.  The caller orders the slabs against its own H2D / D2H copies with events on the
// context's stream; aocl_gpu_finish() then returns the stream's total or the first error.
static void launch_decode_range(aocl_gpu_ctx_t c, int32_t codec, const void* d_in, void* d_out, PartDesc* parts,
                                uint32_t first, uint32_t count, uint64_t out_origin) {
    const uint8_t* in = (const uint8_t*)d_in;
    uint8_t* out = (uint8_t*)d_out;
    if (c->decoder_mode == 1) {
        LLC_LAUNCH(decode_parts_kernel, c->decode_blocks, 128, 0, c->stream, codec, in, out, parts, c->d_res, first, count, out_origin);
        return;
    }
    // The number of partitions is only known on the device: in auto mode both organisations are launched and the
    // one whose regime it is not returns at once (decode_range_info).  A range of `count` partitions can never
    // reach the row decoder's regime when count itself is below the threshold, so that launch is skipped.
    const bool tile = c->decoder_mode != 5, rowq = c->decoder_mode != 4 && (c->decoder_mode == 5 || count >= c->rowq_min_units);
    const uint32_t thr = c->decoder_mode == 5 ? 0u : c->rowq_min_units;
    const uint32_t tile_max = rowq ? (thr ? thr - 1u : 0u) : 0xffffffffu;
    if (tile && !(rowq && thr == 0)) {
        if (codec == AOCL_GPU_LZ4)
            LLC_LAUNCH((decode_parts_tile_kernel<TileLz4, false>), 2 * c->sm_count, kTThreads, sizeof(TileShared<TileLz4>), c->stream,
                       in, out, parts, c->d_res, first, count, out_origin, tile_max);
        else
            LLC_LAUNCH((decode_parts_tile_kernel<TileSnappy, true>), 2 * c->sm_count, kTThreads, sizeof(TileShared<TileSnappy>), c->stream,
                       in, out, parts, c->d_res, first, count, out_origin, tile_max);
    }
    if (rowq) {
        if (codec == AOCL_GPU_LZ4)
            LLC_LAUNCH((decode_parts_rowq_kernel<false>), c->sm_count, kQThreads, sizeof(QShared), c->stream,
                       in, out, parts, c->d_res, first, count, out_origin, thr);
        else
            LLC_LAUNCH((decode_parts_rowq_kernel<true>), c->sm_count, kQThreads, sizeof(QShared), c->stream,
                       in, out, parts, c->d_res, first, count, out_origin, thr);
    }
}

extern "C" int32_t aocl_gpu_decompress_open_async(aocl_gpu_ctx_t c, int32_t codec, const void* d_in, size_t n, size_t out_cap) {
    if (!c) return -5;
    begin_call(c);
    if ((codec != AOCL_GPU_LZ4 && codec != AOCL_GPU_SNAPPY) || !d_in || n == 0 || n > 0xffffffffull) { c->last_rc = -2; return -2; }
    if (!ensure_ws(c, sizeof(PartDesc) * kMaxPartitions)) { c->last_rc = -2; return -2; }
    LLC_LAUNCH(rap_parse_kernel, 1, 1024, 0, c->stream, codec, (const uint8_t*)d_in, (uint64_t)n, (uint64_t)out_cap, 1,
               reinterpret_cast<PartDesc*>(c->ws), c->d_res);
    return 0;
}
extern "C" int32_t aocl_gpu_decompress_slab_async(aocl_gpu_ctx_t c, int32_t codec, const void* d_in, void* d_out,
                                                  uint32_t first, uint32_t count) {
    if (!c) return -5;
    if (c->last_rc) return c->last_rc;
    cudaMemsetAsync(&c->d_res->next, 0, sizeof(unsigned int), c->stream);
    launch_decode_range(c, codec, d_in, d_out, reinterpret_cast<PartDesc*>(c->ws), first, count, 0);
    return 0;
}
extern "C" int32_t aocl_gpu_decompress_close_async(aocl_gpu_ctx_t c) {
    if (!c) return -5;
    if (c->last_rc) return c->last_rc;
    end_call(c);
    return 0;
}

extern "C" int32_t aocl_gpu_decompress_async(aocl_gpu_ctx_t c, int32_t codec, const void* d_in, size_t n, void* d_out,
                                             size_t out_cap) {
    return aocl_gpu_decompress_range_async(c, codec, d_in, n, d_out, out_cap, 0, 0xffffffffu, 0);
}

extern "C" int64_t aocl_gpu_decompress(aocl_gpu_ctx_t c, int32_t codec, const void* d_in, size_t n, void* d_out, size_t out_cap) {
    aocl_gpu_decompress_async(c, codec, d_in, n, d_out, out_cap);
    return aocl_gpu_finish(c);
}

// ---------------------------------------------------------------------------------- compress
extern "C" int32_t aocl_gpu_compress_async(aocl_gpu_ctx_t c, int32_t codec, const void* d_in, size_t n, void* d_out,
                                           size_t out_cap) {
    if (!c) return -5;
    begin_call(c);
    const uint32_t* in_flag = c->in_flag;                      // one-shot
    c->in_flag = nullptr;
    if ((codec != AOCL_GPU_LZ4 && codec != AOCL_GPU_SNAPPY) || (!d_in && n) || !d_out || out_cap == 0 || n > 0x7E000000ull) {
        c->last_rc = -2; return -2;
    }
    const uint8_t* src = (const uint8_t*)d_in;
    uint8_t* dst = (uint8_t*)d_out;
    if (codec == AOCL_GPU_LZ4) {
        const uint32_t T = c->lz4_frameless ? 1u : frame_parts(c, n, kLz4Window);
        if (T == 1) {                                          // frame-less block, lz4.c:2674-2677 / 2485-2541
            const uint64_t bound = n + n / 255 + 16;
            LLC_LAUNCH(lz4_encode_single_kernel, 1, 32, 16384, c->stream, src, (uint32_t)n, dst,
                       out_cap >= bound ? -1ll : (long long)out_cap, c->d_res);
        } else {
            const uint64_t pmax = n / T + n % T;
            const uint64_t slot = align_up(pmax + pmax / 255 + 32, 256);
            int stab = c->stab_ctas_per_sm, gtab = c->gtab_ctas_per_sm;
            // (fastparse packs positions in 19 bits: the larger partitions of an imitated host layout take the exact encoder)
            const bool fastparse = c->fastparse && pmax < (1u << 19);
            if (fastparse) { stab = 0; if (gtab <= 0) gtab = 32; }      // one flavour: tables in L2, every partition resident
            if (stab < 0 && gtab < 0) {                        // auto: one wave if at all possible
                if (T <= (uint32_t)c->sm_count * (uint32_t)kStabMax) { stab = kStabMax; gtab = 0; } else { stab = 0; gtab = 32; }
            }
            if (stab < 0) stab = gtab > 0 ? 0 : kStabMax;
            if (gtab < 0) gtab = 0;
            if (stab == 0 && gtab == 0) stab = kStabMax;
            const int g_ctas = gtab * c->sm_count;
            const size_t o_rec = 0, o_plan = align_up(o_rec + sizeof(Lz4Rec) * T + 256, 256);
            const size_t o_tab = align_up(o_plan + sizeof(Lz4Plan) * T, 256);
            const size_t o_scr = align_up(o_tab + (size_t)g_ctas * 16384, 256);
            if (!ensure_ws(c, o_scr + slot * T)) { c->last_rc = -2; return -2; }
            Lz4Rec* rec = reinterpret_cast<Lz4Rec*>(c->ws + o_rec);
            uint32_t* ticket = reinterpret_cast<uint32_t*>(c->ws + o_rec + sizeof(Lz4Rec) * T);
            Lz4Plan* plan = reinterpret_cast<Lz4Plan*>(c->ws + o_plan);
            uint32_t* tables = reinterpret_cast<uint32_t*>(c->ws + o_tab);
            uint8_t* scratch = c->ws + o_scr;
            cudaMemsetAsync(ticket, 0, sizeof(uint32_t), c->stream);
            const uint32_t a_cap = (uint32_t)c->sm_count * (uint32_t)stab;
            const int a_grid = (int)(T < a_cap ? T : a_cap);
            const int enc_slot = prof_begin(c, fastparse ? "lz4_fastparse_parts_kernel" : "lz4_encode_parts_kernel");   // brackets both flavours (fork .. join)
            if (g_ctas > 0 && T > (uint32_t)a_grid) {
                // fork: the global-table flavour shares the ticket and fills the idle warp slots
                cudaEventRecord(c->ev_fork, c->stream);
                cudaStreamWaitEvent(c->side, c->ev_fork, 0);
                if (c->l2_persist_bytes) {                    // keep the hash tables resident in L2
                    const size_t used = (size_t)(T < (uint32_t)g_ctas ? T : (uint32_t)g_ctas) * 16384;
                    cudaStreamAttrValue av = {};
                    av.accessPolicyWindow.base_ptr = tables;
                    av.accessPolicyWindow.num_bytes = used < c->l2_window_bytes ? used : c->l2_window_bytes;
                    const double ratio = (double)c->l2_persist_bytes / (double)av.accessPolicyWindow.num_bytes;
                    av.accessPolicyWindow.hitRatio = ratio > 1.0 ? 1.0f : (float)ratio;
                    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                    cudaStreamSetAttribute(c->side, cudaStreamAttributeAccessPolicyWindow, &av);
                }
                if (fastparse)
                    lz4_fastparse_parts_kernel<<<(int)(T < (uint32_t)g_ctas ? T : (uint32_t)g_ctas), 32, 0, c->side>>>(
                        src, Lz4Range{(uint64_t)n, T, 0u, T}, scratch, slot, rec, ticket, tables, in_flag, c->d_res);
                else
                    lz4_encode_parts_gtab_kernel<<<(int)(T < (uint32_t)g_ctas ? T : (uint32_t)g_ctas), 32, 0, c->side>>>(
                        src, Lz4Range{(uint64_t)n, T, 0u, T}, scratch, slot, rec, ticket, tables, in_flag, c->d_res);
                g_launches.fetch_add(1, std::memory_order_relaxed);
                cudaEventRecord(c->ev_join, c->side);
            }
            if (a_grid > 0) {
                lz4_encode_parts_kernel<<<a_grid, 32, 16384, c->stream>>>(src, Lz4Range{(uint64_t)n, T, 0u, T}, scratch, slot, rec, ticket, in_flag, c->d_res);
                g_launches.fetch_add(1, std::memory_order_relaxed);
            }
            if (g_ctas > 0 && T > (uint32_t)a_grid) cudaStreamWaitEvent(c->stream, c->ev_join, 0);
            prof_end(c, enc_slot);
            LLC_LAUNCH(lz4_stitch_plan_kernel, 1, 1024, 0, c->stream, rec, (uint64_t)n, T, dst, (uint64_t)out_cap, plan, c->d_res);
            LLC_LAUNCH(lz4_compact_kernel, T, 256, 0, c->stream, src, (uint64_t)0, (const uint8_t*)nullptr, (uint64_t)0, scratch, slot, rec,
                       plan, 0u, dst, (uint64_t)0, c->d_res);
        }
    } else {
        if (out_cap < 32 + n + n / 6) { c->last_rc = -2; return -2; }   // api/codec.cpp:262-265
        const uint32_t T = frame_parts(c, n, kSnappyBlock);
        const SnappyGeom g = snappy_geom(n, T);
        const uint32_t F = g.frags_total;
        const uint64_t slot = 76544;                           // >= 32 + 65536 + 65536/6, multiple of 256
        // Fragments are serial chains, so what counts is how many are resident at once.  A shared-memory-table
        // warp costs 33.5 KiB of shared memory (32 KiB table + claim bits), a global-table warp 1.5 KiB plus 32 KiB
        // of L2 (beyond ~24 per SM = 116 MB the tables outgrow L2).  Small frames run shared-memory warps only (6 per
        // SM), larger ones global-table warps only.  Both flavours CAN run side by side on two streams sharing the
        // fragment ticket (AOCL_GPU_SNAPPY_STAB_CTAS), but measured on B200 24 + 5 per SM is no faster than 24 + 0
        // (28.1 ms per GiB either way): the two kernels need different shared-memory carve-outs, so they do not share
        // an SM, and forcing the large carve-out on the global-table warps costs them the L1 that serves their
        // candidate fetches (29 -> 40 ms).
        int sn_gtab = c->snappy_gtab_ctas_per_sm, sn_stab = c->snappy_stab_ctas_per_sm;
        const bool small = F <= (uint32_t)c->sm_count * 6u;
        if (sn_gtab < 0) sn_gtab = small ? 0 : 24;   // measured (lean encoder, global tables only): 12 -> 40.3 ms, 16 -> 32.8, 20 -> 31.3, 24 -> 29.2, 28 -> 45.6
        if (sn_stab < 0) sn_stab = small ? 6 : (sn_gtab > 0 ? 0 : 6);
        if (sn_stab > 6) sn_stab = 6;
        if (sn_gtab == 0 && sn_stab == 0) sn_stab = 6;
        const uint32_t a_grid = (uint32_t)sn_stab * (uint32_t)c->sm_count < F ? (uint32_t)sn_stab * (uint32_t)c->sm_count : F;
        const uint32_t g_want = F - a_grid < (uint32_t)sn_gtab * (uint32_t)c->sm_count ? F - a_grid : (uint32_t)sn_gtab * (uint32_t)c->sm_count;
        const uint32_t g_grid = g_want;
        const size_t o_len = 0, o_off = align_up(o_len + sizeof(uint32_t) * (F + 2), 256);
        const size_t o_tab = align_up(o_off + sizeof(uint64_t) * (F + 1), 256);
        const size_t o_scr = align_up(o_tab + (size_t)g_grid * 32768, 256);
        if (!ensure_ws(c, o_scr + slot * (F + 1))) { c->last_rc = -2; return -2; }
        uint32_t* frag_len = reinterpret_cast<uint32_t*>(c->ws + o_len);
        uint32_t* ticket = frag_len + F + 1;
        uint64_t* frag_off = reinterpret_cast<uint64_t*>(c->ws + o_off);
        uint16_t* tables = reinterpret_cast<uint16_t*>(c->ws + o_tab);
        uint8_t* scratch = c->ws + o_scr;
        cudaMemsetAsync(ticket, 0, sizeof(uint32_t), c->stream);
        const int enc_slot = prof_begin(c, "snappy_encode_frags_kernel");   // brackets both flavours (fork .. join)
        if (g_grid) {
            // fork: the global-table flavour runs on the side stream
            cudaEventRecord(c->ev_fork, c->stream);
            cudaStreamWaitEvent(c->side, c->ev_fork, 0);
            if (c->l2_persist_bytes) {                         // keep the hash tables resident in L2
                cudaStreamAttrValue av = {};
                av.accessPolicyWindow.base_ptr = tables;
                const size_t used = (size_t)g_grid * 32768;
                av.accessPolicyWindow.num_bytes = used < c->l2_window_bytes ? used : c->l2_window_bytes;
                const double ratio = (double)c->l2_persist_bytes / (double)av.accessPolicyWindow.num_bytes;
                av.accessPolicyWindow.hitRatio = ratio > 1.0 ? 1.0f : (float)ratio;
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                cudaStreamSetAttribute(c->side, cudaStreamAttributeAccessPolicyWindow, &av);
            }
            snappy_encode_frags_gtab_kernel<<<g_grid, 32, 0, c->side>>>(src, g, scratch, slot, frag_len, ticket, tables, in_flag, c->d_res);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            cudaEventRecord(c->ev_join, c->side);
        }
        if (a_grid) {
            snappy_encode_frags_kernel<<<a_grid, 32, 32768, c->stream>>>(src, g, scratch, slot, frag_len, ticket, in_flag, c->d_res);
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        if (g_grid) cudaStreamWaitEvent(c->stream, c->ev_join, 0);
        prof_end(c, enc_slot);
        LLC_LAUNCH(snappy_plan_kernel, 1, 1024, 0, c->stream, g, frag_len, frag_off, dst, (uint64_t)out_cap, c->d_res);
        if (F) LLC_LAUNCH(snappy_compact_kernel, F, 256, 0, c->stream, scratch, slot, frag_len, frag_off, 0u, dst, (uint64_t)0, c->d_res);
    }
    end_call(c);
    return 0;
}

extern "C" int64_t aocl_gpu_compress(aocl_gpu_ctx_t c, int32_t codec, const void* d_in, size_t n, void* d_out, size_t out_cap) {
    aocl_gpu_compress_async(c, codec, d_in, n, d_out, out_cap);
    return aocl_gpu_finish(c);
}

// ---------------------------------------------------------------------------------- pages
extern "C" int32_t aocl_gpu_decompress_batch_async(aocl_gpu_ctx_t c, int32_t codec, const void* const* d_in_ptrs,
                                                   const uint32_t* d_in_sizes, void* const* d_out_ptrs,
                                                   const uint32_t* d_out_caps, int64_t* d_status, size_t count) {
    if (!c) return -5;
    begin_call(c);
    c->batch_mode = true;
    if ((codec != AOCL_GPU_LZ4 && codec != AOCL_GPU_SNAPPY) || (count && (!d_in_ptrs || !d_in_sizes || !d_out_ptrs || !d_out_caps || !d_status))) {
        c->last_rc = -2; return -2;
    }
    if (count) {
        const bool want_rowq = c->decoder_mode == 5 || (c->decoder_mode == 0 && count >= c->rowq_min_units);
        if (c->decoder_mode == 1) {
            const uint64_t blocks = (count + 3) / 4;
            const int grid = (int)(blocks < (uint64_t)c->decode_blocks * 4 ? blocks : (uint64_t)c->decode_blocks * 4);
            LLC_LAUNCH(decode_pages_kernel, grid, 128, 0, c->stream, codec, (const uint8_t* const*)d_in_ptrs, d_in_sizes,
                       (uint8_t* const*)d_out_ptrs, d_out_caps, (long long*)d_status, (uint64_t)count, c->d_res);
        } else if (want_rowq && count <= 0xffffffffull) {
            const uint64_t ctas = (count + kQSlots - 1) / kQSlots;
            const int grid = (int)(ctas < (uint64_t)c->sm_count ? ctas : (uint64_t)c->sm_count);
            if (codec == AOCL_GPU_LZ4)
                LLC_LAUNCH((decode_pages_rowq_kernel<false>), grid, kQThreads, sizeof(QShared), c->stream,
                           (const uint8_t* const*)d_in_ptrs, d_in_sizes, (uint8_t* const*)d_out_ptrs, d_out_caps, (long long*)d_status,
                           (uint32_t)count, c->d_res);
            else
                LLC_LAUNCH((decode_pages_rowq_kernel<true>), grid, kQThreads, sizeof(QShared), c->stream,
                           (const uint8_t* const*)d_in_ptrs, d_in_sizes, (uint8_t* const*)d_out_ptrs, d_out_caps, (long long*)d_status,
                           (uint32_t)count, c->d_res);
        } else {
            const int grid = (int)(count < (uint64_t)c->sm_count * 2 ? count : (uint64_t)c->sm_count * 2);
            if (codec == AOCL_GPU_LZ4)
                LLC_LAUNCH((decode_pages_tile_kernel<TileLz4, false>), grid, kTThreads, sizeof(TileShared<TileLz4>), c->stream,
                           (const uint8_t* const*)d_in_ptrs, d_in_sizes, (uint8_t* const*)d_out_ptrs, d_out_caps, (long long*)d_status,
                           (uint64_t)count, c->d_res);
            else
                LLC_LAUNCH((decode_pages_tile_kernel<TileSnappy, true>), grid, kTThreads, sizeof(TileShared<TileSnappy>), c->stream,
                           (const uint8_t* const*)d_in_ptrs, d_in_sizes, (uint8_t* const*)d_out_ptrs, d_out_caps, (long long*)d_status,
                           (uint64_t)count, c->d_res);
        }
    }
    end_call(c);
    return 0;
}

extern "C" int32_t aocl_gpu_compress_batch_async(aocl_gpu_ctx_t c, int32_t codec, const void* const* d_in_ptrs,
                                                 const uint32_t* d_in_sizes, void* const* d_out_ptrs,
                                                 const uint32_t* d_out_caps, int64_t* d_status, size_t count) {
    if (!c) return -5;
    begin_call(c);
    c->batch_mode = true;
    if ((codec != AOCL_GPU_LZ4 && codec != AOCL_GPU_SNAPPY) || (count && (!d_in_ptrs || !d_in_sizes || !d_out_ptrs || !d_out_caps || !d_status))) {
        c->last_rc = -2; return -2;
    }
    if (count) {
        const size_t smem = codec == AOCL_GPU_LZ4 ? 16384 : 32768;
        const uint64_t max_grid = (uint64_t)c->sm_count * (codec == AOCL_GPU_LZ4 ? kStabMax : 6) * 4;
        const int grid = (int)(count < max_grid ? count : max_grid);
        LLC_LAUNCH(encode_pages_kernel, grid, 32, smem, c->stream, codec, (const uint8_t* const*)d_in_ptrs, d_in_sizes,
                   (uint8_t* const*)d_out_ptrs, d_out_caps, (long long*)d_status, (uint64_t)count, c->d_res);
    }
    end_call(c);
    return 0;
}

#include "llc_shard.cuh"
