// llc_api.cu -- the reference-facing unified API (include/aocl_llc.h) on top of the device
// layer.  Mirrors api/api.cpp:45-189 and the LZ4 / Snappy adapters api/codec.cpp:118-158,
// 253-304 of the reference: same argument checks, same return values, same statistics.
//
// Buffers may live on the host (pageable or pinned) or on the device; host buffers are staged
// through HBM with cudaMemcpyAsync on the context's stream, so measureStats timings include
// the PCIe transfers exactly like the reference's timings include its memcpy epilogues.
#include "../../include/aocl_llc.h"
#include "../../include/aocl_llc_gpu.h"

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <time.h>

#ifndef AOCL_LLC_BUILD_TAG
#define AOCL_LLC_BUILD_TAG "dev"
#endif

namespace {

constexpr uint64_t kRapMagic = 0x434C4C5F4C434F41ULL;

struct Global {
    std::mutex mu;                 // one in-flight call per process-wide context
    aocl_gpu_ctx_t ctx = nullptr;
    bool ctx_failed = false;
    void* d_in = nullptr;  size_t d_in_bytes = 0;     // staging for host inputs
    void* d_out = nullptr; size_t d_out_bytes = 0;    // staging for host outputs
    bool lz4_setup_done = false;   // setup is once-only until destroy (lz4.c:4999-5016)
    bool lz4_frameless = false;
    bool snappy_setup_done = false;
};
Global g;

bool ensure_ctx() {
    if (g.ctx) return true;
    if (g.ctx_failed) return false;
    int dev = -1;
    if (const char* e = getenv("AOCL_GPU_DEVICE")) dev = atoi(e);
    if (aocl_gpu_ctx_create(&g.ctx, dev, nullptr) != 0) {
        g.ctx_failed = true;
        fprintf(stderr, "[aocl-llc-b200] no usable CUDA device: the B200 library has no CPU fallback\n");
        return false;
    }
    return true;
}

bool grow(void** buf, size_t* have, size_t need) {
    if (need <= *have) return true;
    if (*buf) { cudaStreamSynchronize((cudaStream_t)aocl_gpu_ctx_stream(g.ctx)); cudaFree(*buf); *buf = nullptr; *have = 0; }
    need = (need + (size_t(1) << 20)) & ~((size_t(1) << 20) - 1);
    if (cudaMalloc(buf, need) != cudaSuccess) { cudaGetLastError(); return false; }
    *have = need;
    return true;
}

bool on_device(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

uint64_t now_ns() {
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);   // utils/utils.h:243-247
    return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

// Returns bytes produced or a negative codec error (the adapters' CODEC_ERROR).
int64_t run_codec(bool compress, int codec, char* in, size_t in_size, char* out, size_t out_size) {
    std::lock_guard<std::mutex> lock(g.mu);
    if (!ensure_ctx()) return -1;
    cudaStream_t s = (cudaStream_t)aocl_gpu_ctx_stream(g.ctx);

    if (compress) {
        if ((in == nullptr && in_size != 0) || out == nullptr || out_size == 0) return -1;   // lz4.c:2656-2657, snappy.cc:2499
        if (codec == SNAPPY && out_size < 32 + in_size + in_size / 6) return -1;             // api/codec.cpp:262-265
    } else {
        if (in == nullptr || in_size == 0) return -1;                                        // lz4.c:3822, 3859
        if (out == nullptr && out_size != 0) return -1;
        if (codec == LZ4 && out == nullptr) return -1;
    }

    const void* d_in = in;
    if (in_size && !on_device(in)) {
        if (!grow(&g.d_in, &g.d_in_bytes, in_size)) return -1;
        if (cudaMemcpyAsync(g.d_in, in, in_size, cudaMemcpyHostToDevice, s) != cudaSuccess) { cudaGetLastError(); return -1; }
        d_in = g.d_in;
    }
    const bool out_dev = on_device(out);
    void* d_out = out;
    size_t stage_cap = out_size;
    if (!out_dev) {
        if (compress) { const size_t b = aocl_gpu_compress_bound(codec, in_size); if (b < stage_cap) stage_cap = b; }
        if (!grow(&g.d_out, &g.d_out_bytes, stage_cap ? stage_cap : 1)) return -1;
        d_out = g.d_out;
    }
    aocl_gpu_set_lz4_frameless(g.ctx, g.lz4_frameless ? 1 : 0);
    if (compress) aocl_gpu_compress_async(g.ctx, codec, d_in, in_size, d_out, out_size);
    else          aocl_gpu_decompress_async(g.ctx, codec, d_in, in_size, d_out, out_size);
    const int64_t r = aocl_gpu_finish(g.ctx);
    if (r < 0) return -1;
    if (!out_dev && r > 0) {
        if ((size_t)r > stage_cap) return -1;
        if (cudaMemcpyAsync(out, d_out, (size_t)r, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) { cudaGetLastError(); return -1; }
    }
    return r;
}

bool served(aocl_compression_type t) { return t == LZ4 || t == SNAPPY; }

}  // namespace

extern "C" int64_t aocl_llc_compress(aocl_compression_desc* h, aocl_compression_type codec_type) {
    if (!h || !served(codec_type)) return ERR_COMPRESSION_FAILED;
    const uint64_t t0 = now_ns();
    int64_t ret = run_codec(true, (int)codec_type, h->inBuf, h->inSize, h->outBuf, h->outSize);
    if (codec_type == LZ4 && ret == 0) ret = -1;              // api/codec.cpp:132-136 (res > 0 else CODEC_ERROR)
    const uint64_t t1 = now_ns();
    if (h->measureStats == 1) {                               // api/api.cpp:69-75
        h->cSize = (uint64_t)ret;
        h->cTime = t1 - t0;
        h->cSpeed = (float)((h->inSize * 1000.0) / h->cTime);
    }
    return ret < 0 ? ERR_COMPRESSION_FAILED : ret;
}

extern "C" int64_t aocl_llc_decompress(aocl_compression_desc* h, aocl_compression_type codec_type) {
    if (!h || !served(codec_type)) return ERR_COMPRESSION_FAILED;
    const uint64_t t0 = now_ns();
    const int64_t ret = run_codec(false, (int)codec_type, h->inBuf, h->inSize, h->outBuf, h->outSize);
    const uint64_t t1 = now_ns();
    if (h->measureStats == 1) {                               // api/api.cpp:110-116
        h->dSize = (uint64_t)ret;
        h->dTime = t1 - t0;
        h->dSpeed = (float)((h->dSize * 1000.0) / h->dTime);
    }
    return ret < 0 ? ERR_COMPRESSION_FAILED : ret;
}

extern "C" int32_t aocl_llc_setup(aocl_compression_desc* h, aocl_compression_type codec_type) {
    if ((int)codec_type < (int)LZ4 || (int)codec_type >= (int)AOCL_COMPRESSOR_ALGOS_NUM) return ERR_UNSUPPORTED_METHOD;   // api/api.cpp:133-138
    if (!h) return ERR_INVALID_INPUT;
    h->optLevel = 4;                                          // utils/utils.cpp:148-172 overwrites the caller's value
    if (!served(codec_type)) return ERR_EXCLUDED_METHOD;      // api/api.cpp:156-162
    std::lock_guard<std::mutex> lock(g.mu);
    if (codec_type == LZ4 && !g.lz4_setup_done) {
        // optOff (or AOCL_DISABLE_OPT=ON, utils/utils.cpp:207-219) selects the reference's
        // single-threaded layout: one frame-less LZ4 block (lz4.c:4927-4932)
        const char* e = getenv("AOCL_DISABLE_OPT");
        g.lz4_frameless = h->optOff != 0 || (e && strcmp(e, "ON") == 0);
        g.lz4_setup_done = true;
    }
    if (codec_type == SNAPPY) g.snappy_setup_done = true;
    h->workBuf = nullptr;                                     // the reference returns NULL for both codecs
    return ensure_ctx() ? 0 : ERR_COMPRESSION_FAILED;
}

extern "C" void aocl_llc_destroy(aocl_compression_desc* h, aocl_compression_type codec_type) {
    (void)h;
    std::lock_guard<std::mutex> lock(g.mu);
    if (codec_type == LZ4) { g.lz4_setup_done = false; g.lz4_frameless = false; }   // lz4.c:5012-5016
    if (codec_type == SNAPPY) g.snappy_setup_done = false;
}

extern "C" const char* aocl_llc_version(void) {
    return "AOCL-Compression 4.2.0 B200 LZ4/Snappy RAP path, Build " AOCL_LLC_BUILD_TAG;
}

// The GPU always emits the saturated layout T = P(n); for int32-sized inputs P(n) <= 8192
// (threads/threads.c:315-318 returns 16 + 12 * omp_get_max_threads()).
extern "C" int32_t aocl_get_rap_frame_bound_mt(void) { return 16 + 12 * 8192; }

extern "C" int32_t aocl_skip_rap_frame_mt(char* src, int32_t src_size) {
    if (src == nullptr) return ERR_INVALID_INPUT;             // threads/threads.c:322-323
    if (src_size < 8) return 0;
    unsigned char head[16] = {0};
    const size_t take = src_size < 16 ? 8 : 16;
    if (on_device(src)) { if (cudaMemcpy(head, src, take, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return 0; } }
    else memcpy(head, src, take);
    uint64_t magic; memcpy(&magic, head, 8);
    if (magic != kRapMagic) return 0;
    uint32_t frame; memcpy(&frame, head + 8, 4);
    return (int32_t)frame;
}
