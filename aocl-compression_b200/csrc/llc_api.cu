// llc_api.cu -- the reference-facing unified API (include/aocl_llc.h) on top of the device
// layer.  Mirrors api/api.cpp:45-189 and the LZ4 / Snappy adapters api/codec.cpp:118-158,
// 253-304 of the reference: same argument checks, same return values, same statistics.
//
// Buffers may live on the host (pageable or pinned) or on the device; host buffers are staged
// through HBM, so measureStats timings include the PCIe transfers exactly like the reference's
// timings include its memcpy epilogues.  Large pinned host buffers are pipelined:
//   compress   the input goes up in pieces on a copy stream while the encoder is already running
//              behind an input watermark (LZ4: stripes across all partitions, because every partition
//              is one serial chain that has to start early; Snappy: in order, fragments are taken in
//              order);
//   decompress the stream is cut into slabs of partitions: H2D of slab k+1 | decode of slab k | D2H
//              of slab k-1 on three streams.
// Pageable buffers and small calls take the plain copy - run - copy sequence.
#include "../../include/aocl_llc.h"
#include "../../include/aocl_llc_gpu.h"
#include "../../include/aocl_llc_native.h"

#include <cuda_runtime.h>
#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>
#include <time.h>

#ifndef AOCL_LLC_BUILD_TAG
#define AOCL_LLC_BUILD_TAG "dev"
#endif

namespace {

constexpr uint64_t kRapMagic = 0x434C4C5F4C434F41ULL;
constexpr int kMaxPieces = 64;
constexpr size_t kPipeMinBytes = size_t(8) << 20;      // below this the plain sequence is as fast

// Everything one call needs on one GPU: a device context (stream, workspace, result block), two staging
// buffers and the copy streams / events of the pipelined host transfers.  Calls from different host threads
// take different HostCtx objects and run concurrently, like the reference (which has no locks on its data
// path: README.md:332-333); the devices of AOCL_GPU_DEVICES are used round robin.
struct HostCtx {
    int device = 0;
    bool busy = false;
    aocl_gpu_ctx_t ctx = nullptr;
    void* d_in = nullptr;  size_t d_in_bytes = 0;     // staging for host inputs
    void* d_out = nullptr; size_t d_out_bytes = 0;    // staging for host outputs
    cudaStream_t up = nullptr, down = nullptr;         // H2D / D2H copy streams
    uint32_t* d_flag = nullptr;                        // input watermark (device)
    uint32_t* h_marks = nullptr;                       // pinned source values for the watermark copies
    cudaEvent_t ev[2 * kMaxPieces + 2] = {};
    bool pipe_ready = false, pipe_failed = false;
    bool lz4_frameless = false;                        // copied from the process-wide setup state at acquire
};

struct Global {
    std::mutex mu;
    std::condition_variable cv;
    bool init_done = false, init_failed = false;
    std::vector<int> devices;          // AOCL_GPU_DEVICES ("0-3", "0,2,5"), else AOCL_GPU_DEVICE, else the current device
    std::vector<HostCtx*> pool;
    int per_device = 4;                // AOCL_GPU_CONTEXTS: concurrent calls served per device (more callers wait)
    unsigned rr = 0;
    bool lz4_setup_done = false;       // setup is once-only until destroy (lz4.c:4999-5016)
    bool lz4_frameless = false;
    bool snappy_setup_done = false;
};
Global g;

// The library never leaves the caller's current device changed.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

void parse_devices_locked() {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return; }
    if (const char* e = getenv("AOCL_GPU_DEVICES")) {
        const char* p = e;
        while (*p) {
            char* end = nullptr;
            long a = strtol(p, &end, 10);
            if (end == p) break;
            long b = a;
            if (*end == '-') { p = end + 1; b = strtol(p, &end, 10); if (end == p) break; }
            for (long d = a; d <= b; d++) if (d >= 0 && d < count) g.devices.push_back((int)d);
            p = end;
            while (*p == ',' || *p == ' ') p++;
        }
    }
    if (g.devices.empty()) {
        int dev = -1;
        if (const char* e = getenv("AOCL_GPU_DEVICE")) dev = atoi(e);
        if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
        if (dev >= 0 && dev < count) g.devices.push_back(dev);
    }
    if (const char* e = getenv("AOCL_GPU_CONTEXTS")) { const int v = atoi(e); if (v >= 1 && v <= 64) g.per_device = v; }
}

// Takes a free context (creating one if its device has fewer than per_device), or waits for one.
HostCtx* acquire() {
    std::unique_lock<std::mutex> lock(g.mu);
    if (!g.init_done) {
        g.init_done = true;
        parse_devices_locked();
        if (g.devices.empty()) {
            g.init_failed = true;
            fprintf(stderr, "[aocl-llc-b200] no usable CUDA device: the B200 library has no CPU fallback\n");
        }
    }
    if (g.init_failed) return nullptr;
    for (;;) {
        const size_t nd = g.devices.size();
        // least-loaded device first, round robin among equals
        int best_dev = -1, best_busy = 1 << 30;
        HostCtx* best_free = nullptr;
        for (size_t k = 0; k < nd; k++) {
            const int dev = g.devices[(g.rr + k) % nd];
            int busy = 0, total = 0;
            HostCtx* free_one = nullptr;
            for (HostCtx* h : g.pool) if (h->device == dev) { total++; if (h->busy) busy++; else if (!free_one) free_one = h; }
            if ((free_one || total < g.per_device) && busy < best_busy) { best_busy = busy; best_dev = dev; best_free = free_one; }
        }
        if (best_dev >= 0) {
            g.rr++;
            HostCtx* h = best_free;
            if (!h) {
                h = new HostCtx();
                h->device = best_dev;
                DeviceGuard guard(best_dev);
                if (aocl_gpu_ctx_create(&h->ctx, best_dev, nullptr) != 0) {
                    delete h;
                    if (g.pool.empty()) {
                        g.init_failed = true;
                        fprintf(stderr, "[aocl-llc-b200] no usable CUDA device: the B200 library has no CPU fallback\n");
                        return nullptr;
                    }
                    g.cv.wait(lock);
                    continue;
                }
                g.pool.push_back(h);
            }
            h->busy = true;
            h->lz4_frameless = g.lz4_frameless;
            return h;
        }
        g.cv.wait(lock);
    }
}
void release(HostCtx* h) {
    { std::lock_guard<std::mutex> lock(g.mu); h->busy = false; }
    g.cv.notify_one();
}
struct Lease {
    HostCtx* h;
    Lease() : h(acquire()) {}
    ~Lease() { if (h) release(h); }
};

bool ensure_pipe(HostCtx& h) {
    if (h.pipe_ready) return true;
    if (h.pipe_failed || getenv("AOCL_GPU_NO_PIPELINE")) return false;
    bool ok = cudaStreamCreateWithFlags(&h.up, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h.down, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&h.d_flag, 256) == cudaSuccess &&
              cudaMallocHost(&h.h_marks, sizeof(uint32_t) * (kMaxPieces + 1)) == cudaSuccess;
    for (auto& e : h.ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { cudaGetLastError(); h.pipe_failed = true; return false; }
    h.pipe_ready = true;
    return true;
}

bool pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

uint32_t rd32(const unsigned char* p) { uint32_t v; memcpy(&v, p, 4); return v; }

bool grow(HostCtx& h, void** buf, size_t* have, size_t need) {
    if (need <= *have) return true;
    if (*buf) { cudaStreamSynchronize((cudaStream_t)aocl_gpu_ctx_stream(h.ctx)); cudaFree(*buf); *buf = nullptr; *have = 0; }
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

// Enqueue the H2D transfer of a pinned host input on the copy stream in pieces, each followed by a
// 4-byte copy that raises the encoder's input watermark (include/aocl_llc_gpu.h), and make the
// compute stream wait only for the watermark reset.  Returns false if this input takes the plain path.
// `in` / `n`: the bytes of this call's piece of the frame -- the whole input, or one GPU's partition range when the call is
// split over several GPUs: then T and common are the partitions of that range and the frame's bytes per partition, and
// mark_base the piece's offset in the frame (Snappy's watermark counts bytes of the frame).
bool stream_piece_up(HostCtx& h, int codec, const char* in, size_t n, void* d_in, cudaStream_t s, uint32_t T, size_t common,
                     uint64_t mark_base);
bool stream_input_up(HostCtx& h, int codec, const char* in, size_t n, void* d_in, cudaStream_t s) {
    const uint32_t T = (uint32_t)aocl_gpu_ctx_partition_count(h.ctx, codec, n);
    if (T < 2 || (codec == LZ4 && h.lz4_frameless)) return false;
    return stream_piece_up(h, codec, in, n, d_in, s, T, n / T, 0);
}
bool stream_piece_up(HostCtx& h, int codec, const char* in, size_t n, void* d_in, cudaStream_t s, uint32_t T, size_t common,
                     uint64_t mark_base) {
    if (n < kPipeMinBytes || !pinned_host(in) || !ensure_pipe(h)) return false;
    bool ok = cudaMemsetAsync(h.d_flag, 0, sizeof(uint32_t), h.up) == cudaSuccess &&
              cudaEventRecord(h.ev[0], h.up) == cudaSuccess && cudaStreamWaitEvent(s, h.ev[0], 0) == cudaSuccess;
    int pieces = 0;
    if (codec == LZ4) {
        // stripe j = bytes [j*w, (j+1)*w) of EVERY partition (a 2-D copy, pitch = partition size)
        const size_t w = 16384;
        pieces = (int)(common / w);
        if (pieces < 1) pieces = 1;
        if (pieces > kMaxPieces) pieces = kMaxPieces;
        for (int j = 0; j < pieces && ok; j++) {
            const size_t lo = (size_t)j * w, width = (j == pieces - 1) ? common - lo : w;
            ok = cudaMemcpy2DAsync((char*)d_in + lo, common, in + lo, common, width, T, cudaMemcpyHostToDevice, h.up) == cudaSuccess;
            if (ok && j == pieces - 1 && n > common * T)     // the last partition's n % T extra bytes
                ok = cudaMemcpyAsync((char*)d_in + common * T, in + common * T, n - common * T, cudaMemcpyHostToDevice, h.up) == cudaSuccess;
            h.h_marks[j] = (j == pieces - 1) ? 0xffffffffu : (uint32_t)(lo + width);
            ok = ok && cudaMemcpyAsync(h.d_flag, &h.h_marks[j], sizeof(uint32_t), cudaMemcpyHostToDevice, h.up) == cudaSuccess;
        }
    } else {
        const size_t piece = size_t(32) << 20;
        pieces = (int)((n + piece - 1) / piece);
        for (int j = 0; j < pieces && ok; j++) {
            const size_t lo = (size_t)j * piece, len = (j == pieces - 1) ? n - lo : piece;
            ok = cudaMemcpyAsync((char*)d_in + lo, in + lo, len, cudaMemcpyHostToDevice, h.up) == cudaSuccess;
            h.h_marks[j] = (j == pieces - 1) ? 0xffffffffu : (uint32_t)(mark_base + lo + len);
            ok = ok && cudaMemcpyAsync(h.d_flag, &h.h_marks[j], sizeof(uint32_t), cudaMemcpyHostToDevice, h.up) == cudaSuccess;
        }
    }
    if (!ok) {                                              // whatever was enqueued must drain before the plain path reuses d_in
        cudaGetLastError();
        cudaStreamSynchronize(h.up);
        return false;
    }
    aocl_gpu_set_input_watermark(h.ctx, h.d_flag);
    return true;
}

// Host-to-host decompress of a well-formed multi-partition RAP stream in slabs.  Returns bytes
// produced, -1 on failure, or -100 when the stream does not qualify (caller takes the plain path,
// which also produces the reference's error behaviour for malformed frames).
// With [p_lo, p_hi) a sub-range of the partitions (one GPU's share of a call that is split over several GPUs) only
// that range is transferred and decoded; the return value is still the whole stream's size.
int64_t decompress_pipelined(HostCtx& h, int codec, const char* in, size_t n, char* out, size_t out_size, cudaStream_t s,
                             uint32_t p_lo = 0, uint32_t p_hi = 0xffffffffu) {
    const unsigned char* u = (const unsigned char*)in;
    uint64_t magic = 0;
    if (n < kPipeMinBytes || n > 0xffffffffull) return -100;
    memcpy(&magic, u, 8);
    if (magic != kRapMagic) return -100;
    const uint32_t frame = rd32(u + 8), T = rd32(u + 12);
    if (T < 64 || T > 65536 || frame != 16 + 12 * (uint64_t)T || frame > n) return -100;
    if (!pinned_host(in) || !pinned_host(out) || !ensure_pipe(h)) return -100;
    const bool ranged = p_lo != 0 || p_hi < T;
    if (p_hi > T) p_hi = T;
    if (p_lo >= p_hi) return -100;
    const uint32_t cnt = p_hi - p_lo;
    // A slab is one wave of the tile decoder (one partition per resident CTA, two CTAs per SM), so the
    // slab decodes cost what the single launch costs; at most 24 slabs.  Entries must be laid out back
    // to back in order.
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h.device);
    uint32_t per = sms > 0 ? 2u * (uint32_t)sms : 256u;
    if (const char* e = getenv("AOCL_GPU_SLAB_PARTS")) { const int v = atoi(e); if (v >= 64) per = (uint32_t)v; }   // tuning knob
    while ((cnt + per - 1) / per > 23) per += sms > 0 ? 2u * (uint32_t)sms : 256u;
    // the first slab is a quarter wave: the download link, which is what the call takes, starts that much earlier
    const uint32_t head = cnt >= 4u * per ? per / 4u : 0u;
    const int K = (int)((cnt - head + per - 1) / per) + (head ? 1 : 0);
    if (K < 2 && !ranged) return -100;
    uint64_t in_end[24], out_end[24];
    uint32_t first[25];
    first[0] = p_lo;
    for (int k = 1; k < K; k++) first[k] = p_lo + (head ? head + (uint32_t)(k - 1) * per : (uint32_t)k * per);
    first[K] = p_hi;
    // every entry is checked (each GPU of a split call checks the whole table), the slabs cover the range
    uint64_t pos = frame, total = 0, in_lo = 0, out_lo = 0, first_off = 0;
    bool seen_first = false;
    int k = 0;
    for (uint32_t i = 0; i < T; i++) {
        if (i == p_lo) { in_lo = pos; out_lo = total; }
        const unsigned char* e = u + 16 + 12 * (size_t)i;
        const uint64_t off = rd32(e), clen = rd32(e + 4), dlen = rd32(e + 8);
        if (off < pos || off + clen > n) return -100;
        if (clen && !seen_first) { first_off = off; seen_first = true; }
        pos = off + clen;
        if (clen) total += dlen;
        if (i >= p_lo && i < p_hi && i + 1 == first[k + 1]) { in_end[k] = (i + 1 == T) ? n : pos; out_end[k] = total; k++; }
    }
    if (total > out_size || total == 0 || !seen_first) return -100;
    if (in_lo < first_off) in_lo = first_off;                // (the frame, and Snappy's varint behind it, travel with the head)
    const uint64_t out_hi = out_end[K - 1];
    if (!grow(h, &h.d_in, &h.d_in_bytes, n) || !grow(h, &h.d_out, &h.d_out_bytes, out_hi > out_lo ? out_hi - out_lo : 1)) return -1;

    // uploads, in order, on the copy stream: the head of the stream (frame, entry table, varint), then the slabs
    bool ok = cudaMemcpyAsync(h.d_in, in, first_off, cudaMemcpyHostToDevice, h.up) == cudaSuccess &&
              cudaEventRecord(h.ev[0], h.up) == cudaSuccess;
    uint64_t lo = in_lo;
    for (int j = 0; j < K && ok; j++) {
        if (in_end[j] > lo) ok = cudaMemcpyAsync((char*)h.d_in + lo, in + lo, in_end[j] - lo, cudaMemcpyHostToDevice, h.up) == cudaSuccess;
        ok = ok && cudaEventRecord(h.ev[1 + j], h.up) == cudaSuccess;
        if (in_end[j] > lo) lo = in_end[j];
    }
    ok = ok && cudaStreamWaitEvent(s, h.ev[0], 0) == cudaSuccess &&
         aocl_gpu_decompress_open_async(h.ctx, codec, h.d_in, n, out_size) == 0;
    // partitions are decoded to their offsets in the stream's output: the staging buffer stands for [out_lo, out_hi) of it
    char* const d_base = reinterpret_cast<char*>(reinterpret_cast<uintptr_t>(h.d_out) - (uintptr_t)out_lo);
    uint64_t olo = out_lo;
    for (int j = 0; j < K && ok; j++) {
        ok = cudaStreamWaitEvent(s, h.ev[1 + j], 0) == cudaSuccess &&
             aocl_gpu_decompress_slab_async(h.ctx, codec, h.d_in, d_base, first[j], first[j + 1] - first[j]) == 0 &&
             cudaEventRecord(h.ev[1 + kMaxPieces + j], s) == cudaSuccess &&
             cudaStreamWaitEvent(h.down, h.ev[1 + kMaxPieces + j], 0) == cudaSuccess;
        if (ok && out_end[j] > olo)
            ok = cudaMemcpyAsync(out + olo, d_base + olo, out_end[j] - olo, cudaMemcpyDeviceToHost, h.down) == cudaSuccess;
        if (out_end[j] > olo) olo = out_end[j];
    }
    if (ok) ok = aocl_gpu_decompress_close_async(h.ctx) == 0;
    const int64_t r = ok ? aocl_gpu_finish(h.ctx) : -1;
    cudaStreamSynchronize(h.up);
    cudaStreamSynchronize(s);
    const bool down_ok = cudaStreamSynchronize(h.down) == cudaSuccess;
    if (!ok) cudaGetLastError();
    if (r < 0 || !down_ok || (uint64_t)r != total) return -1;
    return r;
}

// ------------------------------------------------------------------------------------------------
// ONE call over several GPUs (AOCL_GPU_SHARD=1 with two or more devices in AOCL_GPU_DEVICES; host buffers).
// The reference forks OpenMP threads over the partitions of a frame (lz4.c:2684-2731, threads/threads.c:121-153);
// here one host thread per GPU takes a contiguous partition range: it uploads only its slice over its own PCIe
// link, runs the sharded device call (csrc/llc_shard.cuh: the GPUs exchange the per-partition records and the
// boundary literals over NCCL) and downloads its piece straight to its final place in the caller's buffer.
// ------------------------------------------------------------------------------------------------
struct ShardGroup {
    std::mutex mu;                     // one sharded call at a time
    std::vector<HostCtx*> rank;        // one dedicated context per device, NCCL communicator joined
    bool ready = false, failed = false;
};
ShardGroup sg;
std::atomic<uint64_t> g_sharded_calls{0};
constexpr size_t kShardMinBytes = size_t(32) << 20;

bool shard_wanted() {
    const char* e = getenv("AOCL_GPU_SHARD");
    return e && atoi(e) != 0;
}

// Creates the per-device contexts and the communicator (once).  Call with sg.mu held.
bool shard_group_init() {
    if (sg.ready) return true;
    if (sg.failed) return false;
    std::vector<int> devs;
    {
        std::lock_guard<std::mutex> lock(g.mu);
        if (!g.init_done) { g.init_done = true; parse_devices_locked(); if (g.devices.empty()) g.init_failed = true; }
        devs = g.devices;
    }
    const int R = (int)devs.size();
    unsigned char id[128];
    if (R < 2 || aocl_gpu_shard_unique_id(id) != 0) { sg.failed = true; return false; }
    sg.rank.assign(R, nullptr);
    std::vector<std::thread> th;
    std::atomic<int> bad{0};
    for (int r = 0; r < R; r++)
        th.emplace_back([&, r] {
            HostCtx* h = new HostCtx();
            h->device = devs[r];
            sg.rank[r] = h;
            DeviceGuard guard(h->device);
            if (aocl_gpu_ctx_create(&h->ctx, h->device, nullptr) != 0 || aocl_gpu_shard_init(h->ctx, id, r, R) != 0) bad++;
        });
    for (auto& t : th) t.join();
    if (bad.load()) { sg.failed = true; return false; }
    sg.ready = true;
    return true;
}

// Returns bytes produced, -1 on failure, -100 when the call does not qualify (the caller takes the one-GPU path).
int64_t run_codec_sharded(bool compress, int codec, char* in, size_t in_size, char* out, size_t out_size) {
    if ((compress ? in_size : out_size) < kShardMinBytes || in_size > 0x7E000000ull) return -100;
    if (compress && codec == LZ4) {                          // optOff / AOCL_DISABLE_OPT: one frame-less block, nothing to shard
        std::lock_guard<std::mutex> lock(g.mu);
        if (g.lz4_frameless) return -100;
    }
    if (compress) { const char* e = getenv("AOCL_GPU_PARTITIONS"); if (e && atoi(e) > 0) return -100; }   // an imitated host layout is not sharded
    std::lock_guard<std::mutex> lock(sg.mu);
    if (!shard_group_init()) return -100;
    const int R = (int)sg.rank.size();
    uint32_t T = 0;
    std::vector<uint64_t> in_lo(R), in_hi(R), out_need(R);
    std::vector<uint32_t> parts(R, 0);
    uint64_t head_bytes = 0;                                  // decompress: frame header (+ varint) every rank needs
    if (compress) {
        T = (uint32_t)aocl_gpu_partition_count(codec, in_size);
        if (T < (uint32_t)R * 2u) return -100;
        for (int r = 0; r < R; r++) {
            uint32_t first, count; uint64_t off, len;
            if (aocl_gpu_shard_range(codec, in_size, r, R, &first, &count, &off, &len) != 0) return -100;
            in_lo[r] = off; in_hi[r] = off + len; parts[r] = count;
            out_need[r] = aocl_gpu_compress_bound(codec, len) + 16 + 12 * (uint64_t)T + 64;
        }
    } else {
        const unsigned char* u = (const unsigned char*)in;
        uint64_t magic = 0;
        if (in_size < 16) return -100;
        memcpy(&magic, u, 8);
        if (magic != kRapMagic) return -100;
        const uint32_t frame = rd32(u + 8);
        T = rd32(u + 12);
        if (T < (uint32_t)R * 2u || T > 65536 || frame != 16 + 12 * (uint64_t)T || frame > in_size) return -100;
        head_bytes = std::min<uint64_t>(in_size, (uint64_t)frame + 8);
        uint64_t pos = frame, total = 0;
        for (int r = 0; r < R; r++) {
            const uint32_t lo = (uint32_t)((uint64_t)T * r / R), hi = (uint32_t)((uint64_t)T * (r + 1) / R);
            uint64_t first_off = 0, need = 0;
            bool seen = false;
            for (uint32_t i = lo; i < hi; i++) {
                const unsigned char* e = u + 16 + 12 * (size_t)i;
                const uint64_t off = rd32(e), clen = rd32(e + 4), dlen = rd32(e + 8);
                if (!clen) continue;
                if (off < pos || off + clen > in_size) return -100;          // not laid out back to back: the plain path decides
                if (!seen) { first_off = off; seen = true; }
                pos = off + clen;
                need += dlen;
            }
            in_lo[r] = seen ? first_off : pos; in_hi[r] = pos; out_need[r] = need;
            total += need;
        }
        if (total > out_size) return -1;
    }
    const bool pipe_decode = !compress && T >= 64 && in_size >= kPipeMinBytes && pinned_host(in) && pinned_host(out);
    std::vector<int64_t> result(R, -1);
    std::vector<uint64_t> piece_off(R, 0), piece_len(R, 0);
    std::atomic<int> alloc_bad{0}, arrived{0};
    std::vector<std::thread> th;
    for (int r = 0; r < R; r++)
        th.emplace_back([&, r] {
            HostCtx& h = *sg.rank[r];
            DeviceGuard guard(h.device);
            cudaStream_t s = (cudaStream_t)aocl_gpu_ctx_stream(h.ctx);
            // all allocations first, then agree: nobody enters a collective unless everybody can
            const size_t need_in = compress ? (size_t)(in_hi[r] - in_lo[r]) : in_size;
            const bool ok_alloc = grow(h, &h.d_in, &h.d_in_bytes, need_in ? need_in : 1) && grow(h, &h.d_out, &h.d_out_bytes, out_need[r] ? out_need[r] : 1);
            if (!ok_alloc) alloc_bad++;
            arrived++;
            while (arrived.load() < R) std::this_thread::yield();
            if (alloc_bad.load()) return;
            uint64_t off = 0, len = 0;
            int64_t tot;
            if (compress) {
                // the slice goes up in stripes behind the encoder's watermark, as in the single-GPU call
                const bool streamed = stream_piece_up(h, codec, in + in_lo[r], (size_t)(in_hi[r] - in_lo[r]), h.d_in, s, parts[r],
                                                      (size_t)(in_size / T), in_lo[r]);
                if (!streamed) cudaMemcpyAsync(h.d_in, in + in_lo[r], in_hi[r] - in_lo[r], cudaMemcpyHostToDevice, s);
                tot = aocl_gpu_compress_sharded(h.ctx, codec, h.d_in, in_size, h.d_out, out_need[r], &off, &len);
                if (streamed) { cudaStreamSynchronize(h.up); aocl_gpu_set_input_watermark(h.ctx, nullptr); }
                if (tot > 0 && off + len <= out_size && cudaMemcpyAsync(out + off, h.d_out, len, cudaMemcpyDeviceToHost, s) == cudaSuccess &&
                    cudaStreamSynchronize(s) == cudaSuccess) result[r] = tot;
                else if (tot > 0) result[r] = -1;
            } else if (pipe_decode) {
                // pinned buffers: every GPU runs the slab pipeline (upload | decode | download) on its partition range;
                // the ranks are threads of this process, so they agree on the outcome here and need no collective
                result[r] = decompress_pipelined(h, codec, in, in_size, out, out_size, s, (uint32_t)((uint64_t)T * r / R),
                                                 (uint32_t)((uint64_t)T * (r + 1) / R));
                cudaGetLastError();
                return;
            } else {
                cudaMemcpyAsync(h.d_in, in, head_bytes, cudaMemcpyHostToDevice, s);
                if (in_hi[r] > in_lo[r]) cudaMemcpyAsync((char*)h.d_in + in_lo[r], in + in_lo[r], in_hi[r] - in_lo[r], cudaMemcpyHostToDevice, s);
                tot = aocl_gpu_decompress_sharded(h.ctx, codec, h.d_in, in_size, h.d_out, out_need[r], &off, &len);
                if (tot >= 0 && off + len <= out_size && (len == 0 || cudaMemcpyAsync(out + off, h.d_out, len, cudaMemcpyDeviceToHost, s) == cudaSuccess) &&
                    cudaStreamSynchronize(s) == cudaSuccess) result[r] = tot;
            }
            piece_off[r] = off; piece_len[r] = len;
            cudaGetLastError();
        });
    for (auto& t : th) t.join();
    if (alloc_bad.load()) return -1;
    if (pipe_decode) for (int r = 0; r < R; r++) if (result[r] == -100) return -100;     // does not qualify after all: the single-GPU path decides
    for (int r = 0; r < R; r++) if (result[r] < 0 || result[r] != result[0]) return -1;
    g_sharded_calls++;
    return result[0];
}

// Returns bytes produced or a negative codec error (the adapters' CODEC_ERROR).
int64_t run_codec(bool compress, int codec, char* in, size_t in_size, char* out, size_t out_size) {
    if (compress) {
        if ((in == nullptr && in_size != 0) || out == nullptr || out_size == 0) return -1;   // lz4.c:2656-2657, snappy.cc:2499
        if (codec == SNAPPY && out_size < 32 + in_size + in_size / 6) return -1;             // api/codec.cpp:262-265
    } else {
        if (in == nullptr || in_size == 0) return -1;                                        // lz4.c:3822, 3859
        if (out == nullptr && out_size != 0) return -1;
        if (codec == LZ4 && out == nullptr) return -1;
    }
    if (shard_wanted() && (in_size ? !on_device(in) : false) && !on_device(out)) {
        const int64_t r = run_codec_sharded(compress, codec, in, in_size, out, out_size);
        if (r != -100) return r;
    }
    Lease lease;                                            // one context for the whole call; other threads take others
    if (!lease.h) return -1;
    HostCtx& h = *lease.h;
    DeviceGuard guard(h.device);                            // allocations, streams and events below belong to h.device
    cudaStream_t s = (cudaStream_t)aocl_gpu_ctx_stream(h.ctx);

    const bool in_dev = in_size ? on_device(in) : true;
    const bool out_dev = on_device(out);
    if (!compress && !in_dev && !out_dev) {
        const int64_t r = decompress_pipelined(h, codec, in, in_size, out, out_size, s);
        if (r != -100) return r;
    }
    // both staging buffers exist before anything is enqueued: no early return may leave a copy in flight
    size_t stage_cap = out_size;
    if (!out_dev) {
        if (compress) { const size_t b = aocl_gpu_compress_bound(codec, in_size); if (b < stage_cap) stage_cap = b; }
        if (!grow(h, &h.d_out, &h.d_out_bytes, stage_cap ? stage_cap : 1)) return -1;
    }
    if (in_size && !in_dev && !grow(h, &h.d_in, &h.d_in_bytes, in_size)) return -1;
    const void* d_in = in;
    bool streamed = false;
    if (in_size && !in_dev) {
        if (compress) streamed = stream_input_up(h, codec, in, in_size, h.d_in, s);
        if (!streamed && cudaMemcpyAsync(h.d_in, in, in_size, cudaMemcpyHostToDevice, s) != cudaSuccess) { cudaGetLastError(); return -1; }
        d_in = h.d_in;
    }
    void* d_out = out_dev ? (void*)out : h.d_out;
    aocl_gpu_set_lz4_frameless(h.ctx, h.lz4_frameless ? 1 : 0);
    if (compress) aocl_gpu_compress_async(h.ctx, codec, d_in, in_size, d_out, out_size);
    else          aocl_gpu_decompress_async(h.ctx, codec, d_in, in_size, d_out, out_size);
    const int64_t r = aocl_gpu_finish(h.ctx);
    if (streamed) {                                         // nothing of this call may still be in flight or armed
        cudaStreamSynchronize(h.up);
        aocl_gpu_set_input_watermark(h.ctx, nullptr);
    }
    if (r < 0) return -1;
    if (!out_dev && r > 0) {
        if ((size_t)r > stage_cap) return -1;
        if (cudaMemcpyAsync(out, d_out, (size_t)r, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) { cudaGetLastError(); return -1; }
    }
    return r;
}

bool ensure_any_ctx() {
    Lease lease;
    return lease.h != nullptr;
}

bool served(aocl_compression_type t) { return t == LZ4 || t == SNAPPY; }
// LZ4HC streams are plain LZ4 blocks; the reference decodes them with the LZ4 decoder (api/codec.h:168 puts
// aocl_lz4_decompress in the lz4hc row, api/codec.cpp:180-188).  The HC encoder itself is not on the path.
bool served_decode(aocl_compression_type t) { return served(t) || t == LZ4HC; }

}  // namespace

extern "C" int64_t aocl_llc_compress(aocl_compression_desc* h, aocl_compression_type codec_type) {
    if (h && codec_type == LZ4HC) return ERR_EXCLUDED_METHOD;   // decode-only codec id (the HC encoder is not built)
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
    if (!h || !served_decode(codec_type)) return ERR_COMPRESSION_FAILED;
    const uint64_t t0 = now_ns();
    const int64_t ret = run_codec(false, codec_type == LZ4HC ? (int)LZ4 : (int)codec_type, h->inBuf, h->inSize, h->outBuf, h->outSize);
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
    if (!served_decode(codec_type)) return ERR_EXCLUDED_METHOD;   // api/api.cpp:156-162
    std::unique_lock<std::mutex> lock(g.mu);
    if (codec_type == LZ4 && !g.lz4_setup_done) {
        // optOff (or AOCL_DISABLE_OPT=ON, utils/utils.cpp:207-219) selects the reference's
        // single-threaded layout: one frame-less LZ4 block (lz4.c:4927-4932)
        const char* e = getenv("AOCL_DISABLE_OPT");
        g.lz4_frameless = h->optOff != 0 || (e && strcmp(e, "ON") == 0);
        g.lz4_setup_done = true;
    }
    if (codec_type == SNAPPY) g.snappy_setup_done = true;
    h->workBuf = nullptr;                                     // the reference returns NULL for both codecs
    lock.unlock();
    return ensure_any_ctx() ? 0 : ERR_COMPRESSION_FAILED;
}

extern "C" void aocl_llc_destroy(aocl_compression_desc* h, aocl_compression_type codec_type) {
    (void)h;
    std::lock_guard<std::mutex> lock(g.mu);
    if (codec_type == LZ4) { g.lz4_setup_done = false; g.lz4_frameless = false; }   // lz4.c:5012-5016
    if (codec_type == SNAPPY) g.snappy_setup_done = false;
}

// Calls that were split over several GPUs so far (AOCL_GPU_SHARD=1); diagnostics / tests.
extern "C" uint64_t aocl_gpu_sharded_host_calls(void) { return g_sharded_calls.load(); }

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

// ------------------------------------------------------------------------------------------------
// The codecs' native C entry points (include/aocl_llc_native.h), thin shims over run_codec().
// ------------------------------------------------------------------------------------------------
extern "C" int LZ4_compressBound(int n) {                      // LZ4_COMPRESSBOUND, lz4.h:258-287
    return (unsigned)n > 0x7E000000u ? 0 : n + n / 255 + 16;
}

extern "C" int LZ4_compress_default(const char* src, char* dst, int src_size, int dst_cap) {
    if (src_size < 0 || dst_cap <= 0) return 0;
    const int64_t r = run_codec(true, (int)LZ4, const_cast<char*>(src), (size_t)src_size, dst, (size_t)dst_cap);
    return r > 0 && r <= 0x7fffffff ? (int)r : 0;               // 0 = failure (lz4.h:196-203)
}

extern "C" int LZ4_decompress_safe(const char* src, char* dst, int csize, int dst_cap) {
    if (csize < 0 || dst_cap < 0) return -1;
    const int64_t r = run_codec(false, (int)LZ4, const_cast<char*>(src), (size_t)csize, dst, (size_t)dst_cap);
    return r >= 0 && r <= 0x7fffffff ? (int)r : -1;             // negative = malformed (lz4.h:222-232)
}

extern "C" size_t snappy_max_compressed_length(size_t n) { return 32 + n + n / 6; }   // snappy.cc:160-182

// varint32 of the uncompressed length; frame-aware (snappy.cc:596-615)
static bool snappy_length_of(const char* p, size_t n, size_t* out) {
    if (!p || !out) return false;
    unsigned char head[16 + 5] = {0};
    size_t skip = 0;
    if (n >= 16) {
        if (on_device(p)) { if (cudaMemcpy(head, p, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return false; } }
        else memcpy(head, p, 16);
        uint64_t magic; memcpy(&magic, head, 8);
        if (magic == kRapMagic) { skip = rd32(head + 8); if (skip > n) return false; }
    }
    const size_t take = n - skip < 5 ? n - skip : 5;
    if (on_device(p)) { if (take && cudaMemcpy(head, p + skip, take, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return false; } }
    else memcpy(head, p + skip, take);
    uint32_t v = 0;
    for (size_t i = 0; i < take; i++) {                         // Varint::Parse32WithLimit, snappy-stubs-internal.h:440-470
        const uint32_t b = head[i];
        if (i == 4 && b > 15) return false;
        v |= (b & 127u) << (7 * i);
        if (b < 128) { *out = v; return true; }
    }
    return false;
}

extern "C" snappy_status snappy_uncompressed_length(const char* compressed, size_t n, size_t* result) {
    return snappy_length_of(compressed, n, result) ? SNAPPY_OK : SNAPPY_INVALID_INPUT;
}

extern "C" snappy_status snappy_compress(const char* input, size_t n, char* compressed, size_t* compressed_length) {
    if (!compressed_length) return SNAPPY_INVALID_INPUT;
    if (*compressed_length < snappy_max_compressed_length(n)) return SNAPPY_BUFFER_TOO_SMALL;   // snappy-c.cc:38-40
    const int64_t r = run_codec(true, (int)SNAPPY, const_cast<char*>(input), n, compressed, *compressed_length);
    if (r < 0) return SNAPPY_INVALID_INPUT;
    *compressed_length = (size_t)r;
    return SNAPPY_OK;
}

extern "C" snappy_status snappy_uncompress(const char* compressed, size_t n, char* uncompressed, size_t* uncompressed_length) {
    size_t real = 0;
    if (!uncompressed_length || !snappy_length_of(compressed, n, &real)) return SNAPPY_INVALID_INPUT;
    if (*uncompressed_length < real) return SNAPPY_BUFFER_TOO_SMALL;                            // snappy-c.cc:55-57
    const int64_t r = run_codec(false, (int)SNAPPY, const_cast<char*>(compressed), n, uncompressed, *uncompressed_length);
    if (r < 0 || (size_t)r != real) return SNAPPY_INVALID_INPUT;
    *uncompressed_length = real;
    return SNAPPY_OK;
}
