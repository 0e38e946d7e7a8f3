// llc_common.cuh -- shared device helpers for the B200 LZ4 / Snappy RAP kernels.
//
// Conventions used by every kernel in this directory:
//  * one warp owns one independent unit (RAP partition, Snappy fragment or page) and
//    runs the codec's serial state machine with warp-uniform control flow; the 32 lanes
//    are used for the data-parallel parts (copies, match extension, speculative probes);
//  * all multi-byte fields are little endian and unaligned (threads/threads.h:46-72 of the
//    reference); unaligned loads are composed from aligned 32-bit loads;
//  * errors are reported through a per-unit status word, never by trapping.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace llc {

constexpr uint64_t kRapMagic = 0x434C4C5F4C434F41ULL;   // "AOCL_LLC", threads/threads.h:79
constexpr int kRapHeaderBytes = 16;                     // magic(8) | frame_len(4) | T(4)
constexpr int kRapEntryBytes = 12;                      // offset | comp_len | decomp_len
constexpr uint32_t kLz4Window = 65567;                  // LZ4_COMPRESS_INPLACE_MARGIN (lz4.c:2667)
constexpr uint32_t kSnappyBlock = 65536;                // kBlockSize (snappy.h:507)
constexpr uint32_t kWindowFactor = 4;
constexpr uint32_t kMaxPartitions = 1u << 16;           // int32-sized inputs give T <= 8192
constexpr unsigned kFull = 0xffffffffu;

// status codes stored per unit / per call (negative = failure)
constexpr int64_t kErrCorrupt = -2;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
// Pull the line that holds p into L1 (no register result)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }

// ---- unaligned little-endian loads built from aligned words ---------------------------
// The aligned words touched always overlap the requested byte range, so they stay inside
// the 4-byte-granular allocation that holds the range.
__device__ __forceinline__ uint32_t ld_u32(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    const unsigned sh = (unsigned)(a & 3) * 8;
    uint32_t lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], sh);
}
__device__ __forceinline__ uint64_t ld_u64(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    const unsigned sh = (unsigned)(a & 3) * 8;
    uint32_t w0 = w[0], w1 = w[1];
    if (sh == 0) return (uint64_t)w0 | ((uint64_t)w1 << 32);
    uint32_t w2 = w[2];
    return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
}
__device__ __forceinline__ uint32_t ld_u16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
__device__ __forceinline__ uint32_t ld_u32_bytes(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ uint64_t ld_u64_bytes(const uint8_t* p) {
    return (uint64_t)ld_u32_bytes(p) | ((uint64_t)ld_u32_bytes(p + 4) << 32);
}
__device__ __forceinline__ void st_u32_bytes(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

// ---- warp-cooperative copies ------------------------------------------------------------
// Non-overlapping copy of len bytes, any alignment.  Long runs whose source and destination
// are congruent mod 16 use 16-byte vectors; everything else goes byte by byte.
__device__ __forceinline__ void warp_copy(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                          uint32_t len, int lane) {
    if (len >= 256 && ((reinterpret_cast<uintptr_t>(dst) ^ reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
        uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
        if (lane < (int)head) dst[lane] = src[lane];
        uint32_t body = (len - head) >> 4;
        const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
        uint4* d4 = reinterpret_cast<uint4*>(dst + head);
        for (uint32_t i = lane; i < body; i += 32) d4[i] = s4[i];
        uint32_t done = head + (body << 4);
        if (done + lane < len) dst[done + lane] = src[done + lane];
        return;
    }
    if (len >= 256 && ((reinterpret_cast<uintptr_t>(dst) ^ reinterpret_cast<uintptr_t>(src)) & 3) == 0) {
        uint32_t head = (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);
        if (lane < (int)head) dst[lane] = src[lane];
        uint32_t body = (len - head) >> 2;
        const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src + head);
        uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + head);
        for (uint32_t i = lane; i < body; i += 32) d4[i] = s4[i];
        uint32_t done = head + (body << 2);
        if (done + lane < len) dst[done + lane] = src[done + lane];
        return;
    }
    for (uint32_t i = lane; i < len; i += 32) dst[i] = src[i];
}

// LZ77 back-reference copy inside one output buffer: out[op .. op+len) = out[op-off ...].
// Caller guarantees 1 <= off <= op and that earlier stores of this warp are ordered
// (__syncwarp) before the call.
__device__ __forceinline__ void warp_match_copy(uint8_t* out, uint64_t op, uint32_t off, uint32_t len, int lane) {
    if (off >= 32) {
        // each 32-byte round only reads bytes written before it (distance >= 32); when the
        // match overlaps its own output the rounds are ordered with __syncwarp
        const bool self_dep = off < len;
        for (uint32_t base = 0; base < len; base += 32) {
            uint32_t k = base + lane;
            uint8_t v = 0;
            if (k < len) v = out[op + k - off];
            if (k < len) out[op + k] = v;
            if (self_dep) __syncwarp();
        }
    } else {
        // periodic pattern: every byte is a copy of one of the `off` bytes before op
        const uint8_t* pat = out + op - off;
        for (uint32_t k = lane; k < len; k += 32) out[op + k] = pat[k % off];
    }
}

// inclusive warp prefix sum
__device__ __forceinline__ uint32_t warp_incl_sum(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// ---- input gate: encoders that start while their input is still arriving over PCIe ---------
// The host enqueues the H2D copies of a host-resident input in pieces on a copy stream and, after
// each piece, a 4-byte copy that raises a watermark word in HBM.  An encoder warp calls wait(need)
// before it touches input bytes below `need`; with a null flag (device-resident input) the gate
// is free.  The watermark counts bytes available in EVERY partition (LZ4: the input is sent in
// stripes across all partitions, because each partition is one serial chain that has to start
// early) or bytes available from the start of the buffer (Snappy: fragments are taken in order).
// A watermark that never arrives (a copy failed) ends the wait after ~2 s and flags the call.
struct InGate {
    const uint32_t* flag;
    int* error;
    uint32_t have;
    __device__ __forceinline__ InGate(const uint32_t* f, int* e) : flag(f), error(e), have(f ? 0u : 0xffffffffu) {}
    __device__ __forceinline__ void wait(uint32_t need) {
        if (need <= have) return;                        // `have` starts at 0xffffffff when there is no flag
        uint32_t spins = 0;
        for (;;) {
            have = *reinterpret_cast<const volatile uint32_t*>(flag);
            if (have >= need) break;
            if (*reinterpret_cast<volatile int*>(error) != 0 || ++spins > (1u << 23)) {
                atomicExch(error, 1);
                have = 0xffffffffu;                      // give up: the call fails, nothing may hang
                break;
            }
            __nanosleep(256);
        }
        __syncwarp();
    }
};

// ---- partition arithmetic (threads/threads.c:55-97 of the reference) ----------------------
__host__ __device__ inline uint32_t partition_count(uint64_t n, uint32_t window) {
    const uint64_t chunk = (uint64_t)window * kWindowFactor;
    if (n < chunk) return 1;
    uint64_t parts = n / chunk;
    if (n % chunk >= (chunk >> 1)) parts++;
    return (uint32_t)parts;
}

// Per-call result block (device copy lives in the context workspace, host copy is pinned).
struct CallResult {
    long long value;       // bytes produced, or negative error
    int error;             // first failing partition + 1, 0 if none
    int parts;             // partitions found in the stream (decompress)
    unsigned int next;     // work counter for persistent kernels
    unsigned int pad;
};

// Normalised partition descriptor produced by the frame-parse kernels.
struct PartDesc {
    uint32_t in_off;       // offset of the partition payload from the stream start
    uint32_t in_len;       // compressed bytes
    uint64_t out_off;      // where its output starts (exclusive scan of decomp_len)
    uint32_t out_len;      // bytes it must produce (capacity for frame-less LZ4)
    uint32_t flags;        // bit0: last partition / frame-less rules, bit1: exact length required
};
constexpr uint32_t kPartLast = 1u;
constexpr uint32_t kPartExact = 2u;

}  // namespace llc
