// in_ring.cuh -- per-warp shared-memory window over a compressed stream, filled by the TMA
// bulk-copy engine (cp.async.bulk.shared.global, SASS UBLKCP) and tracked with mbarriers.
//
// The decoders read tokens, length bytes, offsets and short literal runs from this window
// (29-cycle LDS) instead of issuing one dependent global load per field.  The window is a ring
// of kStages chunks of kChunk bytes; the byte at stream position p (relative to the 16-byte
// aligned base) lives at sm[p % (kStages*kChunk)].  Chunks are prefetched kStages ahead of the
// parse position; long literal runs bypass the ring (global->global copy) and the ring simply
// restarts behind them.
#pragma once
#include "llc_common.cuh"

namespace llc {

constexpr uint32_t kChunk = 512;
constexpr uint32_t kChunkLog = 9;
constexpr uint32_t kStages = 4;
constexpr uint32_t kRingBytes = kChunk * kStages;       // 2 KiB per warp
constexpr uint32_t kRingMask = kRingBytes - 1;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// explicit shared-space byte load (keeps ring reads on the LDS path with 32-bit addresses)
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

struct RingStorage {
    alignas(kRingBytes) uint8_t data[kRingBytes];       // ring-size aligned: address = base | (pos & mask)
    alignas(8) uint64_t bar[kStages];
};

struct Ring {
    uint8_t* sm;
    uint64_t* bar;
    const uint8_t* gbase;     // 16-byte aligned, <= first stream byte
    uint32_t total;           // bytes from gbase to the end of the stream
    uint32_t nchunks;
    uint32_t w0, w1, wr;      // live chunks [w0, w1), of which [w0, wr) have landed
    uint32_t par;             // per-slot parity of the next phase to wait for (persists across streams)

    // once per warp, before the first stream
    __device__ __forceinline__ void init(RingStorage* st, int lane) { init(st->data, st->bar, lane); }
    __device__ __forceinline__ void init(uint8_t* data, uint64_t* bars, int lane) {
        sm = data; bar = bars; par = 0;
        if (lane == 0) {
            for (uint32_t s = 0; s < kStages; s++) mbar_init(&bar[s], 1);
            fence_mbar_init();
        }
        __syncwarp();
    }
    // bind to a new stream; returns the position (relative to gbase) of its first byte
    __device__ __forceinline__ uint32_t open(const uint8_t* in, uint32_t clen) {
        const uint32_t pad = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 15);
        gbase = in - pad;
        total = clen + pad;
        nchunks = (total + kChunk - 1) >> kChunkLog;
        w0 = w1 = wr = 0;
        return pad;
    }
    __device__ __forceinline__ void issue(int lane) {
        const uint32_t c = w1, slot = c & (kStages - 1);
        const uint32_t left = total - (c << kChunkLog);
        const uint32_t bytes = left >= kChunk ? kChunk : ((left + 15u) & ~15u);
        if (lane == 0) {
            mbar_expect_tx(&bar[slot], bytes);
            bulk_g2s(sm + slot * kChunk, gbase + ((size_t)c << kChunkLog), bytes, &bar[slot]);
        }
        w1++;
    }
    __device__ __forceinline__ void wait_next() {
        const uint32_t slot = wr & (kStages - 1);
        const uint32_t parity = (par >> slot) & 1u;
        while (!mbar_try_wait(&bar[slot], parity)) {}
        par ^= 1u << slot;
        wr++;
    }
    // Slide the window so that it starts at the chunk holding `pos`, and keep it kStages deep.
    __device__ __forceinline__ void advance(uint32_t pos, int lane) {
        const uint32_t c = pos >> kChunkLog;
        if (c >= w1) {                        // jumped past everything in flight: drain, restart at c
            while (wr < w1) wait_next();
            w0 = w1 = wr = c;
        } else if (c > w0) {
            w0 = c;
            if (wr < w0) { while (wr < w0) wait_next(); }
        }
        const uint32_t want = min(nchunks, w0 + kStages);
        if (w1 < want) {
            __syncwarp();                     // every lane is done reading the slots being recycled
            fence_proxy_async();
            while (w1 < want) issue(lane);
        }
    }
    // Bytes [.., end) must have landed (end is clipped to the stream).  Caller keeps end within the window.
    __device__ __forceinline__ void ensure(uint32_t end) {
        const uint32_t need = min(nchunks, (end + kChunk - 1) >> kChunkLog);
        while (wr < need && wr < w1) wait_next();
    }
    // finish with the current stream: nothing may stay in flight into the next one
    __device__ __forceinline__ void close() {
        while (wr < w1) wait_next();
        __syncwarp();
    }
    __device__ __forceinline__ uint32_t byte(uint32_t pos) const { return sm[pos & kRingMask]; }
};

}  // namespace llc
