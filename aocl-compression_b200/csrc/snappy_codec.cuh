// snappy_codec.cuh -- warp-per-unit Snappy tag-stream decoder and exact fragment encoder.
//
// Decoder: SnappyDecompressor::DecompressAllTags + SnappyArrayWriter of the reference
// (algos/snappy/snappy.cc:1466-1570, 2131-2209).  Encoder: AOCL_CompressFragment
// (algos/snappy/snappy.cc:846-1046) with the emitters at :436-568; byte-exact with it.
#pragma once
#include "llc_common.cuh"

namespace llc {

// Decode the tag stream [in, in+clen) into exactly `expect` bytes at out.
// Returns expect or kErrCorrupt.  Copies may not reach before out[0] (snappy.cc:2190-2191).
__device__ inline int64_t snappy_decode_warp(const uint8_t* __restrict__ in, uint32_t clen, uint8_t* out,
                                             uint64_t expect, int lane) {
    uint32_t ip = 0;
    uint64_t op = 0;
    while (ip < clen) {
        const uint32_t tag = in[ip++];
        uint32_t len, off;
        if ((tag & 3) == 0) {                               // literal, snappy.cc:1492-1527
            len = (tag >> 2) + 1;
            if (len > 60) {
                const uint32_t nb = len - 60;
                if (ip + nb > clen) return kErrCorrupt;
                uint32_t v = 0;
                for (uint32_t k = 0; k < nb; k++) v |= (uint32_t)in[ip + k] << (8 * k);
                ip += nb;
                if (v == 0xffffffffu) return kErrCorrupt;
                len = v + 1;
            }
            if (len > clen - ip || (uint64_t)len > expect - op) return kErrCorrupt;
            warp_copy(out + op, in + ip, len, lane);
            ip += len; op += len;
            continue;
        }
        const uint32_t kind = tag & 3;                      // char_table, snappy-internal.h:406-439
        if (kind == 1) {
            if (ip + 1 > clen) return kErrCorrupt;
            len = 4 + ((tag >> 2) & 7); off = ((tag >> 5) << 8) | in[ip]; ip += 1;
        } else if (kind == 2) {
            if (ip + 2 > clen) return kErrCorrupt;
            len = 1 + (tag >> 2); off = ld_u16(in + ip); ip += 2;
        } else {
            if (ip + 4 > clen) return kErrCorrupt;
            len = 1 + (tag >> 2); off = ld_u32_bytes(in + ip); ip += 4;
        }
        if (off == 0 || (uint64_t)off > op || (uint64_t)len > expect - op) return kErrCorrupt;   // snappy.cc:2185-2199
        __syncwarp();
        warp_match_copy(out, op, off, len, lane);
        __syncwarp();
        op += len;
    }
    return op == expect ? (int64_t)op : kErrCorrupt;        // snappy.cc:1715
}

// varint32 (snappy-stubs-internal.h:440-470); returns bytes written / consumed (0 = malformed)
__device__ __host__ inline uint32_t put_varint32(uint8_t* p, uint32_t v) {
    uint32_t k = 0;
    while (v >= 128) { p[k++] = (uint8_t)(v | 128); v >>= 7; }
    p[k++] = (uint8_t)v;
    return k;
}
__device__ __host__ inline uint32_t varint32_len(uint32_t v) {
    uint32_t k = 1;
    while (v >= 128) { v >>= 7; k++; }
    return k;
}
__device__ __host__ inline uint32_t get_varint32(const uint8_t* p, uint64_t n, uint32_t* out) {
    uint32_t v = 0;
    for (uint32_t i = 0; i < 5; i++) {
        if (i >= n) return 0;
        const uint32_t b = p[i];
        if (i == 4 && b > 15) return 0;
        v |= (b & 127u) << (7 * i);
        if (b < 128) { *out = v; return i + 1; }
    }
    return 0;
}

// ---- emitters (warp-uniform arguments; returns the new output offset) -----------------------
__device__ __forceinline__ uint32_t snappy_put_literal(uint8_t* dst, uint32_t op, const uint8_t* lit, uint32_t len, int lane) {
    const uint32_t nm1 = len - 1;                           // snappy.cc:436-476
    uint32_t hdr;
    if (nm1 < 60) {
        if (lane == 0) dst[op] = (uint8_t)(nm1 << 2);
        hdr = 1;
    } else {
        const uint32_t count = nm1 < (1u << 8) ? 1 : nm1 < (1u << 16) ? 2 : nm1 < (1u << 24) ? 3 : 4;
        if (lane == 0) dst[op] = (uint8_t)((59 + count) << 2);
        if (lane < (int)count) dst[op + 1 + lane] = (uint8_t)(nm1 >> (8 * lane));
        hdr = 1 + count;
    }
    warp_copy(dst + op + hdr, lit, len, lane);
    return op + hdr + len;
}
__device__ __forceinline__ uint32_t snappy_put_copy(uint8_t* dst, uint32_t op, uint32_t off, uint32_t len, int lane) {
    // snappy.cc:540-568: 64-byte copies while len >= 68, one 60-byte copy if 64 < len < 68, then the rest
    uint32_t n64 = 0;
    if (len >= 68) n64 = (len - 68) / 64 + 1;
    uint32_t rest = len - 64 * n64;
    const uint32_t n60 = rest > 64 ? 1 : 0;
    rest -= 60 * n60;
    const uint32_t pre = n64 + n60;                         // 3-byte COPY_2 elements in front
    for (uint32_t j = lane; j < pre; j += 32) {
        const uint32_t l = j < n64 ? 64 : 60;
        uint8_t* q = dst + op + 3 * j;
        q[0] = (uint8_t)(2 | ((l - 1) << 2)); q[1] = (uint8_t)off; q[2] = (uint8_t)(off >> 8);
    }
    op += 3 * pre;
    if (rest < 12 && off < 2048) {                          // snappy.cc:479-505
        if (lane == 0) { dst[op] = (uint8_t)(1 | ((rest - 4) << 2) | ((off >> 8) << 5)); dst[op + 1] = (uint8_t)off; }
        return op + 2;
    }
    if (lane == 0) { dst[op] = (uint8_t)(2 | ((rest - 1) << 2)); dst[op + 1] = (uint8_t)off; dst[op + 2] = (uint8_t)(off >> 8); }
    return op + 3;
}

// Encode one fragment (n <= 65536) with a u16 hash table of `tsize` entries in shared memory.
// Same speculative 32-probe search as the LZ4 encoder; the probe schedule is
// stride = skip >> 5, skip += stride, starting at skip = 32 (snappy.cc:903-974; the reference's
// unrolled 16-probe prologue is the same walk).
__device__ inline uint32_t snappy_encode_fragment_warp(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst,
                                                       uint16_t* tab, int lane) {
    uint32_t tsize = 256;                                   // snappy.cc:619-632
    if (n > 16384) tsize = 16384; else while (tsize < n) tsize <<= 1;
    const int shift = 32 - (31 - __clz(tsize));
    for (uint32_t i = lane; i < tsize / 2; i += 32) reinterpret_cast<uint32_t*>(tab)[i] = 0;
    __syncwarp();
    uint32_t op = 0, ip = 0;
    if (n >= 15) {
        const uint32_t ip_limit = n - 15;
        bool done = false;
        while (!done) {
            const uint32_t next_emit = ip++;
            uint32_t skip = 32, cand = 0;
            bool found = false;
            for (;;) {
                // stride of the j-th upcoming probe: the skip counter grows by its own stride, so the
                // per-lane value is produced by a short serial recurrence on lane 0's state
                uint32_t my_stride, s = skip;
                if (skip <= 32) {
                    s = skip + lane;                        // first round of a search: stride 1 throughout
                    my_stride = 1;
                } else {
                    // lanes need skip_k = skip after k probes; each lane replays the recurrence
                    // (at most 31 cheap steps; only reached after 32 fruitless probes)
                    for (int k = 0; k < lane; k++) s += s >> 5;
                    my_stride = s >> 5;
                }
                const uint32_t incl = warp_incl_sum(my_stride, lane);
                const uint32_t cur = ip + incl - my_stride;
                const uint32_t nxt = ip + incl;
                const bool valid = nxt <= ip_limit;
                uint32_t h = 0x80000000u | lane, seq4 = 0, c = 0;
                if (valid) {
                    seq4 = ld_u32(src + cur);
                    h = (seq4 * 0x1e35a7bdU) >> shift;      // snappy.cc:152-158
                    c = tab[h];
                }
                const unsigned peers = __match_any_sync(kFull, h);
                const unsigned before = peers & ((1u << lane) - 1u);
                const int from = before ? (31 - __clz(before)) : lane;
                const uint32_t peer_pos = __shfl_sync(kFull, cur, from);
                if (before) c = peer_pos;
                bool hit = false;
                if (valid) hit = ld_u32(src + c) == seq4;
                const unsigned hits = __ballot_sync(kFull, hit);
                const unsigned events = hits | __ballot_sync(kFull, !valid);
                const int win = events ? (__ffs(events) - 1) : 32;
                const bool win_is_match = (win < 32) && ((hits >> win) & 1u);
                const unsigned commit = (win_is_match ? (win == 31 ? kFull : ((2u << win) - 1u))
                                                      : (win == 0 ? 0u : (win >= 32 ? kFull : ((1u << win) - 1u))));
                if (valid && ((commit >> lane) & 1u)) {
                    const unsigned mine = peers & commit;
                    if ((31 - __clz(mine)) == lane) tab[h] = (uint16_t)cur;
                }
                __syncwarp();
                if (win < 32) {
                    if (win_is_match) { ip = __shfl_sync(kFull, cur, win); cand = __shfl_sync(kFull, c, win); found = true; }
                    break;
                }
                ip = __shfl_sync(kFull, nxt, 31);
                skip = __shfl_sync(kFull, s + (s >> 5), 31);
            }
            if (!found) { ip = next_emit; break; }          // -> emit_remainder
            op = snappy_put_literal(dst, op, src + next_emit, ip - next_emit, lane);   // snappy.cc:980
            for (;;) {                                      // snappy.cc:995-1032
                // match length: 4 + common prefix, bounded by the fragment end
                uint32_t mc = 0;
                {
                    const uint32_t delta = ip - cand;
                    uint32_t base = ip + 4;
                    for (;;) {
                        const uint32_t pa = base + 4u * lane;
                        uint32_t cc = 0;
                        if (pa < n) {
                            const uint32_t avail = min(4u, n - pa);
                            uint32_t x;
                            if (avail == 4) x = ld_u32(src + pa) ^ ld_u32(src + pa - delta);
                            else {
                                x = 0;
                                for (uint32_t b = 0; b < avail; b++) x |= (uint32_t)(src[pa + b] ^ src[pa - delta + b]) << (8 * b);
                            }
                            cc = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
                            cc = min(cc, avail);
                        }
                        const unsigned partial = __ballot_sync(kFull, cc < 4);
                        if (partial) {
                            const int first = __ffs(partial) - 1;
                            mc += 4u * first + __shfl_sync(kFull, cc, first);
                            break;
                        }
                        mc += 128; base += 128;
                    }
                }
                const uint32_t len = 4 + mc;
                op = snappy_put_copy(dst, op, ip - cand, len, lane);
                ip += len;
                if (ip >= ip_limit) { done = true; break; }
                const uint32_t prev4 = ld_u32(src + ip - 1), cur4 = ld_u32(src + ip);
                tab[(prev4 * 0x1e35a7bdU) >> shift] = (uint16_t)(ip - 1);   // snappy.cc:1016-1021
                const uint32_t h = (cur4 * 0x1e35a7bdU) >> shift;
                cand = tab[h];
                tab[h] = (uint16_t)ip;
                if (ld_u32(src + cand) != cur4) break;
            }
            __syncwarp();
        }
    }
    if (ip < n) op = snappy_put_literal(dst, op, src + ip, n - ip, lane);   // snappy.cc:1039-1043
    return op;
}

}  // namespace llc
