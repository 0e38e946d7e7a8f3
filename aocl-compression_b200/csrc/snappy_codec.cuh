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


}  // namespace llc
