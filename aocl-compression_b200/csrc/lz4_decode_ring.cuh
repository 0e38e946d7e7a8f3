// lz4_decode_ring.cuh -- pipelined warp-per-partition LZ4 decoder.
//
// Same accept/reject behaviour as lz4_decode_warp (lz4_codec.cuh; reference decoder
// algos/lz4/lz4.c:3806-4305) but organised to keep global-memory latency off the
// per-sequence dependency chain:
//   * the compressed stream is read through a TMA-filled shared-memory ring (in_ring.cuh): one
//     lane-parallel 32-byte window load yields the token, the short literal run and the match
//     offset of a typical sequence;
//   * literals go window -> HBM directly;
//   * match copies of <= 32 bytes are software pipelined four deep: the back-reference load of
//     sequence k is issued into a register and only stored when sequence k+4 needs the slot, so
//     up to four L2 round trips overlap with the parsing of the following sequences.  A match whose
//     source overlaps a still-pending destination flushes the pipeline first.
#pragma once
#include "in_ring.cuh"

namespace llc {

// Reads a 255-terminated length extension starting at ring position p (lz4.c:3330-3352).
// Returns the added value and advances p; bad is set on truncated input.
__device__ __forceinline__ uint32_t lz4_ring_ext(Ring& r, uint32_t& p, uint32_t iend, bool& bad, int lane) {
    uint32_t add = 0;
    for (;;) {
        r.ensure(p + 32);
        const uint32_t q = p + lane;
        const uint32_t b = q < iend ? r.byte(q) : 0u;          // a 0 past the end stops the scan there
        const unsigned stop = __ballot_sync(kFull, b != 255u);
        const int first = stop ? (__ffs(stop) - 1) : 32;
        if (first < 32) {
            if (p + first >= iend) { bad = true; return add; }
            add += 255u * first + __shfl_sync(kFull, b, first);
            p += first + 1;
            return add;
        }
        add += 255u * 32u;
        p += 32;
        if (add > 0x7fff0000u) { bad = true; return add; }
        r.advance(p, lane);
    }
}

#define LLC_PEND_STORE(S)                                                         \
    if (pmask & (1u << S)) { if ((uint32_t)lane < pl##S) out[pd##S + lane] = (uint8_t)pv##S; }
#define LLC_FLUSH()                                                               \
    do { LLC_PEND_STORE(0) LLC_PEND_STORE(1) LLC_PEND_STORE(2) LLC_PEND_STORE(3)  \
         pmask = 0; pend_lo = 0xffffffffu; } while (0)

// One sequence; S is the pipeline slot this sequence's match uses.
#define LLC_LZ4_SEQ(S)                                                                                      \
    {                                                                                                       \
        if (ip >= iend) goto corrupt;                                                                       \
        if ((ip >> kChunkLog) != r.w0 || r.w1 == r.w0) r.advance(ip, lane);                                 \
        r.ensure(ip + 40);                                                                                  \
        const uint32_t w = r.byte(ip + lane);                                                               \
        const uint32_t tok = __shfl_sync(kFull, w, 0);                                                      \
        uint32_t ll = tok >> 4;                                                                             \
        uint32_t p = ip + 1;                                                                                \
        bool bad = false;                                                                                   \
        if (ll == 15) { ll += lz4_ring_ext(r, p, iend, bad, lane); if (bad) goto corrupt; }                 \
        if (ll > iend - p || ll > cap - op) goto corrupt;                                                   \
        const bool closing = ((uint64_t)op + ll + 12 > cap) || ((uint64_t)p + ll + 8 > iend);               \
        if (closing && last && p + ll != iend) goto corrupt;                                                \
        bool in_window = (p == ip + 1) && (ll <= 29);                                                       \
        if (in_window) {                                                                                    \
            if (lane >= 1 && (uint32_t)lane <= ll) out[op + lane - 1] = (uint8_t)w;                         \
        } else if (ll <= 512) {                                                                             \
            r.ensure(p + ll + 8);                                                                           \
            for (uint32_t k = lane; k < ll; k += 32) out[op + k] = (uint8_t)r.byte(p + k);                  \
        } else {                                                                                            \
            warp_copy(out + op, r.gbase + p, ll, lane);                                                     \
        }                                                                                                   \
        op += ll;                                                                                           \
        const uint32_t q = p + ll;          /* position of the match offset */                              \
        if (closing && (last || op == cap)) { ip = q; goto done; }                                          \
        if (q == iend) { ip = q; goto done; }                                                               \
        if (q + 2 > iend) goto corrupt;                                                                     \
        uint32_t off;                                                                                       \
        if (in_window) {                                                                                    \
            off = __shfl_sync(kFull, w, (int)(q - ip)) | (__shfl_sync(kFull, w, (int)(q - ip) + 1) << 8);   \
        } else {                                                                                            \
            if ((q >> kChunkLog) != r.w0) r.advance(q, lane);                                               \
            r.ensure(q + 40);                                                                               \
            off = r.byte(q) | (r.byte(q + 1) << 8);                                                         \
        }                                                                                                   \
        ip = q + 2;                                                                                         \
        uint32_t ml = tok & 15;                                                                             \
        if (ml == 15) { ml += lz4_ring_ext(r, ip, iend, bad, lane); if (bad) goto corrupt; }                \
        ml += 4;                                                                                            \
        if (off == 0 || off > op) goto corrupt;                                                             \
        if (ml > cap - op) goto corrupt;                                                                    \
        if (last && (uint64_t)op + ml + 5 > cap) goto corrupt;                                              \
        if (ml <= 32) {                                                                                     \
            const uint32_t src = op - off;                                                                  \
            if (src + min(ml, off) > pend_lo) LLC_FLUSH();                                                  \
            LLC_PEND_STORE(S)                                                                               \
            pmask &= ~(1u << S);                                                                            \
            __syncwarp();                   /* earlier stores of this warp are ordered before the load */  \
            uint32_t k = lane;              /* overlapping match: periodic pattern of period off */      \
            if (off < ml) k = lane - off * __float2uint_rz(__fdividef((float)lane + 0.5f, (float)off));     \
            pv##S = ((uint32_t)lane < ml) ? out[src + k] : 0;                                               \
            pd##S = op; pl##S = ml;                                                                         \
            if (pmask == 0) pend_lo = op;                                                                   \
            else pend_lo = (pmask >> ((S + 1) & 3) & 1u) ? LLC_PD((S + 1) & 3)                              \
                         : (pmask >> ((S + 2) & 3) & 1u) ? LLC_PD((S + 2) & 3) : LLC_PD((S + 3) & 3);       \
            pmask |= 1u << S;                                                                               \
        } else {                                                                                            \
            LLC_FLUSH();                                                                                    \
            __syncwarp();                                                                                   \
            warp_match_copy(out, op, off, ml, lane);                                                        \
            __syncwarp();                                                                                   \
        }                                                                                                   \
        op += ml;                                                                                           \
        if (!last && (op == cap || ip >= iend)) goto done;                                                  \
    }

#define LLC_PD(i) ((i) == 0 ? pd0 : (i) == 1 ? pd1 : (i) == 2 ? pd2 : pd3)

// `out` points at the partition's first output byte; offsets inside the partition fit 32 bits.
__device__ inline int64_t lz4_decode_warp_ring(Ring& r, const uint8_t* __restrict__ in, uint32_t clen, uint8_t* out,
                                               uint32_t cap, bool last, int lane) {
    if (clen == 0) return kErrCorrupt;
    if (cap == 0) return (clen == 1 && in[0] == 0) ? 0 : kErrCorrupt;   // lz4.c:3854-3858
    uint32_t ip = r.open(in, clen);
    const uint32_t iend = r.total;
    uint32_t op = 0;
    uint32_t pmask = 0, pend_lo = 0xffffffffu;
    uint32_t pv0 = 0, pv1 = 0, pv2 = 0, pv3 = 0;
    uint32_t pd0 = 0, pd1 = 0, pd2 = 0, pd3 = 0;
    uint32_t pl0 = 0, pl1 = 0, pl2 = 0, pl3 = 0;
    for (;;) {
        LLC_LZ4_SEQ(0)
        LLC_LZ4_SEQ(1)
        LLC_LZ4_SEQ(2)
        LLC_LZ4_SEQ(3)
    }
done:
    LLC_FLUSH();
    r.close();
    return (int64_t)op;
corrupt:
    r.close();
    return kErrCorrupt;
}

}  // namespace llc
