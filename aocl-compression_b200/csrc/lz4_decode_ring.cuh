// lz4_decode_ring.cuh -- pipelined warp-per-partition LZ4 decoder.
//
// Same accept/reject behaviour as lz4_decode_warp (lz4_codec.cuh; reference decoder
// algos/lz4/lz4.c:3806-4305), organised so that neither global-memory latency nor instruction
// count sits on the per-sequence dependency chain:
//   * the compressed stream is read through a TMA-filled shared-memory ring (in_ring.cuh);
//   * FAST PATH (a sequence whose token, optional single length bytes, <= 27 literals and offset fit
//     one 32-byte window, match <= 32 bytes, no self-overlap, far from both buffer ends): three
//     shared loads give the window byte per lane, the token and the first length byte; the next
//     sequence's loads are issued as soon as the token tells where it starts (software prefetch),
//     so the serial chain per sequence is LDS -> a few ALU ops -> LDS; literals go window -> HBM;
//     the match is a single byte-per-lane load that is only stored four sequences later
//     (register pipeline, hazard-checked against pending destinations), so up to four L2 round
//     trips overlap with parsing;
//   * everything else (length runs, long literals / matches, overlapping matches, the last bytes
//     of a partition, malformed input) takes lz4_slow_sequence(): flush the pipeline and run the
//     fully checked general code once.
#pragma once
#include "in_ring.cuh"

namespace llc {

// Reads a 255-terminated length extension starting at ring position p (lz4.c:3330-3352).
__device__ __forceinline__ uint32_t lz4_ring_ext(Ring& r, uint32_t& p, uint32_t iend, bool& bad, int lane) {
    uint32_t add = 0;
    for (;;) {
        if ((p >> kChunkLog) != r.w0) r.advance(p, lane);
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
    }
}

// One fully checked sequence without pipelining.  status 0: continue, 1: stream finished, -1: corrupt.
// State goes in and out by value so that the caller's hot-loop variables stay in registers.
struct Lz4Step { int status; uint32_t ip, op; };
__device__ __noinline__ Lz4Step lz4_slow_sequence(Ring* rp, uint32_t ip, uint32_t op, uint8_t* out, uint32_t iend,
                                                  uint32_t cap, bool last, int lane) {
    Ring& r = *rp;
    if (ip >= iend) return Lz4Step{-1, ip, op};
    if ((ip >> kChunkLog) != r.w0 || r.w1 == r.w0) r.advance(ip, lane);
    r.ensure(ip + 40);
    const uint32_t tok = r.byte(ip);
    uint32_t ll = tok >> 4;
    uint32_t p = ip + 1;
    bool bad = false;
    if (ll == 15) { ll += lz4_ring_ext(r, p, iend, bad, lane); if (bad) return Lz4Step{-1, ip, op}; }
    if (ll > iend - p || ll > cap - op) return Lz4Step{-1, ip, op};
    // end-of-block parsing restrictions, lz4.c:4104-4164
    const bool closing = ((uint64_t)op + ll + 12 > cap) || ((uint64_t)p + ll + 8 > iend);
    if (closing && last && p + ll != iend) return Lz4Step{-1, ip, op};
    if (ll <= 512) {
        if ((p >> kChunkLog) != r.w0) r.advance(p, lane);
        r.ensure(p + ll + 8);
        for (uint32_t k = lane; k < ll; k += 32) out[op + k] = (uint8_t)r.byte(p + k);
    } else {
        warp_copy(out + op, r.gbase + p, ll, lane);
    }
    op += ll;
    const uint32_t q = p + ll;
    if (closing && (last || op == cap)) return Lz4Step{1, q, op};
    if (q == iend) return Lz4Step{1, q, op};
    if (q + 2 > iend) return Lz4Step{-1, q, op};
    if ((q >> kChunkLog) != r.w0) r.advance(q, lane);
    r.ensure(q + 40);
    const uint32_t off = r.byte(q) | (r.byte(q + 1) << 8);
    ip = q + 2;
    uint32_t ml = tok & 15;
    if (ml == 15) { ml += lz4_ring_ext(r, ip, iend, bad, lane); if (bad) return Lz4Step{-1, ip, op}; }
    ml += 4;
    if (off == 0 || off > op) return Lz4Step{-1, ip, op};              // lz4.c:4196-4197
    if (ml > cap - op) return Lz4Step{-1, ip, op};
    if (last && (uint64_t)op + ml + 5 > cap) return Lz4Step{-1, ip, op};   // lz4.c:4262-4264
    __syncwarp();
    warp_match_copy(out, op, off, ml, lane);
    __syncwarp();
    op += ml;
    if (!last && (op == cap || ip >= iend)) return Lz4Step{1, ip, op};  // lz4.c:4285-4288
    return Lz4Step{0, ip, op};
}

// Keeps the ring window starting at the chunk of `low` (the sequence being executed; the ring never
// moves backwards) and waits until a 40-byte window at `need` has landed.  Returns the two
// thresholds the fast loop tests: a current position >= *chunk_end means chunks can be recycled,
// a prefetch position > *safe_end means data may still be in flight.
__device__ __noinline__ uint64_t lz4_ring_maintain(Ring* rp, uint32_t low, uint32_t need, int lane) {
    Ring& r = *rp;
    if ((low >> kChunkLog) != r.w0 || r.w1 == r.w0) r.advance(low, lane);
    r.ensure(need + 40);
    const uint32_t chunk_end = (r.w0 + 1) << kChunkLog;
    const uint32_t safe_end = (r.wr >= r.nchunks) ? 0xffffffffu : ((r.wr << kChunkLog) - 40u);
    return (uint64_t)chunk_end | ((uint64_t)safe_end << 32);
}
#define LLC_MAINTAIN(low, need)                                                              \
    do { const uint64_t th_ = lz4_ring_maintain(&r, (low), (need), lane);                    \
         chunk_end = (uint32_t)th_; safe_end = (uint32_t)(th_ >> 32); } while (0)

#define LLC_PEND_STORE(S) { if ((uint32_t)lane < pl##S) out[pd##S + lane] = (uint8_t)pv##S; }
#define LLC_FLUSH()                                                                          \
    do { LLC_PEND_STORE(0) LLC_PEND_STORE(1) LLC_PEND_STORE(2) LLC_PEND_STORE(3)             \
         pl0 = pl1 = pl2 = pl3 = 0; pd0 = pd1 = pd2 = pd3 = 0xffffffffu; pend_lo = 0xffffffffu; } while (0)
#define LLC_LOAD_WINDOW(pos)                                                                 \
    do { w = lds_u8(sbase + (((pos) + lane) & kRingMask)); tok = lds_u8(sbase + ((pos) & kRingMask));       \
         e1 = lds_u8(sbase + (((pos) + 1) & kRingMask)); } while (0)

#define LLC_LZ4_SEQ(S, A, B, C)                                                                              \
    {                                                                                                        \
        const uint32_t nibL = tok >> 4, nibM = tok & 15u;                                                    \
        const uint32_t extL = (nibL == 15u) ? 1u : 0u, extM = (nibM == 15u) ? 1u : 0u;                       \
        const uint32_t ll = nibL + (extL ? e1 : 0u);                                                         \
        const uint32_t hdr = 1u + extL;                                                                      \
        const uint32_t qpos = hdr + ll;                     /* window index of the offset */                 \
        const uint32_t ipn = ip + qpos + 2u + extM;                                                          \
        const uint32_t w_cur = w;                                                                            \
        /* prefetch the next sequence's window while this one is being executed */                          \
        if (ip >= chunk_end || ipn > safe_end) LLC_MAINTAIN(ip, ipn);                                        \
        LLC_LOAD_WINDOW(ipn);                                                                                \
        const uint32_t off = __shfl_sync(kFull, w_cur, (int)qpos) | (__shfl_sync(kFull, w_cur, (int)qpos + 1) << 8); \
        const uint32_t e2 = __shfl_sync(kFull, w_cur, (int)qpos + 2);                                        \
        const uint32_t ml = 4u + nibM + (extM ? e2 : 0u);                                                    \
        const uint32_t op2 = op + ll;                                                                        \
        const bool fast = (ll <= 27u) & (ml <= 32u) & (off >= ml) & (off <= op2) & (ip <= fast_i) & (op <= fast_o); \
        if (!fast) {                                                                                         \
            LLC_FLUSH();                                                                                     \
            const Lz4Step st = lz4_slow_sequence(&r, ip, op, out, iend, cap, last, lane);                    \
            ip = st.ip; op = st.op;                                                                          \
            if (st.status != 0) { status = st.status; break; }                                               \
            LLC_MAINTAIN(ip, ip);                                                                            \
            LLC_LOAD_WINDOW(ip);                                                                             \
            continue;                                                                                        \
        }                                                                                                    \
        if ((uint32_t)lane - hdr < ll) out[op + lane - hdr] = (uint8_t)w_cur;                                \
        const uint32_t src = op2 - off;                                                                      \
        if (src + ml > pend_lo) LLC_FLUSH();                                                                 \
        LLC_PEND_STORE(S)                                                                                    \
        __syncwarp();                       /* earlier stores of this warp are ordered before the load */   \
        pv##S = ((uint32_t)lane < ml) ? out[src + lane] : 0;                                                 \
        pd##S = op2; pl##S = ml;                                                                             \
        pend_lo = min(min(pd##A, pd##B), min(pd##C, op2));                                                   \
        op = op2 + ml;                                                                                       \
        ip = ipn;                                                                                            \
    }

// `out` points at the partition's first output byte; offsets inside the partition fit 32 bits.
__device__ inline int64_t lz4_decode_warp_ring(Ring& r, const uint8_t* __restrict__ in, uint32_t clen, uint8_t* out,
                                               uint32_t cap, bool last, int lane) {
    if (clen == 0) return kErrCorrupt;
    if (cap == 0) return (clen == 1 && in[0] == 0) ? 0 : kErrCorrupt;   // lz4.c:3854-3858
    uint32_t ip = r.open(in, clen);
    const uint32_t iend = r.total;
    const uint32_t sbase = smem_u32(r.sm);
    const uint32_t fast_i = iend >= 40u ? iend - 40u : 0u;
    const uint32_t fast_o = cap >= 72u ? cap - 72u : 0u;
    const bool any_fast = iend >= 40u + ip && cap >= 72u;
    uint32_t op = 0;
    uint32_t pend_lo = 0xffffffffu;
    uint32_t pv0 = 0, pv1 = 0, pv2 = 0, pv3 = 0;
    uint32_t pd0 = 0xffffffffu, pd1 = 0xffffffffu, pd2 = 0xffffffffu, pd3 = 0xffffffffu;
    uint32_t pl0 = 0, pl1 = 0, pl2 = 0, pl3 = 0;
    uint32_t chunk_end = 0, safe_end = 0;
    uint32_t w, tok, e1;
    int status = 0;
    LLC_MAINTAIN(ip, ip);
    LLC_LOAD_WINDOW(ip);
    if (!any_fast) {
        // tiny stream: every sequence through the checked path
        do {
            const Lz4Step st = lz4_slow_sequence(&r, ip, op, out, iend, cap, last, lane);
            ip = st.ip; op = st.op; status = st.status;
        } while (status == 0);
    } else {
        for (;;) {
            LLC_LZ4_SEQ(0, 1, 2, 3)
            LLC_LZ4_SEQ(1, 2, 3, 0)
            LLC_LZ4_SEQ(2, 3, 0, 1)
            LLC_LZ4_SEQ(3, 0, 1, 2)
        }
    }
    LLC_FLUSH();
    r.close();
    return status > 0 ? (int64_t)op : kErrCorrupt;
}

}  // namespace llc
