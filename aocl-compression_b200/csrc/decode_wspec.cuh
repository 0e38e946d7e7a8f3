// decode_wspec.cuh -- warp-specialised LZ4 / Snappy decoder: one CTA of two warps per unit
// (RAP partition or page).
//
//   warp 0  PARSER   walks the token chain of the compressed stream, which is inherently serial
//                    (token -> lengths -> next token).  It reads the stream through the TMA-filled
//                    shared-memory ring (in_ring.cuh), performs EVERY validity check of the reference
//                    decoder (input/output bounds, offset range, end-of-block rules; lz4.c:3806-4305,
//                    snappy.cc:1466-1570, 2185-2199) and emits one 16-byte record per sequence
//                    {literal position, literal length, match offset, match length} into a
//                    shared-memory queue.  No data bytes are touched, so the serial chain is
//                    LDS -> three ALU ops -> LDS.
//   warp 1  COPIER   consumes 32 records at a time, ONE LANE PER SEQUENCE: a warp scan turns the
//                    lengths into output offsets, every lane copies its own literal run
//                    (compressed stream -> output), then the matches are executed in dependency
//                    rounds: a match is ready when its source ends before the destination of the
//                    first unfinished match of the group.  Long runs (> 32 literal / > 64 match
//                    bytes) are copied cooperatively by the whole warp.
//
// The two warps overlap: while the copier moves the bytes of group g the parser is already
// validating group g+1..g+7.  Queue hand-off uses two single-writer counters in shared memory.
#pragma once
#include "in_ring.cuh"
#include "snappy_codec.cuh"

namespace llc {

constexpr uint32_t kQCap = 128;                 // records in flight per CTA (2 KiB)
constexpr uint32_t kWin = 320;                  // bytes the fast parse loops may look ahead without re-checking the ring
constexpr uint32_t kQMask = kQCap - 1;

// The ring window must sit on a 2 KiB boundary of the *shared address space* so that ring addresses
// are base | (pos & mask) (one LOP3 on the parse chain).  The static shared segment does not start
// on such a boundary, so the CTA reserves 2 * kRingBytes and ring() picks the aligned half.
struct WsShared {
    uint8_t ring_area[2 * kRingBytes];
    uint4 q[kQCap];
    uint64_t ring_bar[kStages];
    volatile uint32_t tail;     // records published by the parser
    volatile uint32_t head;     // records retired by the copier
    volatile uint32_t done;     // parser finished (all records published)
    uint32_t unit;              // ticket broadcast
};
#define LLC_WS_SHARED(name)                                                         \
    __shared__ __align__(16) uint8_t name##_raw[sizeof(WsShared)];                  \
    WsShared& name = *reinterpret_cast<WsShared*>(name##_raw)
__device__ __forceinline__ uint8_t* ws_ring_data(WsShared& s) {
    const uint32_t a = smem_u32(s.ring_area);
    return s.ring_area + (((a + kRingBytes - 1u) & ~kRingMask) - a);
}

// ------------------------------------------------------------------------------------------ parser side
// The queue slot of record number t lives at qbase + (t % kQCap) * 16 (shared-space address).
// Every lane holds the same values, so all lanes issue the same store (one wavefront, no branch).
__device__ __forceinline__ void ws_store_record(uint32_t qbase, uint32_t t, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(qbase + ((t & kQMask) << 4)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Make records [.., tail) visible to the copier; block while the queue is nearly full.
__device__ __noinline__ void ws_publish(WsShared* s, uint32_t tail, int lane) {
    __syncwarp();
    if (lane == 0) { __threadfence_block(); s->tail = tail; }
    while (tail - s->head > kQCap - 32) __nanosleep(64);
}
__device__ __forceinline__ void ws_finish(WsShared* s, uint32_t tail, int lane) {
    __syncwarp();
    if (lane == 0) { __threadfence_block(); s->tail = tail; __threadfence_block(); s->done = 1; }
}

// Ring bookkeeping: keeps the window over `pos`, waits for a kWin-byte window there and returns the
// exclusive position limit below which the fast loops may read kWin bytes ahead without further
// checks (min of: end of the current first chunk -> recycle trigger, landed bytes - kWin).
__device__ __noinline__ uint32_t ws_ring_maintain(Ring* rp, uint32_t pos, int lane) {
    Ring& r = *rp;
    if ((pos >> kChunkLog) != r.w0 || r.w1 == r.w0) r.advance(pos, lane);
    r.ensure(pos + kWin);
    const uint32_t chunk_end = (r.w0 + 1) << kChunkLog;
    const uint32_t safe_ex = (r.wr >= r.nchunks) ? 0xffffffffu : ((r.wr << kChunkLog) - (kWin - 1u));
    return min(chunk_end, safe_ex);
}

// 255-terminated LZ4 length extension read through the ring (lz4.c:3330-3352).
// Returns (new position << 32) | added length; added length 0xffffffff means truncated input.
__device__ __noinline__ uint64_t ws_lz4_ext(Ring* rp, uint32_t p, uint32_t iend, int lane) {
    Ring& r = *rp;
    uint32_t add = 0;
    for (;;) {
        if ((p >> kChunkLog) != r.w0 || r.w1 == r.w0) r.advance(p, lane);
        r.ensure(p + 32);
        const uint32_t q = p + lane;
        const uint32_t b = q < iend ? r.byte(q) : 0u;
        const unsigned stop = __ballot_sync(kFull, b != 255u);
        const int first = stop ? (__ffs(stop) - 1) : 32;
        if (first < 32) {
            if (p + first >= iend) return ((uint64_t)p << 32) | 0xffffffffu;
            add += 255u * first + __shfl_sync(kFull, b, first);
            return ((uint64_t)(p + first + 1) << 32) | add;
        }
        add += 255u * 32u;
        p += 32;
        if (add > 0x7fff0000u) return ((uint64_t)p << 32) | 0xffffffffu;
    }
}

// Parses one LZ4 block; *tail_io counts emitted records.  Returns bytes the stream produces, or
// kErrCorrupt.  The fast loop is branch-lean: one combined branch to the general code, one every 32
// records to publish, one loop back edge; a bad offset neutralises its record (ml = 0) and is
// reported at the next publish point, so the copier never sees an out-of-range source.
__device__ inline int64_t ws_parse_lz4(Ring& r, WsShared* s, uint32_t* tail_io, const uint8_t* in, uint32_t clen,
                                       uint32_t cap, bool last, int lane) {
    if (clen == 0) return kErrCorrupt;
    if (cap == 0) return (clen == 1 && in[0] == 0) ? 0 : kErrCorrupt;   // lz4.c:3854-3858
    uint32_t ip = r.open(in, clen);
    const uint32_t pad = ip, iend = r.total;
    const uint32_t sbase = smem_u32(r.sm);                              // 2 KiB aligned: address = sbase | (pos & mask)
    const uint32_t qbase = smem_u32(s->q);
    // fast region: a whole sequence with at most one length byte each (ll <= 269, ml <= 273) stays
    // inside the look-ahead window and cannot reach either end-of-block rule
    const bool any_fast = iend >= kWin + pad && cap >= 560u;
    const uint32_t fast_i_ex = any_fast ? iend - (kWin - 1u) : 0u;      // exclusive bounds of the fast region
    const uint32_t fast_o_ex = any_fast ? cap - 559u : 0u;
    uint32_t op = 0, tail = 0, lim = 0, bad = 0;
    for (;;) {
        const uint32_t tok = lds_u8(sbase | (ip & kRingMask));
        const uint32_t e1 = lds_u8(sbase | ((ip + 1u) & kRingMask));
        const uint32_t nibL = tok >> 4, nibM = tok & 15u;
        const bool extL = nibL == 15u, extM = nibM == 15u;
        const uint32_t ll = nibL + (extL ? e1 : 0u);
        const uint32_t lp = ip + 1u + (extL ? 1u : 0u);                 // first literal byte
        const uint32_t q = lp + ll;
        const uint32_t o0 = lds_u8(sbase | (q & kRingMask)), o1 = lds_u8(sbase | ((q + 1u) & kRingMask));
        const uint32_t e2 = lds_u8(sbase | ((q + 2u) & kRingMask));
        if ((extL & (e1 == 255u)) | (extM & (e2 == 255u)) | (ip >= lim) | (op >= fast_o_ex)) {
            // ---- general code: window upkeep, length bytes, block tail, tiny streams
            if (bad) break;
            if (ip >= iend) { bad = 1; break; }
            const bool window_only = ip >= lim;             // values read above may predate the refill
            lim = min(ws_ring_maintain(&r, ip, lane), fast_i_ex);
            if (window_only & (ip < lim) & (op < fast_o_ex)) continue;   // retry on the fast path with a valid window
            const uint32_t tk = r.byte(ip);
            uint32_t ll = tk >> 4, p = ip + 1;
            if (ll == 15) {
                const uint64_t e = ws_lz4_ext(&r, p, iend, lane);
                if ((uint32_t)e == 0xffffffffu) { bad = 1; break; }
                ll += (uint32_t)e; p = (uint32_t)(e >> 32);
            }
            if (ll > iend - p || ll > cap - op) { bad = 1; break; }
            const bool closing = ((uint64_t)op + ll + 12 > cap) || ((uint64_t)p + ll + 8 > iend);   // lz4.c:4104-4164
            if (closing && last && p + ll != iend) { bad = 1; break; }
            const uint32_t lit_pos = p - pad;
            op += ll;
            const uint32_t qq = p + ll;
            if ((closing && (last || op == cap)) || qq == iend) { ws_store_record(qbase, tail, lit_pos, ll, 0, 0); tail++; break; }
            if (qq + 2 > iend) { bad = 1; break; }
            ws_ring_maintain(&r, qq, lane);
            const uint32_t off = r.byte(qq) | (r.byte(qq + 1) << 8);
            ip = qq + 2;
            uint32_t ml = tk & 15u;
            if (ml == 15) {
                const uint64_t e = ws_lz4_ext(&r, ip, iend, lane);
                if ((uint32_t)e == 0xffffffffu) { bad = 1; break; }
                ml += (uint32_t)e; ip = (uint32_t)(e >> 32);
            }
            ml += 4;
            if (off == 0 || off > op || ml > cap - op) { bad = 1; break; }          // lz4.c:4196-4197
            if (last && (uint64_t)op + ml + 5 > cap) { bad = 1; break; }            // lz4.c:4262-4264
            ws_store_record(qbase, tail, lit_pos, ll, off, ml); tail++;
            if ((tail & 31u) == 0) ws_publish(s, tail, lane);
            op += ml;
            if (!last && (op == cap || ip >= iend)) break;                          // lz4.c:4285-4288
            lim = 0;                                                                // re-validate the window next time
            continue;
        }
        // ---- fast path: at most one length byte each, far from both ends -> no end-of-block rule can fire
        const uint32_t off = o0 | (o1 << 8);
        const uint32_t ml = nibM + 4u + (extM ? e2 : 0u);
        const uint32_t op2 = op + ll;
        const bool ok = (off - 1u) < op2;                                           // 1 <= off <= op2 (lz4.c:4196-4197)
        bad |= ok ? 0u : 1u;
        ws_store_record(qbase, tail, lp - pad, ll, off, ok ? ml : 0u);
        tail++;
        op = op2 + ml;
        ip = q + 2u + (extM ? 1u : 0u);
        if ((tail & 31u) == 0) ws_publish(s, tail, lane);
    }
    *tail_io = tail;
    return bad ? kErrCorrupt : (int64_t)op;
}

// Parses one Snappy tag stream that must produce exactly `expect` bytes; one record per element.
__device__ inline int64_t ws_parse_snappy(Ring& r, WsShared* s, uint32_t* tail_io, const uint8_t* in, uint32_t clen,
                                          uint32_t expect, int lane) {
    uint32_t ip = r.open(in, clen);
    const uint32_t pad = ip, iend = r.total;
    const uint32_t sbase = smem_u32(r.sm);
    const uint32_t qbase = smem_u32(s->q);
    const uint32_t fast_i_ex = iend >= kWin + pad ? iend - (kWin - 1u) : 0u;
    uint32_t op = 0, tail = 0, lim = 0, bad = 0;
    while (ip < iend) {
        const uint32_t tag = lds_u8(sbase | (ip & kRingMask));
        const uint32_t b1 = lds_u8(sbase | ((ip + 1u) & kRingMask)), b2 = lds_u8(sbase | ((ip + 2u) & kRingMask));
        const uint32_t kind = tag & 3u, hi = tag >> 2;
        if ((ip >= lim) | (kind == 3u) | ((kind == 0u) & (hi >= 60u))) {
            // ---- general code: window upkeep, long literals, 4-byte offsets, stream tail
            if (bad) break;
            const bool window_only = ip >= lim;
            lim = min(ws_ring_maintain(&r, ip, lane), fast_i_ex);
            if (window_only & (ip < lim)) continue;
            const uint32_t tg = r.byte(ip);
            const uint32_t kd = tg & 3u;
            if (kd == 0) {                                  // literal, snappy.cc:1492-1527
                uint32_t len = (tg >> 2) + 1u, p = ip + 1u;
                if (len > 60u) {
                    const uint32_t nb = len - 60u;
                    if (p + nb > iend) { bad = 1; break; }
                    uint32_t v = 0;
                    for (uint32_t k = 0; k < nb; k++) v |= r.byte(p + k) << (8 * k);
                    if (v == 0xffffffffu) { bad = 1; break; }
                    len = v + 1u; p += nb;
                }
                if (len > iend - p || len > expect - op) { bad = 1; break; }
                ws_store_record(qbase, tail, p - pad, len, 0, 0); tail++;
                op += len; ip = p + len;
            } else {
                uint32_t len, off;                          // char_table, snappy-internal.h:406-439
                if (kd == 1) {
                    if (ip + 2 > iend) { bad = 1; break; }
                    len = 4u + ((tg >> 2) & 7u); off = ((tg >> 5) << 8) | r.byte(ip + 1); ip += 2;
                } else if (kd == 2) {
                    if (ip + 3 > iend) { bad = 1; break; }
                    len = 1u + (tg >> 2); off = r.byte(ip + 1) | (r.byte(ip + 2) << 8); ip += 3;
                } else {
                    if (ip + 5 > iend) { bad = 1; break; }
                    len = 1u + (tg >> 2);
                    off = r.byte(ip + 1) | (r.byte(ip + 2) << 8) | (r.byte(ip + 3) << 16) | (r.byte(ip + 4) << 24);
                    ip += 5;
                }
                if (off == 0 || off > op || len > expect - op) { bad = 1; break; }   // snappy.cc:2185-2199
                ws_store_record(qbase, tail, 0, 0, off, len); tail++;
                op += len;
            }
            if ((tail & 31u) == 0) ws_publish(s, tail, lane);
            lim = 0;
            continue;
        }
        // ---- fast path: short literal or 1/2-byte-offset copy, at least 40 bytes before the end
        const bool is_lit = kind == 0u;
        const uint32_t lit_len = hi + 1u;
        const uint32_t cp_len = (kind == 1u) ? 4u + (hi & 7u) : 1u + hi;
        const uint32_t cp_off = (kind == 1u) ? (((tag >> 5) << 8) | b1) : (b1 | (b2 << 8));
        const uint32_t len = is_lit ? lit_len : cp_len;
        const bool ok = (len <= expect - op) & (is_lit | ((cp_off - 1u) < op));      // snappy.cc:2185-2199
        bad |= ok ? 0u : 1u;
        ws_store_record(qbase, tail, ip + 1u - pad, (is_lit & ok) ? lit_len : 0u, cp_off, (!is_lit & ok) ? cp_len : 0u);
        tail++;
        op += ok ? len : 0u;
        ip += is_lit ? 1u + lit_len : (kind == 1u ? 2u : 3u);
        if ((tail & 31u) == 0) ws_publish(s, tail, lane);
    }
    *tail_io = tail;
    if (bad) return kErrCorrupt;
    return op == expect ? (int64_t)op : kErrCorrupt;        // snappy.cc:1715
}

// ------------------------------------------------------------------------------------------ copier side
__device__ inline void ws_copy_records(WsShared* s, const uint8_t* __restrict__ in, uint8_t* out, int lane) {
    uint32_t head = 0, op_base = 0;
    for (;;) {
        uint32_t avail, fin;
        for (;;) {
            fin = s->done;
            __threadfence_block();
            avail = s->tail - head;
            if (avail >= 32u || fin) break;
            __nanosleep(32);
        }
        const uint32_t n = min(32u, avail);
        if (n == 0) break;
        uint4 rec = make_uint4(0, 0, 0, 0);
        if ((uint32_t)lane < n) rec = s->q[(head + lane) & kQMask];
        const uint32_t lit_pos = rec.x, ll = rec.y, off = rec.z, ml = rec.w;
        const uint32_t len = ll + ml;
        const uint32_t incl = warp_incl_sum(len, lane);
        const uint32_t dstL = op_base + incl - len;
        const uint32_t dstM = dstL + ll;

        // ---- literals: lane-per-run for short runs (16 bytes in flight per lane, then stored),
        //      whole warp for long ones
        {
            const uint32_t ll_s = ll <= 32u ? ll : 0u;
            const uint32_t maxll = __reduce_max_sync(kFull, ll_s);
            const uint8_t* src = in + lit_pos;
            uint8_t* dst = out + dstL;
            for (uint32_t base = 0; base < maxll; base += 8) {
                uint32_t v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) if (base + j < ll_s) v[j] = src[base + j];
#pragma unroll
                for (int j = 0; j < 8; j++) if (base + j < ll_s) dst[base + j] = (uint8_t)v[j];
            }
            unsigned big = __ballot_sync(kFull, ll > 32u);
            while (big) {
                const int k = __ffs(big) - 1;
                big &= big - 1;
                warp_copy(out + __shfl_sync(kFull, dstL, k), in + __shfl_sync(kFull, lit_pos, k), __shfl_sync(kFull, ll, k), lane);
            }
        }
        __syncwarp();

        // ---- matches in dependency rounds
        unsigned pending = __ballot_sync(kFull, ml != 0u);
        while (pending) {
            const int first = __ffs(pending) - 1;
            const uint32_t frontier = __shfl_sync(kFull, dstM, first);
            const uint32_t f_ml = __shfl_sync(kFull, ml, first);
            if (f_ml > 64u) {
                warp_match_copy(out, frontier, __shfl_sync(kFull, off, first), f_ml, lane);
                __syncwarp();
                pending &= ~(1u << first);
                continue;
            }
            const bool mine = (pending >> lane) & 1u;
            const bool ready = mine && ml <= 64u && (lane == first || dstM - off + min(ml, off) <= frontier);
            const uint32_t maxml = __reduce_max_sync(kFull, ready ? ml : 0u);
            {
                // every source byte lies in the `off` bytes before the destination (periodic pattern
                // for self-overlapping matches), so all loads of a lane are independent
                const uint8_t* src = out + (dstM - off);
                uint8_t* dst = out + dstM;
                const uint32_t my = ready ? ml : 0u;
                uint32_t k = 0;
                for (uint32_t base = 0; base < maxml; base += 8) {
                    uint32_t v[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (base + j < my) v[j] = src[k];
                        k = (k + 1u == off) ? 0u : k + 1u;
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) if (base + j < my) dst[base + j] = (uint8_t)v[j];
                }
            }
            __syncwarp();
            pending &= ~__ballot_sync(kFull, ready);
        }
        op_base += __shfl_sync(kFull, incl, 31);
        head += n;
        __syncwarp();
        if (lane == 0) s->head = head;
    }
}

// One unit (partition / page) through the two-warp pipeline.  Called by all 64 threads of the CTA.
// Returns (on warp 0, lane 0 meaningful) bytes produced or kErrCorrupt.
__device__ inline int64_t ws_decode_unit(WsShared* s, Ring& ring, int codec, const uint8_t* in, uint32_t clen, uint8_t* out,
                                         uint32_t cap, bool last) {
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s->tail = 0; s->head = 0; s->done = 0; }
    __syncthreads();
    int64_t got = 0;
    if (warp == 0) {
        uint32_t tail = 0;
        if (codec == 0) got = ws_parse_lz4(ring, s, &tail, in, clen, cap, last, lane);
        else            got = ws_parse_snappy(ring, s, &tail, in, clen, cap, lane);
        ws_finish(s, tail, lane);
        ring.close();
    } else {
        ws_copy_records(s, in, out, lane);
    }
    __syncthreads();
    return got;
}

}  // namespace llc
