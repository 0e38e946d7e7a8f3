// decode_rowq.cuh -- row decoder: LANE PARSERS feeding THREAD-PER-BYTE ROW COPIERS through
// shared-memory queues.  One CTA keeps kQSlots independent units (RAP partitions or pages) in
// flight at once; nothing in it is CTA-wide (no __syncthreads after start-up, no pointer jumping).
//
// Why.  A unit is one serial token chain.  The tile decoder (decode_tile.cuh) breaks the chain with
// a data-parallel parse (pointer doubling over every byte position) and resolves in-group match
// sources with pointer jumping: 46 warp instructions per 10-byte sequence, 4 % of the HBM roofline.
// With >= ~5 units per SM in flight there is a cheaper organisation:
//
//   warp 0        PARSERS, one LANE per unit.  Each lane walks its own token chain with scalar code:
//                 one unaligned 8-byte window per sequence (two aligned 8-byte loads through L1,
//                 issued as soon as the token tells where the next sequence starts), every validity
//                 check of the reference decoders (lz4.c:3806-4305; snappy.cc:1466-1570, 2185-2199),
//                 and one 16-byte record {literal position, literal length, offset, match length}
//                 pushed into the unit's queue.  One warp instruction advances kQSlots chains.
//   warps 1..28   COPIERS, one warp per unit.  A copier takes 32 records at a time (lane per record:
//                 a warp scan of the lengths gives the output positions), then walks the output in
//                 aligned 32-byte ROWS, lane per byte: byte x finds its record from the start bits
//                 of the row (REDUX.OR + popc), fetches the record's fields with three shuffles and
//                 takes its value from the TMA-staged input ring (literal), from the unit's 4 KiB
//                 output ring in shared memory (match source in the last 4 KiB), from L2 (older
//                 source) or -- a source inside the same row -- by pointer doubling over shuffles
//                 (<= 5 rounds, only in rows that have such a byte).  Rows are executed in order, so
//                 every other source is final when it is read; the row goes to the output ring and
//                 to HBM (one full 32-byte sector per store instruction).
//
// The compressed stream reaches the copier through the TMA bulk-copy ring of in_ring.cuh
// (cp.async.bulk + mbarrier, 4 x 512 B per unit); the parser lanes read the same lines through L1.
// Queues are single-producer / single-consumer rings; unit boundaries travel in-band as marker
// records, so a parser lane starts its next unit (atomic ticket) while the copier is still busy.
//
// Accept / reject behaviour and produced bytes are identical to the warp decoders
// (lz4_decode_ring.cuh, snappy_codec.cuh) and through them to the reference.
#pragma once
#include "in_ring.cuh"
#include "snappy_codec.cuh"

namespace llc {

constexpr int kQSlots = 28;                                  // units in flight per CTA = copier warps
constexpr int kQThreads = 1024;                              // 32 warps: 1 parser, 28 copiers, 3 that leave at once (see rowq_run)
constexpr uint32_t kQCap = 64, kQMask = kQCap - 1;           // 32-bit entries per unit queue
constexpr uint32_t kQStride = kQCap + 1;                     // +1 entry: parser lanes hit different banks
constexpr uint32_t kQRoom = 8;                               // free entries a lane wants before it pushes a marker
constexpr uint32_t kORing = 4096, kORingMask = kORing - 1;   // output bytes kept in shared memory per unit
// Queue entries: a stream position (of a token / element the fast step accepted) or a marker followed by its payload.
constexpr uint32_t kMarkBase = 0xfffffff0u;
constexpr uint32_t kMarkBegin = kMarkBase + 0;               // + unit index
constexpr uint32_t kMarkSeq = kMarkBase + 1;                 // + literal position, literal length, offset, match length
constexpr uint32_t kMarkEnd = kMarkBase + 2;                 // + failed (0 / 1), bytes produced
constexpr uint32_t kMarkQuit = kMarkBase + 3;
constexpr uint32_t kQCapMax = 0xfffe0000u;                   // output positions + ring size must not wrap
constexpr uint32_t kQInMax = 0xffffff00u;                    // stream positions stay below the marker codes
constexpr uint32_t kPChunkLog = 9, kPChunk = 1u << kPChunkLog;   // parser window: two 512-byte chunks per unit
constexpr uint32_t kPChunks = 2, kPWin = kPChunks * kPChunk, kPWinMask = kPWin - 1;
constexpr uint32_t kQReach = 84;                             // a fast step touches stream bytes [ip, ip + kQReach)
constexpr uint32_t kQPass = 16;                              // fast steps per parser pass (queue bookkeeping runs once per pass)
constexpr uint32_t kQSpinMax = 1u << 20;                     // copier watchdog (~1 s of sleeping polls)
constexpr uint32_t kUnitBad = 0x100u;                        // QUnit.flags: the unit header itself is malformed

struct QShared {
    alignas(128) uint8_t oring[kQSlots][kORing];
    alignas(128) uint8_t idata[kQSlots][kRingBytes];
    alignas(128) uint8_t pwin[kQSlots][kPWin];               // parser lanes' own view of their streams (TMA fed)
    uint32_t q[kQSlots * kQStride];
    alignas(8) uint64_t ibar[kQSlots][kStages];
    alignas(8) uint64_t pbar[kQSlots][kPChunks];
    volatile uint32_t tail[kQSlots];                         // entries published by the parser lane
    volatile uint32_t head[kQSlots];                         // entries retired by the copier
    volatile uint32_t done[kQSlots];                         // the parser lane has published its last entry
    volatile uint32_t abort;                                 // watchdog: a queue made no progress, everybody leaves
};
static_assert(sizeof(QShared) <= 227 * 1024, "row decoder shared memory");

struct QUnit { const uint8_t* in; uint8_t* out; uint32_t clen, cap, flags; };

// Units = partitions [first, first + n) of a parsed RAP frame.
struct QPartsSource {
    const uint8_t* in; uint8_t* out; const PartDesc* parts; CallResult* res; uint32_t first; uint64_t origin;
    __device__ __forceinline__ bool open(uint32_t i, QUnit& u) const {
        const PartDesc d = parts[first + i];
        u.in = in + d.in_off; u.out = out + (d.out_off - origin); u.clen = d.in_len; u.cap = d.out_len; u.flags = d.flags;
        return d.in_len != 0;                                // zero-length partitions are skipped (threads.c:264-268)
    }
    __device__ __forceinline__ void report(uint32_t i, int64_t got, const QUnit& u) const {
        if (got < 0 || ((u.flags & kPartExact) && (uint64_t)got != u.cap)) atomicCAS(&res->error, 0, (int)(first + i) + 1);
        else if (!(u.flags & kPartExact)) res->value = got;  // frame-less LZ4: size is whatever was produced
    }
    __device__ __forceinline__ void fail() const { atomicCAS(&res->error, 0, 0x7fffffff); }
};
// Units = independent frame-less pages.
template <bool SNAPPY>
struct QPagesSource {
    const uint8_t* const* in_ptrs; const uint32_t* in_sizes; uint8_t* const* out_ptrs; const uint32_t* out_caps;
    long long* status; CallResult* res;
    __device__ __forceinline__ bool open(uint32_t i, QUnit& u) const {
        u.in = in_ptrs[i]; u.out = out_ptrs[i]; u.clen = in_sizes[i]; u.cap = out_caps[i]; u.flags = kPartLast;
        if (SNAPPY) {
            uint32_t total = 0;
            const uint32_t vb = get_varint32(u.in, u.clen, &total);
            if (vb == 0 || total > u.cap) u.flags |= kUnitBad;
            else { u.in += vb; u.clen -= vb; u.cap = total; u.flags |= kPartExact; }
        }
        return true;
    }
    __device__ __forceinline__ void report(uint32_t i, int64_t got, const QUnit&) const {
        status[i] = got;
        if (got < 0) atomicAdd(&res->error, 1);
    }
    __device__ __forceinline__ void fail() const { atomicAdd(&res->error, 1); }
};

// ----------------------------------------------------------------------------------------- parsers
// Per-lane parser state (registers).  Positions are offsets from gbase, the 16-byte aligned address at or below
// the unit's first byte (the unit starts at position `pad`).
//
// The lane reads its stream through a private four-chunk window in shared memory that it fills itself with the
// TMA bulk-copy engine: 28 lanes that walk 28 different streams through L1 in lock step pay somebody's cache
// miss in almost every iteration (measured: ~1 us per sequence); from shared memory a step costs one 30-cycle
// load.  Chunks [.., w_issued) have been requested, chunks [.., w_landed) are known to have arrived; stream
// bytes below wlim may be read.
struct QLane {
    const uint8_t* gbase;
    uint32_t ip, iend, op, cap;        // iend: stream end (relative to gbase)
    uint32_t fast_i_ex, fast_o_ex;     // exclusive bounds of the region where no end rule can fire
    uint32_t pad;
    uint32_t wbase;                    // shared-space address of the window
    uint32_t bar0;                     // shared-space address of its first mbarrier
    uint32_t w_issued, w_landed, wlim, wpar;   // wpar: bit b = parity of the next phase of buffer b
    uint32_t pf0;                      // the stream byte at ip, requested one step ahead (valid inside a pass of fast steps)
    bool last, bad;
};

__device__ __forceinline__ void qwin_issue(QLane& s, uint32_t c) {
    const uint32_t base = c << kPChunkLog, b = c & (kPChunks - 1u);
    const uint32_t left = s.iend - base;                    // caller: base < iend
    const uint32_t bytes = left >= kPChunk ? kPChunk : ((left + 15u) & ~15u);
    const uint32_t bar = s.bar0 + 8u * b;
    fence_proxy_async();                                     // my earlier reads of this buffer come first
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s.wbase + b * kPChunk), "l"(s.gbase + base), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void qwin_wait(QLane& s, uint32_t c) {
    const uint32_t b = c & (kPChunks - 1u), bar = s.bar0 + 8u * b, parity = (s.wpar >> b) & 1u;
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
    s.wpar ^= 1u << b;
}
// Keep the chunk of ip and the following kPChunks - 1 requested, and up to (chunk of ip) + 1 landed.  A chunk goes into
// buffer (chunk mod kPChunks), i.e. over chunk - kPChunks, which lies behind ip.
__device__ __forceinline__ void qwin_fill(QLane& s) {
    const uint32_t c = s.ip >> kPChunkLog;
    while (s.w_issued < c + kPChunks && (s.w_issued << kPChunkLog) < s.iend) { qwin_issue(s, s.w_issued); s.w_issued++; }
    // wait only for what the next steps can touch: the following chunk was requested a whole chunk ago and is
    // waited for when ip comes close to it (waiting for a chunk right after requesting it stalls all 28 lanes)
    const uint32_t want = min(((s.ip + 2u * kQReach) >> kPChunkLog) + 1u, s.w_issued);
    while (s.w_landed < want) { qwin_wait(s, s.w_landed); s.w_landed++; }
    s.wlim = s.w_landed << kPChunkLog;
}
// wait for everything in flight (before the buffers are reused for another place of the stream / another unit)
__device__ __forceinline__ void qwin_drain(QLane& s) {
    while (s.w_landed < s.w_issued) { qwin_wait(s, s.w_landed); s.w_landed++; }
}
// (re)start the window at the chunk of s.ip
__device__ __forceinline__ void qwin_open(QLane& s) {
    qwin_drain(s);
    s.w_issued = s.w_landed = s.ip >> kPChunkLog;
    qwin_fill(s);
}
__device__ __forceinline__ uint32_t qwin_u8(const QLane& s, uint32_t p) { return lds_u8(s.wbase + (p & kPWinMask)); }

// FAST STEP, LZ4: accept the sequence at ip knowing only where it ends; its fields are read again, lane per
// sequence, by the copier.  Precondition: ip < fast_i_ex and op < fast_o_ex -- at least 319 stream bytes and 559
// output bytes ahead, so a sequence with at most one length byte each (ll <= 269, ml <= 273) cannot trigger an
// end-of-block rule -- and ip + kQReach <= wlim.  Straight-line code, three independent-address loads.
// Returns false (nothing consumed) for a 255 length byte or a literal run beyond the reach: general step.
__device__ __forceinline__ bool qfast_lz4(QLane& s, uint32_t* q, uint32_t& tail) {
    const uint32_t ip = s.ip;
    const uint32_t tok = s.pf0;                              // the token at ip, requested by the previous step
    const uint32_t nibL = tok >> 4, nibM = tok & 15u;
    const bool extL = nibL == 15u, extM = nibM == 15u;
    uint32_t e1 = 0;
    if (extL) e1 = qwin_u8(s, ip + 1u);                     // (a literal run of 15 or more: one more load on the chain)
    const uint32_t ll = nibL + e1;
    const uint32_t rel = ll + (extL ? 2u : 1u);             // the offset field, relative to ip
    const uint32_t ipn = ip + rel + 2u + (extM ? 1u : 0u);
    // The chain token -> next token is all that is serial.  The next token is requested FIRST, the rest of this
    // step (length byte of the match, checks, queue entry) runs while it is on its way.
    s.pf0 = qwin_u8(s, ipn);
    const uint32_t e2 = qwin_u8(s, ip + rel + 2u);          // (garbage when rel is out of reach: rejected below)
    const uint32_t ml = nibM + 4u + (extM ? e2 : 0u);
    if ((extL & (e1 == 255u)) | (extM & (e2 == 255u)) | (rel + 5u > kQReach)) return false;   // (pf0 is stale now: the caller reloads)
    q[tail & kQMask] = ip;
    tail++;
    s.op += ll + ml;
    s.ip = ipn;
    return true;
}
// FAST STEP, Snappy.  Precondition: ip < fast_i_ex (319 stream bytes ahead), ip + kQReach <= wlim.  Returns false
// for a 4-byte-offset copy, a literal with length bytes, or an element that does not fit the output: general step.
__device__ __forceinline__ bool qfast_snappy(QLane& s, uint32_t* q, uint32_t& tail) {
    const uint32_t ip = s.ip;
    const uint32_t tag = s.pf0;
    const uint32_t kind = tag & 3u, hi = tag >> 2;
    const bool is_lit = kind == 0u;
    const uint32_t len = (kind == 1u) ? 4u + (hi & 7u) : hi + 1u;
    const uint32_t adv = is_lit ? 1u + len : (kind == 1u ? 2u : 3u);
    s.pf0 = qwin_u8(s, ip + adv);                           // (an element the fast step refuses is at most 64 bytes ahead)
    if ((kind == 3u) | (is_lit & (hi >= 60u)) | (len > s.cap - s.op)) return false;
    q[tail & kQMask] = ip;
    tail++;
    s.op += len;
    s.ip = ip + adv;
    return true;
}

// Marker with a fully parsed sequence (the general steps below; positions relative to the unit's first byte)
template <class S>
__device__ __forceinline__ void qpush_seq(const S& s, uint32_t* q, uint32_t& tail, uint32_t lp, uint32_t ll, uint32_t off, uint32_t ml) {
    q[tail & kQMask] = kMarkSeq; q[(tail + 1u) & kQMask] = lp + s.pad; q[(tail + 2u) & kQMask] = ll;
    q[(tail + 3u) & kQMask] = off; q[(tail + 4u) & kQMask] = ml;
    tail += 5u;
}

// One LZ4 sequence through the general code (bytes straight from global memory): length runs, block tail, tiny
// units -- every check of the reference.  Returns false when the unit is finished (ok or bad).  Positions in the
// lane state are relative to gbase; the arithmetic below is relative to the unit's first byte.
// State goes in and out BY VALUE: a noinline function that takes the lane state by reference forces the whole state
// into local memory, and the fast loop then pays a dozen local loads / stores per pass.
struct QSlow { uint32_t ip, op, tail; bool more, bad; };
struct QSlowIn { const uint8_t* gbase; uint32_t pad, ip, iend, op, cap; bool last; };
__device__ __noinline__ QSlow qlane_step_lz4(const QSlowIn a, uint32_t* q, uint32_t tail_in) {
    struct { const uint8_t* gbase; uint32_t pad, ip, iend, op, cap; bool last, bad; } s = {a.gbase, a.pad, a.ip, a.iend, a.op, a.cap, a.last, false};
    uint32_t tail = tail_in;
    const bool more = [&]() -> bool {
    const uint8_t* in = s.gbase + s.pad;
    const uint32_t ip = s.ip - s.pad;
    const uint32_t iend = s.iend - s.pad, cap = s.cap;
    if (ip >= iend) { s.bad = true; return false; }
    const uint32_t tok = in[ip];
    const uint32_t nibL = tok >> 4, nibM = tok & 15u;
    const bool last = s.last;
    uint32_t p = ip + 1, ll = nibL;
    if (ll == 15u) {
        uint32_t b;
        do { if (p >= iend) { s.bad = true; return false; } b = in[p++]; ll += b; } while (b == 255u && ll < 0x7fff0000u);
    }
    if (ll > iend - p || ll > cap - s.op) { s.bad = true; return false; }
    const bool closing = ((uint64_t)s.op + ll + 12 > cap) || ((uint64_t)p + ll + 8 > iend);   // lz4.c:4104-4164
    if (closing && last && p + ll != iend) { s.bad = true; return false; }
    const uint32_t lit_pos = p;
    s.op += ll;
    const uint32_t qq = p + ll;
    if ((closing && (last || s.op == cap)) || qq == iend) {
        if (ll) qpush_seq(s, q, tail, lit_pos, ll, 0, 0);
        s.ip = qq + s.pad;
        return false;
    }
    if (qq + 2 > iend) { s.bad = true; return false; }
    const uint32_t off = (uint32_t)in[qq] | ((uint32_t)in[qq + 1] << 8);
    p = qq + 2;
    uint32_t ml = nibM;
    if (ml == 15u) {
        uint32_t b;
        do { if (p >= iend) { s.bad = true; return false; } b = in[p++]; ml += b; } while (b == 255u && ml < 0x7fff0000u);
    }
    ml += 4;
    if (off == 0 || off > s.op || ml > cap - s.op) { s.bad = true; return false; }   // lz4.c:4196-4197
    if (last && (uint64_t)s.op + ml + 5 > cap) { s.bad = true; return false; }       // lz4.c:4262-4264
    qpush_seq(s, q, tail, lit_pos, ll, off, ml);
    s.op += ml;
    s.ip = p + s.pad;
    if (!last && (s.op == cap || p >= iend)) return false;                           // lz4.c:4285-4288
    return true;
    }();
    return QSlow{s.ip, s.op, tail, more, s.bad};
}

// One Snappy element through the general code (cap is the exact size the stream must produce).
__device__ __noinline__ QSlow qlane_step_snappy(const QSlowIn a, uint32_t* q, uint32_t tail_in) {
    struct { const uint8_t* gbase; uint32_t pad, ip, iend, op, cap; bool last, bad; } s = {a.gbase, a.pad, a.ip, a.iend, a.op, a.cap, a.last, false};
    uint32_t tail = tail_in;
    const bool more = [&]() -> bool {
    const uint8_t* in = s.gbase + s.pad;
    const uint32_t ip = s.ip - s.pad, iend = s.iend - s.pad, expect = s.cap;
    if (ip >= iend) return false;
    const uint32_t tag = in[ip];
    const uint32_t kind = tag & 3u, hi = tag >> 2;
    uint32_t len, nip;
    if (kind == 0u) {                                   // literal, snappy.cc:1492-1527
        uint32_t p = ip + 1u;
        len = hi + 1u;
        if (len > 60u) {
            const uint32_t nb = len - 60u;
            if (p + nb > iend) { s.bad = true; return false; }
            uint32_t v = 0;
            for (uint32_t k = 0; k < nb; k++) v |= (uint32_t)in[p + k] << (8 * k);
            if (v == 0xffffffffu) { s.bad = true; return false; }
            len = v + 1u; p += nb;
        }
        if (len > iend - p || len > expect - s.op) { s.bad = true; return false; }
        qpush_seq(s, q, tail, p, len, 0, 0);
        nip = p + len;
    } else {                                            // char_table, snappy-internal.h:406-439
        uint32_t adv, off;
        if (kind == 1u) {
            if (ip + 2 > iend) { s.bad = true; return false; }
            len = 4u + (hi & 7u); off = ((tag >> 5) << 8) | in[ip + 1]; adv = 2;
        } else if (kind == 2u) {
            if (ip + 3 > iend) { s.bad = true; return false; }
            len = 1u + hi; off = (uint32_t)in[ip + 1] | ((uint32_t)in[ip + 2] << 8); adv = 3;
        } else {
            if (ip + 5 > iend) { s.bad = true; return false; }
            len = 1u + hi;
            off = (uint32_t)in[ip + 1] | ((uint32_t)in[ip + 2] << 8) | ((uint32_t)in[ip + 3] << 16) | ((uint32_t)in[ip + 4] << 24);
            adv = 5;
        }
        if (off == 0 || off > s.op || len > expect - s.op) { s.bad = true; return false; }   // snappy.cc:2185-2199
        qpush_seq(s, q, tail, ip + 1u, 0, off, len);
        nip = ip + adv;
    }
    s.op += len;
    s.ip = nip + s.pad;
    return true;
    }();
    return QSlow{s.ip, s.op, tail, more, s.bad};
}

// The parser warp: lane l owns slot l.  A lane that finishes a unit pushes the END marker (the copier reports the
// result: it validates the offsets of the fast sequences) and draws the next unit: the first unit of every slot is
// assigned statically (interleaved over the grid, so that a frame with fewer units than slots spreads over all
// SMs), later ones come from the atomic ticket.
template <bool SNAPPY, class Src>
__device__ inline void rowq_parse(QShared& sh, const Src& src, uint32_t nunits, unsigned int* ticket, int lane) {
    const int slot = lane < kQSlots ? lane : 0;
    uint32_t* const q = sh.q + slot * kQStride;
    QLane s;
    QUnit u;
    s.gbase = nullptr; s.ip = s.iend = s.op = s.cap = s.fast_i_ex = s.fast_o_ex = s.pad = 0;
    s.wbase = smem_u32(sh.pwin[slot]); s.bar0 = smem_u32(&sh.pbar[slot][0]);
    s.w_issued = s.w_landed = s.wlim = s.wpar = 0;
    s.pf0 = 0;
    s.last = false; s.bad = false;
    u.in = nullptr; u.out = nullptr; u.clen = u.cap = u.flags = 0;
    uint32_t tail = 0, published = 0, head_c = 0;
    bool active = false, alive = lane < kQSlots, first_fetch = true;
    while (__any_sync(kFull, alive)) {
        if (sh.abort) break;
        // ---- fast steps: up to 8 sequences per lane back to back (the bookkeeping below runs once per pass)
        bool slow = !active;                                 // this lane needs the general code
        // What bounds the pass, taken once: free queue entries, how far the stream may be read, and how many steps
        // certainly stay inside the output region where no end rule can fire (a step produces < 544 bytes).
        const uint32_t ip_stop = min(s.fast_i_ex, s.wlim > kQReach ? s.wlim - kQReach : 0u);
        uint32_t left = 0;
        if (active) {
            left = min(kQPass, kQCap - (tail - head_c));
            if (!SNAPPY) left = min(left, s.fast_o_ex > s.op ? (s.fast_o_ex - s.op + 543u) / 544u : 0u);
            if (s.ip >= ip_stop) left = 0;
            if (left) s.pf0 = qwin_u8(s, s.ip);
        }
#pragma unroll 1
        for (uint32_t it = 0; it < kQPass; it++) {
            const bool can = it < left && s.ip < ip_stop;
            if (!__any_sync(kFull, can)) break;
            if (can && !(SNAPPY ? qfast_snappy(s, q, tail) : qfast_lz4(s, q, tail))) { slow = true; left = 0u; }
        }
        bool moved = false;
        if (alive) {
            if (tail - head_c + kQRoom > kQCap) head_c = sh.head[lane];
            const bool room = tail - head_c + kQRoom <= kQCap;
            moved = room;
            if (active && !slow) {
                if (s.ip < s.fast_i_ex && (SNAPPY || s.op < s.fast_o_ex)) qwin_fill(s);   // (usually nothing to do)
                else slow = true;                                                        // left the fast region
            }
            if (room && slow) {
                if (!active) {
                    uint32_t i;
                    if (first_fetch) { i = (uint32_t)lane * gridDim.x + blockIdx.x; first_fetch = false; }
                    else i = gridDim.x * (uint32_t)kQSlots + atomicAdd(ticket, 1u);
                    if (i >= nunits) {
                        q[tail & kQMask] = kMarkQuit; tail++;
                        alive = false;
                    } else if (src.open(i, u)) {
                        s.pad = (uint32_t)(reinterpret_cast<uintptr_t>(u.in) & 15);
                        s.gbase = u.in - s.pad;
                        s.iend = u.clen + s.pad; s.cap = min(u.cap, kQCapMax);
                        s.ip = s.pad; s.op = 0; s.bad = false;
                        s.last = (u.flags & kPartLast) != 0;
                        const bool any_fast = u.clen >= 320u && (SNAPPY || s.cap >= 560u);
                        s.fast_i_ex = any_fast ? s.iend - 319u : 0u;
                        s.fast_o_ex = (any_fast && !SNAPPY) ? s.cap - 559u : 0u;
                        bool run = true;
                        if ((u.flags & kUnitBad) || u.clen > kQInMax) { s.bad = true; run = false; }
                        else if (!SNAPPY) {
                            if (u.clen == 0) { s.bad = true; run = false; }
                            else if (s.cap == 0) { s.bad = !(u.clen == 1 && u.in[0] == 0); run = false; }   // lz4.c:3854-3858
                        }
                        q[tail & kQMask] = kMarkBegin; q[(tail + 1u) & kQMask] = i; tail += 2u;
                        if (run) {
                            if (any_fast) qwin_open(s);
                            active = true;
                        } else {                             // nothing to decode: BEGIN is followed by END right away
                            q[tail & kQMask] = kMarkEnd; q[(tail + 1u) & kQMask] = s.bad ? 1u : 0u; q[(tail + 2u) & kQMask] = 0u; tail += 3u;
                        }
                    }
                } else {
                    bool more = false;
                    if (!s.bad) {
                        const QSlowIn a{s.gbase, s.pad, s.ip, s.iend, s.op, s.cap, s.last};
                        const QSlow r = SNAPPY ? qlane_step_snappy(a, q, tail) : qlane_step_lz4(a, q, tail);
                        s.ip = r.ip; s.op = r.op; tail = r.tail; s.bad = r.bad; more = r.more;
                    }
                    if (!more) {
                        const bool failed = s.bad || (SNAPPY && s.op != s.cap);            // snappy.cc:1715
                        q[tail & kQMask] = kMarkEnd; q[(tail + 1u) & kQMask] = failed ? 1u : 0u; q[(tail + 2u) & kQMask] = s.op; tail += 3u;
                        active = false;
                    } else if (s.ip < s.fast_i_ex) {
                        // back to the fast steps: restart the window unless ip is still inside what it holds
                        if ((s.ip >> kPChunkLog) + 1u >= s.w_landed || (s.ip >> kPChunkLog) + kPChunks <= s.w_issued) qwin_open(s);
                        else qwin_fill(s);
                    }
                }
            }
            if (tail - published >= 16u || (!(active && !slow) && tail != published)) {
                __threadfence_block();
                sh.tail[lane] = tail;
                published = tail;
                if (!alive) { __threadfence_block(); sh.done[lane] = 1u; }
            }
        }
        if (!__any_sync(kFull, moved)) __nanosleep(200);    // every queue is full: the copiers are the bottleneck
    }
    qwin_drain(s);                                           // nothing may stay in flight
}

// ----------------------------------------------------------------------------------------- copiers
// Per-batch constants of a copier (uniform across the warp unless noted).
struct QBatch {
    uint32_t op, total;          // output span of the batch: [op, op + total)
    uint32_t sbase, obase;       // shared-space addresses of the input ring and of the output ring
    const uint8_t* gin;          // input-ring base in global memory (positions are relative to it)
    uint8_t* gout;               // 32-byte aligned output base (positions are relative to it)
    uint32_t D, M, L, O;         // per lane = per record: first byte, first match byte, literal position - first byte, offset
    uint32_t lemask;             // per lane: bits 1 .. lane
};

// One 32-byte row [x0, x0 + 32), lane per byte.  FULL: every byte of the row belongs to the batch.
template <bool USE_RING, bool FULL>
__device__ __forceinline__ void rowq_row(const QBatch& B, uint32_t x0, int lane) {
    // which record covers byte x: the records that start at or before x0, plus the starts inside the row up to x
    const uint32_t rel = B.D - x0;
    uint32_t bit;
    asm("shl.b32 %0, 1, %1;" : "=r"(bit) : "r"(rel));        // 0 when rel >= 32: PTX clamps the shift amount
    const uint32_t bits = __reduce_or_sync(kFull, bit);
    const uint32_t cnt0 = (uint32_t)__popc(__ballot_sync(kFull, B.D <= x0));
    const uint32_t k = cnt0 - 1u + (uint32_t)__popc(bits & B.lemask);
    const uint32_t m = __shfl_sync(kFull, B.M, (int)k), l = __shfl_sync(kFull, B.L, (int)k), o = __shfl_sync(kFull, B.O, (int)k);
    const uint32_t x = x0 + (uint32_t)lane;
    const bool live = FULL ? true : (x - B.op) < B.total;
    const bool is_lit = x < m;
    const bool mat = live && !is_lit;
    // source x - o: inside this row (not written yet) / in the output ring / older than the ring (read back from L2)
    const bool dep = FULL ? (mat && o <= (uint32_t)lane) : (mat && (x - o) >= max(x0, B.op));
    const bool far = mat && o > (uint32_t)lane + (kORing - 32u);
    uint32_t v = 0;
    if (live && is_lit) v = USE_RING ? lds_u8(B.sbase + ((x + l) & kRingMask)) : (uint32_t)B.gin[x + l];
    if (mat && !dep && !far) v = lds_u8(B.obase + ((x - o) & kORingMask));
    if (__any_sync(kFull, far)) { if (far) v = __ldcg(B.gout + (x - o)); }
    if (__any_sync(kFull, dep)) {
        // pointer doubling over the lanes of the row: a byte whose source lane is still unknown adopts that
        // lane's source; every round at least halves the chains (<= 5 rounds)
        uint32_t j = (uint32_t)lane - o;
        bool need = dep;
        do {
            const uint32_t vj = __shfl_sync(kFull, v, (int)j);
            const uint32_t jj = __shfl_sync(kFull, j, (int)j);
            const bool nj = __shfl_sync(kFull, need ? 1 : 0, (int)j) != 0;
            if (need) { if (!nj) { v = vj; need = false; } else j = jj; }
        } while (__any_sync(kFull, need));
    }
    if (live) {
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(B.obase + (x & kORingMask)), "r"(v) : "memory");
        B.gout[x] = (uint8_t)v;
    }
    __syncwarp();
}

template <bool USE_RING>
__device__ __forceinline__ void rowq_rows(const QBatch& B, int lane) {
    const uint32_t op_end = B.op + B.total;
    uint32_t x0 = B.op & ~31u;
    if (x0 != B.op) { rowq_row<USE_RING, false>(B, x0, lane); x0 += 32u; }         // the batch starts inside a row
    for (; x0 + 32u <= op_end; x0 += 32u) rowq_row<USE_RING, true>(B, x0, lane);
    if (x0 < op_end) rowq_row<USE_RING, false>(B, x0, lane);                       // ... and ends inside one
}

// Executes the n (1..32) records held lane-per-record in `rec` = {literal position, literal length, offset,
// match length}; lanes >= n hold nothing.  `a` is the position of the unit's first output byte: a match may not
// reach before it (lz4.c:4196-4197, snappy.cc:2190-2191) -- the records of the fast steps are validated here,
// lane per record.  Returns false when a record is bad (the records before it have been executed).
__device__ __forceinline__ bool rowq_batch(Ring& ring, uint32_t obase, uint8_t* gout, uint32_t a, uint32_t& op_io,
                                           const uint4 rec, uint32_t n, int lane) {
    QBatch B;
    B.op = op_io;
    bool valid = (uint32_t)lane < n;
    const uint32_t ll = valid ? rec.y : 0u, ml = valid ? rec.w : 0u;
    const uint32_t len = ll + ml;
    const uint32_t incl = warp_incl_sum(len, lane);
    const uint32_t d = B.op + incl - len;                    // first output byte of my record
    B.M = d + ll;
    B.O = rec.z;
    const unsigned badmask = __ballot_sync(kFull, valid && ml != 0u && (B.O - 1u) >= B.M - a);
    if (badmask) {
        n = (uint32_t)__ffs(badmask) - 1u;
        if (n == 0) return false;
        valid = (uint32_t)lane < n;
    }
    B.total = __shfl_sync(kFull, incl, (int)n - 1) ;         // (all of it: lanes >= the original n add nothing)
    B.D = valid ? d : 0xffffffffu;
    const uint32_t lpos = rec.x;                             // literals of my record in ring coordinates
    B.L = lpos - d;                                          // literal byte x of my record sits at ring position x + L
    B.lemask = ((2u << lane) - 1u) & ~1u;
    B.obase = obase; B.gout = gout;
    // literal window of the batch (positions grow with the lane): the TMA ring when it fits, else global loads
    const uint32_t lit_lo = __shfl_sync(kFull, lpos, 0);
    const uint32_t lit_hi = __shfl_sync(kFull, lpos + ll, (int)n - 1);
    const bool use_ring = lit_hi <= (lit_lo & ~(kChunk - 1u)) + kRingBytes;
    B.sbase = smem_u32(ring.sm);
    B.gin = ring.gbase;
    if (use_ring) {
        ring.advance(lit_lo, lane);
        ring.ensure(lit_hi);
        rowq_rows<true>(B, lane);
    } else {
        rowq_rows<false>(B, lane);
    }
    op_io = B.op + B.total;
    return badmask == 0;
}

// Lane per entry: the fields of the sequence / element whose token sits at stream position p (the parser's fast
// step accepted it: at most one length byte each, everything within kQReach bytes, inside the stream).
template <bool SNAPPY>
__device__ __forceinline__ uint4 rowq_fields(const Ring& ring, uint32_t p) {
    if (!SNAPPY) {
        const uint32_t tok = ring.byte(p), e1 = ring.byte(p + 1u);
        const uint32_t nibL = tok >> 4, nibM = tok & 15u;
        const bool extL = nibL == 15u, extM = nibM == 15u;
        const uint32_t ll = extL ? 15u + e1 : nibL;
        const uint32_t lit = p + (extL ? 2u : 1u), qq = lit + ll;
        const uint32_t off = ring.byte(qq) | (ring.byte(qq + 1u) << 8), e2 = ring.byte(qq + 2u);
        return make_uint4(lit, ll, off, nibM + 4u + (extM ? e2 : 0u));
    } else {
        const uint32_t tag = ring.byte(p), b1 = ring.byte(p + 1u), b2 = ring.byte(p + 2u);
        const uint32_t kind = tag & 3u, hi = tag >> 2;
        if (kind == 0u) return make_uint4(p + 1u, hi + 1u, 0u, 0u);
        if (kind == 1u) return make_uint4(p + 1u, 0u, ((tag >> 5) << 8) | b1, 4u + (hi & 7u));
        return make_uint4(p + 1u, 0u, b1 | (b2 << 8), hi + 1u);
    }
}

// A copier warp: consumes the queue of its slot until the parser's QUIT marker.
template <bool SNAPPY, class Src>
__device__ inline void rowq_copy(QShared& sh, const Src& src, int slot, int lane) {
    Ring ring;
    ring.init(sh.idata[slot], sh.ibar[slot], lane);
    const uint32_t obase = smem_u32(sh.oring[slot]);
    const uint32_t* const q = sh.q + slot * kQStride;
    uint32_t head = 0, op = 0, a = 0, unit = 0;
    uint8_t* gout = nullptr;
    QUnit u;
    u.in = nullptr; u.out = nullptr; u.clen = u.cap = u.flags = 0;
    bool open = false, ubad = false;
    for (;;) {
        // wait for a full batch (or for whatever is left once the parser lane is done); sleeping, not spinning:
        // a polling copier takes issue slots from the parser warp it is waiting for
        uint32_t avail, spins = 0, ns = 64;
        for (;;) {
            avail = sh.tail[slot] - head;
            if (avail >= 32u) break;
            if (sh.done[slot]) { __threadfence_block(); avail = sh.tail[slot] - head; if (avail) break; }
            if (sh.abort || ++spins > kQSpinMax) {           // never hang: flag the call and leave
                if (lane == 0 && !sh.abort) { sh.abort = 1u; src.fail(); }
                avail = 0;
                break;
            }
            __nanosleep(ns);
            if (ns < 1024u) ns <<= 1;
        }
        if (avail == 0) break;
        __threadfence_block();
        uint32_t n = min(avail, 32u);
        const uint32_t e = (uint32_t)lane < n ? q[(head + lane) & kQMask] : 0u;
        const unsigned marks = __ballot_sync(kFull, (uint32_t)lane < n && e >= kMarkBase);
        uint32_t m = marks ? (uint32_t)__ffs(marks) - 1u : n;   // plain positions in front of the first marker
        if (m) {
            // keep what the input ring can hold at once (a fast sequence spans < kQReach bytes)
            const uint32_t p0 = __shfl_sync(kFull, e, 0);
            const uint32_t fit = (uint32_t)__popc(__ballot_sync(kFull, (uint32_t)lane < m && e + kQReach <= (p0 & ~(kChunk - 1u)) + kRingBytes));
            if (fit < m) { m = fit; }
            if (!ubad) {
                ring.advance(p0, lane);
                ring.ensure(__shfl_sync(kFull, e, (int)m - 1) + kQReach);
                uint4 rec = make_uint4(0, 0, 0, 0);
                if ((uint32_t)lane < m) rec = rowq_fields<SNAPPY>(ring, e);
                if (!rowq_batch(ring, obase, gout, a, op, rec, m, lane)) ubad = true;
            }
            head += m;
        } else {
            const uint32_t code = __shfl_sync(kFull, e, 0);
            // the parser publishes a marker together with its payload
            const uint32_t w1 = q[(head + 1u) & kQMask], w2 = q[(head + 2u) & kQMask], w3 = q[(head + 3u) & kQMask], w4 = q[(head + 4u) & kQMask];
            if (code == kMarkSeq) {
                uint4 rec = make_uint4(0, 0, 0, 0);
                if (lane == 0) rec = make_uint4(w1, w2, w3, w4);
                if (!ubad && !rowq_batch(ring, obase, gout, a, op, rec, 1u, lane)) ubad = true;
                head += 5u;
            } else if (code == kMarkBegin) {
                unit = w1;
                src.open(unit, u);
                a = (uint32_t)(reinterpret_cast<uintptr_t>(u.out) & 31);
                gout = u.out - a;                            // rows are aligned 32-byte sectors of the output
                op = a;
                ubad = false;
                if (open) ring.close();
                ring.open(u.in, u.clen);
                open = true;
                head += 2u;
            } else if (code == kMarkEnd) {
                if (lane == 0) src.report(unit, (w1 != 0u || ubad) ? kErrCorrupt : (long long)w2, u);
                head += 3u;
            } else {                                         // kMarkQuit
                break;
            }
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); sh.head[slot] = head; }
    }
    if (open) ring.close();                                  // nothing may stay in flight
}

template <bool SNAPPY, class Src>
__device__ __forceinline__ void rowq_run(QShared& sh, const Src& src, uint32_t nunits, unsigned int* ticket) {
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x < kQSlots) { sh.tail[threadIdx.x] = 0; sh.head[threadIdx.x] = 0; sh.done[threadIdx.x] = 0; }
    if (threadIdx.x == 0) sh.abort = 0;
    if (threadIdx.x < kQSlots) {                             // parser windows: one mbarrier per buffer
        for (uint32_t b = 0; b < kPChunks; b++) mbar_init(&sh.pbar[threadIdx.x][b], 1);
        fence_mbar_init();
    }
    __syncthreads();
    // Warp w issues from scheduler w mod 4.  The parser warp is a single dependent chain that sets the pace of the
    // whole CTA, so it gets scheduler 0 almost to itself: warps 4, 8 and 12 leave at once, scheduler 0 keeps the
    // parser and 4 copiers, the other three schedulers take 8 copiers each.
    if (warp == 0) rowq_parse<SNAPPY>(sh, src, nunits, ticket, lane);
    else if (warp & 3) rowq_copy<SNAPPY>(sh, src, (warp >> 2) * 3 + (warp & 3) - 1, lane);
    else if (warp >= 16) rowq_copy<SNAPPY>(sh, src, 24 + ((warp - 16) >> 2), lane);
}

}  // namespace llc
