// decode_bundle.cuh -- bundle decoder: one CTA decodes up to 32 independent units (RAP partitions
// or pages) at once.
//
// Why: a unit is one serial token chain (token -> lengths -> next token), so a warp that parses one
// unit keeps 31 lanes idle and the whole GPU ends up instruction-issue bound at ~100 warp
// instructions per 10-byte sequence.  Here the chains run on LANES:
//
//   warp 0        32 PARSERS, one lane per unit.  Each lane walks its own token chain with plain
//                 scalar code (bytes through L1, next 128-byte line prefetched), performs every
//                 validity check of the reference decoders (lz4.c:3806-4305; snappy.cc:1466-1570,
//                 2185-2199) and pushes one 16-byte record per sequence {literal position, literal
//                 length, match offset, match length} into that unit's shared-memory queue.
//                 One warp instruction advances 32 chains.
//   warps 1..16   COPIERS, each serving two units.  A copier takes 32 records of one unit at a
//                 time, ONE LANE PER SEQUENCE: warp scan of the lengths -> output offsets, every lane
//                 copies its own literal run, then the matches run in dependency rounds (a match is
//                 ready when its source ends before the destination of the first unfinished match of
//                 the group).  Runs > 32 literal / > 64 match bytes are copied by the whole warp.
//
// Queues are single-producer / single-consumer rings with counters in shared memory.
#pragma once
#include "llc_common.cuh"
#include "snappy_codec.cuh"

namespace llc {

constexpr int kBSlots = 32;                  // units per CTA
constexpr int kBCopiers = 16;                // copier warps per CTA (two units each)
constexpr int kBThreads = 32 * (1 + kBCopiers);
constexpr uint32_t kBQCap = 64;              // records per unit queue
constexpr uint32_t kBQMask = kBQCap - 1;
constexpr uint32_t kBQStride = kBQCap + 1;   // +1 record of padding: parser lanes hit different banks

struct BSlot {
    const uint8_t* in;          // compressed unit
    uint8_t* out;               // where its output goes
    uint32_t clen, cap;         // compressed bytes, output capacity / expected size
    uint32_t flags;             // bit0 last/frame-less rules, bit1 exact size required, bit2 snappy, bit3 occupied
    volatile uint32_t tail;     // records published
    volatile uint32_t head;     // records retired
    volatile uint32_t done;     // parser finished this unit
    long long result;           // bytes produced or kErrCorrupt (parser)
};
constexpr uint32_t kBLast = 1u, kBExact = 2u, kBSnappy = 4u, kBUsed = 8u;

struct BShared {
    uint4 q[kBSlots * kBQStride];
    BSlot slot[kBSlots];
    uint32_t bundle;
};

// Copier loads do not allocate in L1: the 28 parsers of a CTA live off L1-resident lines of their
// compressed streams, and the copiers' match reads (random 64 KiB windows) would evict them.
__device__ __forceinline__ uint32_t ld_na_u8(const uint8_t* p) {
    uint32_t v;
    asm volatile("ld.global.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// ----------------------------------------------------------------------------------------- parsers
// Per-lane parser state (registers).  All positions are offsets from the unit's first byte.
struct LaneParse {
    const uint8_t* in;
    uint32_t ip, iend, op, cap, tail, flags;
    uint32_t fast_i_ex, fast_o_ex;     // exclusive bounds of the region where no end rule can fire
    bool active, bad;
};

// One LZ4 sequence for this lane.  Returns false when the unit is finished (ok or bad).
__device__ __forceinline__ bool lane_step_lz4(LaneParse& s, uint4* q) {
    const uint8_t* in = s.in;
    const uint32_t ip = s.ip;
    if (ip >= s.iend) { s.bad = true; return false; }
    const uint32_t tok = in[ip];
    const uint32_t nibL = tok >> 4, nibM = tok & 15u;
    if ((ip < s.fast_i_ex) & (s.op < s.fast_o_ex)) {
        const uint32_t e1 = in[ip + 1];
        const bool extL = nibL == 15u, extM = nibM == 15u;
        const uint32_t ll = nibL + (extL ? e1 : 0u);
        const uint32_t lp = ip + 1u + (extL ? 1u : 0u);
        const uint32_t qq = lp + ll;
        const uint32_t o0 = in[qq], o1 = in[qq + 1], e2 = in[qq + 2];
        if (!((extL & (e1 == 255u)) | (extM & (e2 == 255u)))) {
            // at most one length byte each and far from both ends: no end-of-block rule can fire
            const uint32_t off = o0 | (o1 << 8);
            const uint32_t ml = nibM + 4u + (extM ? e2 : 0u);
            const uint32_t op2 = s.op + ll;
            if ((off - 1u) >= op2) { s.bad = true; return false; }       // lz4.c:4196-4197
            q[s.tail & kBQMask] = make_uint4(lp, ll, off, ml);
            s.tail++;
            s.op = op2 + ml;
            const uint32_t ipn = qq + 2u + (extM ? 1u : 0u);
            if ((ipn >> 7) != (ip >> 7)) prefetch_l1(in + ((ipn >> 7) + 2u) * 128u);
            s.ip = ipn;
            return true;
        }
    }
    // general case: length runs, block tail, tiny units -- every check of the reference
    const bool last = (s.flags & kBLast) != 0;
    const uint32_t iend = s.iend, cap = s.cap;
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
        q[s.tail & kBQMask] = make_uint4(lit_pos, ll, 0, 0); s.tail++;
        s.ip = qq;
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
    q[s.tail & kBQMask] = make_uint4(lit_pos, ll, off, ml); s.tail++;
    s.op += ml;
    s.ip = p;
    prefetch_l1(in + ((p >> 7) + 1u) * 128u);
    if (!last && (s.op == cap || p >= iend)) return false;                           // lz4.c:4285-4288
    return true;
}

// One Snappy element for this lane (cap is the exact size the stream must produce).
__device__ __forceinline__ bool lane_step_snappy(LaneParse& s, uint4* q) {
    const uint8_t* in = s.in;
    const uint32_t ip = s.ip, iend = s.iend, expect = s.cap;
    if (ip >= iend) return false;
    const uint32_t tag = in[ip];
    const uint32_t kind = tag & 3u, hi = tag >> 2;
    if (ip < s.fast_i_ex && kind != 3u && !(kind == 0u && hi >= 60u)) {
        const uint32_t b1 = in[ip + 1], b2 = in[ip + 2];
        const bool is_lit = kind == 0u;
        const uint32_t lit_len = hi + 1u;
        const uint32_t cp_len = (kind == 1u) ? 4u + (hi & 7u) : 1u + hi;
        const uint32_t cp_off = (kind == 1u) ? (((tag >> 5) << 8) | b1) : (b1 | (b2 << 8));
        const uint32_t len = is_lit ? lit_len : cp_len;
        if (len > expect - s.op || (!is_lit && (cp_off - 1u) >= s.op)) { s.bad = true; return false; }   // snappy.cc:2185-2199
        q[s.tail & kBQMask] = make_uint4(ip + 1u, is_lit ? lit_len : 0u, cp_off, is_lit ? 0u : cp_len);
        s.tail++;
        s.op += len;
        const uint32_t ipn = ip + (is_lit ? 1u + lit_len : (kind == 1u ? 2u : 3u));
        if ((ipn >> 7) != (ip >> 7)) prefetch_l1(in + ((ipn >> 7) + 2u) * 128u);
        s.ip = ipn;
        return true;
    }
    if (kind == 0u) {                                   // literal, snappy.cc:1492-1527
        uint32_t len = hi + 1u, p = ip + 1u;
        if (len > 60u) {
            const uint32_t nb = len - 60u;
            if (p + nb > iend) { s.bad = true; return false; }
            uint32_t v = 0;
            for (uint32_t k = 0; k < nb; k++) v |= (uint32_t)in[p + k] << (8 * k);
            if (v == 0xffffffffu) { s.bad = true; return false; }
            len = v + 1u; p += nb;
        }
        if (len > iend - p || len > expect - s.op) { s.bad = true; return false; }
        q[s.tail & kBQMask] = make_uint4(p, len, 0, 0); s.tail++;
        s.op += len; s.ip = p + len;
        prefetch_l1(in + ((s.ip >> 7) + 1u) * 128u);
        return true;
    }
    uint32_t len, off, adv;                             // char_table, snappy-internal.h:406-439
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
    q[s.tail & kBQMask] = make_uint4(0, 0, off, len); s.tail++;
    s.op += len; s.ip = ip + adv;
    return true;
}

// The parser warp: lane l owns slot l.  Runs until every occupied slot is finished.
__device__ inline void bundle_parse(BShared* sh, int lane) {
    BSlot* slot = &sh->slot[lane];
    uint4* q = sh->q + lane * kBQStride;
    LaneParse s;
    s.flags = slot->flags;
    s.active = (s.flags & kBUsed) != 0;
    s.in = slot->in; s.iend = slot->clen; s.cap = slot->cap;
    s.ip = 0; s.op = 0; s.tail = 0; s.bad = false;
    const bool snappy = (s.flags & kBSnappy) != 0;
    // fast region: a sequence with at most one length byte each (ll <= 269, ml <= 273) fits in 320
    // input bytes and 560 output bytes, so no end-of-block rule can apply inside it
    const bool any_fast = s.iend >= 320u && (snappy || s.cap >= 560u);
    s.fast_i_ex = any_fast ? s.iend - 319u : 0u;
    s.fast_o_ex = (any_fast && !snappy) ? s.cap - 559u : 0u;
    if (s.active && !snappy) {
        if (s.iend == 0) { s.bad = true; s.active = false; }
        else if (s.cap == 0) { s.bad = !(s.iend == 1 && s.in[0] == 0); s.active = false; }   // lz4.c:3854-3858
    }
    if (s.active) { prefetch_l1(s.in + 128); prefetch_l1(s.in + 256); }
    uint32_t published = 0;
    while (__any_sync(kFull, s.active)) {
        if (s.active) {
            if (s.tail - slot->head < kBQCap) {          // room in the queue (else skip a turn: back-pressure)
                const bool more = snappy ? lane_step_snappy(s, q) : lane_step_lz4(s, q);
                if (!more) s.active = false;
            }
            if (s.tail - published >= 16u || !s.active) {
                __threadfence_block();
                slot->tail = s.tail;
                published = s.tail;
            }
        }
    }
    if (s.flags & kBUsed) {
        long long r = s.bad ? kErrCorrupt : (long long)s.op;
        if (!s.bad && snappy && s.op != s.cap) r = kErrCorrupt;             // snappy.cc:1715
        slot->result = r;
        __threadfence_block();
        slot->done = 1;
    }
}

// ----------------------------------------------------------------------------------------- copiers
// Executes up to 32 records of one slot.  Returns the number of records retired (0 = nothing to do).
__device__ inline uint32_t bundle_copy_group(BShared* sh, int s_idx, uint32_t& head, uint32_t& op_base, int lane) {
    BSlot* slot = &sh->slot[s_idx];
    const uint32_t fin = slot->done;
    __threadfence_block();
    const uint32_t avail = slot->tail - head;
    if (avail < 32u && !(fin && avail)) return 0;
    const uint32_t n = min(32u, avail);
    const uint4* q = sh->q + s_idx * kBQStride;
    const uint8_t* __restrict__ in = slot->in;
    uint8_t* out = slot->out;
    uint4 rec = make_uint4(0, 0, 0, 0);
    if ((uint32_t)lane < n) rec = q[(head + lane) & kBQMask];
    const uint32_t lit_pos = rec.x, ll = rec.y, off = rec.z, ml = rec.w;
    const uint32_t len = ll + ml;
    const uint32_t incl = warp_incl_sum(len, lane);
    const uint32_t dstL = op_base + incl - len;
    const uint32_t dstM = dstL + ll;
    // records are in registers: give the slots back to the parser right away
    head += n;
    op_base += __shfl_sync(kFull, incl, 31);
    __syncwarp();
    if (lane == 0) slot->head = head;

    // ---- literals: lane-per-run for short runs, whole warp for long ones
    {
        const uint32_t ll_s = ll <= 32u ? ll : 0u;
        const uint32_t maxll = __reduce_max_sync(kFull, ll_s);
        const uint8_t* src = in + lit_pos;
        uint8_t* dst = out + dstL;
        for (uint32_t base = 0; base < maxll; base += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) if (base + j < ll_s) v[j] = ld_na_u8(src + base + j);
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
            // every source byte lies in the `off` bytes before the destination (periodic pattern for
            // self-overlapping matches), so all loads of a lane are independent
            const uint8_t* src = out + (dstM - off);
            uint8_t* dst = out + dstM;
            const uint32_t my = ready ? ml : 0u;
            uint32_t k = 0;
            for (uint32_t base = 0; base < maxml; base += 8) {
                uint32_t v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (base + j < my) v[j] = ld_na_u8(src + k);
                    k = (k + 1u == off) ? 0u : k + 1u;
                }
#pragma unroll
                for (int j = 0; j < 8; j++) if (base + j < my) dst[base + j] = (uint8_t)v[j];
            }
        }
        __syncwarp();
        pending &= ~__ballot_sync(kFull, ready);
    }
    return n;
}

// A copier warp serves slots first, first + kBCopiers, ... until all of them are drained.
__device__ inline void bundle_copy(BShared* sh, int first_slot, int lane) {
    constexpr int kPer = kBSlots / kBCopiers;
    uint32_t head[kPer], op_base[kPer];
    bool live[kPer];
#pragma unroll
    for (int k = 0; k < kPer; k++) {
        head[k] = 0; op_base[k] = 0;
        live[k] = (sh->slot[first_slot + k * kBCopiers].flags & kBUsed) != 0;
    }
    for (;;) {
        bool any_live = false, progressed = false;
#pragma unroll
        for (int k = 0; k < kPer; k++) {
            if (!live[k]) continue;
            const int s_idx = first_slot + k * kBCopiers;
            const uint32_t got = bundle_copy_group(sh, s_idx, head[k], op_base[k], lane);
            if (got) progressed = true;
            else {
                BSlot* slot = &sh->slot[s_idx];
                const uint32_t fin = slot->done;
                __threadfence_block();
                if (fin && slot->tail == head[k]) live[k] = false;
            }
            any_live |= live[k];
        }
        if (!any_live) break;
        if (!progressed) __nanosleep(100);
    }
}

}  // namespace llc
