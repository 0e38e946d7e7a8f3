// lz4_encode_lean.cuh -- the LZ4 partition encoder the RAP path runs (hash5 / byU32 tables, units
// below 512 KiB, which covers every RAP partition: part < 1.5 * 262,268 + T).
//
// Same exact greedy parse as lz4_encode_warp_fused (lz4_codec.cuh), which follows the reference's
// AOCL_LZ4_compress_generic_validated_mt (algos/lz4/lz4.c:1853-2350, acceleration 1, noDict) bit for
// bit: one warp evaluates the next 32 probe slots of the serial schedule at once and commits the
// table writes of the slots the serial algorithm would have executed.  What changed is what a round
// costs.  An ncu capture of the fused encoder on the 1 GiB text workload showed 56 GB of DRAM reads
// (52x the input), 28 % of all stall samples on the 32 per-lane candidate-window loads, 11 % on
// MATCH.ANY and 260 warp instructions per sequence.  Here:
//
//   * a table entry is pos (19 bits) | check (13 bits), check = 13 bits of a hash of the 4 bytes at pos.
//     A slot whose check differs from the probe's cannot verify (the reference compares exactly those
//     4 bytes, lz4.c:2048-2057), so it is rejected without touching memory.  The table starts out as
//     "position 0 with position 0's check", which is what the reference's zeroed table means
//     (slot value 0 is a real candidate, lz4.c:3047-3055);
//   * only the FIRST surviving slot in serial order is verified, by the whole warp at once: lane 0
//     compares the 4 bytes before the two positions (catch-up, lz4.c:2098), lane 1 the 4 bytes the
//     reference verifies, lanes 2..31 the next 120 bytes (LZ4_count, lz4.c:656-679).  One round trip
//     yields verification, catch-up and match length; a failed verification (a 13-bit check
//     collision) just moves on to the next surviving slot;
//   * same-hash slots inside one round are detected with a 4 KiB shared-memory owner table (one byte
//     store and one byte load per lane); MATCH.ANY only runs in the rounds that have such a pair;
//   * the "insert ip-2" that follows a match (lz4.c:2230) is lane 0 of the next round, an insert-only
//     slot in front of the probe slots, so every lane runs the same code;
//   * positions are 32-bit offsets from a word-aligned base pointer.
#pragma once
#include "lz4_codec.cuh"

namespace llc {

constexpr uint32_t kLeanMaxUnit = 1u << 19;       // positions must fit 19 bits
constexpr uint32_t kLeanPosMask = kLeanMaxUnit - 1u;
constexpr uint32_t kLeanOwnBytes = 4096;          // one owner byte per hash bucket

__device__ __forceinline__ uint32_t lean_check(uint32_t seq4) { return (seq4 * 2654435761U) >> 19; }   // 13 bits

// Byte position p of the unit -> unaligned little-endian words, through a 4-byte aligned base.
struct LeanSrc {
    const uint32_t* w;     // src rounded down to a word
    uint32_t so;           // src & 3
    __device__ __forceinline__ LeanSrc(const uint8_t* src) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
        so = (uint32_t)(a & 3);
    }
    // 4 bytes at p (only words that overlap [p, p+4) are touched)
    __device__ __forceinline__ uint32_t u32(uint32_t p) const {
        const uint32_t q = p + so, i = q >> 2, sh = (q & 3u) * 8u;
        const uint32_t w0 = w[i];
        if (sh == 0) return w0;
        return __funnelshift_r(w0, w[i + 1], sh);
    }
    // 5 bytes at p: lo = bytes 0..3, b4 = byte 4.  They always lie inside two aligned words.
    __device__ __forceinline__ void u40(uint32_t p, uint32_t& lo, uint32_t& b4) const {
        const uint32_t q = p + so, i = q >> 2, sh = (q & 3u) * 8u;
        const uint32_t w0 = w[i], w1 = w[i + 1];
        lo = __funnelshift_r(w0, w1, sh);
        b4 = (w1 >> sh) & 0xffu;
    }
};

// What one finished sequence needs to be written out.  Emission is deferred: the next round first issues
// its table gather and only then writes the previous sequence, so token / literal / offset emission runs
// in the shadow of that gather instead of in front of it (the gather wait went from 19 % to 4 % of the
// stall samples).
struct LeanPending {
    bool valid, post, from_search;
    uint32_t anchor, mpos, mcand, mc, eqb, lit;
};

// Writes the pending sequence at dst + op (lz4.c:2098-2219) and returns false if the limitedOutput checks
// refuse it (lz4.c:2104-2107, 2177-2204).
__device__ __forceinline__ bool lean_emit(const LeanPending& q, const uint8_t* __restrict__ src, uint8_t* dst, uint32_t& op,
                                          bool limited, int64_t cap, int lane) {
    // backward catch-up, bounded by the anchor and by position 0 (lz4.c:2098); eqb = equal bytes among the 4 before
    uint32_t back = 0;
    uint32_t ip = q.mpos, m = q.mcand;
    if (q.from_search) {
        const uint32_t bwin = (q.mpos >= 4u && q.mcand >= 4u) ? 4u : 0u;
        const uint32_t roomb = min(q.mpos - q.anchor, q.mcand);
        back = min(q.eqb, roomb);
        ip -= back; m -= back;
        if ((q.eqb == bwin) && (roomb > bwin)) {            // rare: catch-up longer than the 4 bytes already compared
            for (;;) {
                const bool can = (ip > q.anchor + lane) && (m > (uint32_t)lane);
                const bool eq = can && (src[ip - 1 - lane] == src[m - 1 - lane]);
                const unsigned ne = __ballot_sync(kFull, !eq);
                const uint32_t cnt = ne ? (uint32_t)(__ffs(ne) - 1) : 32u;
                ip -= cnt; m -= cnt; back += cnt;
                if (cnt < 32) break;
            }
        }
    }
    const uint32_t ll = ip - q.anchor;
    const uint32_t code = q.mc + back;                      // match length - 4, counted from the caught-up start
    if (limited) {
        const uint32_t ll_ext = ll >= 15 ? (ll - 15) / 255 + 1 : 0;
        if (q.from_search && (int64_t)op + 1 + ll + 8 + ll / 255 > cap) return false;
        if ((int64_t)op + 1 + ll_ext + ll + 2 + 6 + (code + 240) / 255 > cap) return false;
    }
    if ((ll < 15u) & (code < 15u)) {
        // whole sequence (token, <= 14 literals, offset) is at most 17 bytes: one byte per lane.
        // In a post round lane t >= 1 probed position anchor + t - 1, so it already held literal t-1.
        const uint32_t offv = ip - m;
        uint32_t v = q.lit;
        if (!q.post && (uint32_t)(lane - 1) < ll) v = src[q.anchor + lane - 1];
        if (lane == 0) v = (ll << 4) | code;
        if ((uint32_t)lane == ll + 1u) v = offv;
        if ((uint32_t)lane == ll + 2u) v = offv >> 8;
        if ((uint32_t)lane <= ll + 2u) dst[op + lane] = (uint8_t)v;
        op += ll + 3u;
    } else {
        // ---- token | literal-length bytes | literals | offset | match-length bytes
        const uint32_t ll_ext = ll >= 15 ? (ll - 15) / 255 + 1 : 0;
        const uint32_t ml_ext = code >= 15 ? (code - 15) / 255 + 1 : 0;
        if (lane == 0) dst[op] = (uint8_t)((min(ll, 15u) << 4) | min(code, 15u));
        if (ll_ext) lz4_put_ext(dst + op + 1, ll - 15, lane);
        if (ll <= 32) { if ((uint32_t)lane < ll) dst[op + 1 + ll_ext + lane] = src[q.anchor + lane]; }
        else warp_copy(dst + op + 1 + ll_ext, src + q.anchor, ll, lane);
        op += 1 + ll_ext + ll;
        if (lane == 0) { dst[op] = (uint8_t)(ip - m); dst[op + 1] = (uint8_t)((ip - m) >> 8); }
        op += 2;
        if (ml_ext) lz4_put_ext(dst + op, code - 15, lane);
        op += ml_ext;
    }
    return true;
}

__device__ inline uint32_t lz4_encode_warp_lean(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, int64_t cap,
                                                bool emit_tail, uint32_t* tail_len, uint32_t* tab, uint8_t* own,
                                                int lane, InGate& gate) {
    const LeanSrc S(src);
    const bool limited = cap >= 0;
    const bool gated = gate.flag != nullptr;
    uint32_t op = 0, anchor = 0;
    bool refused = false;

    if (n >= 13) {                                          // LZ4_minLength, lz4.c:1926
        gate.wait(min(n, 128u));
        // zeroed table == every slot holds position 0; position 0 itself is inserted first (lz4.c:1929)
        const uint32_t init = lean_check(S.u32(0)) << 19;
        for (int i = lane; i < 4096; i += 32) tab[i] = init;
        __syncwarp();
        const uint32_t mfl1 = n - 11;                       // mflimitPlusOne, lz4.c:1887
        const uint32_t mlimit = n - 5;                      // matchlimit, lz4.c:1888
        bool post = false;                                  // lane 0 = insert of base-2, lane 1 = probe of base
        uint32_t base = 0;                                  // post rounds: first position after the match
        uint32_t fwd = 1, step = 1, nb = 64;                // search schedule (lz4.c:1991-1997)
        LeanPending pend;
        pend.valid = false;
        for (;;) {
            // ---------------- one round of 32 slots in serial order ----------------
            uint32_t cur, my_step = 1, bound;
            bool valid, probe = true;
            if (post) {
                // lane 0: insert base-2 (lz4.c:2230); lane 1: probe of base (lz4.c:2233-2288, always
                // executed); lane L >= 2: search probe number L-1 at base+L-1, step 1 (lz4.c:1991-2001)
                cur = base + (uint32_t)lane - (lane == 0 ? 2u : 1u);
                probe = lane != 0;
                valid = (lane <= 1) || (cur < mfl1);        // next position <= mflimitPlusOne
                bound = base + 31u;
            } else {
                my_step = (lane == 0) ? step : ((nb + lane - 1) >> 6);
                const uint32_t incl = warp_incl_sum(my_step, lane);
                cur = fwd + incl - my_step;
                valid = fwd + incl <= mfl1;                 // lz4.c:2001
                bound = fwd + 32u * ((nb + 31u) >> 6);
            }
            // everything this round reads (probe windows, the 128-byte verify span) lies below bound + 140
            if (gated) gate.wait(min(n, bound + 256u));
            uint32_t lo = 0, b4 = 0, h = 0, chk = 0, entry = 0;
            if (valid) {
                S.u40(cur, lo, b4);
                h = lz4_hash5((uint64_t)lo | ((uint64_t)b4 << 32));
                chk = lean_check(lo);
                entry = tab[h];
                own[h] = (uint8_t)lane;
            }
            // ---- the previous sequence is written out while the gather is in flight
            if (pend.valid) {
                pend.valid = false;
                if (!lean_emit(pend, src, dst, op, limited, cap, lane)) { refused = true; break; }
            }
            __syncwarp();
            const bool clashed = __any_sync(kFull, valid && own[h] != (uint8_t)lane);
            uint32_t cand = entry & kLeanPosMask, cchk = entry >> 19;
            unsigned peers = 0;
            if (clashed) {                                  // two slots of this round share a bucket
                peers = __match_any_sync(kFull, valid ? h : (0x80000000u | (uint32_t)lane));
                const unsigned before = peers & ((1u << lane) - 1u);
                const int f = before ? (31 - __clz(before)) : lane;
                const uint32_t ppos = __shfl_sync(kFull, cur, f), pchk = __shfl_sync(kFull, chk, f);
                if (before) { cand = ppos; cchk = pchk; }   // that slot would have overwritten the bucket
            }
            const bool maybe = valid && probe && cchk == chk && cur - cand <= 65535u;   // lz4.c:2048-2057
            unsigned mb = __ballot_sync(kFull, maybe);
            const unsigned inv = __ballot_sync(kFull, !valid);

            // ---------------- first surviving slot: verify, count ----------------
            int win;
            bool win_is_match = false;
            uint32_t mpos = 0, mcand = 0, mc = 0, eqb = 0;
            for (;;) {
                const unsigned events = mb | inv;
                win = events ? (__ffs(events) - 1) : 32;
                if (!((mb >> (win & 31)) & 1u) || win >= 32) break;
                mpos = __shfl_sync(kFull, cur, win);
                mcand = __shfl_sync(kFull, cand, win);
                const uint32_t delta = mpos - mcand;
                // lane j looks at bytes [4j-4, 4j) relative to the two positions: lane 0 the catch-up bytes
                // (lz4.c:2098), lane 1 the four bytes the reference verifies, lanes 2..31 the next 120 (LZ4_count)
                const uint32_t pa = mpos + 4u * (uint32_t)lane - 4u;
                const bool look = lane == 0 ? (!(post && win == 1) && mpos >= 4u && mcand >= 4u) : (lane == 1 || pa < mlimit);
                uint32_t c = 0;                             // equal bytes in this lane's word
                if (look) {
                    const uint32_t x = S.u32(pa) ^ S.u32(pa - delta);
                    c = (uint32_t)__clz(lane == 0 ? x : __brev(x)) >> 3;     // from the top for lane 0, from the bottom otherwise
                    if (lane >= 2) c = min(c, mlimit - pa);
                }
                const unsigned part = __ballot_sync(kFull, c < 4u);
                if (part & 2u) { mb &= ~(1u << win); continue; }   // 13-bit check collision: not a match
                win_is_match = true;
                eqb = __shfl_sync(kFull, c, 0);
                // forward: lanes 2..31 hold bytes +4 .. +123
                const unsigned fpart = part & ~3u;
                if (fpart) {
                    const int first = __ffs(fpart) - 1;
                    mc = 4u * (uint32_t)(first - 2) + __shfl_sync(kFull, c, first);
                } else {                                    // longer than 124: 128 bytes per extra round
                    mc = 120;
                    uint32_t pb = mpos + 124u;
                    for (;;) {
                        gate.wait(min(n, pb + 208u));
                        const uint32_t pc = pb + 4u * lane;
                        uint32_t cc = 0;
                        if (pc < mlimit) {
                            const uint32_t x = S.u32(pc) ^ S.u32(pc - delta);
                            cc = min((uint32_t)__clz(__brev(x)) >> 3, mlimit - pc);
                        }
                        const unsigned partial = __ballot_sync(kFull, cc < 4);
                        if (partial) {
                            const int first = __ffs(partial) - 1;
                            mc += 4u * first + __shfl_sync(kFull, cc, first);
                            break;
                        }
                        mc += 128; pb += 128;
                    }
                }
                break;
            }

            // ---------------- commit the table writes the serial algorithm would have made ----------------
            // slots up to the winning match, or up to (not including) the slot that ran into the end of the block
            const int limit = win_is_match ? win : win - 1;
            bool wr = valid && lane <= limit;
            if (clashed && wr) {                            // last writer per bucket
                const unsigned upto = limit >= 31 ? kFull : ((2u << limit) - 1u);
                const unsigned later = lane >= 31 ? 0u : (peers & upto & ~((2u << lane) - 1u));
                if (later) wr = false;
            }
            if (wr) tab[h] = cur | (chk << 19);
            __syncwarp();
            if (win >= 32) {                                // nothing happened: next 32 probes of the same search
                if (post) { fwd = base + 31; step = 1; nb = 64 + 30; post = false; }
                else { fwd = __shfl_sync(kFull, cur + my_step, 31); step = (nb + 31) >> 6; nb += 32; }
                continue;
            }
            if (!win_is_match) break;                       // search ran into the end of the block -> closing literals

            // ---------------- the match: remember it, move on ----------------
            pend.valid = true; pend.post = post; pend.from_search = !(post && win == 1);
            pend.anchor = anchor; pend.mpos = mpos; pend.mcand = mcand; pend.mc = mc; pend.eqb = eqb;
            pend.lit = lo & 0xffu;
            base = mpos + 4 + mc;                           // first position after the match
            anchor = base;
            if (base >= mfl1) break;                        // lz4.c:2227
            post = true;
        }
        if (pend.valid && !refused && !lean_emit(pend, src, dst, op, limited, cap, lane)) refused = true;
    }
    if (refused) return 0;
    gate.wait(n);                                           // the closing literals are read by this warp or by the stitch
    const uint32_t run = n - anchor;
    if (!emit_tail) { if (tail_len) *tail_len = run; return op; }        // lz4.c:2333-2338
    if (tail_len) *tail_len = 0;
    if (limited && (int64_t)op + run + 1 + (run + 255 - 15) / 255 > cap) return 0;   // lz4.c:2299-2311
    const uint32_t ext = run >= 15 ? (run - 15) / 255 + 1 : 0;
    if (lane == 0) dst[op] = (uint8_t)(min(run, 15u) << 4);
    if (ext) lz4_put_ext(dst + op + 1, run - 15, lane);
    warp_copy(dst + op + 1 + ext, src + anchor, run, lane);
    return op + 1 + ext + run;
}

// Dispatch: every unit that fits 19-bit positions and uses the byU32 / hash5 table takes the lean
// encoder; small units (byU16 / hash4, n < 65547) and oversized frame-less blocks keep the fused one.
__device__ inline uint32_t lz4_encode_unit(const uint8_t* src, uint32_t n, uint8_t* dst, int64_t cap, bool emit_tail,
                                           uint32_t* tail_len, uint32_t* tab_mem, uint8_t* own, int lane, InGate& gate) {
#ifndef LLC_LZ4_ENCODER_FUSED
    if (n >= 65547u && n < kLeanMaxUnit) return lz4_encode_warp_lean(src, n, dst, cap, emit_tail, tail_len, tab_mem, own, lane, gate);
#endif
    return lz4_encode_warp(src, n, dst, cap, emit_tail, tail_len, tab_mem, lane, gate);
}

}  // namespace llc
