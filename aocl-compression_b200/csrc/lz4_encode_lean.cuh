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
    // 4 bytes at p
    __device__ __forceinline__ uint32_t u32(uint32_t p) const {
        const uint32_t q = p + so, i = q >> 2, sh = (q & 3u) * 8u;
        const uint32_t w0 = w[i];
        if (sh == 0) return w0;
        return __funnelshift_r(w0, w[i + 1], sh);
    }
    // 8 bytes at p (reads three aligned words; the third always overlaps [p, p+8) or is the word after
    // a word that does, so it stays inside the unit's 4-byte-granular allocation plus one word; callers
    // only use it at positions at least 12 bytes before the end of the unit)
    __device__ __forceinline__ void u64(uint32_t p, uint32_t& lo, uint32_t& hi) const {
        const uint32_t q = p + so, i = q >> 2, sh = (q & 3u) * 8u;
        const uint32_t w0 = w[i], w1 = w[i + 1], w2 = w[i + 2];
        lo = __funnelshift_r(w0, w1, sh);
        hi = __funnelshift_r(w1, w2, sh);
    }
};

__device__ inline uint32_t lz4_encode_warp_lean(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, int64_t cap,
                                                bool emit_tail, uint32_t* tail_len, uint32_t* tab, uint8_t* own,
                                                int lane, InGate& gate) {
    const LeanSrc S(src);
    const bool limited = cap >= 0;
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
        for (;;) {
            // ---------------- one round of 32 slots in serial order ----------------
            uint32_t cur, nxt;
            bool valid, probe = true;
            if (post) {
                // lane 0: insert base-2 (lz4.c:2230); lane 1: probe of base (lz4.c:2233-2288, always
                // executed); lane L >= 2: search probe number L-1 at base+L-1, step 1 (lz4.c:1991-2001)
                cur = base + lane - (lane == 0 ? 0u : 1u) - (lane == 0 ? 2u : 0u);
                nxt = cur + 1;
                probe = lane != 0;
                valid = (lane <= 1) || (nxt <= mfl1);
            } else {
                const uint32_t my_step = (lane == 0) ? step : ((nb + lane - 1) >> 6);
                const uint32_t incl = warp_incl_sum(my_step, lane);
                cur = fwd + incl - my_step; nxt = fwd + incl;
                valid = nxt <= mfl1;                        // lz4.c:2001
            }
            gate.wait(min(n, __shfl_sync(kFull, cur, 31) + 96u));
            uint32_t lo = 0, hi = 0, h = 0, chk = 0, entry = 0;
            if (valid) {
                S.u64(cur, lo, hi);
                h = lz4_hash5((uint64_t)lo | ((uint64_t)hi << 32));
                chk = lean_check(lo);
                entry = tab[h];
                own[h] = (uint8_t)lane;
            }
            __syncwarp();
            const bool clash = valid && own[h] != (uint8_t)lane;
            uint32_t cand = entry & kLeanPosMask, cchk = entry >> 19;
            unsigned peers = 1u << lane;
            if (__any_sync(kFull, clash)) {                 // two slots of this round share a bucket
                peers = __match_any_sync(kFull, valid ? h : (0x80000000u | (uint32_t)lane));
                const unsigned before = peers & ((1u << lane) - 1u);
                const int from = before ? (31 - __clz(before)) : lane;
                const uint32_t ppos = __shfl_sync(kFull, cur, from), pchk = __shfl_sync(kFull, chk, from);
                if (before) { cand = ppos; cchk = pchk; }    // that slot would have overwritten the bucket
            }
            const bool maybe = valid && probe && cchk == chk && cur - cand <= 65535u;   // lz4.c:2048-2057
            unsigned mb = __ballot_sync(kFull, maybe);
            const unsigned inv = __ballot_sync(kFull, !valid);

            // ---------------- first surviving slot: verify, catch up, count ----------------
            int win;
            bool win_is_match = false, go_b = false;
            uint32_t mpos = 0, mcand = 0, back = 0, mc = 0;
            for (;;) {
                const unsigned events = mb | inv;
                win = events ? (__ffs(events) - 1) : 32;
                if (win >= 32 || !((mb >> win) & 1u)) break;
                mpos = __shfl_sync(kFull, cur, win);
                mcand = __shfl_sync(kFull, cand, win);
                const uint32_t delta = mpos - mcand;
                const bool from_search = !(post && win == 1);
                gate.wait(min(n, mpos + 192u));
                // lane j looks at bytes [4j-4, 4j) relative to the two positions
                uint32_t c = 0;                             // equal bytes in this lane's word
                if (lane == 0) {
                    if (from_search && mpos >= 4u && mcand >= 4u) {
                        const uint32_t x = S.u32(mpos - 4u) ^ S.u32(mcand - 4u);
                        c = x ? ((uint32_t)__clz(x) >> 3) : 4u;
                    }
                } else {
                    const uint32_t pa = mpos + 4u * (uint32_t)(lane - 1);
                    if (lane == 1 || pa < mlimit) {
                        const uint32_t x = S.u32(pa) ^ S.u32(pa - delta);
                        c = x ? ((uint32_t)(__ffs(x) - 1) >> 3) : 4u;
                        if (lane != 1) c = min(c, mlimit - pa);
                    }
                }
                const unsigned part = __ballot_sync(kFull, c < 4u);
                if (part & 2u) { mb &= ~(1u << win); continue; }   // check collision: not a match
                win_is_match = true;
                // backward, bounded by the anchor and by position 0 (lz4.c:2098)
                back = 0; go_b = false;
                if (from_search) {
                    const uint32_t bwin = (mpos >= 4u && mcand >= 4u) ? 4u : 0u;
                    const uint32_t eqb = __shfl_sync(kFull, c, 0);
                    const uint32_t roomb = min(mpos - anchor, mcand);
                    back = min(eqb, roomb);
                    go_b = (eqb == bwin) && (roomb > bwin);
                }
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
                        const uint32_t pa = pb + 4u * lane;
                        uint32_t cc = 0;
                        if (pa < mlimit) {
                            const uint32_t x = S.u32(pa) ^ S.u32(pa - delta);
                            cc = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
                            cc = min(cc, mlimit - pa);
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
            const unsigned commit = (win_is_match ? (win == 31 ? kFull : ((2u << win) - 1u))
                                                  : (win == 0 ? 0u : (win >= 32 ? kFull : ((1u << win) - 1u))));
            if (valid && ((commit >> lane) & 1u)) {
                const unsigned mine = peers & commit;
                if ((31 - __clz(mine)) == lane) tab[h] = cur | (chk << 19);
            }
            __syncwarp();
            if (win >= 32) {                                // nothing happened: next 32 probes of the same search
                if (post) { fwd = base + 31; step = 1; nb = 64 + 30; post = false; }
                else { fwd = __shfl_sync(kFull, nxt, 31); step = (nb + 31) >> 6; nb += 32; }
                continue;
            }
            if (!win_is_match) break;                       // search ran into the end of the block -> closing literals

            // ---------------- the match ----------------
            uint32_t ip = mpos - back, m = mcand - back;
            if (go_b) {                                     // rare: catch-up longer than 4 bytes
                for (;;) {
                    const bool can = (ip > anchor + lane) && (m > (uint32_t)lane);
                    const bool eq = can && (src[ip - 1 - lane] == src[m - 1 - lane]);
                    const unsigned ne = __ballot_sync(kFull, !eq);
                    const uint32_t cnt = ne ? (uint32_t)(__ffs(ne) - 1) : 32u;
                    ip -= cnt; m -= cnt; back += cnt;
                    if (cnt < 32) break;
                }
            }
            const bool from_search = !(post && win == 1);
            const uint32_t ll = ip - anchor;
            const uint32_t code = mc + back;                // match length - 4, counted from the caught-up start
            // ---- emit: token | literal-length bytes | literals | offset | match-length bytes
            const uint32_t ll_ext = ll >= 15 ? (ll - 15) / 255 + 1 : 0;
            const uint32_t ml_ext = code >= 15 ? (code - 15) / 255 + 1 : 0;
            if (limited) {
                // lz4.c:2104-2107 (literals) and lz4.c:2177-2204 (match length)
                if (from_search && (int64_t)op + 1 + ll + 8 + ll / 255 > cap) { refused = true; break; }
                if ((int64_t)op + 1 + ll_ext + ll + 2 + 6 + (code + 240) / 255 > cap) { refused = true; break; }
            }
            if ((ll < 15u) & (code < 15u)) {
                // whole sequence (token, <= 14 literals, offset) is at most 17 bytes: one byte per lane.
                // In a post round lane t >= 1 probed position anchor + t - 1, so it already holds literal t-1.
                const uint32_t offv = ip - m;
                uint32_t v = lo & 0xffu;
                if (!post && lane >= 1 && (uint32_t)lane <= ll) v = src[anchor + lane - 1];
                if (lane == 0) v = (ll << 4) | code;
                if ((uint32_t)lane == ll + 1u) v = offv & 0xffu;
                if ((uint32_t)lane == ll + 2u) v = offv >> 8;
                if ((uint32_t)lane <= ll + 2u) dst[op + lane] = (uint8_t)v;
                op += ll + 3u;
            } else {
                if (lane == 0) dst[op] = (uint8_t)((min(ll, 15u) << 4) | min(code, 15u));
                if (ll_ext) lz4_put_ext(dst + op + 1, ll - 15, lane);
                if (ll <= 32) { if ((uint32_t)lane < ll) dst[op + 1 + ll_ext + lane] = src[anchor + lane]; }
                else warp_copy(dst + op + 1 + ll_ext, src + anchor, ll, lane);
                op += 1 + ll_ext + ll;
                if (lane == 0) { dst[op] = (uint8_t)(ip - m); dst[op + 1] = (uint8_t)((ip - m) >> 8); }
                op += 2;
                if (ml_ext) lz4_put_ext(dst + op, code - 15, lane);
                op += ml_ext;
            }

            base = mpos + 4 + mc;                           // first position after the match
            anchor = base;
            if (base >= mfl1) break;                        // lz4.c:2227
            post = true;
        }
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
