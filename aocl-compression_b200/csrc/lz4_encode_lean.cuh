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
        uint32_t v = q.lit;                                  // post rounds: already the byte this lane writes
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

// Verification of one candidate, split in two so that other work can be placed between the loads and their
// first use.  Lane j looks at bytes [4j-4, 4j) relative to the two positions: lane 0 the catch-up bytes
// (lz4.c:2098), lane 1 the four bytes the reference verifies (lz4.c:2048-2057), lanes 2..30 the next 116
// (LZ4_count, lz4.c:656-679).
struct LeanProbe { uint32_t xa, xb, pa; bool look; };
__device__ __forceinline__ LeanProbe lean_verify_issue(const LeanSrc& S, uint32_t mpos, uint32_t mcand, bool from_search,
                                                       uint32_t mlimit, uint32_t n, int lane) {
    LeanProbe v;
    v.pa = mpos + 4u * (uint32_t)lane - 4u;
    if (mpos >= 4u && mcand >= 4u) {
        // The 32 lanes read 128 contiguous bytes of each stream: one aligned word per lane; the following word
        // comes from the next lane, in lean_verify_finish -- a shuffle here would be the loads' first use and put
        // their whole latency in front of the emission that is meant to run in its shadow (20 % of all stall
        // samples sat on these two shuffles).  Lane 31 has no next lane; it does not take part.
        const uint32_t qa = S.so + mpos - 4u, qb = S.so + mcand - 4u;
        const uint32_t ia = (qa >> 2) + (uint32_t)lane, ib = (qb >> 2) + (uint32_t)lane;
        const uint32_t last = (S.so + n - 1u) >> 2;         // last aligned word that holds bytes of the unit
        v.xa = ia <= last ? S.w[ia] : 0u;
        v.xb = ib <= last ? S.w[ib] : 0u;
        v.look = lane == 0 ? from_search : (lane == 1 || (lane < 31 && v.pa < mlimit));
    } else {                                                // within 4 bytes of the start of the unit: per-lane loads
        v.look = lane == 0 ? false : (lane == 1 || (lane < 31 && v.pa < mlimit));
        v.xa = v.xb = 0;
        if (v.look) { v.xa = S.u32(v.pa); v.xb = S.u32(v.pa - (mpos - mcand)); }
    }
    return v;
}
// Returns false on a 13-bit check collision (the four bytes differ).  mc = equal bytes after the first four,
// eqb = equal bytes among the four before the two positions.
__device__ __forceinline__ bool lean_verify_finish(const LeanSrc& S, const LeanProbe& v, uint32_t mpos, uint32_t mcand,
                                                   uint32_t mlimit, uint32_t n, int lane, InGate& gate, uint32_t& mc, uint32_t& eqb) {
    uint32_t xa = v.xa, xb = v.xb;
    if (mpos >= 4u && mcand >= 4u) {                        // (uniform) aligned words -> the four bytes at this lane's offset
        const uint32_t wa1 = __shfl_down_sync(kFull, xa, 1), wb1 = __shfl_down_sync(kFull, xb, 1);
        xa = __funnelshift_r(xa, wa1, ((S.so + mpos - 4u) & 3u) * 8u);
        xb = __funnelshift_r(xb, wb1, ((S.so + mcand - 4u) & 3u) * 8u);
    }
    uint32_t c = 0;                                         // equal bytes in this lane's word
    if (v.look) {
        const uint32_t x = xa ^ xb;
        c = (uint32_t)__clz(lane == 0 ? x : __brev(x)) >> 3;             // from the top for lane 0, from the bottom otherwise
        if (lane >= 2) c = min(c, mlimit - v.pa);
    }
    const unsigned part = __ballot_sync(kFull, c < 4u);
    if (part & 2u) return false;
    eqb = __shfl_sync(kFull, c, 0);
    const unsigned fpart = part & 0x7ffffffcu;             // lanes 2..30 hold bytes +4 .. +119
    if (fpart) {
        const int first = __ffs(fpart) - 1;
        mc = 4u * (uint32_t)(first - 2) + __shfl_sync(kFull, c, first);
    } else {                                                // longer than 120: 128 bytes per extra round
        const uint32_t delta = mpos - mcand;
        mc = 116;
        uint32_t pb = mpos + 120u;
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
        bool finished = false;
        while (!finished) {
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
            const uint32_t lit = lo & 0xffu;

            // ---------------- sequences of this round ----------------
            // A post round keeps going after a match: the slots behind the match are the insert / probe / search
            // slots of the NEXT sequence, and their table entries are already here -- untouched by the slots
            // committed so far unless two slots share a bucket, in which case the candidates are worked out again
            // (below).  So one gather serves every sequence that starts inside the 32-position window (2-3 for text).
            unsigned wmask = post ? 1u : 0u;                // slots whose table write is committed (lane 0: insert of base-2)
            int s_lane = post ? 1 : 0;                      // slot of the current sequence's first probe
            bool next_post = false;
            for (;;) {
                const unsigned scope = ~((1u << s_lane) - 1u);
                const unsigned events = (mb | inv) & scope;
                const int win = events ? (__ffs(events) - 1) : 32;
                if (win >= 32) {                            // no event: the search goes on in the next round
                    wmask |= scope;
                    if (post) { fwd = base + 31u; step = 1; nb = 64u + 31u - (uint32_t)s_lane; post = false; }
                    else { fwd = __shfl_sync(kFull, cur + my_step, 31); step = (nb + 31) >> 6; nb += 32; }
                    break;
                }
                const unsigned upto_win = scope & ((2u << win) - 1u);     // slots s_lane .. win   (win <= 31)
                if (!((mb >> win) & 1u)) {                  // the search ran into the end of the block -> closing literals
                    wmask |= upto_win & ~(1u << win);
                    finished = true;
                    break;
                }
                const uint32_t mpos = __shfl_sync(kFull, cur, win);
                const uint32_t mcand = __shfl_sync(kFull, cand, win);
                const bool from_search = !(post && win == s_lane);      // the probe right after a match takes no catch-up
                const LeanProbe vp = lean_verify_issue(S, mpos, mcand, from_search, mlimit, n, lane);
                // ---- the sequence before this one is written out while the verify loads are in flight
                if (pend.valid) {
                    pend.valid = false;
                    if (!lean_emit(pend, src, dst, op, limited, cap, lane)) { refused = true; finished = true; break; }
                }
                uint32_t mc = 0, eqb = 0;
                if (!lean_verify_finish(S, vp, mpos, mcand, mlimit, n, lane, gate, mc, eqb)) { mb &= ~(1u << win); continue; }
                wmask |= upto_win;
                // ---- the match: remember it
                pend.valid = true; pend.post = post; pend.from_search = from_search;
                pend.anchor = anchor; pend.mpos = mpos; pend.mcand = mcand; pend.mc = mc; pend.eqb = eqb;
                // post rounds: lane t >= 1 of the emission needs the byte at anchor + t - 1, held by slot s_lane + t - 1
                pend.lit = post ? __shfl_sync(kFull, lit, (lane + s_lane - 1) & 31) : 0u;
                const uint32_t nbase = mpos + 4u + mc;      // first position after the match
                anchor = nbase;
                if (nbase >= mfl1) { finished = true; break; }          // lz4.c:2227
                const uint32_t nl = nbase - base + 1u;      // slot of nbase in this round's layout
                if (!post || nl > 31u) { base = nbase; next_post = true; break; }
                wmask |= 1u << (nl - 2u);                   // insert of nbase-2 (lz4.c:2230)
                s_lane = (int)nl;
                if (clashed) {
                    // Slots that share a bucket: a slot's candidate is the nearest earlier slot of its bucket that the
                    // serial algorithm EXECUTES, else the table entry.  The slots inside the match were not executed, so
                    // the slots behind it look again -- among the committed slots (wmask) and the slots from nbase on.
                    const unsigned live = (wmask | ~((1u << s_lane) - 1u)) & peers & ((1u << lane) - 1u);
                    const int f = live ? (31 - __clz(live)) : lane;
                    const uint32_t ppos = __shfl_sync(kFull, cur, f), pchk = __shfl_sync(kFull, chk, f);
                    cand = live ? ppos : (entry & kLeanPosMask);
                    cchk = live ? pchk : (entry >> 19);
                    mb = __ballot_sync(kFull, valid && probe && cchk == chk && cur - cand <= 65535u);
                }
            }

            // ---------------- commit the table writes the serial algorithm would have made ----------------
            bool wr = valid && ((wmask >> lane) & 1u);
            if (clashed && wr && lane < 31 && (peers & wmask & ~((2u << lane) - 1u))) wr = false;   // last writer per bucket
            if (wr) tab[h] = cur | (chk << 19);
            __syncwarp();
            if (next_post) post = true;
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
    // (keeps the unit's input and output addresses in register pairs: under the register cap of the partition kernels
    //  the compiler would rather rebuild them from the kernel parameters in front of an access)
    asm volatile("" : "+l"(src), "+l"(dst));
#ifndef LLC_LZ4_ENCODER_FUSED
    if (n >= 65547u && n < kLeanMaxUnit) return lz4_encode_warp_lean(src, n, dst, cap, emit_tail, tail_len, tab_mem, own, lane, gate);
#endif
    return lz4_encode_warp(src, n, dst, cap, emit_tail, tail_len, tab_mem, lane, gate);
}

}  // namespace llc
