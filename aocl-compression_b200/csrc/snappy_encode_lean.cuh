// snappy_encode_lean.cuh -- the Snappy fragment encoder the RAP path runs.
//
// Follows the reference's AOCL_CompressFragment (algos/snappy/snappy.cc:846-1046; emitters :436-568) bit for
// bit: one warp evaluates the next 32 probe slots of the serial schedule at once (stride = skip >> 5,
// skip += stride, starting at skip = 32: snappy.cc:903-974; the reference's unrolled 16-probe prologue is the
// same walk) and commits the table writes of the slots the serial algorithm would have executed.  Organised
// the way the LZ4 partition encoder is (lz4_encode_lean.cuh), after its ncu captures:
//   * the "insert ip-1, probe ip" that follows a copy (snappy.cc:1016-1032) is lanes 0 and 1 of the next
//     round, in front of the first 30 probes of the next search (stride 1: skip runs 32..61), so every
//     lane runs the same code;
//   * same-bucket slots inside a round are detected with a 512-byte claim bitmap in shared memory (one
//     bit per value of the low 12 hash bits: conservative; ATOMS.OR returns whether another slot of the
//     round already claimed the bit); MATCH.ANY only runs in rounds that have such a pair.  The bitmap
//     replaced a 4 KiB owner-byte table so that shared-memory-table and global-table warps fit one SM
//     together (llc_device.cu);
//   * every lane verifies its own candidate (the u16 table has no room for check bits), so after one
//     round the hit / no-hit answer of all 32 slots is known: a round without bucket clashes serves
//     every element pair that starts inside its 32-position window -- the slots behind a copy are the
//     insert / probe / search slots of the next search and their table entries cannot have been touched
//     by the slots committed so far;
//   * the match length is counted by the whole warp from coalesced loads (one aligned word per lane,
//     the next word from the next lane), 124 bytes per round trip;
//   * emission of a literal + copy pair is deferred until the loads of the next step have been issued.
#pragma once
#include "snappy_codec.cuh"
#include "lz4_encode_lean.cuh"

namespace llc {

constexpr uint32_t kSnappyClaimBytes = 512;             // one claim bit per value of the low 12 hash bits

struct SnappyPending { bool valid; uint32_t lit_from, mpos, off, len; };

__device__ __forceinline__ void snappy_lean_emit(const SnappyPending& q, const uint8_t* __restrict__ src, uint8_t* dst,
                                                 uint32_t& op, int lane) {
    if (q.mpos > q.lit_from) op = snappy_put_literal(dst, op, src + q.lit_from, q.mpos - q.lit_from, lane);   // snappy.cc:980
    op = snappy_put_copy(dst, op, q.off, q.len, lane);                                                         // snappy.cc:1004
}

// bytes equal from p / p - delta on, bounded by n (FindMatchLength, snappy-internal.h:228-352): the 32 lanes read 128
// contiguous bytes of each stream, one aligned word per lane (the next word comes from the next lane), 124 bytes per
// round trip.  Split in two so that the pending emission runs between the first loads and their first use.
struct SnappyCount { uint32_t wa, wb; };
__device__ __forceinline__ SnappyCount snappy_lean_count_issue(const LeanSrc& S, uint32_t p, uint32_t delta, uint32_t n, int lane) {
    const uint32_t last = (S.so + n - 1u) >> 2;             // last aligned word that holds bytes of the fragment
    const uint32_t qa = S.so + p, qb = qa - delta;
    const uint32_t ia = (qa >> 2) + (uint32_t)lane, ib = (qb >> 2) + (uint32_t)lane;
    SnappyCount c;
    c.wa = ia <= last ? S.w[ia] : 0u;
    c.wb = ib <= last ? S.w[ib] : 0u;
    return c;
}
__device__ __forceinline__ uint32_t snappy_lean_count_finish(const LeanSrc& S, SnappyCount ld, uint32_t p, uint32_t delta, uint32_t n,
                                                             int lane) {
    uint32_t total = 0;
    for (;;) {
        const uint32_t qa = S.so + p, qb = qa - delta;
        const uint32_t wa1 = __shfl_down_sync(kFull, ld.wa, 1), wb1 = __shfl_down_sync(kFull, ld.wb, 1);
        const uint32_t x = __funnelshift_r(ld.wa, wa1, (qa & 3u) * 8u) ^ __funnelshift_r(ld.wb, wb1, (qb & 3u) * 8u);
        const uint32_t pa = p + 4u * (uint32_t)lane;
        uint32_t c = 0;                                     // lane 31 has no next word: it only ends the round
        if (lane < 31 && pa < n) c = min((uint32_t)__clz(__brev(x)) >> 3, n - pa);
        const unsigned part = __ballot_sync(kFull, c < 4u);
        const int first = __ffs(part) - 1;                  // part != 0: lane 31 always reports
        total += 4u * (uint32_t)first + __shfl_sync(kFull, c, first);
        if (first < 31) return total;
        p += 124u;
        ld = snappy_lean_count_issue(S, p, delta, n, lane);
    }
}

__device__ inline uint32_t snappy_encode_fragment_lean(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst,
                                                       uint16_t* tab, uint32_t* claim, int lane) {
    // (keeps the fragment's input and output addresses in register pairs: under the register cap of the fragment
    //  kernels the compiler would rather rebuild them from the kernel parameters in front of an access)
    asm volatile("" : "+l"(src), "+l"(dst));
    const LeanSrc S(src);
    uint32_t tsize = 256;                                   // snappy.cc:619-632
    if (n > 16384) tsize = 16384; else while (tsize < n) tsize <<= 1;
    const int shift = 32 - (31 - __clz(tsize));
    for (uint32_t i = lane; i < tsize / 2; i += 32) reinterpret_cast<uint32_t*>(tab)[i] = 0;
    for (uint32_t i = lane; i < kSnappyClaimBytes / 4; i += 32) claim[i] = 0;     // all zero between rounds
    __syncwarp();
    uint32_t op = 0, anchor = 0;                            // anchor = next_emit: first byte not yet emitted
    if (n >= 15) {
        const uint32_t ip_limit = n - 15;
        bool post = false, finished = false;
        uint32_t base = 0;                                  // post rounds: first position after the copy
        uint32_t ip = 1, skip = 32;                         // search rounds: next probe position and skip counter (snappy.cc:903-974)
        SnappyPending pend;
        pend.valid = false;
        while (!finished) {
            // ---------------- one round of 32 slots in serial order ----------------
            uint32_t cur, stride = 1, s = skip;
            bool valid, probe = true;
            if (post) {
                // lane 0: insert base-1; lane 1: probe of base (always executed, snappy.cc:1016-1032);
                // lane L >= 2: search probe L-2 at base+L-1, stride 1 (skip = 32 + L - 2 < 64)
                cur = base + (uint32_t)lane - 1u;
                probe = lane != 0;
                valid = (lane <= 1) || (cur + 1u <= ip_limit);
            } else {
                if (skip <= 32) s = skip + lane;            // first round of a search: stride 1 throughout
                else {
                    // lanes need skip after k probes; each lane replays the recurrence (at most 31 cheap steps;
                    // only reached after 32 fruitless probes)
                    for (int k = 0; k < lane; k++) s += s >> 5;
                    stride = s >> 5;
                }
                const uint32_t incl = warp_incl_sum(stride, lane);
                cur = ip + incl - stride;
                valid = ip + incl <= ip_limit;
            }
            uint32_t seq4 = 0, h = 0, cand = 0;
            bool seen = false;
            if (valid) {
                seq4 = S.u32(cur);
                h = (seq4 * 0x1e35a7bdU) >> shift;          // snappy.cc:152-158
                cand = tab[h];
                const uint32_t bit = 1u << (h & 31u);
                seen = (atomicOr(&claim[(h >> 5) & (kSnappyClaimBytes / 4 - 1u)], bit) & bit) != 0;
            }
            // ---- the previous literal + copy is written out while the table gather is in flight
            if (pend.valid) { pend.valid = false; snappy_lean_emit(pend, src, dst, op, lane); }
            const bool clashed = __any_sync(kFull, seen);
            __syncwarp();
            if (valid) claim[(h >> 5) & (kSnappyClaimBytes / 4 - 1u)] = 0;   // every claimant clears its word
            unsigned peers = 0;
            if (clashed) {                                  // two slots of this round (may) share a bucket
                peers = __match_any_sync(kFull, valid ? h : (0x80000000u | (uint32_t)lane));
                const unsigned before = peers & ((1u << lane) - 1u);
                const int f = before ? (31 - __clz(before)) : lane;
                const uint32_t ppos = __shfl_sync(kFull, cur, f);
                if (before) cand = ppos;                    // that slot would have overwritten the bucket
            }
            const bool hit = valid && probe && S.u32(cand) == seq4;
            const unsigned hits = __ballot_sync(kFull, hit);
            const unsigned inv = __ballot_sync(kFull, !valid);

            // ---------------- elements of this round ----------------
            unsigned wmask = post ? 1u : 0u;                // slots whose table write is committed (lane 0: insert of base-1)
            int s_lane = post ? 1 : 0;                      // slot of the current search's first probe
            bool next_post = false;
            for (;;) {
                const unsigned scope = ~((1u << s_lane) - 1u);
                const unsigned events = (hits | inv) & scope;
                const int win = events ? (__ffs(events) - 1) : 32;
                if (win >= 32) {                            // no event: the search goes on in the next round
                    wmask |= scope;
                    if (post) { ip = base + 31u; skip = 32u + 31u - (uint32_t)s_lane; post = false; }
                    else { ip = __shfl_sync(kFull, cur + stride, 31); skip = __shfl_sync(kFull, s + (s >> 5), 31); }
                    break;
                }
                const unsigned upto_win = scope & ((2u << win) - 1u);     // slots s_lane .. win
                if (!((hits >> win) & 1u)) {                // the search ran into the end of the fragment -> emit_remainder
                    wmask |= upto_win & ~(1u << win);
                    finished = true;
                    break;
                }
                wmask |= upto_win;
                const uint32_t mpos = __shfl_sync(kFull, cur, win);
                const uint32_t mcand = __shfl_sync(kFull, cand, win);
                const SnappyCount ld = snappy_lean_count_issue(S, mpos + 4u, mpos - mcand, n, lane);
                if (pend.valid) snappy_lean_emit(pend, src, dst, op, lane);          // ... in the shadow of those loads
                const uint32_t len = 4u + snappy_lean_count_finish(S, ld, mpos + 4u, mpos - mcand, n, lane);   // snappy.cc:995-1003
                pend.valid = true; pend.lit_from = anchor; pend.mpos = mpos; pend.off = mpos - mcand; pend.len = len;
                const uint32_t nbase = mpos + len;
                anchor = nbase;
                if (nbase >= ip_limit) { finished = true; break; }          // snappy.cc:1009
                const uint32_t nl = nbase - base + 1u;      // slot of nbase in this round's layout
                if (!post || clashed || nl > 31u) { base = nbase; next_post = true; break; }
                wmask |= 1u << (nl - 1u);                   // insert of nbase-1 (snappy.cc:1016-1021)
                s_lane = (int)nl;
            }

            // ---------------- commit the table writes the serial algorithm would have made ----------------
            bool wr = valid && ((wmask >> lane) & 1u);
            if (clashed && wr && lane < 31 && (peers & wmask & ~((2u << lane) - 1u))) wr = false;   // last writer per bucket
            if (wr) tab[h] = (uint16_t)cur;
            __syncwarp();
            if (next_post) post = true;
        }
        if (pend.valid) snappy_lean_emit(pend, src, dst, op, lane);
    }
    if (anchor < n) op = snappy_put_literal(dst, op, src + anchor, n - anchor, lane);   // snappy.cc:1039-1043
    return op;
}

}  // namespace llc
