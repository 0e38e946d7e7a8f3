// lz4_fastparse.cuh -- the separately named `fastparse` compress mode (never the default).
//
// Same RAP frame, same partitions, same stitch as the exact mode -- only the partition encoder differs, so the
// streams decode with any LZ4 decoder (the reference's included) but are NOT byte-identical to the reference's.
// The exact encoder reproduces LZ4_compress_fast's serial greedy parse (lz4.c:1853-2350): ~270 warp instructions per
// sequence, because everything about a sequence (probe schedule, table commits, catch-up, emission) is decided one
// sequence at a time.  Here the parse is POSITION PARALLEL, 64 positions per round:
//   * every lane hashes its own positions and reads the table (the state BEFORE this round), then all lanes insert
//     with atomicMax -- the table ends up holding the highest position per bucket whatever the order of the lanes,
//     so the output is deterministic; repeats closer than a round are found with MATCH.ANY on the four bytes;
//   * the round's bytes come from one coalesced load of 20 aligned words, cut to size with shuffles; every lane
//     fetches its candidate's twelve bytes in one go, verifies it (4 equal bytes, distance <= 65535) and knows its
//     match length up to 12; longer matches are extended by the whole warp, 128 bytes per step, only if selected;
//   * the greedy selection (first match at or after the end of the previous one) is worked out for all lanes at
//     once by pointer jumping over "the match that follows mine"; literals, token and offset of the selected
//     sequences are written lane parallel.
// Deviations from the reference's parse that cost ratio (reported by bench.py as `ratio_delta_vs_exact`): matches
// against positions of the same round only inside one 32-position half, no backward catch-up
// (lz4.c:2098), every position of a round is inserted (the reference skips the inside of matches and accelerates
// through incompressible data: lz4.c:1991-1997).  Precedent for trading ratio for speed in the reference itself:
// AOCL_LZ4_MATCH_SKIP_OPT_LDS_STRAT1/2 (lz4.c:1447-1450, 1572-1584), AOCL_SNAPPY_MATCH_SKIP_OPT (snappy.cc:939-969).
#pragma once
#include "lz4_encode_lean.cuh"

namespace llc {

constexpr uint32_t kFpTabLog = 12;                           // 4096 x u32, the size of the reference's table
constexpr uint32_t kFpLaneMax = 12;                          // match bytes a lane extends on its own; longer ones by the warp
constexpr int kFpHalves = 2;                                 // a round = kFpHalves x 32 positions

// Lane l < 20 loads aligned word (byte_off >> 2) + l of the input (0 where the word starts behind the unit's end).
__device__ __forceinline__ uint32_t fp_window(const uint32_t* words, uint32_t byte_off, uint32_t end_off, int lane) {
    const uint32_t wi = (byte_off >> 2) + (uint32_t)lane;
    return (lane < 20 && wi * 4u < end_off) ? words[wi] : 0u;
}

// Encodes src[0, n) as LZ4 sequences at dst.  Same contract as lz4_encode_unit (unlimited output): returns the
// body length; a non-final unit leaves its trailing literals to the stitch (*tail_len), the final one writes them.
// `tab`: 4096 words of shared or global memory owned by this warp; an entry is (position + 1) << 13 | 13 check bits
// of the four bytes at that position (0: empty) -- atomicMax keeps the highest position per bucket, and a candidate
// whose check bits differ is dropped without fetching it.
//
// A round covers kFpHalves x 32 consecutive positions, lane l taking positions t + l, t + 32 + l: the loads of the
// halves are independent, so their latencies (table word from L2, candidate bytes from L2 / HBM, the words of the
// match extension) overlap.  All halves read the table as it was BEFORE the round and insert afterwards; a repeat
// closer than that is found inside a half with MATCH.ANY on the four bytes (the nearest earlier lane with the same
// bytes), which is what catches runs and the short periods of columnar data.
// Per half: the greedy selection by pointer jumping (every lane learns whether it is a match start, a literal of
// which sequence, or covered; the warp only loops over matches it has to extend); then the
// sequences of the half are written LANE PARALLEL -- a warp scan of their sizes gives every sequence its place, the
// start lanes write token / length bytes / offset, every literal lane writes its own byte.
__device__ inline uint32_t lz4_fastparse_unit(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, bool emit_tail,
                                              uint32_t* tail_len, uint32_t* tab, int lane, InGate& gate) {
    // (keeps the unit's input and output addresses in register pairs: under the register cap the compiler would
    //  rather rebuild them from the kernel parameters in front of every access, five instructions each)
    asm volatile("" : "+l"(src), "+l"(dst));
    const LeanSrc S(src);
    // the input as aligned words: src[i] is byte i + mis of words[]
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    const uint32_t* const words = reinterpret_cast<const uint32_t*>(src - mis);
    uint32_t op = 0, anchor = 0;
    if (n >= 13u) {                                          // lz4.c:1926: shorter inputs are all literals
        for (uint32_t i = lane; i < (1u << kFpTabLog); i += 32) tab[i] = 0;
        __syncwarp();
        const uint32_t last_start = n - 12u;                 // a match starts at least 12 bytes before the end (MFLIMIT)
        const uint32_t mlimit = n - 5u;                      // ... and ends at least 5 bytes before it (LASTLITERALS)
        const uint32_t lower = (1u << lane) - 1u;            // lanes below mine
        uint32_t t = 0, W = 0, w_at = 0xffffffffu;            // W: the window of words at position w_at
        while (t <= last_start) {
            gate.wait(min(n, t + 32u * kFpHalves + 64u + 136u));
            // ONE coalesced load per round: lane l holds the aligned word l of the window (20 words cover the 64
            // positions and the kFpLaneMax bytes behind the last one, whatever the alignment); every position's
            // twelve bytes (its four, and the two words a match is extended over) are cut out of it with shuffles.
            const uint32_t bo = t + mis;
            if (t != w_at) W = fp_window(words, bo, n + mis, lane);
            // (the next round's window is fetched now: the round after this one starts 64 positions on unless a match
            //  carries it further -- which, on text, happens in two rounds out of three: a fixed stride that made the
            //  bet safe, and the next round's table words fetched ahead as well, were both measured and both lost,
            //  profiles/r2_fastparse.txt)
            w_at = t + 32u * kFpHalves;
            const uint32_t Wn = fp_window(words, w_at + mis, n + mis, lane);
            uint32_t v[kFpHalves], v4[kFpHalves], v8[kFpHalves], h[kFpHalves], e[kFpHalves], c[kFpHalves], ml[kFpHalves];
            bool inb[kFpHalves], ok[kFpHalves], cap[kFpHalves], cand[kFpHalves];
#pragma unroll
            for (int k = 0; k < kFpHalves; k++) {
                const uint32_t p = t + 32u * k + (uint32_t)lane;
                const uint32_t o = (bo & 3u) + 32u * k + (uint32_t)lane, wi = o >> 2, sh = (o & 3u) * 8u;
                const uint32_t a0 = __shfl_sync(kFull, W, (int)wi), a1 = __shfl_sync(kFull, W, (int)wi + 1);
                const uint32_t a2 = __shfl_sync(kFull, W, (int)wi + 2), a3 = __shfl_sync(kFull, W, (int)wi + 3);
                v[k] = __funnelshift_r(a0, a1, sh); v4[k] = __funnelshift_r(a1, a2, sh); v8[k] = __funnelshift_r(a2, a3, sh);
                inb[k] = p <= last_start;
                h[k] = 0; e[k] = 0;
                if (inb[k]) { h[k] = (v[k] * 2654435761U) >> (32u - kFpTabLog); e[k] = tab[h[k]]; }
            }
            __syncwarp();                                    // every lane has read the old state
#pragma unroll
            for (int k = 0; k < kFpHalves; k++) {
                const uint32_t p = t + 32u * k + (uint32_t)lane;
                const uint32_t chk = (v[k] * 0x9E3779B1u) >> 19;           // 13 bits, independent of the bucket bits
                if (inb[k]) atomicMax(&tab[h[k]], ((p + 1u) << 13) | chk);
                // the nearest earlier lane of this half with the same four bytes, else the table's candidate
                const unsigned same = __match_any_sync(kFull, inb[k] ? v[k] : (0x80000000u | (uint32_t)lane) ^ v[k]) & lower;
                c[k] = (e[k] >> 13) - 1u;                    // < t: entries come from earlier rounds
                cand[k] = inb[k] && e[k] != 0u && (e[k] & 0x1fffu) == chk && (p - c[k]) <= 65535u;
                if (inb[k] && same) { c[k] = t + 32u * k + (31u - (uint32_t)__clz(same)); cand[k] = true; }
            }
            // the candidates' twelve bytes, all halves in flight together (c + 12 < p + 12 <= n: the four words start
            // inside the unit)
            uint32_t b[kFpHalves][4];
#pragma unroll
            for (int k = 0; k < kFpHalves; k++) {
                const uint32_t ci = cand[k] ? (c[k] + mis) >> 2 : 0u;
#pragma unroll
                for (int j = 0; j < 4; j++) b[k][j] = words[ci + j];
            }
#pragma unroll
            for (int k = 0; k < kFpHalves; k++) {
                const uint32_t p = t + 32u * k + (uint32_t)lane;
                ok[k] = false; ml[k] = 0;
                if (cand[k]) {
                    const uint32_t cs = ((c[k] + mis) & 3u) * 8u;
                    ok[k] = __funnelshift_r(b[k][0], b[k][1], cs) == v[k];
                    const uint32_t x1 = __funnelshift_r(b[k][1], b[k][2], cs) ^ v4[k], x2 = __funnelshift_r(b[k][2], b[k][3], cs) ^ v8[k];
                    uint32_t m = 4u;
                    if (p + 8u <= mlimit) {
                        if (x1) m = 4u + ((uint32_t)(__ffs(x1) - 1) >> 3);
                        else if (p + 12u > mlimit) m = 8u;
                        else m = x2 ? 8u + ((uint32_t)(__ffs(x2) - 1) >> 3) : 12u;
                    }
                    if (ok[k]) {
                        // the last bytes in front of the limit, where no whole word fits (only in a unit's last round)
                        if (m < kFpLaneMax && p + m + 4u > mlimit) while (p + m < mlimit && src[p + m] == src[c[k] + m]) m++;
                        ml[k] = m;
                    }
                }
                // still equal after kFpLaneMax bytes and room for more: the warp extends it if it is selected
                cap[k] = ok[k] && ml[k] >= kFpLaneMax && p + ml[k] + 4u <= mlimit;
            }
#pragma unroll
            for (int k = 0; k < kFpHalves; k++) {
                const uint32_t tk = t + 32u * k;
                const uint32_t p = tk + (uint32_t)lane;
                // ---- greedy selection in position order (first match at or after the end of the previous one), done
                //      for all lanes at once: J = the match that follows mine if mine is selected; three rounds of
                //      pointer jumping give every match the set M of matches selected from it on (at most eight:
                //      a match covers four positions or more).  A match the warp still has to extend ends a chain.
                const unsigned okm = __ballot_sync(kFull, ok[k]);
                const uint32_t anchor0 = anchor;
                unsigned sel = 0;
                uint32_t mlen = ml[k];
                if (okm) {
                    const unsigned capm = __ballot_sync(kFull, cap[k]);
                    const uint32_t er = (uint32_t)lane + ml[k];
                    const unsigned after = er < 32u ? okm >> er : 0u;
                    uint32_t J = (ok[k] && !cap[k] && after) ? er + (uint32_t)(__ffs(after) - 1) : 32u;
                    unsigned M = ok[k] ? 1u << lane : 0u;
#pragma unroll
                    for (int it = 0; it < 3; it++) {
                        const unsigned Mj = __shfl_sync(kFull, M, (int)(J & 31u));
                        const uint32_t Jj = __shfl_sync(kFull, J, (int)(J & 31u));
                        if (J < 32u) { M |= Mj; J = Jj; }
                    }
                    for (;;) {
                        const uint32_t arel = anchor > tk ? anchor - tk : 0u;     // first lane of this half that may start a match
                        if (arel >= 32u) break;
                        const unsigned rest = okm >> arel;
                        if (!rest) break;
                        const int s = (int)arel + __ffs(rest) - 1;
                        const unsigned Ms = __shfl_sync(kFull, M, s);
                        sel |= Ms;
                        const int z = 31 - __clz(Ms);                              // the chain's last match
                        uint32_t mls = __shfl_sync(kFull, ml[k], z);
                        const uint32_t ps = tk + (uint32_t)z;
                        if ((capm >> z) & 1u) {
                            // lane j compares the word at +4j, 128 bytes per step
                            const uint32_t cs = __shfl_sync(kFull, c[k], z);
                            for (;;) {
                                const uint32_t q = mls + 4u * (uint32_t)lane;
                                uint32_t eq = 0;                              // equal bytes my word contributes
                                if (ps + q + 4u <= mlimit) {
                                    const uint32_t x = S.u32(ps + q) ^ S.u32(cs + q);
                                    eq = x ? ((uint32_t)(__ffs(x) - 1) >> 3) : 4u;
                                } else {
                                    while (eq < 4u && ps + q + eq < mlimit && src[ps + q + eq] == src[cs + q + eq]) eq++;
                                }
                                const unsigned part = __ballot_sync(kFull, eq < 4u);
                                if (part) {
                                    const int j = __ffs(part) - 1;
                                    mls += 4u * (uint32_t)j + __shfl_sync(kFull, eq, j);
                                    break;
                                }
                                mls += 128u;
                            }
                            if (lane == z) mlen = mls;
                            anchor = ps + mls;                                 // ... and the selection goes on behind it
                        } else {
                            anchor = ps + mls;                                 // the chain ran out of the half
                            break;
                        }
                    }
                }
                // what the emission needs: start = I begin a selected sequence (mlen, lit0 = where its literals begin),
                // seq = the lane whose sequence I am a literal of (32: none)
                const bool start = (sel >> lane) & 1u;
                const unsigned below = sel & lower, above = lane < 31 ? (sel >> (lane + 1)) : 0u;
                const uint32_t endp = __shfl_sync(kFull, p + mlen, (31 - __clz(below)) & 31);
                const uint32_t lit0 = below ? endp : anchor0;
                const uint32_t seq = (!start && above && p >= lit0) ? (uint32_t)lane + (uint32_t)__ffs(above) : 32u;
                // ---- lane-parallel emission of the half's sequences
                if (sel) {
                    const uint32_t ll = p - lit0, code = mlen - 4u;
                    const uint32_t ll_ext = (start && ll >= 15u) ? (ll - 15u) / 255u + 1u : 0u;
                    const uint32_t ml_ext = (start && code >= 15u) ? (code - 15u) / 255u + 1u : 0u;
                    const uint32_t size = start ? 1u + ll_ext + ll + 2u + ml_ext : 0u;
                    const uint32_t incl = warp_incl_sum(size, lane);
                    const uint32_t at = op + incl - size;                     // my sequence starts here
                    const uint32_t litbase = at + 1u + ll_ext - lit0;         // literal at position q goes to dst[litbase + q]
                    if (start) {
                        dst[at] = (uint8_t)((min(ll, 15u) << 4) | min(code, 15u));
                        if (ll_ext) { for (uint32_t j = 0; j + 1u < ll_ext; j++) dst[at + 1u + j] = 255; dst[at + ll_ext] = (uint8_t)((ll - 15u) % 255u); }
                        const uint32_t o = at + 1u + ll_ext + ll, offv = p - c[k];
                        dst[o] = (uint8_t)offv; dst[o + 1u] = (uint8_t)(offv >> 8);
                        if (ml_ext) { for (uint32_t j = 0; j + 1u < ml_ext; j++) dst[o + 2u + j] = 255; dst[o + 1u + ml_ext] = (uint8_t)((code - 15u) % 255u); }
                    }
                    // literals inside the half: every such lane writes its own byte (the low byte of its four)
                    const uint32_t lb = __shfl_sync(kFull, litbase, (int)(seq & 31u));
                    if (seq < 32u) dst[lb + p] = (uint8_t)v[k];
                    // literals in front of the half (left over by earlier rounds) belong to the first sequence
                    const int s0 = __ffs(sel) - 1;
                    const uint32_t b0 = __shfl_sync(kFull, litbase, s0);
                    if (anchor0 < tk) {
                        const uint32_t pend = tk - anchor0;
                        if (pend <= 32u) { if ((uint32_t)lane < pend) dst[(uint32_t)(b0 + anchor0 + lane)] = src[anchor0 + lane]; }
                        else warp_copy(dst + (uint32_t)(b0 + anchor0), src + anchor0, pend, lane);   // (b0 is a difference: wrap in 32 bits first)
                    }
                    op += __shfl_sync(kFull, incl, 31);
                }
            }
            t = max(t + 32u * kFpHalves, anchor);            // the inside of a match that leaves the round is skipped
            W = Wn;
        }
    }
    gate.wait(n);
    const uint32_t run = n - anchor;
    if (!emit_tail) { if (tail_len) *tail_len = run; return op; }        // lz4.c:2333-2338
    if (tail_len) *tail_len = 0;
    const uint32_t ext = run >= 15 ? (run - 15) / 255 + 1 : 0;
    if (lane == 0) dst[op] = (uint8_t)(min(run, 15u) << 4);
    if (ext) lz4_put_ext(dst + op + 1, run - 15, lane);
    warp_copy(dst + op + 1 + ext, src + anchor, run, lane);
    return op + 1 + ext + run;
}

}  // namespace llc
