// lz4_fastparse.cuh -- the separately named `fastparse` compress mode (never the default).
//
// Same RAP frame, same partitions, same stitch as the exact mode -- only the partition encoder differs, so the
// streams decode with any LZ4 decoder (the reference's included) but are NOT byte-identical to the reference's.
// The exact encoder reproduces LZ4_compress_fast's serial greedy parse (lz4.c:1853-2350): ~270 warp instructions per
// sequence, because everything about a sequence (probe schedule, table commits, catch-up, emission) is decided one
// sequence at a time.  Here the parse is POSITION PARALLEL, 32 positions per round:
//   * every lane hashes its own position and reads the table (the state BEFORE this round: candidates come from
//     earlier rounds only), then all lanes insert with atomicMax -- the table ends up holding the highest position
//     per bucket whatever the order of the lanes, so the output is deterministic;
//   * every lane verifies its candidate (4 equal bytes, distance <= 65535) and extends its own match, word by
//     word, up to 36 bytes; longer matches are extended by the whole warp, 128 bytes per step, only if selected;
//   * the greedy selection walks the round's matches in position order (first match at or after the end of the
//     previous one); literals, token and offset of each selected sequence are written by the warp.
// Deviations from the reference's parse that cost ratio (reported by bench.py as `ratio_delta_vs_exact`): no
// matches against positions of the same round (distance < 32 only via earlier rounds), no backward catch-up
// (lz4.c:2098), every position of a round is inserted (the reference skips the inside of matches and accelerates
// through incompressible data: lz4.c:1991-1997).  Precedent for trading ratio for speed in the reference itself:
// AOCL_LZ4_MATCH_SKIP_OPT_LDS_STRAT1/2 (lz4.c:1447-1450, 1572-1584), AOCL_SNAPPY_MATCH_SKIP_OPT (snappy.cc:939-969).
#pragma once
#include "lz4_encode_lean.cuh"

namespace llc {

constexpr uint32_t kFpTabLog = 12;                           // 4096 x u32, the size of the reference's table
constexpr uint32_t kFpLaneMax = 36;                          // match bytes a lane extends on its own

// One sequence, written by the whole warp: token | literal-length bytes | literals | offset | match-length bytes
__device__ __forceinline__ void fp_emit(const uint8_t* __restrict__ src, uint8_t* dst, uint32_t& op, uint32_t anchor,
                                        uint32_t ps, uint32_t offv, uint32_t ml, int lane) {
    const uint32_t ll = ps - anchor, code = ml - 4u;
    if ((ll < 15u) & (code < 15u)) {                         // at most 17 bytes: one byte per lane
        uint32_t v = 0;
        if ((uint32_t)(lane - 1) < ll) v = src[anchor + lane - 1];
        if (lane == 0) v = (ll << 4) | code;
        if ((uint32_t)lane == ll + 1u) v = offv;
        if ((uint32_t)lane == ll + 2u) v = offv >> 8;
        if ((uint32_t)lane <= ll + 2u) dst[op + lane] = (uint8_t)v;
        op += ll + 3u;
        return;
    }
    const uint32_t ll_ext = ll >= 15 ? (ll - 15) / 255 + 1 : 0;
    const uint32_t ml_ext = code >= 15 ? (code - 15) / 255 + 1 : 0;
    if (lane == 0) dst[op] = (uint8_t)((min(ll, 15u) << 4) | min(code, 15u));
    if (ll_ext) lz4_put_ext(dst + op + 1, ll - 15, lane);
    if (ll <= 32) { if ((uint32_t)lane < ll) dst[op + 1 + ll_ext + lane] = src[anchor + lane]; }
    else warp_copy(dst + op + 1 + ll_ext, src + anchor, ll, lane);
    op += 1 + ll_ext + ll;
    if (lane == 0) { dst[op] = (uint8_t)offv; dst[op + 1] = (uint8_t)(offv >> 8); }
    op += 2;
    if (ml_ext) lz4_put_ext(dst + op, code - 15, lane);
    op += ml_ext;
}

// Encodes src[0, n) as LZ4 sequences at dst.  Same contract as lz4_encode_unit (unlimited output): returns the
// body length; a non-final unit leaves its trailing literals to the stitch (*tail_len), the final one writes them.
// `tab`: 4096 words of shared or global memory owned by this warp.
__device__ inline uint32_t lz4_fastparse_unit(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, bool emit_tail,
                                              uint32_t* tail_len, uint32_t* tab, int lane, InGate& gate) {
    const LeanSrc S(src);
    uint32_t op = 0, anchor = 0;
    if (n >= 13u) {                                          // lz4.c:1926: shorter inputs are all literals
        for (uint32_t i = lane; i < (1u << kFpTabLog); i += 32) tab[i] = 0;
        __syncwarp();
        const uint32_t last_start = n - 12u;                 // a match starts at least 12 bytes before the end (MFLIMIT)
        const uint32_t mlimit = n - 5u;                      // ... and ends at least 5 bytes before it (LASTLITERALS)
        uint32_t t = 0;
        while (t <= last_start) {
            gate.wait(min(n, t + 32u + 64u + 136u));
            const uint32_t p = t + (uint32_t)lane;
            const bool inb = p <= last_start;
            uint32_t v = 0, h = 0, e = 0;
            if (inb) { v = S.u32(p); h = (v * 2654435761U) >> (32u - kFpTabLog); e = tab[h]; }
            __syncwarp();                                    // every lane has read the old state
            if (inb) atomicMax(&tab[h], p + 1u);             // entry = position + 1 (0: empty)
            const uint32_t c = e - 1u;                       // < t: entries come from earlier rounds
            const bool ok = inb && e != 0u && (p - c) <= 65535u && S.u32(c) == v;
            uint32_t ml = 0;
            if (ok) {
                ml = 4u;
                while (ml < kFpLaneMax && p + ml + 4u <= mlimit) {
                    const uint32_t x = S.u32(p + ml) ^ S.u32(c + ml);
                    if (x) { ml += (uint32_t)(__ffs(x) - 1) >> 3; break; }
                    ml += 4u;
                }
                if (ml < kFpLaneMax && p + ml + 4u > mlimit)   // the last bytes before the limit, one at a time
                    while (p + ml < mlimit && src[p + ml] == src[c + ml]) ml++;
            }
            unsigned okm = __ballot_sync(kFull, ok);
            while (okm) {
                const uint32_t arel = anchor > t ? anchor - t : 0u;       // first lane of this round that may start a match
                if (arel >= 32u) break;
                okm &= ~((1u << arel) - 1u);
                if (!okm) break;
                const int s = __ffs(okm) - 1;
                okm &= okm - 1u;
                const uint32_t ps = t + (uint32_t)s;
                const uint32_t cs = __shfl_sync(kFull, c, s);
                uint32_t mls = __shfl_sync(kFull, ml, s);
                if (mls >= kFpLaneMax) {
                    // a long match: the warp extends it, lane j compares the word at +4j, 128 bytes per step
                    for (;;) {
                        const uint32_t q = mls + 4u * (uint32_t)lane;
                        uint32_t eq = 0;                                  // equal bytes my word contributes
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
                }
                fp_emit(src, dst, op, anchor, ps, ps - cs, mls, lane);
                anchor = ps + mls;
            }
            t = max(t + 32u, anchor);                        // the inside of a match that leaves the round is skipped
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
