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
//   * every lane verifies its candidate (4 equal bytes, distance <= 65535) and extends its own match, word by
//     word, up to 36 bytes; longer matches are extended by the whole warp, 128 bytes per step, only if selected;
//   * the greedy selection walks the round's matches in position order (first match at or after the end of the
//     previous one); literals, token and offset of each selected sequence are written by the warp.
// Deviations from the reference's parse that cost ratio (reported by bench.py as `ratio_delta_vs_exact`): matches
// against positions of the same round only inside one 32-position half, no backward catch-up
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
//
// A round covers kFpHalves x 32 consecutive positions, lane l taking positions t + l, t + 32 + l, ...: the loads of
// the halves are independent, so their latencies (table word from L2, candidate bytes from L2 / HBM, the words of
// the match extension) overlap -- the encoder is latency bound, one round is one dependent chain of ~6 memory
// round trips.  All halves read the table as it was BEFORE the round and insert afterwards (atomicMax: the result
// does not depend on the order of the lanes); a repeat closer than the round is found inside a half with
// MATCH.ANY on the four bytes (the nearest earlier lane with the same bytes), which is what catches runs and the
// short periods of columnar data.
constexpr int kFpHalves = 2;
__device__ inline uint32_t lz4_fastparse_unit(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, bool emit_tail,
                                              uint32_t* tail_len, uint32_t* tab, int lane, InGate& gate) {
    const LeanSrc S(src);
    uint32_t op = 0, anchor = 0;
    if (n >= 13u) {                                          // lz4.c:1926: shorter inputs are all literals
        for (uint32_t i = lane; i < (1u << kFpTabLog); i += 32) tab[i] = 0;
        __syncwarp();
        const uint32_t last_start = n - 12u;                 // a match starts at least 12 bytes before the end (MFLIMIT)
        const uint32_t mlimit = n - 5u;                      // ... and ends at least 5 bytes before it (LASTLITERALS)
        const uint32_t lower = (1u << lane) - 1u;            // lanes below mine
        uint32_t t = 0;
        while (t <= last_start) {
            gate.wait(min(n, t + 32u * kFpHalves + 64u + 136u));
            uint32_t v[kFpHalves], h[kFpHalves], e[kFpHalves], c[kFpHalves], ml[kFpHalves];
            bool inb[kFpHalves], ok[kFpHalves];
#pragma unroll
            for (int k = 0; k < kFpHalves; k++) {
                const uint32_t p = t + 32u * k + (uint32_t)lane;
                inb[k] = p <= last_start;
                v[k] = 0; h[k] = 0; e[k] = 0;
                if (inb[k]) { v[k] = S.u32(p); h[k] = (v[k] * 2654435761U) >> (32u - kFpTabLog); e[k] = tab[h[k]]; }
            }
            __syncwarp();                                    // every lane has read the old state
#pragma unroll
            for (int k = 0; k < kFpHalves; k++) {
                const uint32_t p = t + 32u * k + (uint32_t)lane;
                if (inb[k]) atomicMax(&tab[h[k]], p + 1u);   // entry = position + 1 (0: empty)
                // the nearest earlier lane of this half with the same four bytes, else the table's candidate
                const unsigned same = __match_any_sync(kFull, inb[k] ? v[k] : (0x80000000u | (uint32_t)lane) ^ v[k]) & lower;
                ok[k] = false;
                c[k] = e[k] - 1u;                            // < t: entries come from earlier rounds
                if (inb[k] && same) {                        // same four bytes by construction: nothing to verify
                    c[k] = t + 32u * k + (31u - (uint32_t)__clz(same)); ok[k] = true;
                } else if (inb[k] && e[k] != 0u && (p - c[k]) <= 65535u) {
                    ok[k] = S.u32(c[k]) == v[k];
                }
                ml[k] = ok[k] ? 4u : 0u;
            }
            // every lane extends its own matches, word by word, up to kFpLaneMax bytes (both halves in one loop)
            bool more[kFpHalves];
#pragma unroll
            for (int k = 0; k < kFpHalves; k++) more[k] = ok[k];
            for (;;) {
                bool any = false;
#pragma unroll
                for (int k = 0; k < kFpHalves; k++) {
                    const uint32_t p = t + 32u * k + (uint32_t)lane;
                    more[k] = more[k] && ml[k] < kFpLaneMax && p + ml[k] + 4u <= mlimit;
                    if (more[k]) {
                        const uint32_t x = S.u32(p + ml[k]) ^ S.u32(c[k] + ml[k]);
                        if (x) { ml[k] += (uint32_t)(__ffs(x) - 1) >> 3; more[k] = false; }
                        else ml[k] += 4u;
                    }
                    any = any || more[k];
                }
                if (!any) break;
            }
#pragma unroll
            for (int k = 0; k < kFpHalves; k++) {
                const uint32_t p = t + 32u * k + (uint32_t)lane;
                if (ok[k] && ml[k] < kFpLaneMax && p + ml[k] + 4u > mlimit)   // the last bytes before the limit, one at a time
                    while (p + ml[k] < mlimit && src[p + ml[k]] == src[c[k] + ml[k]]) ml[k]++;
            }
            // greedy selection in position order
#pragma unroll
            for (int k = 0; k < kFpHalves; k++) {
                const uint32_t tk = t + 32u * k;
                unsigned okm = __ballot_sync(kFull, ok[k]);
                while (okm) {
                    const uint32_t arel = anchor > tk ? anchor - tk : 0u;     // first lane of this half that may start a match
                    if (arel >= 32u) break;
                    okm &= ~((1u << arel) - 1u);
                    if (!okm) break;
                    const int s = __ffs(okm) - 1;
                    okm &= okm - 1u;
                    const uint32_t ps = tk + (uint32_t)s;
                    const uint32_t cs = __shfl_sync(kFull, c[k], s);
                    uint32_t mls = __shfl_sync(kFull, ml[k], s);
                    if (mls >= kFpLaneMax) {
                        // a long match: the warp extends it, lane j compares the word at +4j, 128 bytes per step
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
                    }
                    fp_emit(src, dst, op, anchor, ps, ps - cs, mls, lane);
                    anchor = ps + mls;
                }
            }
            t = max(t + 32u * kFpHalves, anchor);            // the inside of a match that leaves the round is skipped
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
