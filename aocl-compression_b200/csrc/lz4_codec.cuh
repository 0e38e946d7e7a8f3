// lz4_codec.cuh -- warp-per-unit LZ4 block decoder and exact greedy encoder.
//
// Behaviour follows the reference's partition decoder AOCL_LZ4_decompress_generic_mt
// (algos/lz4/lz4.c:3806-4305) and partition encoder AOCL_LZ4_compress_generic_validated_mt
// (algos/lz4/lz4.c:1853-2350, acceleration 1, noDict); the encoder is byte-exact with it.
#pragma once
#include "llc_common.cuh"

namespace llc {

// -------------------------------------------------------------------------------------------
// Decoder.  One warp decodes [in, in+clen) into [out, out+cap).  `last` selects the vanilla
// end-of-block rules (final RAP partition or frame-less block); non-final partitions may end
// right after a match (lz4.c:4285-4288).  Returns bytes produced or kErrCorrupt; never reads
// outside the input range nor writes outside the output range.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lz4_read_ext(const uint8_t* in, uint32_t& ip, uint32_t clen, bool& bad) {
    uint32_t add = 0, b;
    do {
        if (ip >= clen) { bad = true; return add; }
        b = in[ip++];
        add += b;
    } while (b == 255 && add < 0x7fff0000u);
    return add;
}

__device__ inline int64_t lz4_decode_warp(const uint8_t* __restrict__ in, uint32_t clen, uint8_t* out,
                                          uint64_t cap, bool last, int lane) {
    uint32_t ip = 0;
    uint64_t op = 0;
    if (clen == 0) return kErrCorrupt;
    if (cap == 0) return (clen == 1 && in[0] == 0) ? 0 : kErrCorrupt;   // lz4.c:3854-3858
    for (;;) {
        if (ip >= clen) return kErrCorrupt;
        const uint32_t tok = in[ip++];
        uint32_t len = tok >> 4;
        bool bad = false;
        if (len == 15) { len += lz4_read_ext(in, ip, clen, bad); if (bad) return kErrCorrupt; }
        if (len > clen - ip || (uint64_t)len > cap - op) return kErrCorrupt;
        // end-of-block parsing restrictions, lz4.c:4104-4164
        const bool closing = (op + len + 12 > cap) || ((uint64_t)ip + len + 8 > clen);
        if (closing && last && ip + len != clen) return kErrCorrupt;
        warp_copy(out + op, in + ip, len, lane);
        ip += len; op += len;
        if (closing && (last || op == cap)) break;
        if (ip == clen) break;
        if (ip + 2 > clen) return kErrCorrupt;
        const uint32_t off = ld_u16(in + ip); ip += 2;
        uint32_t ml = tok & 15;
        if (ml == 15) { ml += lz4_read_ext(in, ip, clen, bad); if (bad) return kErrCorrupt; }
        ml += 4;
        if (off == 0 || (uint64_t)off > op) return kErrCorrupt;          // lz4.c:4196-4197
        if ((uint64_t)ml > cap - op) return kErrCorrupt;
        if (last && op + ml + 5 > cap) return kErrCorrupt;               // lz4.c:4262-4264
        __syncwarp();
        warp_match_copy(out, op, off, ml, lane);
        __syncwarp();
        op += ml;
        if (!last && (op == cap || ip >= clen)) break;                   // lz4.c:4285-4288
    }
    return (int64_t)op;
}

// -------------------------------------------------------------------------------------------
// Encoder.  One warp runs the greedy single-probe parse over src[0..n) with the hash table in
// shared memory (16 KiB: 4096 x u32 positions for n >= 65547, 8192 x u16 below; lz4.c:2557-2563).
//
// The parse is inherently serial, but while no match is found the probe positions follow a
// fixed schedule (lz4.c:1991-1997).  The warp therefore evaluates 32 future probes at once:
// lane k takes the k-th upcoming position, its candidate is the position of the nearest
// earlier lane with the same hash (that lane would have overwritten the slot) or else the
// table entry; the first lane whose candidate verifies wins; table writes are committed for
// lanes up to the winner only.  This reproduces the serial algorithm bit for bit.
//
// emit_tail: write the closing literal run (last partition / frame-less block); otherwise
// stop before it and report its length in *tail_len (lz4.c:2333-2338).
// cap < 0: unbounded output; cap >= 0: the reference's limitedOutput checks (returns 0).
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lz4_hash5(uint64_t v) { return (uint32_t)(((v << 24) * 889523592379ULL) >> 52); }
__device__ __forceinline__ uint32_t lz4_hash4(uint32_t v) { return (v * 2654435761U) >> 19; }

template <bool WIDE>
struct Lz4Table {
    uint32_t* t32;
    __device__ __forceinline__ uint32_t get(uint32_t h) const {
        if (WIDE) return t32[h];
        return reinterpret_cast<const uint16_t*>(t32)[h];
    }
    __device__ __forceinline__ void put(uint32_t h, uint32_t pos) const {
        if (WIDE) t32[h] = pos;
        else reinterpret_cast<uint16_t*>(t32)[h] = (uint16_t)pos;
    }
    __device__ __forceinline__ uint32_t hash(const uint8_t* p) const {
        return WIDE ? lz4_hash5(ld_u64(p)) : lz4_hash4(ld_u32(p));
    }
};

// Writes `count` length-extension bytes for value v (already reduced by 15): 255 ... 255, v % 255.
__device__ __forceinline__ void lz4_put_ext(uint8_t* dst, uint32_t v, int lane) {
    const uint32_t count = v / 255 + 1;
    for (uint32_t j = lane; j < count; j += 32) dst[j] = (j + 1 < count) ? (uint8_t)255 : (uint8_t)(v % 255);
}

// -------------------------------------------------------------------------------------------
// Encoder, fused-round version (units below 65,547 bytes -- the byU16 / hash4 table -- and frame-less blocks of
// 512 KiB and more; RAP partitions take lz4_encode_lean.cuh).  One sequence costs two dependent global round trips:
//   * every lane loads a 12-byte window [p-4, p+8) around its probe position and, after the table
//     lookup, around its candidate; from the two windows it derives locally whether the candidate
//     verifies, how far the match extends backwards (catch-up, up to 4 bytes) and forwards (up to 8
//     bytes).  Only longer extensions take the extra compare rounds;
//   * the probe of the position right after a match (lz4.c:2230-2288: insert ip-2, test ip) is
//     slot 0 of the next 32-probe round, followed by the first 31 probes of the next search.
// -------------------------------------------------------------------------------------------
struct Lz4Win { uint32_t back, lo, hi; };      // bytes [p-4,p), [p,p+4), [p+4,p+8); back is 0 when p < 4
__device__ __forceinline__ Lz4Win lz4_load_win(const uint8_t* src, uint32_t pos) {
    const uint32_t o = pos >= 4u ? 4u : 0u;
    const uint8_t* p = src + pos - o;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    const unsigned sh = (unsigned)(a & 3) * 8;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
    const uint32_t x = __funnelshift_r(w0, w1, sh), y = __funnelshift_r(w1, w2, sh), z = __funnelshift_r(w2, w3, sh);
    Lz4Win r;
    if (o) { r.back = x; r.lo = y; r.hi = z; } else { r.back = 0; r.lo = x; r.hi = y; }
    return r;
}

template <bool WIDE>
__device__ inline uint32_t lz4_encode_warp_fused(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, int64_t cap,
                                                 bool emit_tail, uint32_t* tail_len, uint32_t* tab_mem, int lane, InGate& gate) {
    Lz4Table<WIDE> tab{tab_mem};
    for (int i = lane; i < 4096; i += 32) tab_mem[i] = 0;
    __syncwarp();
    gate.wait(min(n, 64u));
    const bool limited = cap >= 0;
    uint32_t op = 0, anchor = 0;
    bool refused = false;

    if (n >= 13) {                                          // LZ4_minLength, lz4.c:1926
        const uint32_t mfl1 = n - 11;                       // mflimitPlusOne, lz4.c:1887
        const uint32_t mlimit = n - 5;                      // matchlimit, lz4.c:1888
        tab.put(tab.hash(src), 0);                          // lz4.c:1929
        __syncwarp();
        bool post = false;                                  // slot 0 is the probe right after a match
        uint32_t base = 0;                                  // post rounds: position of slot 0
        uint32_t fwd = 1, step = 1, nb = 64;                // search schedule (lz4.c:1991-1997)
        for (;;) {
            // ---------------- one round of 32 probe slots in serial order ----------------
            uint32_t cur, nxt;
            bool valid;
            if (post) { cur = base + lane; nxt = cur + 1; valid = (lane == 0) || (nxt <= mfl1); }
            else {
                const uint32_t my_step = (lane == 0) ? step : ((nb + lane - 1) >> 6);
                const uint32_t incl = warp_incl_sum(my_step, lane);
                cur = fwd + incl - my_step; nxt = fwd + incl;
                valid = nxt <= mfl1;                        // lz4.c:2001
            }
            gate.wait(min(n, __shfl_sync(kFull, cur, 31) + 32u));   // every window of this round lies below cur[31] + 16
            Lz4Win cw = {0, 0, 0};
            uint32_t h = 0x80000000u | lane;
            if (valid) {
                cw = lz4_load_win(src, cur);
                h = WIDE ? lz4_hash5((uint64_t)cw.lo | ((uint64_t)cw.hi << 32)) : lz4_hash4(cw.lo);
            }
            if (post) {                                     // lz4.c:2230: insert ip-2 before anything else
                if (lane == 0) {
                    // bytes [base-2, base+6) out of the window [base-4, base+8)
                    const uint64_t v = ((uint64_t)(cw.back >> 16)) | ((uint64_t)cw.lo << 16) | ((uint64_t)cw.hi << 48);
                    tab.put(WIDE ? lz4_hash5(v) : lz4_hash4((uint32_t)v), base - 2);
                }
                __syncwarp();
            }
            uint32_t cand = valid ? tab.get(h) : 0u;
            const unsigned peers = __match_any_sync(kFull, h);
            const unsigned before = peers & ((1u << lane) - 1u);
            const int from = before ? (31 - __clz(before)) : lane;
            const uint32_t peer_pos = __shfl_sync(kFull, cur, from);
            if (before) cand = peer_pos;
            bool hit = false;
            uint32_t bk = 0, fw = 0;                        // locally known backward / forward extension
            bool more_b = false, more_f = false;
            if (valid) {
                const Lz4Win mw = lz4_load_win(src, cand);
                hit = (mw.lo == cw.lo) && (!WIDE || cur - cand <= 65535u);      // lz4.c:2048-2057
                // forward: bytes [p+4, p+8), bounded by matchlimit
                const uint32_t xf = cw.hi ^ mw.hi;
                const uint32_t eqf = xf ? ((uint32_t)(__ffs(xf) - 1) >> 3) : 4u;
                const uint32_t roomf = mlimit > cur + 4u ? mlimit - (cur + 4u) : 0u;
                fw = min(eqf, roomf);
                more_f = (eqf == 4u) && (roomf > 4u);
                // backward: bytes [p-4, p) from the top, bounded by the anchor and by position 0 (lz4.c:2098)
                const uint32_t bwin = (cur >= 4u && cand >= 4u) ? 4u : 0u;
                const uint32_t xb = cw.back ^ mw.back;
                const uint32_t eqb = bwin ? (xb ? ((uint32_t)__clz(xb) >> 3) : 4u) : 0u;
                const uint32_t roomb = min(cur - anchor, cand);
                bk = min(eqb, roomb);
                more_b = (eqb == bwin) && (roomb > bwin);
            }
            const unsigned hits = __ballot_sync(kFull, hit);
            const unsigned events = hits | __ballot_sync(kFull, !valid);
            const int win = events ? (__ffs(events) - 1) : 32;
            const bool win_is_match = (win < 32) && ((hits >> win) & 1u);
            const unsigned commit = (win_is_match ? (win == 31 ? kFull : ((2u << win) - 1u))
                                                  : (win == 0 ? 0u : (win >= 32 ? kFull : ((1u << win) - 1u))));
            if (valid && ((commit >> lane) & 1u)) {
                const unsigned mine = peers & commit;
                if ((31 - __clz(mine)) == lane) tab.put(h, cur);
            }
            __syncwarp();
            if (win >= 32) {                                // nothing happened: next 32 probes of the same search
                if (post) { fwd = base + 32; step = 1; nb = 64 + 31; post = false; }
                else { fwd = __shfl_sync(kFull, nxt, 31); step = (nb + 31) >> 6; nb += 32; }
                continue;
            }
            if (!win_is_match) break;                       // search ran into the end of the block -> closing literals

            // ---------------- the winner's match ----------------
            const uint32_t mpos = __shfl_sync(kFull, cur, win);
            const uint32_t mcand = __shfl_sync(kFull, cand, win);
            uint32_t back = __shfl_sync(kFull, bk, win);
            uint32_t mc = __shfl_sync(kFull, fw, win);
            const bool go_b = __shfl_sync(kFull, (int)more_b, win) != 0;
            const bool go_f = __shfl_sync(kFull, (int)more_f, win) != 0;
            const bool from_search = !(post && win == 0);   // slot 0 of a post round: zero literals, no catch-up
            uint32_t ip = mpos - back, m = mcand - back;
            if (go_b && from_search) {                      // rare: catch-up longer than the window
                for (;;) {
                    const bool can = (ip > anchor + lane) && (m > (uint32_t)lane);
                    const bool eq = can && (src[ip - 1 - lane] == src[m - 1 - lane]);
                    const unsigned ne = __ballot_sync(kFull, !eq);
                    const uint32_t cnt = ne ? (uint32_t)(__ffs(ne) - 1) : 32u;
                    ip -= cnt; m -= cnt; back += cnt;
                    if (cnt < 32) break;
                }
            }
            if (go_f) {                                     // match longer than 8: 128 bytes per extra round (lz4.c:656-679)
                const uint32_t delta = mpos - mcand;
                uint32_t pb = mpos + 8;
                for (;;) {
                    gate.wait(min(n, pb + 144u));
                    const uint32_t pa = pb + 4u * lane;
                    uint32_t c = 0;
                    if (pa < mlimit) {
                        const uint32_t avail = min(4u, mlimit - pa);
                        const uint32_t x = ld_u32(src + pa) ^ ld_u32(src + pa - delta);
                        c = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
                        c = min(c, avail);
                    }
                    const unsigned partial = __ballot_sync(kFull, c < 4);
                    if (partial) {
                        const int first = __ffs(partial) - 1;
                        mc += 4u * first + __shfl_sync(kFull, c, first);
                        break;
                    }
                    mc += 128; pb += 128;
                }
            }
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
                // In a post round lane j probed position anchor + j, so it already holds literal j.
                const uint32_t offv = ip - m;
                uint32_t v = __shfl_up_sync(kFull, cw.lo & 0xffu, 1);
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

__device__ inline uint32_t lz4_encode_warp(const uint8_t* src, uint32_t n, uint8_t* dst, int64_t cap, bool emit_tail,
                                           uint32_t* tail_len, uint32_t* tab_mem, int lane, InGate& gate) {
    if (n == 0) { if (lane == 0) dst[0] = 0; if (tail_len) *tail_len = 0; return 1; }   // lz4.c:2418-2428
    if (n >= 65547) return lz4_encode_warp_fused<true>(src, n, dst, cap, emit_tail, tail_len, tab_mem, lane, gate);
    return lz4_encode_warp_fused<false>(src, n, dst, cap, emit_tail, tail_len, tab_mem, lane, gate);
}

}  // namespace llc
