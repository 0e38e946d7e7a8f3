// decode_tile.cuh -- tile decoder: one 512-thread CTA per unit (RAP partition or page), LANE PER
// SEQUENCE for the parse and THREAD PER BYTE for the copies, with the last 32 KiB of output kept in a
// shared-memory ring.
//
// Why.  A unit is one serial token chain, and the warp-per-unit decoders (lz4_decode_ring.cuh) spend
// ~160 warp instructions per 10-byte sequence: the whole GPU ends up issue bound at ~1 % of the HBM
// roofline.  Byte-granular copies against global memory cost one L1 wavefront per lane, so handing
// sequences to lanes only pays if the bytes live in shared memory.  Here:
//
//   * the compressed stream arrives in 4 KiB chunks through the TMA bulk-copy engine
//     (cp.async.bulk + mbarrier, double buffered);
//   * PARSE is data parallel: every byte position of the chunk is treated as a potential token and
//     gets a "next token" link (n1); six rounds of pointer doubling give 8-hop and 64-hop links; one
//     thread chases the 64-hop links from the known chunk entry (~25 dependent shared loads per chunk)
//     and 8-hop/1-hop links expand the anchors into the list of real sequence starts;
//   * a block scan of the sequence lengths gives every sequence its output position; offsets and
//     capacity are validated there (the first irregular sequence truncates the group);
//   * COPY is THREAD PER OUTPUT BYTE.  Byte x of the group finds its sequence (a bitmap of sequence
//     starts + the sequence that covers the first byte of every 32-byte row), and then either takes
//     its literal from the chunk buffer, or its match source from the ring (sources older than the
//     group are final; sources that have left the 32 KiB ring are read back from HBM/L2), or -- a
//     match whose source lies inside the group -- records a POINTER to that byte.  Text-like data
//     chains every occurrence of a string to the previous one, so these pointers form chains hundreds
//     of links deep: they are resolved by pointer jumping, P[x] = P[P[x]], until every chain has
//     reached a byte that is known.  A table entry holds either a pointer (< 16384) or 0xFF00 | value,
//     so one 16-bit load tells a reader both whether the byte is known and what it is; concurrent
//     updates are benign (any entry read is an ancestor of x with the same final value).  A first
//     version copied lane-per-sequence in dependency rounds with source forwarding: 62.5 % of the
//     group time (profiles/r1_v10_*), most of it spent polling for producers;
//   * the finished span is flushed ring -> HBM with aligned 16-byte stores, so HBM only sees
//     coalesced traffic: C bytes in through TMA, U bytes out through STG.128.
//
// Anything irregular -- length runs >= 270 (a 255 extension byte), the closing sequences of a stream,
// sequences that would cross the output capacity, malformed input -- is executed one sequence at a
// time by the fully checked slow step, which mirrors lz4_slow_sequence()/snappy_decode_warp() and
// through them the reference decoders (algos/lz4/lz4.c:3806-4305; algos/snappy/snappy.cc:1466-1570,
// 2185-2199).  Accept/reject behaviour and produced bytes are identical to the warp decoders.
#pragma once
#include "in_ring.cuh"
#include "snappy_codec.cuh"

namespace llc {

constexpr int kTThreads = 512;
constexpr int kTWarps = kTThreads / 32;
constexpr uint32_t kTRing = 32768u, kTRingMask = kTRing - 1u;
constexpr uint32_t kTChunkLog = 12, kTChunk = 1u << kTChunkLog;
constexpr uint32_t kTMargin = 384;                       // a regular sequence reads < 280 bytes past its token
constexpr uint32_t kTBuf = kTChunk + kTMargin;           // multiple of 16
constexpr uint32_t kTSpan = 16384;                       // output bytes per group (truncated beyond)
constexpr uint32_t kTRows = kTSpan / 32;                 // 32-byte rows of a group span: one warp handles one row at a time
constexpr uint32_t kTKnown = 0xFF00u;                    // P entry >= kTKnown: the byte is known, value in the low 8 bits
constexpr uint32_t kTPtr = 0x4000u;                      // P entry in [kTPtr, 2 kTPtr): pointer to span index (entry - kTPtr)
constexpr uint32_t kTPiece = 16384;                      // slow-step copy granule (<= half the ring)
static_assert(kTRows == (uint32_t)kTThreads && kTSpan <= kTPtr && 2 * kTPtr <= kTKnown && kTSpan + kTPiece <= kTRing, "tile geometry");
constexpr uint32_t kTNone = 0xffffu;
constexpr uint32_t kTCapMax = 0xffffff00u;

template <class Fmt>
struct TileShared {
    static constexpr int MAXSEQ = Fmt::kMaxSeq;
    alignas(128) uint8_t ring[kTRing];
    alignas(128) uint8_t inbuf[2][kTBuf];
    uint16_t n1[kTChunk], ta[kTChunk], tb[kTChunk];
    alignas(16) uint16_t P[kTSpan];                      // per byte of the group span: 0 = not written yet, kTPtr | index of its
                                                         // source byte, or kTKnown | value
    uint2 sq[MAXSEQ];                                    // x: span index of the sequence | literal length << 16;
                                                         // y: match offset | chunk index of the literals << 16
    uint16_t seq_start[MAXSEQ];
    uint32_t startbits[kTRows + 1];                      // bit x: a sequence starts at byte x of the group span
    uint16_t row2seq[kTRows + 32];                       // 1 + the sequence covering the first byte of every 32-byte row (0: none)
    uint32_t batch_tot[MAXSEQ / 32 + 1];
    uint16_t anchors[MAXSEQ / 64 + 2];                   // 64-sequence anchors (the last one starts the tail of < 64)
    alignas(8) uint64_t bar[2];
    uint32_t nanch, ntail, end_kind, end_pos, first_bad, unit, abort, span_end;
};

struct TSeq { uint32_t nxt, lit, ll, ml, off; };         // chunk-relative indices; nxt == kTNone: irregular

// Per-unit decoder state.  Everything is uniform across the CTA (kept redundantly in registers).
struct TileState {
    const uint8_t* gin;      // 16-byte aligned, <= first stream byte
    uint8_t* gout;           // 16-byte aligned, <= first output byte; positions below are relative to it
    uint32_t iend;           // stream end (relative to gin)
    uint32_t a;              // first output position (out & 15)
    uint32_t cap;            // output end (a + capacity)
    uint32_t ip, op, flushed;
    uint32_t par;            // mbarrier parities (bit b = next phase of buffer b); persists across units
    uint32_t pend;           // bit b: a load into buffer b is in flight
    int32_t bufc[2];         // chunk held by / in flight into each buffer
    bool last;
};

__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) { return __reduce_max_sync(kFull, v); }

// Debug counters (AOCL_GPU tile decoder bring-up): [0..15] cycles per phase summed over CTAs (thread 0),
// [16..23] event counts, [24..31] watchdog codes.  Compiled in only with -DLLC_TILE_PROF.
__device__ unsigned long long g_tile_prof[32];
#ifdef LLC_TILE_PROF
#define TP_DECL long long tp_t0 = clock64();
#define TP(i) do { if (threadIdx.x == 0) { const long long tp_t1 = clock64(); atomicAdd(&g_tile_prof[i], (unsigned long long)(tp_t1 - tp_t0)); tp_t0 = tp_t1; } } while (0)
#define TC(i, v) do { if (threadIdx.x == 0) atomicAdd(&g_tile_prof[16 + (i)], (unsigned long long)(v)); } while (0)
#else
#define TP_DECL
#define TP(i) do {} while (0)
#define TC(i, v) do {} while (0)
#endif
#define TWATCH(code) do { atomicAdd(&g_tile_prof[24 + (code)], 1ull); } while (0)
constexpr uint32_t kTSpinMax = 1u << 22;

// ------------------------------------------------------------------------------------------------
// Formats
// ------------------------------------------------------------------------------------------------
struct TileLz4 {
    static constexpr int kMaxSeq = 1024;                 // sequences per group (a 4 KiB chunk of text holds ~700)
    static constexpr bool kHasLit = true;                // a sequence carries literals and a match
    static constexpr bool kSetupPairs = false;           // per-byte source step: one row per trip
    static constexpr uint32_t kEndSlack = 12;            // regular sequences end >= 12 bytes before the capacity

    __device__ static __forceinline__ TSeq parse(const uint8_t* bp, uint32_t i) {
        TSeq s;
        const uint32_t tok = bp[i], e1 = bp[i + 1];
        const uint32_t nibL = tok >> 4, nibM = tok & 15u;
        const bool extL = nibL == 15u, extM = nibM == 15u;
        s.ll = nibL + (extL ? e1 : 0u);
        s.lit = i + 1u + (extL ? 1u : 0u);
        const uint32_t q = s.lit + s.ll;
        s.off = (uint32_t)bp[q] | ((uint32_t)bp[q + 1] << 8);
        const uint32_t e2 = bp[q + 2];
        s.ml = 4u + nibM + (extM ? e2 : 0u);
        const bool special = (extL && e1 == 255u) || (extM && e2 == 255u);
        s.nxt = special ? kTNone : q + 2u + (extM ? 1u : 0u);
        return s;
    }
    // A sequence is regular only if it is not one of the closing ones: its literals end at least 8 bytes
    // before the end of the stream (lz4.c:4104-4164), which also keeps its offset / length bytes inside.
    __device__ static __forceinline__ bool ends_inside(const TSeq& s, uint32_t cbase, uint32_t iend) {
        return cbase + s.lit + s.ll + 8u <= iend;
    }
};

struct TileSnappy {
    static constexpr int kMaxSeq = 1280;                 // elements per group (a 4 KiB chunk of text holds ~1200); sized to keep 2 CTAs per SM
    static constexpr bool kHasLit = false;               // an element is either literals or a copy
    static constexpr bool kSetupPairs = true;            // per-byte source step: two rows per trip
    static constexpr uint32_t kEndSlack = 0;

    // one element = one "sequence" with either literals or a copy
    __device__ static __forceinline__ TSeq parse(const uint8_t* bp, uint32_t i) {
        TSeq s;
        const uint32_t tag = bp[i], b1 = bp[i + 1], b2 = bp[i + 2];
        const uint32_t kind = tag & 3u;
        s.ll = 0; s.ml = 0; s.off = 0; s.lit = i + 1u; s.nxt = kTNone;
        if (kind == 0) {
            const uint32_t v = tag >> 2;
            if (v < 60u) { s.ll = v + 1u; s.nxt = i + 1u + s.ll; }
            else if (v == 60u) { s.ll = b1 + 1u; s.lit = i + 2u; s.nxt = i + 2u + s.ll; }
        } else if (kind == 1) {
            s.ml = 4u + ((tag >> 2) & 7u); s.off = ((tag >> 5) << 8) | b1; s.nxt = i + 2u;
        } else if (kind == 2) {
            s.ml = 1u + (tag >> 2); s.off = b1 | (b2 << 8); s.nxt = i + 3u;
        }
        return s;
    }
    // an element is regular only if all of it lies inside the stream
    __device__ static __forceinline__ bool ends_inside(const TSeq& s, uint32_t cbase, uint32_t iend) {
        return cbase + s.nxt <= iend;
    }
};

static_assert(sizeof(TileShared<TileLz4>) <= 115712 && sizeof(TileShared<TileSnappy>) <= 115712,
              "two CTAs per SM need <= 113 KiB of shared memory each");

// ------------------------------------------------------------------------------------------------
// Ring -> HBM.  Writes [st.flushed, hi) with aligned 16-byte stores; a partial first unit is written
// byte-wise, a partial last unit only when `final` (otherwise it stays in the ring for the next flush).
// ------------------------------------------------------------------------------------------------
template <class Fmt>
__device__ __forceinline__ void tile_flush(TileShared<Fmt>& sh, TileState& st, uint32_t hi, bool final) {
    uint32_t lo = st.flushed;
    const uint32_t tid = threadIdx.x;
    if (lo >= hi) return;
    if (lo & 15u) {
        const uint32_t n = min((lo + 15u) & ~15u, hi) - lo;
        if (tid < n) st.gout[lo + tid] = sh.ring[(lo + tid) & kTRingMask];
        lo += n;
    }
    const uint32_t hi_al = hi & ~15u;
    if (!(lo & 15u) && hi_al > lo) {
        const uint4* r4 = reinterpret_cast<const uint4*>(sh.ring);
        uint4* g4 = reinterpret_cast<uint4*>(st.gout);
        for (uint32_t u = (lo >> 4) + tid; u < (hi_al >> 4); u += kTThreads) g4[u] = r4[u & (kTRingMask >> 4)];
        lo = hi_al;
    }
    if (final && hi > lo) {
        if (tid < hi - lo) st.gout[lo + tid] = sh.ring[(lo + tid) & kTRingMask];
        lo = hi;
    }
    st.flushed = lo;
}

// ------------------------------------------------------------------------------------------------
// Slow step helpers: CTA-cooperative copies through the ring, flushed piece by piece.
// ------------------------------------------------------------------------------------------------
template <class Fmt>
__device__ __forceinline__ void tile_slow_literals(TileShared<Fmt>& sh, TileState& st, uint32_t p, uint32_t ll) {
    for (uint32_t done = 0; done < ll; done += kTPiece) {
        const uint32_t n = min(kTPiece, ll - done);
        const uint32_t d = st.op + done;
        for (uint32_t j = threadIdx.x; j < n; j += kTThreads) sh.ring[(d + j) & kTRingMask] = st.gin[p + done + j];
        __syncthreads();
        tile_flush(sh, st, d + n, false);
        __syncthreads();                                             // the next piece reuses ring slots the flush may still be reading
    }
    st.op += ll;
}
// out[op .. op+ml) = out[op-off ..]; a match that overlaps itself is periodic with period `off`.
template <class Fmt>
__device__ __forceinline__ void tile_slow_match(TileShared<Fmt>& sh, TileState& st, uint32_t off, uint32_t ml) {
    const uint32_t s = st.op - off;
    const bool periodic = off < ml;
    for (uint32_t done = 0; done < ml; done += kTPiece) {
        const uint32_t n = min(kTPiece, ml - done);
        const uint32_t d = st.op + done;
        const uint32_t hi = d + n;
        const uint32_t ring_lo = hi > kTRing ? hi - kTRing : 0u;      // older bytes have left the ring (they are flushed)
        for (uint32_t j = threadIdx.x; j < n; j += kTThreads) {
            const uint32_t k = done + j;
            const uint32_t sp = s + (periodic ? k % off : k);
            const uint8_t v = sp < ring_lo ? st.gout[sp] : sh.ring[sp & kTRingMask];
            sh.ring[(d + j) & kTRingMask] = v;
        }
        __syncthreads();
        tile_flush(sh, st, hi, false);
        __syncthreads();                                             // flushed bytes may be read back by the next piece
    }
    st.op += ml;
}

// One fully checked LZ4 sequence at st.ip (mirrors lz4_slow_sequence, lz4_decode_ring.cuh).
// Returns 0: continue, 1: stream finished, -1: corrupt.  Uniform across the CTA.
template <class Fmt>
__device__ __forceinline__ int tile_slow_step_lz4(TileShared<Fmt>& sh, TileState& st) {
    const uint8_t* in = st.gin;
    const uint32_t iend = st.iend, cap = st.cap;
    uint32_t ip = st.ip;
    if (ip >= iend) return -1;
    const uint32_t tok = in[ip];
    uint32_t ll = tok >> 4, p = ip + 1;
    if (ll == 15) {
        uint32_t b;
        do {
            if (p >= iend) return -1;
            b = in[p++]; ll += b;
        } while (b == 255 && ll < 0x7fff0000u);
    }
    if (ll > iend - p || ll > cap - st.op) return -1;
    // end-of-block parsing restrictions, lz4.c:4104-4164
    const bool closing = ((uint64_t)st.op + ll + 12 > cap) || ((uint64_t)p + ll + 8 > iend);
    if (closing && st.last && p + ll != iend) return -1;
    tile_slow_literals(sh, st, p, ll);
    const uint32_t q = p + ll;
    st.ip = q;
    if (closing && (st.last || st.op == cap)) return 1;
    if (q == iend) return 1;
    if (q + 2 > iend) return -1;
    const uint32_t off = (uint32_t)in[q] | ((uint32_t)in[q + 1] << 8);
    ip = q + 2;
    uint32_t ml = tok & 15;
    if (ml == 15) {
        uint32_t b;
        do {
            if (ip >= iend) return -1;
            b = in[ip++]; ml += b;
        } while (b == 255 && ml < 0x7fff0000u);
    }
    ml += 4;
    st.ip = ip;
    if (off == 0 || off > st.op - st.a) return -1;                   // lz4.c:4196-4197
    if (ml > cap - st.op) return -1;
    if (st.last && (uint64_t)st.op + ml + 5 > cap) return -1;        // lz4.c:4262-4264
    tile_slow_match(sh, st, off, ml);
    if (!st.last && (st.op == cap || ip >= iend)) return 1;          // lz4.c:4285-4288
    return 0;
}

// One fully checked Snappy element at st.ip (mirrors snappy_decode_warp, snappy_codec.cuh).
template <class Fmt>
__device__ __forceinline__ int tile_slow_step_snappy(TileShared<Fmt>& sh, TileState& st) {
    const uint8_t* in = st.gin;
    const uint32_t iend = st.iend, cap = st.cap;
    uint32_t ip = st.ip;
    if (ip >= iend) return st.op == cap ? 1 : -1;                    // snappy.cc:1715
    const uint32_t tag = in[ip++];
    uint32_t len, off;
    if ((tag & 3) == 0) {                                            // literal, snappy.cc:1492-1527
        len = (tag >> 2) + 1;
        if (len > 60) {
            const uint32_t nb = len - 60;
            if (ip + nb > iend) return -1;
            uint32_t v = 0;
            for (uint32_t k = 0; k < nb; k++) v |= (uint32_t)in[ip + k] << (8 * k);
            ip += nb;
            if (v == 0xffffffffu) return -1;
            len = v + 1;
        }
        if (len > iend - ip || len > cap - st.op) return -1;
        tile_slow_literals(sh, st, ip, len);
        st.ip = ip + len;
        return 0;
    }
    const uint32_t kind = tag & 3;                                   // char_table, snappy-internal.h:406-439
    if (kind == 1) {
        if (ip + 1 > iend) return -1;
        len = 4 + ((tag >> 2) & 7); off = ((tag >> 5) << 8) | in[ip]; ip += 1;
    } else if (kind == 2) {
        if (ip + 2 > iend) return -1;
        len = 1 + (tag >> 2); off = (uint32_t)in[ip] | ((uint32_t)in[ip + 1] << 8); ip += 2;
    } else {
        if (ip + 4 > iend) return -1;
        len = 1 + (tag >> 2); off = ld_u32_bytes(in + ip); ip += 4;
    }
    st.ip = ip;
    if (off == 0 || off > st.op - st.a || len > cap - st.op) return -1;     // snappy.cc:2185-2199
    tile_slow_match(sh, st, off, len);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Input chunks (TMA bulk copies, double buffered)
// ------------------------------------------------------------------------------------------------
template <class Fmt>
__device__ __forceinline__ void tile_wait_buf(TileShared<Fmt>& sh, TileState& st, int b) {
    if (st.pend & (1u << b)) {
        uint32_t spins = 0;
        while (!mbar_try_wait(&sh.bar[b], (st.par >> b) & 1u)) {
            if (++spins > kTSpinMax) { TWATCH(0); sh.abort = 1; break; }
        }
        st.par ^= 1u << b;
        st.pend &= ~(1u << b);
    }
}
// Caller guarantees (with a __syncthreads) that nobody still reads buffer b.
template <class Fmt>
__device__ __forceinline__ void tile_issue_buf(TileShared<Fmt>& sh, TileState& st, int b, int32_t chunk, uint32_t elected = 0) {
    if (st.pend & (1u << b)) {
        // At most one load per buffer in flight.  EVERY thread has to observe the old phase before thread 0
        // re-arms the barrier: a warp that polls late would otherwise find the barrier two phases on, read
        // its parity as "not completed yet" and spin until the watchdog (seen as an intermittent failure on
        // streams whose long literal runs make every group jump over an unconsumed prefetch).
        tile_wait_buf(sh, st, b);
        __syncthreads();
    }
    const uint32_t base = (uint32_t)chunk << kTChunkLog;
    st.bufc[b] = chunk;
    if (base >= st.iend) return;
    const uint32_t left = st.iend - base;
    const uint32_t bytes = left >= kTBuf ? kTBuf : ((left + 15u) & ~15u);
    if (threadIdx.x == elected) {
        fence_proxy_async();
        mbar_expect_tx(&sh.bar[b], bytes);
        bulk_g2s(sh.inbuf[b], st.gin + base, bytes, &sh.bar[b]);
    }
    st.pend |= 1u << b;
}

// ------------------------------------------------------------------------------------------------
// One group: the regular sequences that start in the current chunk, from st.ip up to the first
// irregular one.  Returns 1 when a slow step has to follow at st.ip, 0 when not, -1 on a watchdog abort.
// ------------------------------------------------------------------------------------------------
template <class Fmt>
__device__ __forceinline__ int tile_group(TileShared<Fmt>& sh, TileState& st, uint32_t fast_i_ex) {
    constexpr int MAXSEQ = Fmt::kMaxSeq;
    constexpr int kRounds = (MAXSEQ / 32 + kTWarps - 1) / kTWarps;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int32_t chunk = (int32_t)(st.ip >> kTChunkLog);
    const int b = chunk & 1;
    const uint32_t cbase = (uint32_t)chunk << kTChunkLog;
    const uint8_t* bp = sh.inbuf[b];

    TP_DECL
    __syncthreads();                                                 // previous phase is done with both buffers
    if (st.bufc[b] != chunk) tile_issue_buf(sh, st, b, chunk);
    tile_wait_buf(sh, st, b);
    TP(0);

    // ---- links: n1 = next token, then 2/4/8/16/32/64-hop links by pointer doubling.  n1, n8 (ta) and n64 (tb)
    //      have their own tables, n2/n4/n16/n32 live in the four quarters of the pointer table P, which is free
    //      until the copies start.  Every thread owns 8 positions; loads are issued together, stores afterwards.
    {
        constexpr int kPer = kTChunk / kTThreads;
        uint16_t* const S0 = sh.P, *const S1 = sh.P + kTChunk, *const S2 = sh.P + 2 * kTChunk, *const S3 = sh.P + 3 * kTChunk;
        uint32_t v[kPer];
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            const uint32_t i = tid + j * kTThreads;
            const TSeq sq = Fmt::parse(bp, i);               // may read stale bytes past the end of the stream: discarded below
            v[j] = (cbase + i < fast_i_ex && sq.nxt != kTNone && Fmt::ends_inside(sq, cbase, fast_i_ex)) ? sq.nxt : kTNone;
        }
#pragma unroll
        for (int j = 0; j < kPer; j++) sh.n1[tid + j * kTThreads] = (uint16_t)v[j];
        __syncthreads();
#define LLC_TILE_DOUBLE(FROM, TO)                                                                     \
        _Pragma("unroll") for (int j = 0; j < kPer; j++) v[j] = v[j] >= kTChunk ? kTNone : FROM[v[j]]; \
        _Pragma("unroll") for (int j = 0; j < kPer; j++) TO[tid + j * kTThreads] = (uint16_t)v[j];     \
        __syncthreads();
        LLC_TILE_DOUBLE(sh.n1, S0)                           // n2
        LLC_TILE_DOUBLE(S0, S1)                              // n4
        LLC_TILE_DOUBLE(S1, sh.ta)                           // n8
        LLC_TILE_DOUBLE(sh.ta, S2)                           // n16
        LLC_TILE_DOUBLE(S2, S3)                              // n32
        LLC_TILE_DOUBLE(S3, sh.tb)                           // n64
#undef LLC_TILE_DOUBLE
        TC(0, 1);
    }
    TP(1);

    // ---- prefetch of the next chunk, issued by a thread of warp 1 (idle during the chase); nobody has read the
    //      other buffer since the previous group
    if (st.bufc[b ^ 1] != chunk + 1 && cbase + kTChunk < st.iend) tile_issue_buf(sh, st, b ^ 1, chunk + 1, 32u);

    // ---- chase: one thread follows the 64-hop links from the chunk entry, then finds how many more regular
    //      sequences there are (< 64) by descending through the 32/16/8/4/2/1-hop links
    if (tid == 0) {
        const uint16_t* const S0 = sh.P, *const S1 = sh.P + kTChunk, *const S2 = sh.P + 2 * kTChunk, *const S3 = sh.P + 3 * kTChunk;
        uint32_t p = st.ip - cbase, na = 0, rem = 0, n;
        while (p < kTChunk && na < MAXSEQ / 64 - 1 && (n = sh.tb[p]) != kTNone) { sh.anchors[na++] = (uint16_t)p; p = n; }
        sh.anchors[na] = (uint16_t)p;                                // the tail starts here
        if (p < kTChunk && (n = S3[p]) != kTNone) { rem += 32; p = n; }
        if (p < kTChunk && (n = S2[p]) != kTNone) { rem += 16; p = n; }
        if (p < kTChunk && (n = sh.ta[p]) != kTNone) { rem += 8; p = n; }
        if (p < kTChunk && (n = S1[p]) != kTNone) { rem += 4; p = n; }
        if (p < kTChunk && (n = S0[p]) != kTNone) { rem += 2; p = n; }
        if (p < kTChunk && (n = sh.n1[p]) != kTNone) { rem += 1; p = n; }
        // kind 0: left the chunk (or the group is full), 1: irregular token at p
        sh.nanch = na; sh.ntail = rem; sh.end_kind = (p < kTChunk && sh.n1[p] == kTNone) ? 1u : 0u; sh.end_pos = p;
        sh.first_bad = 0xffffffffu;
    }
    sh.startbits[tid] = 0;                                           // kTRows == kTThreads
    if (tid == 0) sh.row2seq[0] = 0;
    __syncthreads();
    const uint32_t nanch = sh.nanch;
    const uint32_t nseq = nanch * 64u + sh.ntail;
    const uint32_t end_ip = cbase + sh.end_pos;
    const bool end_special = sh.end_kind != 0;
    TP(2);
    TC(1, 1); TC(2, nseq);
    if (sh.abort) return -1;
    if (nseq == 0) { st.ip = end_ip; return 1; }                     // irregular token right at st.ip
    TP(3);

    // ---- fields + lengths, scanned to output positions
    const uint32_t nbatch = (nseq + 31u) >> 5;
    uint32_t f_len[kRounds], f_src[kRounds], f_dl[kRounds];          // ll | ml<<16, off | lit<<16, output position
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
        const uint32_t bt = warp + r * kTWarps, k = bt * 32u + lane;
        uint32_t ll = 0, ml = 0, off = 0, lit = 0;
        if (k < nseq) {
            // sequence k = anchor (k / 64) + (k % 64) hops, taken as at most one 32/16/8/4/2/1-hop link each
            const uint16_t* const S0 = sh.P, *const S1 = sh.P + kTChunk, *const S2 = sh.P + 2 * kTChunk, *const S3 = sh.P + 3 * kTChunk;
            uint32_t p = sh.anchors[k >> 6];
            if (k & 32u) p = S3[p];
            if (k & 16u) p = S2[p];
            if (k & 8u) p = sh.ta[p];
            if (k & 4u) p = S1[p];
            if (k & 2u) p = S0[p];
            if (k & 1u) p = sh.n1[p];
            sh.seq_start[k] = (uint16_t)p;
            const TSeq s = Fmt::parse(bp, p); ll = s.ll; ml = s.ml; off = s.off; lit = s.lit;
        }
        const uint32_t len = ll + ml;
        const uint32_t incl = warp_incl_sum(len, lane);
        f_len[r] = ll | (ml << 16); f_src[r] = off | (lit << 16); f_dl[r] = incl - len;
        if (lane == 31 && bt < nbatch) sh.batch_tot[bt] = incl;
    }
    __syncthreads();
    if (warp == 0) {
        const uint32_t v0 = lane < nbatch ? sh.batch_tot[lane] : 0u;
        const uint32_t v1 = lane + 32u < nbatch ? sh.batch_tot[lane + 32] : 0u;
        const uint32_t s0 = warp_incl_sum(v0, lane);
        const uint32_t t0 = __shfl_sync(kFull, s0, 31);
        const uint32_t s1 = warp_incl_sum(v1, lane) + t0;
        if (lane < nbatch) sh.batch_tot[lane] = s0 - v0;
        if (lane + 32u < nbatch) sh.batch_tot[lane + 32] = s1 - v1;
    }
    __syncthreads();
    // the link tables are done with: P back to "nothing written"
#pragma unroll
    for (uint32_t j = 0; j < kTSpan / (8u * kTThreads); j++) *reinterpret_cast<uint4*>(&sh.P[(tid + j * kTThreads) * 8u]) = make_uint4(0, 0, 0, 0);
    // Span indices are relative to `base`, the 16-byte unit that holds the first byte of the group, so that
    // aligned units of the span are aligned units of the ring and of the output.
    const uint32_t op0 = st.op, base = op0 & ~15u, a0 = op0 - base;
    const uint32_t lim_o = st.cap >= op0 + Fmt::kEndSlack ? min(st.cap - Fmt::kEndSlack, base + kTSpan) : 0u;
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
        const uint32_t bt = warp + r * kTWarps, k = bt * 32u + lane;
        if (k < nseq) {
            const uint32_t ll = f_len[r] & 0xffffu, ml = f_len[r] >> 16, off = f_src[r] & 0xffffu;
            const uint32_t dlk = op0 + sh.batch_tot[bt] + f_dl[r];
            const uint32_t dm = dlk + ll, end = dm + ml;
            const uint32_t rel = dlk - base, erel = end - base;
            sh.sq[k] = make_uint2((rel & 0xffffu) | (ll << 16), f_src[r]);
            if (k == nseq - 1) sh.span_end = erel;
            const bool bad = (ml != 0 && (off == 0 || off > dm - st.a)) || end > lim_o;
            if (bad) atomicMin(&sh.first_bad, k);
            if (rel < kTSpan) {                                      // sequences beyond the span are never executed
                atomicOr(&sh.startbits[rel >> 5], 1u << (rel & 31u));
                for (uint32_t R = (rel + 31u) >> 5; (R << 5) < erel && R < kTRows; R++) sh.row2seq[R] = (uint16_t)(k + 1u);
            }
        }
    }
    __syncthreads();
    TP(4);
    const uint32_t nexec = min(nseq, sh.first_bad);
    TC(3, nexec);
    if (nexec == 0) return 1;                                        // first sequence is irregular: slow step at st.ip
    const uint32_t gend = nexec < nseq ? (sh.sq[nexec].x & 0xffffu) : sh.span_end;   // span index one past the group's last byte
    const uint32_t g_hi = base + gend;
    const uint32_t ring_lo = g_hi > kTRing ? g_hi - kTRing : 0u;     // older bytes have left the ring (they are flushed)
    const uint32_t nrows = (gend + 31u) >> 5;

    // ---- every output byte finds its source.  Warp w takes rows w, w + 16, ...; lane = byte of the row.
    //      The ring is only read here (sources older than the group); the group's own bytes collect in P and
    //      move to the ring in one piece afterwards.  A source inside the group is looked at right away: the
    //      warps walk the span front to back, so most earlier rows have been written, and whatever their
    //      entry holds by now -- a value or a pointer further back -- is as good as the source itself.
    //      No data-dependent branches: the literal and the old-source byte come through one load with a selected
    //      address, the look at P goes to entry a0 (harmless) when the byte does not need it.
    uint32_t unres = 0;                                              // bit i: my byte of row warp + 16 i still follows a pointer
    if constexpr (Fmt::kSetupPairs) {
        // two rows per trip, loads before stores (Snappy: 5.30 -> 5.11 ms per GiB; LZ4 prefers one row, whose
        // look-through sees more of the rows in front of it: 6.42 vs 6.66 ms)
        const uint32_t le_mask = (2u << lane) - 2u;
        uint32_t bit = 1u;
        for (uint32_t row0 = warp; row0 < nrows; row0 += 2u * kTWarps, bit <<= 2) {
            uint32_t xs[2], ks[2], es[2], ps[2], vs[2], pas[2];
            bool live[2], ing[2], far[2];
            uint2 qs[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const uint32_t row = row0 + (uint32_t)u * kTWarps;
                xs[u] = (row << 5) + lane;
                live[u] = row < nrows && xs[u] >= a0 && xs[u] < gend;
                ks[u] = 0;
                if (row < nrows) ks[u] = (uint32_t)sh.row2seq[row] + (uint32_t)__popc(sh.startbits[row] & le_mask) - 1u;
                if (!live[u]) ks[u] = 0;
            }
#pragma unroll
            for (int u = 0; u < 2; u++) qs[u] = sh.sq[ks[u]];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const uint32_t r = xs[u] - (qs[u].x & 0xffffu), ll = qs[u].x >> 16;
                pas[u] = base + xs[u] - (qs[u].y & 0xffffu);
                const bool is_lit = r < ll;
                ing[u] = live[u] && !is_lit && pas[u] >= op0;
                far[u] = live[u] && !is_lit && pas[u] < ring_lo;
                const uint8_t* vp = is_lit ? bp + (qs[u].y >> 16) + r : sh.ring + (pas[u] & kTRingMask);
                vs[u] = *vp;
                ps[u] = ing[u] ? pas[u] - base : a0;
                es[u] = sh.P[ps[u]];
            }
            if (__any_sync(kFull, far[0] || far[1])) {
                if (far[0]) vs[0] = __ldcg(st.gout + pas[0]);
                if (far[1]) vs[1] = __ldcg(st.gout + pas[1]);
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                uint32_t e = es[u];
                if (e == 0) e = kTPtr | ps[u];
                if (!ing[u]) e = kTKnown | vs[u];
                if (live[u]) sh.P[xs[u]] = (uint16_t)e;
                if (live[u] && e < kTKnown) unres |= bit << u;
            }
        }
    } else {
        const uint32_t le_mask = (2u << lane) - 2u;                  // bits 1 .. lane
        uint32_t bit = 1u;
        for (uint32_t row = warp; row < nrows; row += kTWarps, bit <<= 1) {
            const uint32_t x = (row << 5) + lane;
            const bool live = x >= a0 && x < gend;
            uint32_t k = (uint32_t)sh.row2seq[row] + (uint32_t)__popc(sh.startbits[row] & le_mask) - 1u;
            if (!live) k = 0;
            const uint2 q = sh.sq[k];
            const uint32_t r = x - (q.x & 0xffffu), ll = q.x >> 16;
            const uint32_t pa = base + x - (q.y & 0xffffu);          // match source (validated: >= st.a)
            const bool is_lit = r < ll;
            const bool in_group = live && !is_lit && pa >= op0;      // (live: pa of a dead lane may have wrapped)
            const bool far = live && !is_lit && pa < ring_lo;        // has left the ring: read back from L2 / HBM
            const uint8_t* vp = is_lit ? bp + (q.y >> 16) + r : sh.ring + (pa & kTRingMask);
            uint32_t v = *vp;
            if (__any_sync(kFull, far)) { if (far) v = __ldcg(st.gout + pa); }
            const uint32_t p = in_group ? pa - base : a0;
            uint32_t e = sh.P[p];
            if (e == 0) e = kTPtr | p;
            if (!in_group) e = kTKnown | v;
            if (live) sh.P[x] = (uint16_t)e;
            if (live && e < kTKnown) unres |= bit;
        }
    }
    TP(5);

    // ---- pointer jumping.  Invariant: byte x has the same final value as the byte P[x] points at, and
    //      pointers only point backwards, so every chain ends at a known byte; each round quarters the chains.
    TC(4, 1);
    while (__syncthreads_or(unres != 0)) {
        TC(4, 1);
        uint32_t m = unres;
        while (m) {
            const uint32_t i = (uint32_t)__ffs(m) - 1u, bit = m & (0u - m);
            m ^= bit;
            uint16_t* px = &sh.P[((warp + i * kTWarps) << 5) + lane];
            uint32_t q = sh.P[*px - kTPtr];
            if (q < kTKnown) q = sh.P[q - kTPtr];           // a second jump in the same round (6.63 -> 6.41 ms per GiB)
            *px = (uint16_t)q;
            if (q >= kTKnown) unres ^= bit;
        }
    }
    TP(6);
    if (sh.abort) return -1;

    // ---- P -> ring and HBM: aligned 16-byte units in one piece (32 bytes of P -> 16 bytes of ring and of the
    //      output), the partial unit at the head byte by byte together with what the previous flush left in the
    //      ring (< 16 bytes); a partial unit at the tail stays in the ring for the next flush.
    {
        const uint32_t u_lo = (a0 + 15u) >> 4, u_hi = gend >> 4;      // full units [u_lo, u_hi)
        uint4* const g4 = reinterpret_cast<uint4*>(st.gout + base);
        for (uint32_t u = u_lo + tid; u < u_hi; u += kTThreads) {
            const uint4 lo = *reinterpret_cast<const uint4*>(&sh.P[u << 4]);
            const uint4 hi = *reinterpret_cast<const uint4*>(&sh.P[(u << 4) + 8u]);
            uint4 o;
            o.x = __byte_perm(lo.x, lo.y, 0x6420); o.y = __byte_perm(lo.z, lo.w, 0x6420);
            o.z = __byte_perm(hi.x, hi.y, 0x6420); o.w = __byte_perm(hi.z, hi.w, 0x6420);
            *reinterpret_cast<uint4*>(&sh.ring[(base + (u << 4)) & kTRingMask]) = o;
            g4[u] = o;
        }
        const bool tail_kept = u_hi >= u_lo;                          // the tail starts at an aligned unit
        if (tid < 16u) {
            const uint32_t xh = tid;                                 // head unit (unit 0 when the group starts inside it)
            if (xh >= a0 && xh < min(gend, u_lo << 4)) {
                const uint8_t v = (uint8_t)sh.P[xh];
                sh.ring[(base + xh) & kTRingMask] = v;
                st.gout[base + xh] = v;
            }
            const uint32_t xt = (u_hi << 4) + tid;                   // tail unit
            if (xt >= max(a0, u_lo << 4) && xt < gend) sh.ring[(base + xt) & kTRingMask] = (uint8_t)sh.P[xt];
        } else if (tid < 32u) {
            const uint32_t pos = st.flushed + (tid - 16u);           // left in the ring by the previous flush
            if (pos < op0) st.gout[pos] = sh.ring[pos & kTRingMask];
        }
        st.flushed = tail_kept ? base + (u_hi << 4) : g_hi;
    }
    __syncthreads();
    st.op = g_hi;
    TP(7);
    if (nexec == nseq) { st.ip = end_ip; return end_special ? 1 : 0; }
    st.ip = cbase + sh.seq_start[nexec];
    return 1;
}

// ------------------------------------------------------------------------------------------------
// One unit.  Returns bytes produced or kErrCorrupt.  `last` selects the vanilla LZ4 end-of-block rules.
// ------------------------------------------------------------------------------------------------
template <class Fmt, bool SNAPPY>
__device__ __forceinline__ int64_t tile_decode_unit(TileShared<Fmt>& sh, uint32_t& par, const uint8_t* in,
                                                    uint32_t clen, uint8_t* out, uint32_t cap, bool last) {
    if (!SNAPPY) {
        if (clen == 0) return kErrCorrupt;
        if (cap == 0) return (clen == 1 && in[0] == 0) ? 0 : kErrCorrupt;      // lz4.c:3854-3858
    }
    TileState st;
    const uint32_t pad = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 15);
    st.gin = in - pad;
    st.iend = clen + pad;
    st.a = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15);
    st.gout = out - st.a;
    st.cap = st.a + min(cap, kTCapMax);
    st.ip = pad; st.op = st.a; st.flushed = st.a;
    st.par = par; st.pend = 0; st.bufc[0] = st.bufc[1] = -1;
    st.last = last;
    // Tokens take the fast path up to the end of the stream; the parse itself keeps the closing sequences
    // (and anything that would read past the end) for the checked step (Fmt::ends_inside).
    const uint32_t fast_i_ex = st.iend;
    int status = 0;
    while (status == 0) {
        int slow = 1;
        if (st.ip < fast_i_ex) slow = tile_group<Fmt>(sh, st, fast_i_ex);
        TP_DECL
        if (slow < 0) { status = -1; break; }
        if (slow) { status = SNAPPY ? tile_slow_step_snappy<Fmt>(sh, st) : tile_slow_step_lz4<Fmt>(sh, st); TC(5, 1); }
        TP(8);
    }
    __syncthreads();
    if (status > 0) tile_flush(sh, st, st.op, true);
    tile_wait_buf(sh, st, 0);                                        // nothing may stay in flight into the next unit
    tile_wait_buf(sh, st, 1);
    par = st.par;
    __syncthreads();
    return status > 0 ? (int64_t)(st.op - st.a) : kErrCorrupt;
}

template <class Fmt>
__device__ __forceinline__ void tile_init(TileShared<Fmt>& sh) {
    if (threadIdx.x == 0) {
        sh.abort = 0;
        mbar_init(&sh.bar[0], 1);
        mbar_init(&sh.bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
}

}  // namespace llc
