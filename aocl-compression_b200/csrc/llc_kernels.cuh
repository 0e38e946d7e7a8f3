// llc_kernels.cuh -- __global__ kernels of the RAP path: frame parse/scan, per-partition
// codec kernels, stitch/plan and compaction.  Launch logic lives in llc_device.cu.
#pragma once
#include "llc_common.cuh"
#include "lz4_codec.cuh"
#include "lz4_encode_lean.cuh"
#include "lz4_fastparse.cuh"
#include "snappy_encode_lean.cuh"
#include "snappy_codec.cuh"
#include "lz4_decode_ring.cuh"
#include "decode_tile.cuh"
#include "decode_rowq.cuh"

namespace llc {

// ------------------------------------------------------------------------------------------
// block-wide exclusive scan of one uint64 per thread (blockDim.x <= 1024, multiple of 32)
// ------------------------------------------------------------------------------------------
__device__ inline uint64_t block_excl_scan(uint64_t v, uint64_t* total, uint64_t* smem /* >= 33 entries */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    uint64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(kFull, x, d);
        if (lane >= d) x += t;
    }
    if (lane == 31) smem[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = lane < nwarp ? smem[lane] : 0;
        uint64_t y = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t t = __shfl_up_sync(kFull, y, d);
            if (lane >= d) y += t;
        }
        smem[lane] = y - w;                 // exclusive warp offsets
        if (lane == 31) smem[32] = y;       // grand total
    }
    __syncthreads();
    const uint64_t r = smem[warp] + x - v;
    if (total) *total = smem[32];
    __syncthreads();
    return r;
}

// ------------------------------------------------------------------------------------------
// Decompress step 1: parse the RAP frame (or recognise a frame-less stream), validate the
// entries and lay the partitions out in the output (exclusive scan of decomp_len).
// Replaces aocl_setup_parallel_decompress_mt / aocl_do_partition_decompress_mt
// (threads/threads.c:174-293) and the serial concatenation epilogues (lz4.c:4863-4881,
// snappy.cc:2351-2366): partitions are decoded straight to their final offsets.
// One CTA of 1024 threads.
// ------------------------------------------------------------------------------------------
// `check_total` = 0 for a ranged decode: the capacity then applies to the range (range_check_kernel), not to the
// whole stream; a frame-less stream is always bounded by out_cap.
__global__ void __launch_bounds__(1024) rap_parse_kernel(int codec, const uint8_t* __restrict__ in, uint64_t n,
                                                         uint64_t out_cap, int check_total, PartDesc* parts, CallResult* res) {
    __shared__ uint64_t sm[40];
    __shared__ uint32_t s_T, s_frame, s_bad;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_bad = 0;
        uint32_t T = 1, frame = 0;
        if (n >= 8 && ld_u64_bytes(in) == kRapMagic) {        // threads/threads.c:194-201
            if (n < 16) s_bad = 1;
            else {
                frame = ld_u32_bytes(in + 8);
                T = ld_u32_bytes(in + 12);
                if (T == 0 || T > kMaxPartitions || (uint64_t)kRapHeaderBytes + (uint64_t)kRapEntryBytes * T > n ||
                    frame > n)
                    s_bad = 1;                                  // threads/threads.c:208-209
            }
        }
        s_T = T; s_frame = frame;
    }
    __syncthreads();
    if (s_bad) { if (tid == 0) { res->error = 1; res->value = kErrCorrupt; res->parts = 0; } return; }
    const uint32_t T = s_T, frame = s_frame;

    if (frame == 0 && T == 1) {                                 // frame-less stream: one partition
        if (tid == 0) {
            PartDesc d;
            d.in_off = 0; d.in_len = (uint32_t)n; d.out_off = 0;
            d.flags = kPartLast;
            if (codec == 0) {
                d.out_len = (uint32_t)min(out_cap, (uint64_t)0xffffffffu);
                res->value = 0;
            } else {
                uint32_t total = 0;
                const uint32_t vb = get_varint32(in, n, &total);
                if (vb == 0 || total > out_cap) { res->error = 1; res->value = kErrCorrupt; res->parts = 0; return; }
                d.in_off = vb; d.in_len = (uint32_t)(n - vb); d.out_len = total; d.flags |= kPartExact;
                res->value = total;
            }
            parts[0] = d;
            res->parts = 1;
        }
        return;
    }
    // A stream with a one-entry frame (T == 1) is handled by the same table walk.
    const uint32_t per = (T + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = min(T, tid * per), hi = min(T, lo + per);
    uint64_t local = 0;
    bool bad = false;
    for (uint32_t i = lo; i < hi; i++) {
        const uint8_t* e = in + kRapHeaderBytes + (uint64_t)kRapEntryBytes * i;
        const uint32_t off = ld_u32_bytes(e), clen = ld_u32_bytes(e + 4), dlen = ld_u32_bytes(e + 8);
        if ((uint64_t)off + clen > n) bad = true;
        if (clen) local += dlen;                                // zero-length partitions are skipped (threads.c:264-268)
    }
    uint64_t total = 0;
    uint64_t base = block_excl_scan(local, &total, sm);
    for (uint32_t i = lo; i < hi; i++) {
        const uint8_t* e = in + kRapHeaderBytes + (uint64_t)kRapEntryBytes * i;
        PartDesc d;
        d.in_off = ld_u32_bytes(e); d.in_len = ld_u32_bytes(e + 4); d.out_len = ld_u32_bytes(e + 8);
        d.out_off = base;
        d.flags = kPartExact | (i == T - 1 ? kPartLast : 0u);
        if (d.in_len) base += d.out_len;
        parts[i] = d;
    }
    if (bad) atomicOr(&s_bad, 1u);
    __syncthreads();
    if (tid == 0) {
        bool fail = s_bad || (check_total && total > out_cap);
        if (codec != 0 && !fail) {                              // Snappy: varint(total) follows the frame
            uint32_t v = 0;
            const uint32_t vb = frame <= n ? get_varint32(in + frame, n - frame, &v) : 0;
            if (vb == 0 || v != total) fail = true;
        }
        res->parts = (int)T;
        res->value = fail ? kErrCorrupt : (long long)total;
        if (fail) res->error = 1;
    }
}

// The decode kernels of a partition range read the same three things from the result block: a sticky error
// (an earlier kernel of the call failed), the partition count, and -- because the host does not know how many
// partitions a device-resident stream has -- which decoder organisation serves this range: both organisations
// are launched and the one whose regime it is not returns at once.  Thread 0 reads, everybody takes the broadcast
// value, so a CTA never exits half way (another CTA of the same launch may raise res->error at any time).
struct RangeInfo { uint32_t first, n; bool run; };
__device__ __forceinline__ RangeInfo decode_range_info(const CallResult* res, uint32_t first, uint32_t count,
                                                       uint32_t min_units, uint32_t max_units) {
    __shared__ uint32_t s_info[2];
    if (threadIdx.x == 0) {
        const uint32_t T = (uint32_t)res->parts;
        const uint32_t end = min(T, first + min(count, T));
        const uint32_t n = end > first ? end - first : 0u;
        s_info[0] = n;
        s_info[1] = (res->error == 0 && n >= min_units && n <= max_units) ? 1u : 0u;
    }
    __syncthreads();
    RangeInfo r;
    r.first = first; r.n = s_info[0]; r.run = s_info[1] != 0;
    return r;
}

// ------------------------------------------------------------------------------------------
// Decompress step 2: persistent grid, one warp per partition, partitions handed out through an
// atomic ticket so long partitions do not serialise a CTA.  Decodes partitions
// [first, first+count) of the table and writes each at out + (out_off - origin).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) decode_parts_kernel(int codec, const uint8_t* __restrict__ in, uint8_t* out,
                                                           const PartDesc* __restrict__ parts, CallResult* res,
                                                           uint32_t first, uint32_t count, uint64_t origin) {
    __shared__ RingStorage ring_mem[4];
    const int lane = lane_id();
    const RangeInfo ri = decode_range_info(res, first, count, 0u, 0xffffffffu);
    if (!ri.run) return;
    Ring ring;
    ring.init(&ring_mem[threadIdx.x >> 5], lane);
    const uint32_t end = first + ri.n;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = first + atomicAdd(&res->next, 1u);
        i = __shfl_sync(kFull, i, 0);
        if (i >= end) break;
        const PartDesc d = parts[i];
        if (d.in_len == 0) continue;
        uint8_t* dst = out + (d.out_off - origin);
        int64_t got;
        if (codec == 0) got = lz4_decode_warp_ring(ring, in + d.in_off, d.in_len, dst, d.out_len, (d.flags & kPartLast) != 0, lane);
        else            got = snappy_decode_warp(in + d.in_off, d.in_len, dst, d.out_len, lane);
        if (lane == 0) {
            if (got < 0 || ((d.flags & kPartExact) && (uint64_t)got != d.out_len)) atomicCAS(&res->error, 0, (int)i + 1);
            else if (!(d.flags & kPartExact)) res->value = got;     // frame-less LZ4: size is whatever was produced
        }
    }
}

// Tile variant (decode_tile.cuh): one 512-thread CTA per partition, lane per sequence for the parse, thread
// per byte for the copies, 32 KiB output window in shared memory; partitions handed out through the atomic ticket.
// Serves ranges of fewer than `max_units + 1` partitions (too few units per SM for the row decoder).
template <class Fmt, bool SNAPPY>
__global__ void __launch_bounds__(kTThreads, 2) decode_parts_tile_kernel(const uint8_t* __restrict__ in, uint8_t* out,
                                                                        const PartDesc* __restrict__ parts, CallResult* res,
                                                                        uint32_t first, uint32_t count, uint64_t origin,
                                                                        uint32_t max_units) {
    extern __shared__ __align__(128) uint8_t tile_smem[];
    TileShared<Fmt>& sh = *reinterpret_cast<TileShared<Fmt>*>(tile_smem);
    const RangeInfo ri = decode_range_info(res, first, count, 0u, max_units);
    if (!ri.run) return;
    tile_init(sh);
    uint32_t par = 0;
    const uint32_t end = first + ri.n;
    for (;;) {
        if (threadIdx.x == 0) sh.unit = first + atomicAdd(&res->next, 1u);
        __syncthreads();
        const uint32_t i = sh.unit;
        __syncthreads();
        if (i >= end) break;
        const PartDesc d = parts[i];
        if (d.in_len == 0) continue;
        const int64_t got = tile_decode_unit<Fmt, SNAPPY>(sh, par, in + d.in_off, d.in_len, out + (d.out_off - origin), d.out_len,
                                                          (d.flags & kPartLast) != 0);
        if (threadIdx.x == 0) {
            if (got < 0 || ((d.flags & kPartExact) && (uint64_t)got != d.out_len)) atomicCAS(&res->error, 0, (int)i + 1);
            else if (!(d.flags & kPartExact)) res->value = got;     // frame-less LZ4: size is whatever was produced
        }
    }
}

template <class Fmt, bool SNAPPY>
__global__ void __launch_bounds__(kTThreads, 2) decode_pages_tile_kernel(const uint8_t* const* __restrict__ in_ptrs,
                                                                        const uint32_t* __restrict__ in_sizes, uint8_t* const* out_ptrs,
                                                                        const uint32_t* __restrict__ out_caps, long long* status,
                                                                        uint64_t count, CallResult* res) {
    extern __shared__ __align__(128) uint8_t tile_smem[];
    TileShared<Fmt>& sh = *reinterpret_cast<TileShared<Fmt>*>(tile_smem);
    tile_init(sh);
    uint32_t par = 0;
    for (uint64_t i = blockIdx.x; i < count; i += gridDim.x) {
        const uint8_t* in = in_ptrs[i];
        const uint32_t n = in_sizes[i], cap = out_caps[i];
        int64_t got;
        if (!SNAPPY) got = tile_decode_unit<Fmt, false>(sh, par, in, n, out_ptrs[i], cap, true);
        else {
            uint32_t total = 0;
            const uint32_t vb = get_varint32(in, n, &total);
            if (vb == 0 || total > cap) got = kErrCorrupt;
            else got = tile_decode_unit<Fmt, true>(sh, par, in + vb, n - vb, out_ptrs[i], total, true);
        }
        if (threadIdx.x == 0) {
            status[i] = got;
            if (got < 0) atomicAdd(&res->error, 1);
        }
        __syncthreads();
    }
}

// Row variant (decode_rowq.cuh): one CTA per SM, 28 partitions in flight per CTA (one parser lane + one copier
// warp each).  Serves ranges of at least `min_units` partitions.
template <bool SNAPPY>
__global__ void __launch_bounds__(kQThreads, 1) decode_parts_rowq_kernel(const uint8_t* __restrict__ in, uint8_t* out,
                                                                        const PartDesc* __restrict__ parts, CallResult* res,
                                                                        uint32_t first, uint32_t count, uint64_t origin,
                                                                        uint32_t min_units) {
    extern __shared__ __align__(128) uint8_t rowq_smem[];
    QShared& sh = *reinterpret_cast<QShared*>(rowq_smem);
    const RangeInfo ri = decode_range_info(res, first, count, min_units, 0xffffffffu);
    if (!ri.run) return;
    QPartsSource src;
    src.in = in; src.out = out; src.parts = parts; src.res = res; src.first = first; src.origin = origin;
    rowq_run<SNAPPY>(sh, src, ri.n, &res->next);
}

template <bool SNAPPY>
__global__ void __launch_bounds__(kQThreads, 1) decode_pages_rowq_kernel(const uint8_t* const* __restrict__ in_ptrs,
                                                                        const uint32_t* __restrict__ in_sizes, uint8_t* const* out_ptrs,
                                                                        const uint32_t* __restrict__ out_caps, long long* status,
                                                                        uint32_t count, CallResult* res) {
    extern __shared__ __align__(128) uint8_t rowq_smem[];
    QShared& sh = *reinterpret_cast<QShared*>(rowq_smem);
    QPagesSource<SNAPPY> src;
    src.in_ptrs = in_ptrs; src.in_sizes = in_sizes; src.out_ptrs = out_ptrs; src.out_caps = out_caps; src.status = status; src.res = res;
    rowq_run<SNAPPY>(sh, src, count, &res->next);
}

// Ranged decode (a rank decoding its share of a frame, aocl_gpu_decompress_range_async): runs BEFORE the decode
// kernels.  The RAP entries are untrusted, so every partition of the range must land inside the caller's buffer
// [out, out + out_cap) once shifted by `origin`; otherwise the call fails and the decode kernels return at once.
// Also replaces the stream total by the byte count of the range.
__global__ void __launch_bounds__(256) range_check_kernel(const PartDesc* __restrict__ parts, CallResult* res, uint32_t first,
                                                          uint32_t count, uint64_t origin, uint64_t out_cap) {
    __shared__ unsigned long long acc;
    __shared__ int bad;
    if (threadIdx.x == 0) { acc = 0; bad = 0; }
    __syncthreads();
    if (res->error == 0) {                                    // (read-only here: uniform for the whole CTA)
        const uint32_t T = (uint32_t)res->parts;
        const uint32_t end = min(T, first + min(count, T));
        unsigned long long sum = 0;
        bool oob = false;
        for (uint32_t i = first + threadIdx.x; i < end; i += blockDim.x) {
            const PartDesc d = parts[i];
            if (!d.in_len) continue;
            sum += d.out_len;
            if (d.out_off < origin || d.out_off - origin > out_cap || (uint64_t)d.out_len > out_cap - (d.out_off - origin)) oob = true;
        }
        for (int k = 16; k; k >>= 1) sum += __shfl_down_sync(kFull, sum, k);
        if ((threadIdx.x & 31) == 0) atomicAdd(&acc, sum);
        if (oob) bad = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0 && res->error == 0) {
        if (bad) { res->error = 1; res->value = kErrCorrupt; }
        else res->value = (long long)acc;
    }
}

// ------------------------------------------------------------------------------------------
// Batched pages: one warp per independent frame-less page.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) decode_pages_kernel(int codec, const uint8_t* const* __restrict__ in_ptrs,
                                                           const uint32_t* __restrict__ in_sizes, uint8_t* const* out_ptrs,
                                                           const uint32_t* __restrict__ out_caps, long long* status,
                                                           uint64_t count, CallResult* res) {
    __shared__ RingStorage ring_mem[4];
    const int lane = lane_id();
    Ring ring;
    ring.init(&ring_mem[threadIdx.x >> 5], lane);
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < count; i += warps) {
        const uint8_t* in = in_ptrs[i];
        const uint32_t n = in_sizes[i], cap = out_caps[i];
        int64_t got;
        if (codec == 0) got = lz4_decode_warp_ring(ring, in, n, out_ptrs[i], cap, true, lane);
        else {
            uint32_t total = 0;
            const uint32_t vb = get_varint32(in, n, &total);
            if (vb == 0 || total > cap) got = kErrCorrupt;
            else got = snappy_decode_warp(in + vb, n - vb, out_ptrs[i], total, lane);
        }
        if (lane == 0) {
            status[i] = got;
            if (got < 0) atomicAdd(&res->error, 1);
        }
    }
}

// One warp per block (the hash table fills the block's dynamic shared memory).
__global__ void __launch_bounds__(32) encode_pages_kernel(int codec, const uint8_t* const* __restrict__ in_ptrs,
                                                          const uint32_t* __restrict__ in_sizes, uint8_t* const* out_ptrs,
                                                          const uint32_t* __restrict__ out_caps, long long* status,
                                                          uint64_t count, CallResult* res) {
    extern __shared__ __align__(16) uint32_t tab_mem[];
    __shared__ __align__(16) uint8_t own_mem[kLeanOwnBytes];   // LZ4: owner bytes; Snappy: claim bits
    const int lane = lane_id();
    for (uint64_t i = blockIdx.x; i < count; i += gridDim.x) {
        const uint8_t* in = in_ptrs[i];
        const uint32_t n = in_sizes[i], cap = out_caps[i];
        uint8_t* dst = out_ptrs[i];
        int64_t got;
        if (codec == 0) {
            const uint64_t bound = (uint64_t)n + n / 255 + 16;
            InGate gate(nullptr, nullptr);
            got = (cap == 0) ? 0 : lz4_encode_unit(in, n, dst, cap >= bound ? -1 : (int64_t)cap, true, nullptr, tab_mem, own_mem, lane, gate);
            if (got == 0) got = kErrCorrupt;
        } else if ((uint64_t)cap < 32ull + n + n / 6) {
            got = kErrCorrupt;                                  // api/codec.cpp:262-265
        } else {
            uint32_t op = 0;
            if (lane == 0) op = put_varint32(dst, n);
            op = __shfl_sync(kFull, op, 0);
            for (uint32_t p = 0; p < n; p += kSnappyBlock) {
                op += snappy_encode_fragment_lean(in + p, min(kSnappyBlock, n - p), dst + op,
                                                  reinterpret_cast<uint16_t*>(tab_mem), reinterpret_cast<uint32_t*>(own_mem), lane);
                __syncwarp();
            }
            got = op;
        }
        if (lane == 0) {
            status[i] = got;
            if (got < 0) atomicAdd(&res->error, 1);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// LZ4 compress.  Step 1: one warp per RAP partition encodes into its scratch slot.
// Replaces the `#pragma omp parallel` region of AOCL_LZ4_compress_fast_mt (lz4.c:2684-2731).
// ------------------------------------------------------------------------------------------
// What the stitch needs to know about a partition.  The first token of the body travels with the sizes, so that the
// stitch plan can be computed by a GPU that does not hold the body (one frame sharded over several GPUs: the records
// are the only thing the ranks exchange).
struct Lz4Rec {
    uint32_t body_len;     // bytes in the scratch slot (0: the partition is all literals)
    uint32_t tail_len;     // trailing literals left to the next partition
    uint32_t first_ll;     // literal length of the first sequence of the body
    uint32_t skip_tok;     // bytes of its token + length bytes | first token << 16
};

// Persistent: CTAs (one warp each) draw partitions from a shared ticket.  Two flavours run
// CONCURRENTLY on two streams: the shared-memory flavour is capped at 14 warps per SM by its
// 16 KiB table, and leaves most issue slots idle because every warp is a serial dependency chain;
// the global-memory flavour keeps its table in an L2-resident workspace slice (no shared memory),
// so its warps fill the remaining warp slots of each SM.  Both produce identical bytes.
// The launch covers partitions [p0, p0 + cnt) of the frame; `src` points at the first byte of partition p0 and
// scratch slot t belongs to partition p0 + t (single GPU: p0 = 0, cnt = T).
struct Lz4Range { uint64_t n; uint32_t T, p0, cnt; };
template <bool FAST>
__device__ __forceinline__ void lz4_encode_parts_loop(const uint8_t* __restrict__ src, Lz4Range g,
                                                      uint8_t* scratch, uint64_t slot, Lz4Rec* rec, uint32_t* ticket,
                                                      uint32_t* tab_mem, uint8_t* own, const uint32_t* in_flag, CallResult* res) {
    const int lane = lane_id();
    InGate gate(in_flag, &res->error);                       // watermark: bytes present in every partition
    const uint64_t common = g.n / g.T, left = g.n % g.T;     // threads/threads.c:91-97,127-135
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u);
        t = __shfl_sync(kFull, t, 0);
        if (t >= g.cnt) break;
        const uint32_t i = g.p0 + t;
        const uint32_t pn = (uint32_t)(common + (i == g.T - 1 ? left : 0));
        uint32_t tail = 0;
        uint8_t* const body_at = scratch + slot * t;
        const uint32_t body = FAST ? lz4_fastparse_unit(src + common * t, pn, body_at, i == g.T - 1, &tail, tab_mem, lane, gate)
                                   : lz4_encode_unit(src + common * t, pn, body_at, -1, i == g.T - 1, &tail, tab_mem, own, lane, gate);
        __syncwarp();                                        // the body bytes (written by all lanes) are ordered before the reads below
        // the first token of the body travels with the sizes.  Warp-uniform on purpose (every lane reads the same
        // bytes): a data-dependent loop under `if (lane == 0)` makes the compiler give up the convergence guarantee
        // for the whole partition loop (88 WARPSYNCs in the encoder, +8 % kernel time).
        uint32_t first_ll = 0, skip_tok = 0;
        if (body) {
            const uint32_t tok = __ldcg(body_at);
            uint32_t ll = tok >> 4, skip = 1;
            if (ll == 15) { uint32_t x; do { x = __ldcg(body_at + skip); skip++; ll += x; } while (x == 255); }
            first_ll = ll; skip_tok = skip | (tok << 16);
        }
        if (lane == 0) { Lz4Rec r; r.body_len = body; r.tail_len = tail; r.first_ll = first_ll; r.skip_tok = skip_tok; rec[i] = r; }
        __syncwarp();
    }
}
__global__ void __launch_bounds__(32, 28) lz4_encode_parts_kernel(const uint8_t* __restrict__ src, Lz4Range g,
                                                              uint8_t* scratch, uint64_t slot, Lz4Rec* rec, uint32_t* ticket,
                                                              const uint32_t* in_flag, CallResult* res) {
    extern __shared__ __align__(16) uint32_t tab_mem[];
    __shared__ uint8_t own_mem[kLeanOwnBytes];
    lz4_encode_parts_loop<false>(src, g, scratch, slot, rec, ticket, tab_mem, own_mem, in_flag, res);
}
__global__ void __launch_bounds__(32, 28) lz4_encode_parts_gtab_kernel(const uint8_t* __restrict__ src, Lz4Range g,
                                                                   uint8_t* scratch, uint64_t slot, Lz4Rec* rec,
                                                                   uint32_t* ticket, uint32_t* tables,
                                                                   const uint32_t* in_flag, CallResult* res) {
    __shared__ uint8_t own_mem[kLeanOwnBytes];
    lz4_encode_parts_loop<false>(src, g, scratch, slot, rec, ticket, tables + (size_t)blockIdx.x * 4096, own_mem, in_flag, res);
}
// The named fastparse mode (lz4_fastparse.cuh): same units, same records, same stitch; the tables live in an
// L2-resident workspace slice so that every partition of a 1 GiB frame is resident at once (the parse is latency bound).
__global__ void __launch_bounds__(32, 28) lz4_fastparse_parts_kernel(const uint8_t* __restrict__ src, Lz4Range g,
                                                                 uint8_t* scratch, uint64_t slot, Lz4Rec* rec,
                                                                 uint32_t* ticket, uint32_t* tables,
                                                                 const uint32_t* in_flag, CallResult* res) {
    lz4_encode_parts_loop<true>(src, g, scratch, slot, rec, ticket, tables + (size_t)blockIdx.x * 4096, nullptr, in_flag, res);
}

// Frame-less block written straight to the destination (T == 1, lz4.c:2674-2677).
__global__ void __launch_bounds__(32) lz4_encode_single_kernel(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst,
                                                               long long cap, CallResult* res) {
    extern __shared__ __align__(16) uint32_t tab_mem[];
    const int lane = lane_id();
    __shared__ uint8_t own_mem[kLeanOwnBytes];
    InGate gate(nullptr, nullptr);
    const uint32_t got = lz4_encode_unit(src, n, dst, cap, true, nullptr, tab_mem, own_mem, lane, gate);
    if (lane == 0) {
        if (got == 0) { res->error = 1; res->value = kErrCorrupt; }
        else res->value = got;
    }
}

// Step 2: stitch plan.  Works out, for every partition, the literal carry it inherits from its
// predecessors (a segmented sum: an all-literal partition forwards its predecessor's carry,
// lz4.c:2808-2822), the re-encoded first token, its final offset and length, and writes the RAP
// frame.  Replaces the serial loop lz4.c:2736-2905.  One CTA of 1024 threads.
struct Lz4Plan {
    uint32_t out_off;     // absolute offset in the destination stream
    uint32_t out_len;     // bytes this partition contributes (0 for an all-literal partition)
    uint32_t skip;        // bytes of the scratch body replaced by the new header
    uint32_t new_ll;      // literal length of the re-encoded first sequence
    uint32_t carry;       // inherited literal bytes
    uint32_t token_low;   // match-length nibble of the first token
    uint64_t carry_src;   // source offset of the inherited literals
};

// `dst` == nullptr: plan only (a rank that does not own the head of the stream).
__global__ void __launch_bounds__(1024) lz4_stitch_plan_kernel(const Lz4Rec* __restrict__ rec, uint64_t n, uint32_t T,
                                                               uint8_t* dst, uint64_t out_cap, Lz4Plan* plan,
                                                               CallResult* res) {
    __shared__ uint64_t sm[40];
    __shared__ uint32_t s_carry_val[1024];
    __shared__ uint32_t s_carry_flag[1024];
    const int tid = threadIdx.x;
    const uint32_t per = (T + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = min(T, tid * per), hi = min(T, lo + per);
    const uint64_t common = n / T, left = n % T;

    // pass A: segmented inclusive sum of tails; a partition with a body starts a new segment.
    // (value, flag) pairs combine as (f2 ? v2 : v1 + v2, f1 | f2).
    uint32_t v = 0, f = 0;
    for (uint32_t i = lo; i < hi; i++) {
        const bool has_body = rec[i].body_len != 0;
        v = has_body ? rec[i].tail_len : v + rec[i].tail_len;
        f |= has_body ? 1u : 0u;
    }
    s_carry_val[tid] = v; s_carry_flag[tid] = f;
    __syncthreads();
    // carry entering this thread's chunk = segmented sum over all earlier chunks (serial over <= 1024
    // chunk summaries is cheap, but do it in log steps anyway)
    for (int d = 1; d < (int)blockDim.x; d <<= 1) {
        uint32_t pv = 0, pf = 0;
        if (tid >= d) { pv = s_carry_val[tid - d]; pf = s_carry_flag[tid - d]; }
        __syncthreads();
        if (tid >= d) {
            if (!s_carry_flag[tid]) s_carry_val[tid] += pv;
            s_carry_flag[tid] |= pf;
        }
        __syncthreads();
    }
    uint32_t carry = tid ? s_carry_val[tid - 1] : 0;         // inclusive result of the previous chunk

    // pass B: per-partition header arithmetic and output length
    uint64_t local = 0;
    for (uint32_t i = lo; i < hi; i++) {
        const Lz4Rec r = rec[i];
        Lz4Plan p;
        p.carry = carry;
        const uint32_t pn = (uint32_t)(common + (i == T - 1 ? left : 0));
        p.carry_src = common * i - carry;                     // tails are contiguous in the source
        if (r.body_len == 0) {                                // all-literal partition
            p.out_len = 0; p.skip = 0; p.new_ll = 0; p.token_low = 0;
            carry += r.tail_len;
        } else {
            const uint32_t tok = r.skip_tok >> 16, skip = r.skip_tok & 0xffffu, ll = r.first_ll;
            const uint32_t nl = ll + carry;
            const uint32_t hdr = 1 + (nl >= 15 ? (nl - 15) / 255 + 1 : 0);
            p.skip = skip; p.new_ll = nl; p.token_low = tok & 15;
            p.out_len = hdr + carry + (r.body_len - skip);
            carry = r.tail_len;
        }
        (void)pn;
        p.out_off = 0;
        plan[i] = p;
        local += p.out_len;
    }
    uint64_t total = 0;
    const uint64_t frame = (uint64_t)kRapHeaderBytes + (uint64_t)kRapEntryBytes * T;
    uint64_t base = frame + block_excl_scan(local, &total, sm);
    total += frame;
    const bool fits = total <= out_cap && total <= 0xffffffffull;

    // pass C: offsets + RAP entries (lz4.c:2763-2780, 2879-2896)
    carry = tid ? s_carry_val[tid - 1] : 0;
    for (uint32_t i = lo; i < hi; i++) {
        const Lz4Rec r = rec[i];
        const uint32_t pn = (uint32_t)(common + (i == T - 1 ? left : 0));
        plan[i].out_off = (uint32_t)base;
        const uint32_t olen = plan[i].out_len;
        uint32_t dlen;
        if (r.body_len == 0) { dlen = 0; carry += r.tail_len; }
        else { dlen = pn - r.tail_len + carry; carry = r.tail_len; }
        if (fits && dst) {
            uint8_t* e = dst + kRapHeaderBytes + (uint64_t)kRapEntryBytes * i;
            st_u32_bytes(e, (uint32_t)base); st_u32_bytes(e + 4, olen); st_u32_bytes(e + 8, dlen);
        }
        base += olen;
    }
    if (tid == 0) {
        if (fits) {
            if (dst) {
                st_u32_bytes(dst, (uint32_t)kRapMagic); st_u32_bytes(dst + 4, (uint32_t)(kRapMagic >> 32));
                st_u32_bytes(dst + 8, (uint32_t)frame); st_u32_bytes(dst + 12, T);    // threads/threads.c:105-110
            }
            res->value = (long long)total;
        } else { res->error = 1; res->value = kErrCorrupt; }
    }
}

// dst-aligned word copy with arbitrary source alignment (block-cooperative)
__device__ inline void block_copy(uint8_t* dst, const uint8_t* src, uint32_t len) {
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t head = min(len, (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3));
    if (tid < head) dst[tid] = src[tid];
    const uint32_t words = (len - head) >> 2;
    uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + head);
    const uint8_t* s = src + head;
    const uintptr_t sa = reinterpret_cast<uintptr_t>(s);
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(sa & ~uintptr_t(3));
    const unsigned sh = (unsigned)(sa & 3) * 8;
    if (sh == 0) for (uint32_t i = tid; i < words; i += nt) d4[i] = s4[i];
    else for (uint32_t i = tid; i < words; i += nt) d4[i] = __funnelshift_r(s4[i], s4[i + 1], sh);
    const uint32_t done = head + (words << 2);
    if (done + tid < len) dst[done + tid] = src[done + tid];
}

// Step 3: compaction.  CTA t writes [re-encoded first token][inherited literals][rest of body] of partition
// p0 + t at its final offset.  Replaces the memcpy chain of lz4.c:2825-2877.  `src` holds the input from byte
// `src_off` of the frame on; inherited literals that start before it (the tail of the previous GPU's last
// partitions) come from `halo`, which holds the `halo_len` input bytes in front of src_off.  The destination
// `dst` corresponds to stream offset `dst_off`.
__global__ void __launch_bounds__(256) lz4_compact_kernel(const uint8_t* __restrict__ src, uint64_t src_off,
                                                          const uint8_t* __restrict__ halo, uint64_t halo_len,
                                                          const uint8_t* __restrict__ scratch,
                                                          uint64_t slot, const Lz4Rec* __restrict__ rec,
                                                          const Lz4Plan* __restrict__ plan, uint32_t p0, uint8_t* dst,
                                                          uint64_t dst_off, const CallResult* res) {
    if (res->error) return;
    const uint32_t t = blockIdx.x, i = p0 + t;
    const Lz4Plan p = plan[i];
    if (p.out_len == 0) return;
    uint8_t* o = dst + (p.out_off - dst_off);
    const uint32_t nl = p.new_ll;
    const uint32_t ext = nl >= 15 ? (nl - 15) / 255 + 1 : 0;
    if (threadIdx.x == 0) o[0] = (uint8_t)((min(nl, 15u) << 4) | p.token_low);
    for (uint32_t j = threadIdx.x; j < ext; j += blockDim.x) o[1 + j] = (j + 1 < ext) ? (uint8_t)255 : (uint8_t)((nl - 15) % 255);
    uint8_t* lit = o + 1 + ext;
    uint32_t before = 0;                                     // carry bytes that lie in front of this GPU's slice
    if (p.carry_src < src_off) {
        before = (uint32_t)min((uint64_t)p.carry, src_off - p.carry_src);
        block_copy(lit, halo + (halo_len - (src_off - p.carry_src)), before);
    }
    block_copy(lit + before, src + (p.carry_src + before - src_off), p.carry - before);
    block_copy(lit + p.carry, scratch + slot * t + p.skip, rec[i].body_len - p.skip);
}

// One frame over several GPUs: what a rank learns from the plan -- the byte range of the final stream it writes
// and how many input bytes in front of its slice its partitions inherit as literals.
struct ShardInfo { unsigned long long halo, out_lo, out_hi, total; };
__global__ void __launch_bounds__(256) lz4_shard_info_kernel(const Lz4Plan* __restrict__ plan, uint64_t n, uint32_t T, uint32_t p0,
                                                             uint32_t cnt, const CallResult* res, ShardInfo* info) {
    __shared__ unsigned long long s_halo;
    if (threadIdx.x == 0) s_halo = 0;
    __syncthreads();
    const uint64_t common = n / T, src_off = common * p0;
    unsigned long long h = 0;
    for (uint32_t t = threadIdx.x; t < cnt; t += blockDim.x) {
        const Lz4Plan p = plan[p0 + t];
        if (p.out_len && p.carry_src < src_off) h = max(h, (unsigned long long)(src_off - p.carry_src));
    }
    atomicMax(&s_halo, h);
    __syncthreads();
    if (threadIdx.x == 0) {
        info->halo = s_halo;
        info->out_lo = p0 == 0 ? 0ull : plan[p0].out_off;
        info->out_hi = (unsigned long long)plan[p0 + cnt - 1].out_off + plan[p0 + cnt - 1].out_len;
        info->total = res->error ? 0ull : (unsigned long long)res->value;
    }
}

// ------------------------------------------------------------------------------------------
// Snappy compress.  Step 1: one warp per <=64 KiB fragment (the unit AOCL_CompressFragment works
// on; snappy.cc:1762-1818 runs them back to back inside each partition).
// ------------------------------------------------------------------------------------------
struct SnappyGeom {
    uint64_t n; uint32_t T; uint32_t frags_common; uint32_t frags_total; uint64_t common; uint64_t left;
    uint32_t f0, fcnt; uint64_t src_off;                     // the fragment range a launch works on
};
__host__ __device__ inline SnappyGeom snappy_geom(uint64_t n, uint32_t T) {
    SnappyGeom g;
    g.n = n; g.T = T; g.common = n / T; g.left = n % T;
    g.frags_common = (uint32_t)((g.common + kSnappyBlock - 1) / kSnappyBlock);
    const uint32_t last = (uint32_t)((g.common + g.left + kSnappyBlock - 1) / kSnappyBlock);
    g.frags_total = (T - 1) * g.frags_common + last;
    g.f0 = 0; g.fcnt = g.frags_total; g.src_off = 0;
    return g;
}
__device__ __forceinline__ void snappy_locate(const SnappyGeom& g, uint32_t f, uint32_t* part, uint64_t* off, uint32_t* len) {
    uint32_t p = g.frags_common ? f / g.frags_common : 0;
    if (p >= g.T) p = g.T - 1;
    const uint32_t j = f - p * g.frags_common;
    const uint64_t pn = g.common + (p == g.T - 1 ? g.left : 0);
    const uint64_t o = (uint64_t)j * kSnappyBlock;
    *part = p; *off = g.common * p + o; *len = (uint32_t)min((uint64_t)kSnappyBlock, pn - o);
}

// The launch covers fragments [g.f0, g.f0 + g.fcnt) of the frame; `src` holds the input from byte g.src_off on and
// scratch slot k belongs to fragment f0 + k (single GPU: f0 = 0, fcnt = frags_total, src_off = 0).
__device__ __forceinline__ void snappy_encode_frags_loop(const uint8_t* __restrict__ src, const SnappyGeom& g, uint8_t* scratch,
                                                         uint64_t slot, uint32_t* frag_len, uint32_t* ticket, uint16_t* tab,
                                                         uint32_t* claim, const uint32_t* in_flag, CallResult* res) {
    const int lane = lane_id();
    InGate gate(in_flag, &res->error);                       // watermark: bytes present from the start of the input
    for (;;) {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(ticket, 1u);
        k = __shfl_sync(kFull, k, 0);
        if (k >= g.fcnt) break;
        const uint32_t f = g.f0 + k;
        uint32_t part, len; uint64_t off;
        snappy_locate(g, f, &part, &off, &len);
        gate.wait((uint32_t)(off + len));
        const uint32_t got = snappy_encode_fragment_lean(src + (off - g.src_off), len, scratch + slot * k, tab, claim, lane);
        if (lane == 0) frag_len[f] = got;
        __syncwarp();
    }
}
__global__ void __launch_bounds__(32) snappy_encode_frags_kernel(const uint8_t* __restrict__ src, SnappyGeom g,
                                                                 uint8_t* scratch, uint64_t slot, uint32_t* frag_len,
                                                                 uint32_t* ticket, const uint32_t* in_flag, CallResult* res) {
    extern __shared__ __align__(16) uint32_t tab_mem[];
    __shared__ uint32_t claim_mem[kSnappyClaimBytes / 4];
    snappy_encode_frags_loop(src, g, scratch, slot, frag_len, ticket, reinterpret_cast<uint16_t*>(tab_mem), claim_mem, in_flag, res);
}
__global__ void __launch_bounds__(32) snappy_encode_frags_gtab_kernel(const uint8_t* __restrict__ src, SnappyGeom g,
                                                                      uint8_t* scratch, uint64_t slot, uint32_t* frag_len,
                                                                      uint32_t* ticket, uint16_t* tables,
                                                                      const uint32_t* in_flag, CallResult* res) {
    __shared__ uint32_t claim_mem[kSnappyClaimBytes / 4];
    snappy_encode_frags_loop(src, g, scratch, slot, frag_len, ticket, tables + (size_t)blockIdx.x * 16384, claim_mem, in_flag, res);
}

// Step 2: offsets of every fragment in the final stream, RAP frame and the leading varint
// (snappy.cc:2567-2651).  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) snappy_plan_kernel(SnappyGeom g, const uint32_t* __restrict__ frag_len,
                                                           uint64_t* frag_off, uint8_t* dst, uint64_t out_cap,
                                                           CallResult* res) {
    __shared__ uint64_t sm[40];
    const int tid = threadIdx.x;
    const uint32_t F = g.frags_total;
    const uint32_t per = (F + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = min(F, tid * per), hi = min(F, lo + per);
    uint64_t local = 0;
    for (uint32_t f = lo; f < hi; f++) local += frag_len[f];
    uint64_t total = 0;
    const bool framed = g.T > 1;
    const uint64_t frame = framed ? (uint64_t)kRapHeaderBytes + (uint64_t)kRapEntryBytes * g.T : 0;
    const uint64_t head = frame + varint32_len((uint32_t)g.n);
    uint64_t base = head + block_excl_scan(local, &total, sm);
    total += head;
    const bool fits = total <= out_cap && total <= 0xffffffffull;
    for (uint32_t f = lo; f < hi; f++) { frag_off[f] = base; base += frag_len[f]; }
    __syncthreads();
    if (!fits) { if (tid == 0) { res->error = 1; res->value = kErrCorrupt; } return; }
    if (!dst) { if (tid == 0) res->value = (long long)total; return; }   // plan only (a rank that does not own the head)
    if (framed) {
        for (uint32_t p = tid; p < g.T; p += blockDim.x) {      // RAP_i = {offset of body_i, |body_i|, part_i}
            const uint32_t f0 = p * g.frags_common;
            const uint32_t f1 = (p == g.T - 1) ? F : f0 + g.frags_common;
            const uint64_t o0 = frag_off[f0];
            const uint64_t o1 = frag_off[f1 - 1] + frag_len[f1 - 1];
            uint8_t* e = dst + kRapHeaderBytes + (uint64_t)kRapEntryBytes * p;
            st_u32_bytes(e, (uint32_t)o0); st_u32_bytes(e + 4, (uint32_t)(o1 - o0));
            st_u32_bytes(e + 8, (uint32_t)(g.common + (p == g.T - 1 ? g.left : 0)));
        }
    }
    if (tid == 0) {
        if (framed) {
            st_u32_bytes(dst, (uint32_t)kRapMagic); st_u32_bytes(dst + 4, (uint32_t)(kRapMagic >> 32));
            st_u32_bytes(dst + 8, (uint32_t)frame); st_u32_bytes(dst + 12, g.T);
        }
        put_varint32(dst + frame, (uint32_t)g.n);               // snappy.cc:2617-2619
        res->value = (long long)total;
    }
}

// Step 3: compaction of the fragment bodies.
__global__ void __launch_bounds__(256) snappy_compact_kernel(const uint8_t* __restrict__ scratch, uint64_t slot,
                                                             const uint32_t* __restrict__ frag_len,
                                                             const uint64_t* __restrict__ frag_off, uint32_t f0, uint8_t* dst,
                                                             uint64_t dst_off, const CallResult* res) {
    if (res->error) return;
    const uint32_t f = f0 + blockIdx.x;
    block_copy(dst + (frag_off[f] - dst_off), scratch + slot * blockIdx.x, frag_len[f]);
}

}  // namespace llc
