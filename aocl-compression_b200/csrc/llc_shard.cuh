// llc_shard.cuh -- ONE RAP frame over several GPUs; included by llc_device.cu.
//
// One process (or host thread) per GPU, each with its own context.  Rank r of R owns the contiguous partition
// range [floor(r*T/R), floor((r+1)*T/R)) of the frame: it holds only that slice of the input, encodes / decodes
// only those partitions, and writes only its own byte range of the result.  What crosses NVLink (NCCL, inside the
// library, on the context's stream):
//   compress    an all-gather of the per-partition records (LZ4: 16 bytes = sizes + first token; Snappy: 4 bytes
//               per 64 KiB fragment) -- every rank then runs the same stitch plan and knows every offset of the
//               final stream, rank 0 writes the RAP frame -- and, LZ4 only, the trailing literals a rank's first
//               partitions inherit from its predecessor's last ones (lz4.c:2808-2877): usually a few bytes, a
//               whole chain of partitions for incompressible data, sent point to point;
//   decompress  an all-gather of {bytes produced, error} so that every rank returns the same total.
// Reference being replaced: the OpenMP fork/join of AOCL_LZ4_compress_fast_mt / _decompress_safe_mt
// (lz4.c:2684-2905, 4785-4890) and snappy::RawCompress / RawUncompress (snappy.cc:2506-2655, 2271-2390), where the
// "ranks" are threads of one address space and the stitch is a serial loop.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a torch process that is the copy torch already
// loaded), so hosts that never shard do not need it.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace shard {

struct Api {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    bool ok = false;
};

static Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (a.lib) break; }
        if (!a.lib) return;
#define LLC_NCCL_SYM(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, sym))
        LLC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); LLC_NCCL_SYM(CommInitRank, "ncclCommInitRank");
        LLC_NCCL_SYM(CommDestroy, "ncclCommDestroy"); LLC_NCCL_SYM(AllGather, "ncclAllGather");
        LLC_NCCL_SYM(Broadcast, "ncclBroadcast"); LLC_NCCL_SYM(Send, "ncclSend"); LLC_NCCL_SYM(Recv, "ncclRecv");
        LLC_NCCL_SYM(GroupStart, "ncclGroupStart"); LLC_NCCL_SYM(GroupEnd, "ncclGroupEnd");
#undef LLC_NCCL_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.Broadcast && a.Send && a.Recv &&
               a.GroupStart && a.GroupEnd;
    });
    return a;
}

struct State {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    uint8_t* halo = nullptr; size_t halo_bytes = 0;          // inherited literals in front of this rank's slice
    ShardInfo* d_info = nullptr;                              // [nranks + 1]: slot 0 mine, 1.. gathered
    ShardInfo* h_info = nullptr;                              // pinned mirror
};

// partition range of a rank ([floor(g*T/G), floor((g+1)*T/G)))
static inline uint32_t part_lo(uint32_t T, int r, int R) { return (uint32_t)((uint64_t)T * (uint64_t)r / (uint64_t)R); }

}  // namespace shard

struct aocl_gpu_shard_s : shard::State {};

extern "C" int32_t aocl_gpu_shard_unique_id(void* id_out) {
    shard::Api& a = shard::api();
    if (!a.ok || !id_out) return -2;
    ncclUniqueId id;
    if (a.GetUniqueId(&id) != ncclSuccess) return -2;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

extern "C" void aocl_gpu_shard_destroy(aocl_gpu_ctx_t c) {
    if (!c || !c->shard) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->shard->comm) shard::api().CommDestroy(c->shard->comm);
    if (c->shard->halo) cudaFree(c->shard->halo);
    if (c->shard->d_info) cudaFree(c->shard->d_info);
    if (c->shard->h_info) cudaFreeHost(c->shard->h_info);
    delete c->shard;
    c->shard = nullptr;
}

extern "C" int32_t aocl_gpu_shard_init(aocl_gpu_ctx_t c, const void* id_bytes, int32_t rank, int32_t nranks) {
    shard::Api& a = shard::api();
    if (!c || !id_bytes || nranks < 1 || rank < 0 || rank >= nranks) return -5;
    if (!a.ok) { if (getenv("AOCL_GPU_VERBOSE")) fprintf(stderr, "[aocl-llc-b200] libnccl.so.2 not found\n"); return -2; }
    aocl_gpu_shard_destroy(c);
    cudaSetDevice(c->device);
    aocl_gpu_shard_s* s = new aocl_gpu_shard_s();
    s->rank = rank; s->nranks = nranks;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    if (a.CommInitRank(&s->comm, nranks, id, rank) != ncclSuccess ||
        cudaMalloc(&s->d_info, sizeof(ShardInfo) * (size_t)(nranks + 1)) != cudaSuccess ||
        cudaMallocHost(&s->h_info, sizeof(ShardInfo) * (size_t)(nranks + 1)) != cudaSuccess) {
        cudaGetLastError();
        c->shard = s;
        aocl_gpu_shard_destroy(c);
        return -2;
    }
    c->shard = s;
    return 0;
}

extern "C" int32_t aocl_gpu_shard_range(int32_t codec, size_t n, int32_t rank, int32_t nranks, uint32_t* first, uint32_t* count,
                                        uint64_t* byte_off, uint64_t* byte_len) {
    if ((codec != AOCL_GPU_LZ4 && codec != AOCL_GPU_SNAPPY) || nranks < 1 || rank < 0 || rank >= nranks) return -5;
    const uint32_t T = partition_count(n, codec == AOCL_GPU_LZ4 ? kLz4Window : kSnappyBlock);
    if (T < (uint32_t)nranks) return -2;                     // fewer partitions than ranks: not worth sharding
    const uint32_t p0 = shard::part_lo(T, rank, nranks), p1 = shard::part_lo(T, rank + 1, nranks);
    const uint64_t common = n / T;
    if (first) *first = p0;
    if (count) *count = p1 - p0;
    if (byte_off) *byte_off = common * p0;
    if (byte_len) *byte_len = (p1 == T ? (uint64_t)n : common * p1) - common * p0;
    return 0;
}

static bool shard_ok(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return true;
    if (getenv("AOCL_GPU_VERBOSE")) fprintf(stderr, "[aocl-llc-b200] NCCL %s failed (%d)\n", what, (int)r);
    return false;
}

// ------------------------------------------------------------------------------------------------ compress
extern "C" int64_t aocl_gpu_compress_sharded(aocl_gpu_ctx_t c, int32_t codec, const void* d_in_slice, size_t n, void* d_out_slice,
                                             size_t out_cap, uint64_t* out_off, uint64_t* out_len) {
    if (!c || !c->shard) return -5;
    shard::Api& a = shard::api();
    aocl_gpu_shard_s& S = *c->shard;
    const int R = S.nranks, me = S.rank;
    uint32_t p0 = 0, cnt = 0;
    uint64_t my_off = 0, my_len = 0;
    if (!d_in_slice || !d_out_slice || n > 0x7E000000ull || aocl_gpu_shard_range(codec, n, me, R, &p0, &cnt, &my_off, &my_len) != 0) return -2;
    begin_call(c);
    const uint32_t* in_flag = c->in_flag;                     // one-shot (aocl_gpu_set_input_watermark): the slice is still arriving
    c->in_flag = nullptr;
    const uint8_t* src = (const uint8_t*)d_in_slice;
    uint8_t* dst = (uint8_t*)d_out_slice;
    cudaStream_t st = c->stream;
    bool ok = true;
    ShardInfo mine = {};
    if (codec == AOCL_GPU_LZ4) {
        const uint32_t T = partition_count(n, kLz4Window);
        const uint64_t common = n / T, pmax = n / T + n % T;
        const uint64_t slot = align_up(pmax + pmax / 255 + 32, 256);
        // every partition of the range resident at once when possible (each one is a serial chain)
        const bool stab = cnt <= (uint32_t)c->sm_count * (uint32_t)kStabMax;
        const uint32_t grid = stab ? cnt : std::min<uint32_t>(cnt, 32u * (uint32_t)c->sm_count);
        const size_t o_rec = 0, o_plan = align_up(o_rec + sizeof(Lz4Rec) * T + 256, 256);
        const size_t o_tab = align_up(o_plan + sizeof(Lz4Plan) * T, 256);
        const size_t o_scr = align_up(o_tab + (stab ? 0 : (size_t)grid * 16384), 256);
        if (!ensure_ws(c, o_scr + slot * cnt)) return -2;
        Lz4Rec* rec = reinterpret_cast<Lz4Rec*>(c->ws + o_rec);
        uint32_t* ticket = reinterpret_cast<uint32_t*>(c->ws + o_rec + sizeof(Lz4Rec) * T);
        Lz4Plan* plan = reinterpret_cast<Lz4Plan*>(c->ws + o_plan);
        uint32_t* tables = reinterpret_cast<uint32_t*>(c->ws + o_tab);
        uint8_t* scratch = c->ws + o_scr;
        cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st);
        const Lz4Range g{(uint64_t)n, T, p0, cnt};
        if (stab) LLC_LAUNCH(lz4_encode_parts_kernel, grid, 32, 16384, st, src, g, scratch, slot, rec, ticket, in_flag, c->d_res);
        else LLC_LAUNCH(lz4_encode_parts_gtab_kernel, grid, 32, 0, st, src, g, scratch, slot, rec, ticket, tables, in_flag, c->d_res);
        // all-gather (in place, ranges differ by at most one partition) of the partition records
        ok = shard_ok(a.GroupStart(), "GroupStart");
        for (int r = 0; r < R && ok; r++) {
            const uint32_t lo = shard::part_lo(T, r, R), hi = shard::part_lo(T, r + 1, R);
            ok = shard_ok(a.Broadcast(rec + lo, rec + lo, sizeof(Lz4Rec) * (size_t)(hi - lo), ncclUint8, r, S.comm, st), "Broadcast(records)");
        }
        ok = shard_ok(a.GroupEnd(), "GroupEnd") && ok;
        // the same plan on every rank; rank 0 owns the head of the stream and writes the RAP frame
        LLC_LAUNCH(lz4_stitch_plan_kernel, 1, 1024, 0, st, rec, (uint64_t)n, T, me == 0 ? dst : (uint8_t*)nullptr, (uint64_t)0xffffffffull, plan, c->d_res);
        LLC_LAUNCH(lz4_shard_info_kernel, 1, 256, 0, st, plan, (uint64_t)n, T, p0, cnt, c->d_res, S.d_info);
        ok = ok && shard_ok(a.AllGather(S.d_info, S.d_info + 1, sizeof(ShardInfo), ncclUint8, S.comm, st), "AllGather(info)");
        cudaMemcpyAsync(S.h_info, S.d_info, sizeof(ShardInfo) * (size_t)(R + 1), cudaMemcpyDeviceToHost, st);
        if (!ok || cudaStreamSynchronize(st) != cudaSuccess) { cudaGetLastError(); c->last_rc = -2; return -2; }
        mine = S.h_info[0];
        bool fail = mine.total == 0;
        for (int r = 0; r < R; r++) fail = fail || S.h_info[1 + r].total == 0;
        if (fail) { c->last_rc = -2; return -2; }            // the plan failed: it is the same plan on every rank, all of them leave here
        // From here on a rank with a LOCAL problem (its piece does not fit, no memory for the inherited literals) must
        // not simply leave: its neighbours are about to exchange boundary literals with it.  It still serves their
        // requests from its input slice, skips its own receive side into a dummy, skips its compaction, and fails.
        bool local_fail = mine.out_hi - mine.out_lo > out_cap;
        if (mine.halo > S.halo_bytes) {
            if (S.halo) cudaFree(S.halo);
            S.halo_bytes = (size_t)align_up((size_t)mine.halo, 1 << 16);
            if (cudaMalloc(&S.halo, S.halo_bytes) != cudaSuccess) { cudaGetLastError(); S.halo = nullptr; S.halo_bytes = 0; local_fail = true; }
        }
        bool any = false;
        for (int gk = 1; gk < R; gk++) any = any || S.h_info[1 + gk].halo != 0;
        if (any) {
            ok = shard_ok(a.GroupStart(), "GroupStart");
            for (int gk = 1; gk < R && ok; gk++) {
                const uint64_t hk = S.h_info[1 + gk].halo;
                if (!hk) continue;
                const uint64_t Sg = common * shard::part_lo(T, gk, R), lo = Sg - hk;      // input bytes [lo, Sg)
                for (int h = 0; h < gk && ok; h++) {
                    const uint64_t Sh = common * shard::part_lo(T, h, R), Sh1 = common * shard::part_lo(T, h + 1, R);
                    const uint64_t x0 = std::max(lo, Sh), x1 = std::min(Sg, Sh1);
                    if (x0 >= x1) continue;
                    if (me == h) ok = shard_ok(a.Send(src + (x0 - my_off), (size_t)(x1 - x0), ncclUint8, gk, S.comm, st), "Send(halo)");
                    if (me == gk) {
                        // (no halo buffer: receive into the scratch area, which the skipped compaction will not read)
                        uint8_t* into = S.halo ? S.halo + (x0 - lo) : scratch;
                        if (!S.halo && (x1 - x0) > slot * (uint64_t)cnt) { ok = false; break; }
                        ok = shard_ok(a.Recv(into, (size_t)(x1 - x0), ncclUint8, h, S.comm, st), "Recv(halo)");
                    }
                }
            }
            ok = shard_ok(a.GroupEnd(), "GroupEnd") && ok;
        }
        if (local_fail) { cudaStreamSynchronize(st); c->last_rc = -2; return -2; }
        LLC_LAUNCH(lz4_compact_kernel, cnt, 256, 0, st, src, my_off, (const uint8_t*)S.halo, (uint64_t)mine.halo, scratch, slot, rec, plan, p0, dst,
                   (uint64_t)mine.out_lo, c->d_res);
    } else {
        const uint32_t T = partition_count(n, kSnappyBlock);
        SnappyGeom g = snappy_geom(n, T);
        const uint32_t F = g.frags_total;
        g.f0 = p0 * g.frags_common;
        g.fcnt = (p0 + cnt == T ? F : (p0 + cnt) * g.frags_common) - g.f0;
        g.src_off = my_off;
        const uint64_t slot = 76544;
        const bool stab = g.fcnt <= (uint32_t)c->sm_count * 6u;
        const uint32_t grid = stab ? g.fcnt : std::min<uint32_t>(g.fcnt, 24u * (uint32_t)c->sm_count);
        const size_t o_len = 0, o_off = align_up(o_len + sizeof(uint32_t) * (F + 2), 256);
        const size_t o_tab = align_up(o_off + sizeof(uint64_t) * (F + 1), 256);
        const size_t o_scr = align_up(o_tab + (stab ? 0 : (size_t)grid * 32768), 256);
        if (!ensure_ws(c, o_scr + slot * ((size_t)g.fcnt + 1))) return -2;
        uint32_t* frag_len = reinterpret_cast<uint32_t*>(c->ws + o_len);
        uint32_t* ticket = frag_len + F + 1;
        uint64_t* frag_off = reinterpret_cast<uint64_t*>(c->ws + o_off);
        uint16_t* tables = reinterpret_cast<uint16_t*>(c->ws + o_tab);
        uint8_t* scratch = c->ws + o_scr;
        cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st);
        if (stab) LLC_LAUNCH(snappy_encode_frags_kernel, grid, 32, 32768, st, src, g, scratch, slot, frag_len, ticket, in_flag, c->d_res);
        else LLC_LAUNCH(snappy_encode_frags_gtab_kernel, grid, 32, 0, st, src, g, scratch, slot, frag_len, ticket, tables, in_flag, c->d_res);
        ok = shard_ok(a.GroupStart(), "GroupStart");
        for (int r = 0; r < R && ok; r++) {
            const uint32_t lo = shard::part_lo(T, r, R) * g.frags_common;
            const uint32_t hi = shard::part_lo(T, r + 1, R) == T ? F : shard::part_lo(T, r + 1, R) * g.frags_common;
            ok = shard_ok(a.Broadcast(frag_len + lo, frag_len + lo, sizeof(uint32_t) * (size_t)(hi - lo), ncclUint8, r, S.comm, st), "Broadcast(fragment sizes)");
        }
        ok = shard_ok(a.GroupEnd(), "GroupEnd") && ok;
        SnappyGeom whole = g;
        whole.f0 = 0; whole.fcnt = F; whole.src_off = 0;
        LLC_LAUNCH(snappy_plan_kernel, 1, 1024, 0, st, whole, frag_len, frag_off, me == 0 ? dst : (uint8_t*)nullptr, (uint64_t)0xffffffffull, c->d_res);
        // my byte range of the stream: rank 0 starts at 0 (frame + varint), the others at their first fragment
        cudaMemcpyAsync(&S.h_info[0].out_lo, frag_off + g.f0, sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(&S.h_info[0].out_hi, frag_off + (g.f0 + g.fcnt - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(&S.h_info[0].halo, frag_len + (g.f0 + g.fcnt - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(&S.h_info[0].total, &c->d_res->value, sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
        S.h_info[0].halo = 0;
        if (!ok || cudaStreamSynchronize(st) != cudaSuccess) { cudaGetLastError(); c->last_rc = -2; return -2; }
        mine.out_lo = me == 0 ? 0 : S.h_info[0].out_lo;
        mine.out_hi = S.h_info[0].out_hi + (S.h_info[0].halo & 0xffffffffull);
        mine.total = S.h_info[0].total;
        if ((long long)mine.total <= 0 || mine.out_hi - mine.out_lo > out_cap) { c->last_rc = -2; return -2; }
        LLC_LAUNCH(snappy_compact_kernel, g.fcnt, 256, 0, st, scratch, slot, frag_len, frag_off, g.f0, dst, (uint64_t)mine.out_lo, c->d_res);
    }
    end_call(c);
    const int64_t r = aocl_gpu_finish(c);
    if (!ok || r < 0) return -2;
    if (out_off) *out_off = mine.out_lo;
    if (out_len) *out_len = mine.out_hi - mine.out_lo;
    return (int64_t)mine.total;
}

// ------------------------------------------------------------------------------------------------ decompress
__global__ void shard_range_info_kernel(const PartDesc* __restrict__ parts, const CallResult* res, int rank, int nranks, ShardInfo* info) {
    const uint32_t T = res->error ? 0u : (uint32_t)res->parts;
    ShardInfo s = {};
    if (T) {
        const uint32_t lo = (uint32_t)((uint64_t)T * rank / nranks), hi = (uint32_t)((uint64_t)T * (rank + 1) / nranks);
        s.halo = ((unsigned long long)lo << 32) | (hi - lo);
        unsigned long long first = 0, last = 0;
        bool seen = false;
        for (uint32_t i = lo; i < hi; i++)
            if (parts[i].in_len) { if (!seen) { first = parts[i].out_off; seen = true; } last = parts[i].out_off + parts[i].out_len; }
        s.out_lo = first; s.out_hi = seen ? last : first;
        s.total = 1;
    }
    *info = s;
}

extern "C" int64_t aocl_gpu_decompress_sharded(aocl_gpu_ctx_t c, int32_t codec, const void* d_stream, size_t n, void* d_out_slice,
                                               size_t out_cap, uint64_t* out_off, uint64_t* out_len) {
    if (!c || !c->shard) return -5;
    shard::Api& a = shard::api();
    aocl_gpu_shard_s& S = *c->shard;
    if ((codec != AOCL_GPU_LZ4 && codec != AOCL_GPU_SNAPPY) || !d_stream || n == 0 || n > 0xffffffffull || (!d_out_slice && out_cap)) return -2;
    begin_call(c);
    cudaStream_t st = c->stream;
    if (!ensure_ws(c, sizeof(PartDesc) * kMaxPartitions)) return -2;
    PartDesc* parts = reinterpret_cast<PartDesc*>(c->ws);
    // every rank parses the frame it holds (header + entry table + its own partitions at their stream offsets)
    LLC_LAUNCH(rap_parse_kernel, 1, 1024, 0, st, codec, (const uint8_t*)d_stream, (uint64_t)n, (uint64_t)out_cap, 0, parts, c->d_res);
    LLC_LAUNCH(shard_range_info_kernel, 1, 1, 0, st, parts, c->d_res, S.rank, S.nranks, S.d_info);
    cudaMemcpyAsync(S.h_info, S.d_info, sizeof(ShardInfo), cudaMemcpyDeviceToHost, st);
    bool ok = cudaStreamSynchronize(st) == cudaSuccess;
    ShardInfo mine = S.h_info[0];
    const uint32_t first = (uint32_t)(mine.halo >> 32), count = (uint32_t)(mine.halo & 0xffffffffull);
    const bool parsed = ok && mine.total != 0;
    if (parsed && count) {
        LLC_LAUNCH(range_check_kernel, 1, 256, 0, st, parts, c->d_res, first, count, (uint64_t)mine.out_lo, (uint64_t)out_cap);
        launch_decode_range(c, codec, d_stream, d_out_slice, parts, first, count, (uint64_t)mine.out_lo);
    }
    // all-gather of {bytes produced, failed}: every rank returns the stream's total or the failure
    mine.total = 0;
    cudaMemcpyAsync(S.d_info, &c->d_res->value, sizeof(long long), cudaMemcpyDeviceToDevice, st);        // halo  := produced (or error value)
    cudaMemcpyAsync(&S.d_info->total, &c->d_res->error, sizeof(int), cudaMemcpyDeviceToDevice, st);     // total := error flag (low word)
    if (!parsed || !count) cudaMemsetAsync(S.d_info, 0, sizeof(unsigned long long), st);
    if (!parsed) cudaMemsetAsync(&S.d_info->total, 0xff, sizeof(int), st);
    ok = shard_ok(a.AllGather(S.d_info, S.d_info + 1, sizeof(ShardInfo), ncclUint8, S.comm, st), "AllGather(result)") && ok;
    cudaMemcpyAsync(S.h_info, S.d_info, sizeof(ShardInfo) * (size_t)(S.nranks + 1), cudaMemcpyDeviceToHost, st);
    end_call(c);
    const int64_t r = aocl_gpu_finish(c);
    (void)r;
    if (!ok) return -2;
    uint64_t total = 0;
    for (int k = 0; k < S.nranks; k++) {
        if ((uint32_t)S.h_info[1 + k].total != 0) return -2;
        total += S.h_info[1 + k].halo;
    }
    if (out_off) *out_off = mine.out_lo;
    if (out_len) *out_len = mine.out_hi - mine.out_lo;
    return (int64_t)total;
}
