#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 LZ4/Snappy RAP path.

A "step" is one pass of the hot path over one batch of synthetic input: compress the
workload buffer into a RAP frame and decompress that frame again (BASELINE.json configs[1]:
LZ4, 1 GiB synthetic text-like data, 256 KiB chunks; `--workload snappy_log` is configs[2]).
`value` is the round-trip throughput in GB/s of uncompressed bytes (2*U per step: U read by
the compressor + U written by the decompressor) with all buffers resident in HBM, timed with
CUDA events on the library's stream.  `detail` splits it into compress and decompress GB/s
(the north-star target is the decompress figure) and the compression ratio.  `e2e` is the same
metric through aocl_llc_compress / aocl_llc_decompress with pinned HOST buffers.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--workload lz4_text|snappy_log] [--size BYTES]

N > 1 is launched with torchrun, one rank per GPU: every rank round-trips its own frame of the
same size (weak scaling; RAP frames are independent units) and the ranks all-gather their RAP
entry tables (the only exchange the path has) over NCCL.
`--impl reference` times the unmodified reference (oracle/_ref/libaocl_ref.so, OpenMP) on the
host cores for the same metric; it is the only mode (with the cpu_baseline leg at N=1) that
loads anything under oracle/.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python"))

LZ4, SNAPPY = 0, 4
WORKLOADS = {
    "lz4_text": dict(codec=LZ4, gen="text_like", seed=2024, size=1 << 30,
                     name="LZ4 RAP compress+decompress, 1 GiB synthetic text-like, 256 KiB partitions (BASELINE configs[1])"),
    "snappy_log": dict(codec=SNAPPY, gen="log_like", seed=2025, size=1 << 30,
                       name="Snappy RAP compress+uncompress, 1 GiB synthetic log-like, 64 KiB blocks (BASELINE configs[2])"),
}


def make_data(gen_name: str, size: int, seed: int) -> np.ndarray:
    """Deterministic synthetic input, generated slab-wise on several host threads."""
    from concurrent.futures import ThreadPoolExecutor
    from llc_b200 import gen
    fn = getattr(gen, gen_name)
    slab = 64 << 20
    nslab = (size + slab - 1) // slab
    out = np.empty(size, dtype=np.uint8)

    def work(i):
        lo = i * slab
        hi = min(size, lo + slab)
        out[lo:hi] = fn(hi - lo, seed=seed * 1000 + i)

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
        list(ex.map(work, range(nslab)))
    return out


# ------------------------------------------------------------------------------- reference arm
class RefDesc(C.Structure):
    _fields_ = [("inBuf", C.c_void_p), ("outBuf", C.c_void_p), ("workBuf", C.c_void_p),
                ("inSize", C.c_size_t), ("outSize", C.c_size_t), ("level", C.c_size_t), ("optVar", C.c_size_t),
                ("numThreads", C.c_int), ("numMPIranks", C.c_int), ("memLimit", C.c_size_t),
                ("measureStats", C.c_int), ("cSize", C.c_uint64), ("dSize", C.c_uint64),
                ("cTime", C.c_uint64), ("dTime", C.c_uint64), ("cSpeed", C.c_float), ("dSpeed", C.c_float),
                ("optOff", C.c_int), ("optLevel", C.c_int)]


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def load_reference():
    path = os.path.join(ROOT, "oracle", "_ref", "libaocl_ref.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm must use every host core
    try:
        omp = C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL)
        omp.omp_set_num_threads(host_threads())
    except OSError:
        pass
    dp = C.POINTER(RefDesc)
    L.aocl_llc_setup.restype, L.aocl_llc_setup.argtypes = C.c_int32, [dp, C.c_int]
    L.aocl_llc_compress.restype, L.aocl_llc_compress.argtypes = C.c_int64, [dp, C.c_int]
    L.aocl_llc_decompress.restype, L.aocl_llc_decompress.argtypes = C.c_int64, [dp, C.c_int]
    return L


def reference_round_trip(L, data: np.ndarray, codec: int, steps: int, warmup: int):
    """Times aocl_llc_compress + aocl_llc_decompress of the unmodified reference (OpenMP, all host
    threads) the way test/codec_bench.c does.  Returns (seconds per step, compress s, decompress s, csize)."""
    n = len(data)
    cap = n + n // 6 + 16384 + 16 + 12 * 8192
    comp = np.empty(cap, dtype=np.uint8)
    back = np.empty(n, dtype=np.uint8)
    d = RefDesc()
    d.optOff, d.optLevel, d.measureStats = 0, -1, 0
    assert L.aocl_llc_setup(C.byref(d), codec) == 0
    tc = td = 0.0
    csize = 0
    for it in range(warmup + steps):
        d.inBuf, d.inSize, d.outBuf, d.outSize = data.ctypes.data, n, comp.ctypes.data, cap
        t0 = time.perf_counter()
        csize = L.aocl_llc_compress(C.byref(d), codec)
        t1 = time.perf_counter()
        assert csize > 0
        d.inBuf, d.inSize, d.outBuf, d.outSize = comp.ctypes.data, csize, back.ctypes.data, n
        r = L.aocl_llc_decompress(C.byref(d), codec)
        t2 = time.perf_counter()
        assert r == n
        if it >= warmup:
            tc += t1 - t0
            td += t2 - t1
    assert np.array_equal(back, data)
    return (tc + td) / steps, tc / steps, td / steps, int(csize)


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L = load_reference()
    if L is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libaocl_ref.so not built"}))
        return
    cores = host_threads()
    sample = min(wl["size"], args.ref_sample)
    data = make_data(wl["gen"], sample, wl["seed"])
    per, tc, td, csize = reference_round_trip(L, data, wl["codec"], args.steps, args.warmup)
    gbps = 2 * sample / per / 1e9
    line = {
        "impl": "reference", "metric": "rap_round_trip_GBps", "value": gbps, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl["name"], "sample_bytes": sample},
        "detail": {"compress_GBps": sample / tc / 1e9, "decompress_GBps": sample / td / 1e9,
                   "ratio": csize / sample, "compressed_bytes": csize},
        "cpu_baseline": {"value": gbps, "unit": "GB/s", "cores": cores, "kind": "reference",
                         "sample": f"first {sample >> 20} MiB of the workload, OpenMP max threads = {cores}"},
        "e2e": {"value": gbps, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi in loop mode (100 ms) for the duration of the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            time.sleep(0.15)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.12)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in out.splitlines():
            self.rows.append([x.strip() for x in line.split(",")])

    def summary(self):
        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[1]) for r in self.rows if len(r) >= 9 and num(r[1]) is not None]
        mx = [num(r[2]) for r in self.rows if len(r) >= 9 and num(r[2]) is not None]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------- B200 arm
def run_b200(args, wl):
    import torch
    import llc_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    os.environ.setdefault("AOCL_GPU_DEVICE", str(local))     # the aocl_llc_* host API opens its context on this rank's GPU
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    L = llc_b200.load()
    codec, U = wl["codec"], wl["size"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured" if "hbm_gbs" in peaks else "fallback"

    # ---- synthetic input: every rank owns one frame of the same size (weak scaling)
    data = make_data(wl["gen"], U, wl["seed"] + rank)
    h_in = torch.from_numpy(data).pin_memory()
    d_in = h_in.cuda(non_blocking=True)
    cap = L.aocl_gpu_compress_bound(codec, U)
    d_comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_back = torch.empty(U, dtype=torch.uint8, device="cuda")
    ctx = llc_b200.GpuContext(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    torch.cuda.synchronize()

    def step_device(record=None):
        """compress -> [allgather of RAP entries when N > 1] -> decompress, all on the library's stream."""
        if record:
            record[0].record(stream)
        ctx.compress_async(codec, d_in, d_comp)
        if record:
            record[1].record(stream)
        csz = ctx.finish()
        assert csz > 0, csz
        if dist is not None:
            T = L.aocl_gpu_partition_count(codec, U)
            table = d_comp[16:16 + 12 * T].view(torch.int32)
            gathered = torch.empty(world * table.numel(), dtype=torch.int32, device="cuda")
            dist.all_gather_into_tensor(gathered, table.contiguous())
        if record:
            record[2].record(stream)
        ctx.decompress_async(codec, d_comp, csz, d_back)
        if record:
            record[3].record(stream)
        got = ctx.finish()
        assert got == U, got
        return csz

    # ---- warm-up + correctness of the round trip (size-independent property at full size)
    for _ in range(max(args.warmup, 3)):
        csz = step_device()
    assert torch.equal(d_back, d_in), "round trip mismatch"
    launches0 = L.aocl_gpu_launch_count()

    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clocks:
        t_wall0 = time.perf_counter()
        for k in range(args.steps):
            step_device(evs[k])
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
    launches = L.aocl_gpu_launch_count() - launches0
    tc = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps       # ms, compress kernels
    td = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps       # ms, decompress kernels
    step_ms = tc + td
    if dist is not None:
        t = torch.tensor([step_ms, tc, td], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, tc, td = [float(x) for x in t.tolist()]
        dist.barrier()

    # ---- per-kernel timing (CUDA events around every launch, on the launching stream)
    ctx.set_profiling(True)
    ctx.compress_async(codec, d_in, d_comp); ctx.finish(); prof_c = ctx.profile()
    ctx.decompress_async(codec, d_comp, csz, d_back); ctx.finish(); prof_d = ctx.profile()
    ctx.set_profiling(False)

    frame = 16 + 12 * L.aocl_gpu_partition_count(codec, U)
    alg_bytes = U + csz                      # read U + write (C+F) for compress; read (C+F) + write U for decompress
    kern = dict(prof_c + prof_d)
    enc_name = "lz4_encode_parts_kernel" if codec == LZ4 else "snappy_encode_frags_kernel"
    dom_name = max(kern, key=lambda k: kern[k])
    dom_ms = kern[dom_name]
    dec_name = next((k for k in kern if "decode_parts" in k), "decode_parts_kernel")
    dec_ms = kern.get(dec_name, td)
    # DRAM traffic of the same kernels from the committed `ncu --set full` captures (profiles/traffic.json)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {})
    except Exception:
        pass

    def traffic_of(name):
        for key, val in traffic.items():
            if key in name:
                return val
        return None

    # ---- end to end through the reference-facing API with pinned host buffers
    e2e = None
    cpu_baseline = None
    if True:
        from llc_b200 import AoclDesc
        h_comp = torch.empty(cap, dtype=torch.uint8).pin_memory()
        h_back = torch.empty(U, dtype=torch.uint8).pin_memory()
        d = AoclDesc()
        d.optOff, d.optLevel, d.measureStats = 0, -1, 1
        assert L.aocl_llc_setup(C.byref(d), codec) == 0
        e2e_steps = max(1, min(args.steps, 5))
        tot = tot_c = 0.0
        for it in range(1 + e2e_steps):
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            d.inBuf, d.inSize, d.outBuf, d.outSize = h_in.data_ptr(), U, h_comp.data_ptr(), cap
            c2 = L.aocl_llc_compress(C.byref(d), codec)
            t1 = time.perf_counter()
            assert c2 == csz, (c2, csz)
            d.inBuf, d.inSize, d.outBuf, d.outSize = h_comp.data_ptr(), c2, h_back.data_ptr(), U
            r2 = L.aocl_llc_decompress(C.byref(d), codec)
            assert r2 == U, r2
            if it >= 1:
                tot += time.perf_counter() - t0
                tot_c += t1 - t0
        assert torch.equal(h_back, h_in), "e2e round trip mismatch"
        e2e_s = tot / e2e_steps
        if dist is not None:
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": world * 2 * U / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": U + csz,
               "d2h_bytes_per_step": csz + U, "ms_per_step": e2e_s * 1e3,
               "compress_ms": tot_c / e2e_steps * 1e3, "decompress_ms": (tot - tot_c) / e2e_steps * 1e3,
               "api": "aocl_llc_compress + aocl_llc_decompress, pinned host buffers; transfers pipelined with the kernels "
                      "(striped H2D behind an input watermark / slab-wise H2D-decode-D2H)"}

    # ---- CPU baseline: the unmodified reference on this box's host cores, bounded sample (rank 0, N=1)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        R = load_reference()
        if R is not None:
            sample = min(U, args.ref_sample)
            per, rtc, rtd, rcs = reference_round_trip(R, data[:sample], codec, 2, 1)
            cores = host_threads()
            cpu_baseline = {"value": 2 * sample / per / 1e9, "unit": "GB/s", "cores": cores, "kind": "reference",
                            "sample": f"first {sample >> 20} MiB of the workload, OpenMP max threads = {cores}",
                            "compress_GBps": sample / rtc / 1e9, "decompress_GBps": sample / rtd / 1e9}

    if rank == 0:
        value = world * 2 * U / (step_ms / 1e3) / 1e9
        line = {
            "metric": "rap_round_trip_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl["name"], "bytes_per_gpu": U, "l2": "inputs (1 GiB) exceed the 126 MB L2; no flush",
                       "partitions_per_frame": L.aocl_gpu_partition_count(codec, U), "parallelism": f"frames x{world}"},
            "detail": {"compress_GBps": world * U / (tc / 1e3) / 1e9, "decompress_GBps": world * U / (td / 1e3) / 1e9,
                       "compress_ms": tc, "decompress_ms": td, "ratio": csz / U, "compressed_bytes": int(csz),
                       "wall_s_timed_region": t_wall,
                       "kernels_ms": {k: round(v, 4) for k, v in kern.items()}},
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": alg_bytes / (dom_ms / 1e3) / 1e9,
                         "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / (dom_ms / 1e3) / 1e9 / hbm_peak,
                         "traffic": traffic_of(dom_name), "peak_kind": peak_kind, "algorithmic_bytes": int(alg_bytes)},
            "roofline_decompress": {"bound": "hbm", "kernel": dec_name, "traffic": traffic_of(dec_name),
                                    "achieved": alg_bytes / (dec_ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                    "frac": alg_bytes / (dec_ms / 1e3) / 1e9 / hbm_peak, "peak_kind": peak_kind,
                                    "user_GBps": U / (dec_ms / 1e3) / 1e9},
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        _ = (frame, enc_name)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lz4_text", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=0, help="override the workload size in bytes (testing)")
    ap.add_argument("--ref-sample", type=int, default=256 << 20, help="bytes of the workload the CPU reference is timed on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.size:
        wl["size"] = args.size
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
