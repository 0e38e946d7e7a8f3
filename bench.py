#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 LZ4/Snappy RAP path.

A "step" is one pass of the hot path over one batch of synthetic input: compress the workload buffer
into a RAP frame and decompress that frame again (BASELINE.json configs[1]: LZ4, 1 GiB synthetic
text-like data, 256 KiB partitions).  `value` is the round-trip throughput in GB/s of uncompressed
bytes (2*U per step: U read by the compressor + U written by the decompressor) with all buffers
resident in HBM, timed with CUDA events on the library's stream.  `detail` splits it into compress
and decompress GB/s (the north-star target is the decompress figure), the compression ratio and --
outside the timed region -- whether the stream is byte-identical to what the unmodified reference
writes for the same input.  `e2e` is the same metric through aocl_llc_compress / aocl_llc_decompress
with pinned HOST buffers.  `detail.configs` carries the other BASELINE configs, each measured the same
way in the same run: [0] the reference's single-thread CPU case, [2] Snappy 1 GiB log-like, [3] LZ4
decode of 16 x 1 GiB frames, [4] batched decode of 1,048,576 independent 64 KiB pages.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--workload lz4_text|snappy_log] [--size BYTES] [--configs all|none|0,2,3,4]

N > 1 is launched with torchrun, one rank per GPU.  Frames are the natural shard unit of a job larger
than the int32-sized API (rank r holds frame r): every rank round-trips its own frame (weak scaling)
and, INSIDE the timed step, the ranks all-gather their RAP entry tables over NCCL -- the one exchange
the path has -- into the archive index (global offsets of every partition of every frame), which the
step checks.  configs[3] / [4] shard their frames / pages over the ranks (strong scaling).
`--impl reference` times the unmodified reference (oracle/_ref/libaocl_ref.so, OpenMP, all host
threads, OMP_PROC_BIND=close) on the host cores for the same metric on the SAME full-size input; it is
the only mode (with the cpu_baseline legs at N=1) that loads anything under oracle/.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
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

    with ThreadPoolExecutor(max_workers=min(32, host_threads())) as ex:
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


def _affinity():
    try:
        return set(os.sched_getaffinity(0))
    except AttributeError:
        return set(range(os.cpu_count() or 1))


# Taken at import: loading libgomp with OMP_PROC_BIND set binds the calling thread to its first place, which would
# shrink every later reading (and the affinity every later thread inherits) to one core.
_CPUS = _affinity()


def host_threads() -> int:
    return max(1, len(_CPUS))


def physical_cores() -> int:
    """Distinct (package, core) pairs among the CPUs this process may run on."""
    try:
        cpus = sorted(_CPUS)
        seen = set()
        for c in cpus:
            base = f"/sys/devices/system/cpu/cpu{c}/topology/"
            seen.add((open(base + "physical_package_id").read().strip(), open(base + "core_id").read().strip()))
        return max(1, len(seen))
    except Exception:
        return host_threads()


_REF = None


def load_reference():
    """The unmodified reference compiled by oracle/Makefile.  OMP_PROC_BIND / OMP_PLACES must be in the
    environment before libgomp initialises."""
    global _REF
    if _REF is not None:
        return _REF
    path = os.path.join(ROOT, "oracle", "_ref", "libaocl_ref.so")
    if not os.path.exists(path):
        return None
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    omp = C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL)
    omp.omp_set_num_threads.argtypes = [C.c_int]
    omp.omp_get_max_threads.restype = C.c_int
    omp.omp_get_max_threads()                                # forces libgomp's start-up (and its binding of this thread) now
    try:
        os.sched_setaffinity(0, _CPUS)                       # the calling thread floats again; OpenMP workers stay bound
    except (AttributeError, OSError):
        pass
    dp = C.POINTER(RefDesc)
    L.aocl_llc_setup.restype, L.aocl_llc_setup.argtypes = C.c_int32, [dp, C.c_int]
    L.aocl_llc_compress.restype, L.aocl_llc_compress.argtypes = C.c_int64, [dp, C.c_int]
    L.aocl_llc_decompress.restype, L.aocl_llc_decompress.argtypes = C.c_int64, [dp, C.c_int]
    _REF = (L, omp)
    return _REF


def rap_threads(stream: np.ndarray) -> int:
    """T from the RAP header (test/codec_bench.c convention); 1 for a frame-less stream."""
    if len(stream) >= 16 and bytes(stream[:8]) == b"AOCL_LLC":
        return int.from_bytes(bytes(stream[12:16]), "little")
    return 1


def reference_round_trip(R, data: np.ndarray, codec: int, threads: int, steps: int, warmup: int, keep=None):
    """Times aocl_llc_compress + aocl_llc_decompress of the unmodified reference at `threads` OpenMP threads the
    way test/codec_bench.c does.  Returns a dict (seconds per step, split, compressed size, T of the frame)."""
    L, omp = R
    omp.omp_set_num_threads(int(threads))
    n = len(data)
    cap = n + n // 6 + 16384 + 16 + 12 * 8192
    comp = np.empty(cap, dtype=np.uint8)
    back = np.empty(n, dtype=np.uint8)
    d = RefDesc()
    d.optOff, d.optLevel, d.measureStats = 0, -1, 0
    assert L.aocl_llc_setup(C.byref(d), codec) == 0
    tc = td = 0.0
    best_c = best_d = 1e30
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
            best_c, best_d = min(best_c, t1 - t0), min(best_d, t2 - t1)
    assert np.array_equal(back, data)
    if keep is not None:
        keep["sha256"] = hashlib.sha256(comp[:csize]).hexdigest()
    return {"s_per_step": (tc + td) / steps, "compress_s": tc / steps, "decompress_s": td / steps,
            "compress_best_s": best_c, "decompress_best_s": best_d, "csize": int(csize), "T": rap_threads(comp[:64]),
            "threads": int(threads)}


def cpu_baseline_sweep(R, data, codec, steps=2, warmup=1):
    """OMP_NUM_THREADS in {1, physical cores, all hardware threads} on the full input."""
    n = len(data)
    sweep = []
    counts = sorted({1, physical_cores(), host_threads()})
    for t in counts:
        r = reference_round_trip(R, data, codec, t, steps if t > 1 else 1, warmup if t > 1 else 0)
        sweep.append({"threads": t, "T_in_frame": r["T"], "round_trip_GBps": 2 * n / r["s_per_step"] / 1e9,
                      "compress_GBps": n / r["compress_s"] / 1e9, "decompress_GBps": n / r["decompress_s"] / 1e9,
                      "compress_best_GBps": n / r["compress_best_s"] / 1e9, "decompress_best_GBps": n / r["decompress_best_s"] / 1e9})
    best = max(sweep, key=lambda s: s["round_trip_GBps"])
    return {"value": best["round_trip_GBps"], "unit": "GB/s", "cores": best["threads"], "kind": "reference",
            "sample": f"the full {n >> 20} MiB workload, aocl_llc_compress + aocl_llc_decompress of the unmodified reference "
                      f"(oracle/_ref), OMP_PROC_BIND={os.environ.get('OMP_PROC_BIND')} OMP_PLACES={os.environ.get('OMP_PLACES')}; "
                      f"value = best of the thread sweep",
            "compress_GBps": best["compress_GBps"], "decompress_GBps": best["decompress_GBps"],
            "host": {"hw_threads": host_threads(), "physical_cores": physical_cores()}, "sweep": sweep}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    n = wl["size"]
    data = make_data(wl["gen"], n, wl["seed"])               # before libgomp exists (see _CPUS)
    R = load_reference()
    if R is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libaocl_ref.so not built"}))
        return
    r = reference_round_trip(R, data, wl["codec"], cores, args.steps, args.warmup)
    gbps = 2 * n / r["s_per_step"] / 1e9
    line = {
        "impl": "reference", "metric": "rap_round_trip_GBps", "value": gbps, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl["name"], "bytes_per_gpu": n, "partitions_per_frame": r["T"],
                   "note": "the reference's OpenMP path on the host cores, same full-size input as the B200 arm; its frame has "
                           "T = min(threads, P) partitions (threads/threads.c:55-88)"},
        "detail": {"compress_GBps": n / r["compress_s"] / 1e9, "decompress_GBps": n / r["decompress_s"] / 1e9,
                   "ratio": r["csize"] / n, "compressed_bytes": r["csize"], "omp_threads": cores, "T_in_frame": r["T"],
                   "OMP_PROC_BIND": os.environ.get("OMP_PROC_BIND"), "physical_cores": physical_cores()},
        "cpu_baseline": {"value": gbps, "unit": "GB/s", "cores": cores, "kind": "reference",
                         "sample": f"the full {n >> 20} MiB workload, OpenMP threads = {cores}"},
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
class Env:
    """Per-process handles shared by the measurement helpers."""

    def __init__(self):
        import torch
        import llc_b200
        self.torch, self.llc = torch, llc_b200
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
        torch.cuda.set_device(self.local)
        os.environ.setdefault("AOCL_GPU_DEVICE", str(self.local))     # the aocl_llc_* host API opens its contexts on this rank's GPU
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.L = llc_b200.load()
        self.ctx = llc_b200.GpuContext(self.local)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=torch.device("cuda", self.local))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        self.peak_kind = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        self.traffic = {}
        try:
            self.traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass

    def traffic_of(self, workload, name):
        for key, val in self.traffic.get(workload, {}).items():
            if key in name:
                return val
        return None

    def max_over_ranks(self, vals):
        if self.dist is None:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def roofline(self, workload, name, alg_bytes, ms, extra=None):
        r = {"bound": "hbm", "kernel": name, "achieved": alg_bytes / (ms / 1e3) / 1e9, "peak": self.hbm_peak, "unit": "GB/s",
             "frac": alg_bytes / (ms / 1e3) / 1e9 / self.hbm_peak, "traffic": self.traffic_of(workload, name),
             "peak_kind": self.peak_kind, "algorithmic_bytes": int(alg_bytes), "kernel_ms": ms}
        if extra:
            r.update(extra)
        return r


def archive_index(env, d_comp, T, U):
    """N > 1: the ranks all-gather their RAP entry tables (12 bytes per partition) and scan them into the archive
    index: where every partition of every frame starts in the concatenated archive and in the uncompressed job.
    Returns (index tensor, total uncompressed bytes): the caller checks the total -- the result is used."""
    torch, dist = env.torch, env.dist
    table = d_comp[16:16 + 12 * T].view(torch.int32).view(T, 3)
    gathered = torch.empty((env.world, T, 3), dtype=torch.int32, device="cuda")
    dist.all_gather_into_tensor(gathered.view(-1), table.reshape(-1).contiguous())
    g = gathered.to(torch.int64) & 0xffffffff
    comp_off = torch.cumsum(g[:, :, 1].reshape(-1), 0)
    plain_off = torch.cumsum(g[:, :, 2].reshape(-1), 0)
    return (comp_off, plain_off), int(plain_off[-1].item())


def round_trip(env, args, wl, workload_key, with_cpu, want_identity):
    """The headline measurement for one workload.  Returns the JSON line (dict) on rank 0, else None."""
    torch, L, ctx, stream, dist = env.torch, env.L, env.ctx, env.stream, env.dist
    codec, U = wl["codec"], wl["size"]
    data = make_data(wl["gen"], U, wl["seed"] + env.rank)
    h_in = torch.from_numpy(data).pin_memory()
    d_in = h_in.cuda(non_blocking=True)
    cap = L.aocl_gpu_compress_bound(codec, U)
    d_comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_back = torch.empty(U, dtype=torch.uint8, device="cuda")
    T = L.aocl_gpu_partition_count(codec, U)
    torch.cuda.synchronize()

    def step_device(record=None):
        """compress -> [all-gather of the RAP tables into the archive index when N > 1] -> decompress; everything
        between record[0] and record[3] is one timed step."""
        if record:
            record[0].record(stream)
        ctx.compress_async(codec, d_in, d_comp)
        if record:
            record[1].record(stream)
        csz = ctx.finish()
        assert csz > 0, csz
        if dist is not None:
            _, total = archive_index(env, d_comp, T, U)
            assert total == env.world * U, (total, env.world * U)
        if record:
            record[2].record(stream)
        ctx.decompress_async(codec, d_comp, csz, d_back)
        if record:
            record[3].record(stream)
        got = ctx.finish()
        assert got == U, got
        return csz

    W = max(args.warmup, 3)
    for _ in range(W):
        csz = step_device()
    assert torch.equal(d_back, d_in), "round trip mismatch"
    launches0 = L.aocl_gpu_launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(env.local) as clocks:
        t_wall0 = time.perf_counter()
        for k in range(args.steps):
            step_device(evs[k])
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
    launches = L.aocl_gpu_launch_count() - launches0
    tc = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps       # ms, compress kernels
    tx = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps       # ms, size read-back + collective (N > 1)
    td = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps       # ms, decompress kernels
    step_ms = sum(e[0].elapsed_time(e[3]) for e in evs) / args.steps if dist is not None else tc + td
    my = [step_ms, tc, td, tx]
    step_ms, tc, td, tx = env.max_over_ranks(my)

    # ---- per-kernel timing (CUDA events around every launch, on the launching stream)
    ctx.set_profiling(True)
    ctx.compress_async(codec, d_in, d_comp); ctx.finish(); prof_c = ctx.profile()
    ctx.decompress_async(codec, d_comp, csz, d_back); ctx.finish(); prof_d = ctx.profile()
    ctx.set_profiling(False)
    kern = {k: v for k, v in dict(prof_c + prof_d).items() if v > 0.0005}
    per_rank = None
    if dist is not None:
        per_rank = [None] * env.world
        dist.all_gather_object(per_rank, {"rank": env.rank, "step_ms": my[0], "compress_ms": my[1], "decompress_ms": my[2],
                                          "exchange_ms": my[3], "kernels_ms": {k: round(v, 4) for k, v in kern.items()}})
    alg_bytes = U + csz                      # read U + write (C+F) for compress; read (C+F) + write U for decompress
    dom_name = max(kern, key=lambda k: kern[k])
    dec_name = max((k for k in kern if "decode_parts" in k), key=lambda k: kern[k], default="decode_parts_kernel")
    dec_ms = kern.get(dec_name, td)

    # ---- the separately named fastparse mode (LZ4 only): speed and ratio delta, never part of `value`
    fast = None
    if codec == LZ4 and env.world == 1:
        assert ctx.set_mode("fastparse") == 0
        try:
            fe = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(3)]
            fsz = 0
            for k in range(4):
                if k:
                    fe[k - 1][0].record(stream)
                ctx.compress_async(codec, d_in, d_comp)
                if k:
                    fe[k - 1][1].record(stream)
                fsz = ctx.finish()
                assert fsz > 0, fsz
            fms = min(a.elapsed_time(b) for a, b in fe)
            assert ctx.decompress(codec, d_comp, fsz, d_back) == U and torch.equal(d_back, d_in), "fastparse round trip mismatch"
            fast = {"mode": "fastparse (AOCL_GPU_MODE=fastparse; never the default)", "compress_ms": fms, "compress_GBps": U / fms / 1e6,
                    "ratio": fsz / U, "ratio_exact": csz / U, "ratio_delta_vs_exact_pct": 100.0 * (fsz - csz) / csz,
                    "speedup_vs_exact": tc / fms, "decodes_bit_exact": True}
        finally:
            ctx.set_mode("exact")
        ctx.compress_async(codec, d_in, d_comp)              # the exact stream again, for the checks below
        assert ctx.finish() == csz

    # ---- is the stream the reference's stream?  (outside the timed region; rank 0, N = 1)
    identity = None
    R = load_reference() if (env.rank == 0 and env.world == 1 and (with_cpu or want_identity)) else None
    if R is not None and want_identity:
        keep = {}
        ref = reference_round_trip(R, data, codec, max(T, 1), 1, 0, keep=keep)   # saturated layout: T = P(n) threads
        mine = hashlib.sha256(d_comp[:csz].cpu().numpy()).hexdigest()
        identity = {"bytes_identical_to_reference": bool(mine == keep["sha256"] and ref["csize"] == csz),
                    "reference_bytes": ref["csize"], "reference_T": ref["T"], "sha256": mine}

    # ---- end to end through the reference-facing API with pinned host buffers
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
    (e2e_s,) = env.max_over_ranks([tot / e2e_steps])
    e2e = {"value": env.world * 2 * U / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": U + csz,
           "d2h_bytes_per_step": csz + U, "ms_per_step": e2e_s * 1e3,
           "compress_ms": tot_c / e2e_steps * 1e3, "decompress_ms": (tot - tot_c) / e2e_steps * 1e3,
           "api": "aocl_llc_compress + aocl_llc_decompress, pinned host buffers; transfers pipelined with the kernels "
                  "(striped H2D behind an input watermark / slab-wise H2D-decode-D2H)"}

    # ---- what the box's PCIe / host memory can do for this step's traffic: every rank moves U up and U down with
    #      plain pinned copies at the same time (aggregate = what N concurrent e2e calls could get at best)
    pc = torch.empty(U, dtype=torch.uint8, device="cuda")
    ph = torch.empty(U, dtype=torch.uint8).pin_memory()
    pe = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    best_up = best_dn = 1e30
    for _ in range(3):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        pe[0].record(); pc.copy_(ph, non_blocking=True); pe[1].record(); ph.copy_(pc, non_blocking=True); pe[2].record()
        torch.cuda.synchronize()
        best_up, best_dn = min(best_up, pe[0].elapsed_time(pe[1])), min(best_dn, pe[1].elapsed_time(pe[2]))
    up_ms, dn_ms = env.max_over_ranks([best_up, best_dn])
    floor_ms = (U + csz) / U * up_ms + (U + csz) / U * dn_ms   # this step's H2D and D2H bytes at those rates, nothing overlapped
    e2e["pcie"] = {"h2d_GBps_aggregate": env.world * U / up_ms / 1e6, "d2h_GBps_aggregate": env.world * U / dn_ms / 1e6,
                   "transfer_floor_ms_per_step": floor_ms, "frac_of_transfer_floor": floor_ms / (e2e_s * 1e3),
                   "note": "measured with concurrent plain pinned copies of 1 GiB per rank; a step cannot be faster than its transfers"}
    del pc, ph
    cpu_baseline = cpu_baseline_sweep(R, data, codec) if (R is not None and with_cpu) else None
    del h_comp, h_back, d_in, d_comp, d_back, h_in
    torch.cuda.empty_cache()
    if env.rank != 0:
        return None
    value = env.world * 2 * U / (step_ms / 1e3) / 1e9
    detail = {"compress_GBps": env.world * U / (tc / 1e3) / 1e9, "decompress_GBps": env.world * U / (td / 1e3) / 1e9,
              "compress_ms": tc, "decompress_ms": td, "exchange_ms": tx if dist is not None else 0.0,
              "ratio": csz / U, "compressed_bytes": int(csz), "wall_s_timed_region": t_wall,
              "decoder": os.environ.get("AOCL_GPU_DECODER", "auto"),
              "kernels_ms": {k: round(v, 4) for k, v in kern.items()}}
    if identity:
        detail.update(identity)
    if fast:
        detail["fastparse"] = fast
    if per_rank:
        detail["per_rank"] = per_rank
    return {
        "metric": "rap_round_trip_GBps", "value": value, "unit": "GB/s", "n_gpus": env.world, "steps": args.steps,
        "warmup": W, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl["name"], "bytes_per_gpu": U, "l2": "inputs (1 GiB) exceed the 126 MB L2; no flush",
                   "partitions_per_frame": T,
                   "parallelism": f"frames x{env.world}" + ("; NCCL all-gather of the RAP entry tables into the archive index inside every timed step" if dist is not None else "")},
        "detail": detail,
        "roofline": env.roofline(workload_key, dom_name, alg_bytes, kern[dom_name]),
        "roofline_decompress": env.roofline(workload_key, dec_name, alg_bytes, dec_ms, {"user_GBps": U / (dec_ms / 1e3) / 1e9}),
        "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }


def sharded_frame(env, args, wl):
    """N > 1: ONE frame over all the ranks, through the library (aocl_gpu_compress_sharded / _decompress_sharded,
    csrc/llc_shard.cuh): rank r holds only its slice of the input, the library all-gathers the per-partition records
    over NCCL, forwards the boundary literals and every rank writes its own byte range of the stream.  Strong
    scaling of a single frame: the exact LZ4 encoder is one serial chain per partition, so its time barely moves --
    this is reported next to the frames-x-N headline, not instead of it.  Wall clock around the blocking collective
    calls (barrier before, device synchronised after), max over ranks."""
    torch, L, ctx, dist, llc = env.torch, env.L, env.ctx, env.dist, env.llc
    codec, U = wl["codec"], wl["size"]
    rng = llc.shard_range(codec, U, env.rank, env.world)
    if rng is None:
        return None
    first, count, boff, blen = rng
    data = make_data(wl["gen"], U, wl["seed"])               # the same frame on every rank; each keeps its slice
    box = [llc.shard_unique_id() if env.rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    assert ctx.shard_init(box[0], env.rank, env.world) == 0
    d_full = torch.from_numpy(data).cuda()
    cap = L.aocl_gpu_compress_bound(codec, U)
    d_ref = torch.empty(cap, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ref_len = ctx.compress(codec, d_full, d_ref)             # the single-GPU stream, to compare the pieces with
    assert ref_len > 0
    d_slice = d_full[boff:boff + blen].clone()
    del d_full
    d_piece = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(blen + (1 << 20), dtype=torch.uint8, device="cuda")
    tc, td = [], []
    same = back = True
    for it in range(4):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        total, off, ln = ctx.compress_sharded(codec, d_slice, U, d_piece)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        assert total == ref_len, (total, ref_len)
        same = same and bool(torch.equal(d_piece[:ln], d_ref[off:off + ln]))
        torch.cuda.synchronize(); dist.barrier()
        t2 = time.perf_counter()
        tot2, ooff, olen = ctx.decompress_sharded(codec, d_ref.data_ptr(), ref_len, d_out)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        assert tot2 == U, tot2
        back = back and bool(d_out[:olen].cpu().numpy().tobytes() == data[ooff:ooff + olen].tobytes()) if it == 0 else back
        if it:
            tc.append((t1 - t0) * 1e3); td.append((t3 - t2) * 1e3)
    cms, dms = env.max_over_ranks([min(tc), min(td)])
    flags = [None] * env.world
    dist.all_gather_object(flags, (same, back, first, count))
    del d_ref, d_piece, d_out, d_slice
    torch.cuda.empty_cache()
    return {"workload": f"ONE {U >> 20} MiB frame over {env.world} GPUs through aocl_gpu_compress_sharded / aocl_gpu_decompress_sharded "
                        f"({wl['name'].split(',')[0]})", "n_gpus": env.world, "scaling": "strong",
            "partitions_per_rank": [f[3] for f in flags],
            "compress_ms": cms, "decompress_ms": dms, "compress_GBps": U / cms / 1e6, "decompress_GBps": U / dms / 1e6,
            "pieces_identical_to_the_single_gpu_stream": all(f[0] for f in flags), "slices_round_trip": all(f[1] for f in flags),
            "collectives": "NCCL inside the library: all-gather of the per-partition records (+ point-to-point boundary literals for LZ4) "
                           "in compress, all-gather of {bytes, error} in decompress; inside the timed calls"}


def frame_variants(env, base, count):
    """`count` distinct 1 GiB frames with the statistics of `base` (BASELINE configs[3]: 16 GiB of text-like data),
    made on the GPU: the 64 MiB slabs of the base frame rotated by the frame index and every byte value sent through a
    per-frame random substitution (seed 3000 + index).  Match structure -- hence ratio and decode work -- is that of
    the base text; the bytes of every frame differ.  (Generating 16 GiB with the numpy generator takes minutes.)"""
    torch = env.torch
    U = base.numel()
    slab = 64 << 20
    for f in range(count):
        g = torch.Generator(device="cpu").manual_seed(3000 + f)
        table = torch.randperm(256, generator=g).to(torch.uint8).cuda()
        rolled = torch.roll(base, shifts=-(f % max(1, U // slab)) * slab) if U >= slab else base
        yield table[rolled.long()] if U <= (256 << 20) else torch.cat([table[p.long()] for p in rolled.split(128 << 20)])


def config3_frames(env, args, frames=16):
    """BASELINE configs[3]: device-resident LZ4 decode of 16 x 1 GiB RAP frames (65,504 partitions); with N ranks,
    rank r owns frames r, r + N, ... and the ranks all-gather the RAP entry tables (NCCL) inside the timed region."""
    torch, L, ctx, stream, dist = env.torch, env.L, env.ctx, env.stream, env.dist
    U = args.size or (1 << 30)
    mine = [f for f in range(frames) if f % env.world == env.rank]
    base = torch.from_numpy(make_data("text_like", U, 2024)).cuda()
    cap = L.aocl_gpu_compress_bound(LZ4, U)
    T = L.aocl_gpu_partition_count(LZ4, U)
    d_tmp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    streams, sizes, originals = [], [], []
    for f, frame in zip(range(frames), frame_variants(env, base, frames)):
        if f not in mine:
            continue
        torch.cuda.synchronize()
        csz = ctx.compress(LZ4, frame, d_tmp)
        assert csz > 0
        streams.append(d_tmp[:csz].clone()); sizes.append(csz)
        originals.append(hashlib.sha256(frame[: 1 << 20].cpu().numpy()).hexdigest())
    del base, d_tmp
    d_out = torch.empty(len(mine) * U, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for it in range(1 + max(2, min(args.steps, 3))):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        if dist is not None:
            tables = torch.stack([s[16:16 + 12 * T].view(torch.int32) for s in streams]) if streams else torch.zeros((0, 3 * T), dtype=torch.int32, device="cuda")
            pad = torch.zeros(((frames + env.world - 1) // env.world, 3 * T), dtype=torch.int32, device="cuda")
            pad[: tables.shape[0]] = tables
            allt = torch.empty((env.world,) + tuple(pad.shape), dtype=torch.int32, device="cuda")
            dist.all_gather_into_tensor(allt.view(-1), pad.view(-1))
            total = int((allt.view(-1, 3)[:, 2].to(torch.int64) & 0xffffffff).sum().item())
            assert total == frames * U, (total, frames * U)
        for k, s in enumerate(streams):
            ctx.decompress_async(LZ4, s, sizes[k], d_out[k * U:(k + 1) * U])
            assert ctx.finish() == U
        e1.record(stream)
        torch.cuda.synchronize()
        if it:
            times.append(e0.elapsed_time(e1))
    for k in range(len(mine)):
        assert hashlib.sha256(d_out[k * U: k * U + (1 << 20)].cpu().numpy()).hexdigest() == originals[k], "frame decode mismatch"
    (ms,) = env.max_over_ranks([min(times)])
    C_total = env.max_over_ranks([float(sum(sizes))])[0] * env.world if dist is not None else float(sum(sizes))
    del d_out, streams
    torch.cuda.empty_cache()
    return {"workload": f"LZ4 RAP decode of {frames} x {U >> 20} MiB synthetic text-like frames ({frames * T} partitions), device-resident "
                        f"(BASELINE configs[3]); frames = one base text, slabs rotated + per-frame byte substitution",
            "n_gpus": env.world, "scaling": "strong", "decompress_ms": ms, "decompress_GBps": frames * U / ms / 1e6,
            "collective": "NCCL all-gather of the RAP entry tables inside the timed region" if dist is not None else None,
            "roofline": {"bound": "hbm", "achieved": (frames * U + C_total) / ms / 1e6, "peak": env.hbm_peak * env.world, "unit": "GB/s",
                         "frac": (frames * U + C_total) / ms / 1e6 / (env.hbm_peak * env.world), "peak_kind": env.peak_kind}}


def config4_pages(env, args, total_pages=1 << 20, distinct=4096):
    """BASELINE configs[4]: batched decode of 1,048,576 independent 64 KiB pages, device-resident.  `distinct` pages are
    generated and compressed (GPU batch encoder, checked against the oracle in tests/); their compressed bytes are
    referenced total/distinct times, every reference with its own 64 KiB output page (SURVEY 8(d) allows it).  With N
    ranks every rank decodes total/N pages."""
    torch, L, ctx, stream, dist = env.torch, env.L, env.ctx, env.stream, env.dist
    from llc_b200 import gen
    PS, P = 65536, distinct
    if args.size:                                            # test mode: scale the job with --size
        total_pages = max(P, (args.size >> 16) * 64)
    N = total_pages // env.world
    pages = gen.pages(P)
    d_in = torch.from_numpy(pages.reshape(-1)).cuda()
    out = {}
    for codec, name in ((LZ4, "lz4"), (SNAPPY, "snappy")):
        bound = (int(L.LZ4_compressBound(PS) if codec == LZ4 else L.snappy_max_compressed_length(PS)) + 255) // 256 * 256
        d_comp = torch.zeros(P * bound, dtype=torch.uint8, device="cuda")
        in_ptrs = torch.arange(P, dtype=torch.int64, device="cuda") * PS + d_in.data_ptr()
        out_ptrs = torch.arange(P, dtype=torch.int64, device="cuda") * bound + d_comp.data_ptr()
        in_sizes = torch.full((P,), PS, dtype=torch.int32, device="cuda")
        out_caps = torch.full((P,), bound, dtype=torch.int32, device="cuda")
        status = torch.zeros(P, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        ctx.compress_batch_async(codec, in_ptrs, in_sizes, out_ptrs, out_caps, status, P)
        assert ctx.finish() == 0
        csz = status.cpu().numpy()
        reps = (N + P - 1) // P
        d_out = torch.empty(N * PS, dtype=torch.uint8, device="cuda")
        big_in = out_ptrs.repeat(reps)[:N].contiguous()
        big_sz = torch.from_numpy(csz.astype(np.int32)).cuda().repeat(reps)[:N].contiguous()
        big_out = torch.arange(N, dtype=torch.int64, device="cuda") * PS + d_out.data_ptr()
        big_caps = torch.full((N,), PS, dtype=torch.int32, device="cuda")
        big_status = torch.zeros(N, dtype=torch.int64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for it in range(3):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record(stream)
            ctx.decompress_batch_async(codec, big_in, big_sz, big_out, big_caps, big_status, N)
            e1.record(stream)
            assert ctx.finish() == 0
            if it:
                times.append(e0.elapsed_time(e1))
        assert bool((big_status == PS).all())
        for r in (0, reps // 2, reps - 1):
            hi = min(N, (r + 1) * P)
            assert torch.equal(d_out[r * P * PS: hi * PS], d_in[: (hi - r * P) * PS]), (name, r)
        (ms,) = env.max_over_ranks([min(times)])
        Cb = float(csz.astype(np.int64).sum()) * N / P
        tot = N * env.world
        out[name] = {"pages": tot, "decompress_ms": ms, "decompress_GBps": tot * PS / ms / 1e6, "pages_per_s": tot / ms * 1e3,
                     "ratio": float(csz.sum()) / (P * PS),
                     "roofline": {"bound": "hbm", "achieved": (N * PS + Cb) * env.world / ms / 1e6, "peak": env.hbm_peak * env.world,
                                  "unit": "GB/s", "frac": (N * PS + Cb) * env.world / ms / 1e6 / (env.hbm_peak * env.world),
                                  "peak_kind": env.peak_kind}}
        del d_out, d_comp, big_in, big_sz, big_out, big_caps, big_status
        torch.cuda.empty_cache()
    return {"workload": f"batched decode of {total_pages} independent 64 KiB columnar-like pages, device-resident (BASELINE configs[4]); "
                        f"{P} distinct pages, compressed bytes referenced {total_pages // P}x, every reference with its own output page",
            "n_gpus": env.world, "scaling": "strong", **out}


def config0_cpu(R, args):
    """BASELINE configs[0]: the reference's own CPU-runnable case -- LZ4 round trip of the 64 MiB mixed-entropy buffer
    through aocl_llc_compress / aocl_llc_decompress on ONE thread (test/codec_bench conventions)."""
    from llc_b200 import gen
    n = 64 << 20
    data = gen.mixed_entropy(n)
    r = reference_round_trip(R, data, LZ4, 1, 2, 1)
    return {"workload": "LZ4 round trip of the 64 MiB synthetic mixed-entropy buffer, reference on 1 host thread (BASELINE configs[0])",
            "compress_MBps": n / r["compress_s"] / 1e6, "decompress_MBps": n / r["decompress_s"] / 1e6,
            "compress_best_MBps": n / r["compress_best_s"] / 1e6, "decompress_best_MBps": n / r["decompress_best_s"] / 1e6,
            "ratio": r["csize"] / n, "compressed_bytes": r["csize"], "T_in_frame": r["T"]}


def run_b200(args, wl):
    env = Env()
    with_cpu = env.world == 1 and not args.no_cpu_baseline
    line = round_trip(env, args, wl, args.workload, with_cpu, want_identity=with_cpu)
    want = [] if args.configs == "none" else (["0", "2", "3", "4"] if args.configs == "all" else args.configs.split(","))
    configs = {}

    def guarded(key, fn):
        try:
            r = fn()
            if env.rank == 0 and r is not None:
                configs[key] = r
        except Exception as e:                               # noqa: BLE001 - a secondary config must not lose the headline
            if env.rank == 0:
                configs[key] = {"error": repr(e)[:300]}

    if "2" in want and args.workload != "snappy_log":
        wl2 = dict(WORKLOADS["snappy_log"])
        if args.size:
            wl2["size"] = args.size
        def snappy_line():
            r = round_trip(env, args, wl2, "snappy_log", with_cpu, want_identity=with_cpu)
            if r is None:
                return None
            return {k: r[k] for k in ("value", "unit", "ms_per_step", "config", "detail", "roofline", "roofline_decompress", "cpu_baseline", "e2e")}
        guarded("2_snappy_log_1GiB", snappy_line)
    if env.world > 1 and args.configs != "none":
        guarded("1b_one_frame_sharded_over_the_ranks", lambda: sharded_frame(env, args, wl))
    if "3" in want:
        guarded("3_lz4_16_frames_decode", lambda: config3_frames(env, args))
    if "4" in want:
        guarded("4_pages_1M_decode", lambda: config4_pages(env, args))
    if "0" in want and with_cpu and env.rank == 0:
        R = load_reference()
        if R is not None:
            guarded("0_cpu_reference_64MiB_mixed_1_thread", lambda: config0_cpu(R, args))
    if env.rank == 0:
        line["detail"]["configs"] = configs
        print(json.dumps(line))
    if env.dist is not None:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lz4_text", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=0, help="override the workload size in bytes (testing)")
    ap.add_argument("--configs", default="all", help="other BASELINE configs reported under detail.configs: all | none | e.g. 2,4")
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
