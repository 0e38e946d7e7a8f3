#!/usr/bin/env python
"""Quick A/B of one library build: device-resident compress / decompress times of a bench workload and a
digest of the compressed stream (every variant must print the same digest).  The input is generated once and
cached under /dev/shm so that a sweep of many variants only pays for it once.

  AOCL_LLC_LIB=.../lib_x/libaocl_compression.so AOCL_GPU_SNAPPY_GTAB_CTAS=18 python tools/enc_sweep.py snappy_log
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python"))
import bench  # noqa: E402


def main():
    import torch
    import llc_b200
    name = sys.argv[1] if len(sys.argv) > 1 else "lz4_text"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    extra = {  # further data shapes for perf sanity (not bench lines)
        "lz4_mixed": dict(codec=0, gen="mixed_entropy", seed=1234, size=1 << 30),
        "snappy_mixed": dict(codec=4, gen="mixed_entropy", seed=1234, size=1 << 30),
        "lz4_log": dict(codec=0, gen="log_like", seed=2025, size=1 << 30),
        "snappy_text": dict(codec=4, gen="text_like", seed=2024, size=1 << 30),
    }
    wl = bench.WORKLOADS.get(name) or extra[name]
    cache = f"/dev/shm/llc_{name}_{wl['size']}.npy"
    if os.path.exists(cache):
        data = np.load(cache)
    else:
        data = bench.make_data(wl["gen"], wl["size"], wl["seed"])
        np.save(cache, data)
    L = llc_b200.load()
    codec, U = wl["codec"], wl["size"]
    d_in = torch.from_numpy(data).cuda()
    d_comp = torch.empty(L.aocl_gpu_compress_bound(codec, U), dtype=torch.uint8, device="cuda")
    d_back = torch.empty(U, dtype=torch.uint8, device="cuda")
    ctx = llc_b200.GpuContext(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
    tc, td = [], []
    csz = 0
    for it in range(reps + 2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(stream); ctx.compress_async(codec, d_in, d_comp); e[1].record(stream)
        csz = ctx.finish()
        assert csz > 0, csz
        e[2].record(stream); ctx.decompress_async(codec, d_comp, csz, d_back); e[3].record(stream)
        got = ctx.finish()
        assert got == U, got
        torch.cuda.synchronize()
        if it >= 2:
            tc.append(e[0].elapsed_time(e[1])); td.append(e[2].elapsed_time(e[3]))
    import ctypes as C
    cnt = (C.c_uint64 * 32)()
    L.aocl_gpu_debug_counters(C.cast(cnt, C.c_void_p), 1)
    ctx.decompress_async(codec, d_comp, csz, d_back); ctx.finish()
    L.aocl_gpu_debug_counters(C.cast(cnt, C.c_void_p), 1)
    c = list(cnt)
    if any(c[:16]):
        tot = sum(c[:10]) or 1
        names = ["wait", "links", "chase", "expand", "fields", "lits", "match", "flush", "slow", "fwd"]
        print("     tile phases % :", " ".join(f"{n}={100*v/tot:.1f}" for n, v in zip(names, c[:10])),
              f"| cyc/group={sum(c[:10])//max(c[17],1)} tables={c[16]} groups={c[17]} seqs={c[18]} exec={c[19]} rounds={c[20]} slow={c[21]}", flush=True)
    ok = bool(torch.equal(d_back, d_in))
    digest = hashlib.sha1(d_comp[:csz].cpu().numpy().tobytes()).hexdigest()[:16]
    tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("AOCL_"))
    print(f"{name} [{tag}] compress_ms min {min(tc):.2f} med {sorted(tc)[len(tc)//2]:.2f} | decompress_ms min {min(td):.2f} "
          f"med {sorted(td)[len(td)//2]:.2f} | csz {csz} sha1 {digest} roundtrip {'ok' if ok else 'MISMATCH'}", flush=True)


if __name__ == "__main__":
    main()
