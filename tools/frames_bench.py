#!/usr/bin/env python
"""BASELINE configs[3]: LZ4 RAP-frame decompress of 16 GiB of synthetic data, frames sharded across the GPUs of
one box (each frame holds 4094 partitions; with 16 frames and N ranks, rank r owns frames r, r+N, ...; the only
exchange is the NCCL all-gather of the RAP entry tables, 12 bytes per partition).

To bound host time every rank generates `--distinct` different 1 GiB text-like frames (seeds 3000 + frame index),
compresses them on its GPU (the encoder is byte-identical to the oracle, tests/test_gpu_parity.py), checks the
round trip of each, and then decodes its share of the 16 frames from those compressed streams (frame f uses
stream f mod distinct), every frame into its own slice of one output buffer.  Prints one JSON line (rank 0).

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/frames_bench.py
"""
import argparse
import json
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--distinct", type=int, default=2)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--codec", default="lz4")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    codec = 0 if args.codec == "lz4" else 4
    gen_name = "text_like" if codec == 0 else "log_like"
    U = 1 << 30
    L = llc_b200.load()
    ctx = llc_b200.GpuContext(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    mine = [f for f in range(args.frames) if f % world == rank]
    cap = L.aocl_gpu_compress_bound(codec, U)
    d_back = torch.empty(len(mine) * U, dtype=torch.uint8, device="cuda")
    comp, sizes = {}, {}
    for f in mine:
        s = f % args.distinct
        if s in comp:
            continue
        data = bench.make_data(gen_name, U, 3000 + s)
        d_in = torch.from_numpy(data).cuda()
        d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        csz = ctx.compress(codec, d_in, d_c)
        assert csz > 0
        assert ctx.decompress(codec, d_c, csz, d_back[:U]) == U and torch.equal(d_back[:U], d_in), "round trip"
        comp[s], sizes[s] = d_c[:csz].clone(), csz
        del d_in, d_c
    T = L.aocl_gpu_partition_count(codec, U)
    best = 1e9
    for it in range(args.iters + 1):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if dist is not None:                                 # the only exchange: every rank learns every frame's layout
            with torch.cuda.stream(stream):
                for j, f in enumerate(mine):
                    table = comp[f % args.distinct][16:16 + 12 * T].view(torch.int32)
                    gathered = torch.empty(world * table.numel(), dtype=torch.int32, device="cuda")
                    dist.all_gather_into_tensor(gathered, table.contiguous())
        for j, f in enumerate(mine):
            s = f % args.distinct
            ctx.decompress_async(codec, comp[s], sizes[s], d_back[j * U:(j + 1) * U])
            got = ctx.finish()
            assert got == U, got
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if it > 0:
            best = min(best, ms)
    if rank == 0:
        total = args.frames * U
        print(json.dumps({"workload": f"{args.frames} x 1 GiB {args.codec} RAP frames ({args.frames * T} partitions) decoded on {world} GPU(s), "
                          f"frames round-robin over ranks, all-gather of the RAP entry tables only; {args.distinct} distinct frames per rank",
                          "n_gpus": world, "decompress_ms": best, "decompress_GBps": total / best / 1e6,
                          "per_gpu_GBps": total / best / 1e6 / world}))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
