#!/usr/bin/env python
"""ONE frame over several GPUs, checked against the oracle (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/shard_check.py

Every rank generates the same input, keeps only its slice (aocl_gpu_shard_range), and calls
aocl_gpu_compress_sharded: the pieces, gathered on rank 0 in rank order, must be byte-identical to the oracle's
stream.  Then every rank decodes its partition range of that stream with aocl_gpu_decompress_sharded and compares
its output slice with the input.  Inputs include incompressible data, whose all-literal partitions hand their
literals across the rank boundary (lz4.c:2808-2822), and sizes that leave ranks with unequal ranges.  With
--bench it also times a 1 GiB frame (strong scaling of one frame; the collectives are inside the timed calls).
Prints one JSON line per case on rank 0; exit code 0 iff everything matched."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import llc_b200
    ap = argparse.ArgumentParser()
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--no-oracle", action="store_true", help="skip the byte comparison with the oracle (timing runs)")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = llc_b200.load()
    ctx = llc_b200.GpuContext(local)
    # the library's own communicator: rank 0's id travels over torch.distributed
    box = [llc_b200.shard_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    assert ctx.shard_init(box[0], rank, world) == 0
    from llc_b200 import gen
    import bench
    ok_all = True

    def one_case(name, data, codec, want):
        nonlocal ok_all
        n = len(data)
        rng = llc_b200.shard_range(codec, n, rank, world)
        assert rng is not None, (name, n)
        first, count, boff, blen = rng
        d_slice = torch.from_numpy(data[boff:boff + blen].copy()).cuda()
        cap = L.aocl_gpu_compress_bound(codec, n)
        d_piece = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        total, off, ln = ctx.compress_sharded(codec, d_slice, n, d_piece)
        torch.cuda.synchronize(); dist.barrier()
        t_c = time.perf_counter() - t0
        assert total > 0, (name, codec, rank, total)
        # assemble the stream on every rank (the check needs it; a real job writes the pieces at their offsets)
        meta = [None] * world
        dist.all_gather_object(meta, (off, ln))
        stream = torch.zeros(total, dtype=torch.uint8, device="cuda")
        for r in range(world):
            o, l = meta[r]
            piece = d_piece[:l].clone() if r == rank else torch.empty(l, dtype=torch.uint8, device="cuda")
            dist.broadcast(piece, src=r)
            stream[o:o + l] = piece
        assert meta[0][0] == 0 and all(meta[r][0] + meta[r][1] == meta[r + 1][0] for r in range(world - 1)) and meta[-1][0] + meta[-1][1] == total
        same = None
        if want is not None:
            same = bool(total == len(want) and stream.cpu().numpy().tobytes() == want)
        # decode: this rank only needs the header and its own partitions; it gets the whole stream here.  Its output
        # range follows the RAP entries (LZ4: literals carried over a partition boundary belong to the later partition,
        # lz4.c:2879-2896), so for incompressible data it can be much more than its share of the input
        d_out = torch.zeros(n if n <= (64 << 20) else blen + (1 << 20), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        tot2, ooff, olen = ctx.decompress_sharded(codec, stream.data_ptr(), total, d_out)
        torch.cuda.synchronize(); dist.barrier()
        t_d = time.perf_counter() - t0
        back_ok = bool(tot2 == n and d_out[:olen].cpu().numpy().tobytes() == data[ooff:ooff + olen].tobytes())
        lens = [None] * world
        dist.all_gather_object(lens, (ooff, olen, back_ok, same))
        covered = lens[0][0] == 0 and all(lens[r][0] + lens[r][1] == lens[r + 1][0] for r in range(world - 1)) and lens[-1][0] + lens[-1][1] == n
        good = covered and all(x[2] for x in lens) and all(x[3] in (None, True) for x in lens)
        ok_all = ok_all and good
        if rank == 0:
            print(json.dumps({"case": name, "codec": codec, "n": n, "ranks": world, "stream_bytes": int(total),
                              "identical_to_oracle": same, "round_trip": all(x[2] for x in lens), "ranges_cover_output": bool(covered),
                              "pieces": meta, "compress_ms": t_c * 1e3, "decompress_ms": t_d * 1e3,
                              "compress_GBps": n / t_c / 1e9, "decompress_GBps": n / t_d / 1e9}), flush=True)

    import oracle_lib as ol
    orc = None if args.no_oracle else ol.Oracle()
    rs = np.random.default_rng(9)
    cases = [
        ("text 48 MiB", gen.text_like(48 << 20, seed=61)),
        ("mixed 5 partitions (unequal ranges)", gen.mixed_entropy(5 * 262272 + 1000)),
        ("random 6 MiB (all-literal partitions: literals cross the rank boundary)", rs.integers(0, 256, 6 << 20, dtype=np.uint8)),
        ("random tail after text (carry chain ends inside a rank)", np.concatenate([gen.text_like(3 << 20, seed=62), rs.integers(0, 256, 3 << 20, dtype=np.uint8), gen.text_like(2 << 20, seed=63)])),
        ("log 40 MiB", gen.log_like(40 << 20, seed=64)),
    ]
    for name, data in cases:
        for codec in (0, 4):
            want = orc.compress(data, codec) if orc is not None else None
            one_case(name, data, codec, want)
    # one rank's piece does not fit its buffer: that rank fails, nobody hangs (the neighbours still get the boundary
    # literals they inherit from it), and the next call on the same communicator works
    data = rs.integers(0, 256, 6 << 20, dtype=np.uint8)
    for codec in (0, 4):
        first, count, boff, blen = llc_b200.shard_range(codec, len(data), rank, world)
        d_slice = torch.from_numpy(data[boff:boff + blen].copy()).cuda()
        small = rank == world // 2
        d_piece = torch.zeros(4096 if small else L.aocl_gpu_compress_bound(codec, len(data)), dtype=torch.uint8, device="cuda")
        total, off, ln = ctx.compress_sharded(codec, d_slice, len(data), d_piece)
        torch.cuda.synchronize(); dist.barrier()
        got = [None] * world
        dist.all_gather_object(got, int(total))
        good = all((g < 0) == (r == world // 2) for r, g in enumerate(got))
        ok_all = ok_all and good
        if rank == 0:
            print(json.dumps({"case": "one rank's buffer too small", "codec": codec, "returns": got, "ok": good}), flush=True)
    one_case("after the failed call", gen.text_like(8 << 20, seed=65), 0, None)
    if args.bench:
        big = bench.make_data("text_like", 1 << 30, 2024)
        for _ in range(3):
            one_case("1 GiB text (timing; third run counts)", big, 0, None)
        big = bench.make_data("log_like", 1 << 30, 2025)
        for _ in range(3):
            one_case("1 GiB log (timing; third run counts)", big, 4, None)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("SHARD CHECK", "OK" if ok_all else "FAILED", flush=True)
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
