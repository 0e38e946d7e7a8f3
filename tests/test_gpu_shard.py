"""One RAP frame over several GPUs: partition ranges per rank, NCCL exchange of the per-partition
records and of the boundary literals inside the library (include/aocl_llc_gpu.h, csrc/llc_shard.cuh)."""
import os
import subprocess
import sys

import pytest

import kat
import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_input():
    """aocl_gpu_shard_range: contiguous, disjoint, covering; equal to the reference's partition arithmetic
    (threads/threads.c:91-97,127-135) cut at floor(r*T/R)."""
    import llc_b200
    from llc_b200 import shard
    L = llc_b200.load()
    for codec, chunk in ((kat.LZ4, 262268), (kat.SNAPPY, 262144)):
        for n in (chunk * 2, chunk * 5 + 1000, 64 << 20, (1 << 30), (1 << 30) + 12345):
            T = L.aocl_gpu_partition_count(codec, n)
            common = n // T
            for R in (1, 2, 3, 4, 8):
                if T < R:
                    assert llc_b200.shard_range(codec, n, 0, R) is None
                    continue
                pos = 0
                for r in range(R):
                    first, count, boff, blen = llc_b200.shard_range(codec, n, r, R)
                    lo, hi = shard.partition_range(T, r, R)
                    assert (first, first + count) == (lo, hi)
                    assert boff == pos == common * lo
                    pos += blen
                assert pos == n
    assert llc_b200.shard_range(kat.LZ4, 1000, 0, 2) is None       # one partition cannot be sharded


@pytest.mark.gpu
def test_one_frame_on_two_gpus_matches_the_oracle():
    """torchrun, 2 ranks: the gathered pieces are byte-identical to the oracle's stream (LZ4 and Snappy, incl. the
    all-literal carry chain across the rank boundary), every rank's decode of its range matches the input."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "shard_check.py")],
                       capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0 and "SHARD CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
