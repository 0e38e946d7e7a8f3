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


_HOST_SPLIT = r"""
import ctypes as C, os, sys
import numpy as np
import torch          # (before the library opens libnccl.so.2: torch must find its own, newer NCCL first)
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import llc_b200, oracle_lib as ol
from llc_b200 import gen
lib = ol.LlcLib(llc_b200.LIB_PATH)
orc = ol.Oracle()
rng = np.random.default_rng(4)
cases = {"text": gen.text_like(72 << 20, seed=91),
         "text+random": np.concatenate([gen.text_like(20 << 20, seed=92), rng.integers(0, 256, 14 << 20, dtype=np.uint8), gen.log_like(8 << 20, seed=93)])}
for name, data in cases.items():
    for codec in (0, 4):
        want = orc.compress(data, codec)
        r, got = lib.compress(data, codec)
        assert r == len(want) and got == want, (name, codec, r, len(want))
        r2, back = lib.decompress(got, codec, len(data))
        assert r2 == len(data) and back == data.tobytes(), (name, codec, r2)
        print("ok", name, codec, r)
# pinned buffers: every GPU's slice goes up in stripes behind its encoder's watermark
data = cases["text+random"]
pin = torch.from_numpy(data).pin_memory().numpy()
for codec in (0, 4):
    r, got = lib.compress(pin, codec)
    assert r > 0 and got == orc.compress(data, codec), ("pinned", codec, r)
    # ... and with both buffers pinned every GPU decodes its partition range through the slab pipeline
    pin_c = torch.frombuffer(bytearray(got), dtype=torch.uint8).pin_memory()
    pin_o = torch.zeros(len(data) + 64, dtype=torch.uint8).pin_memory()
    d = lib.new_desc(codec)
    d.inBuf, d.inSize, d.outBuf, d.outSize = pin_c.data_ptr(), len(got), pin_o.data_ptr(), len(data)
    r2 = lib.L.aocl_llc_decompress(C.byref(d), codec)
    assert r2 == len(data) and pin_o[:len(data)].numpy().tobytes() == data.tobytes() and int(pin_o[len(data):].sum()) == 0, ("pinned", codec, r2)
    pin_c[len(got) // 2] ^= 0x55                                            # a damaged stream fails (or decodes to something else), nothing hangs
    r3 = lib.L.aocl_llc_decompress(C.byref(d), codec)
    assert r3 < 0 or pin_o[:len(data)].numpy().tobytes() != data.tobytes(), ("pinned corrupt", codec, r3)
    print("ok pinned", codec, r, r3)
L = C.CDLL(llc_b200.LIB_PATH)
L.aocl_gpu_sharded_host_calls.restype = C.c_uint64
n_split = L.aocl_gpu_sharded_host_calls()
assert n_split in (12, 13, 14), n_split     # 2 inputs x 2 codecs x (compress + decompress) + 2 x (pinned compress + decompress) [+ failed ones do not count]
print("HOST SPLIT OK", n_split)
"""


@pytest.mark.gpu
def test_one_host_call_over_two_gpus():
    """AOCL_GPU_DEVICES=0,1 AOCL_GPU_SHARD=1: aocl_llc_compress / aocl_llc_decompress split ONE host call over both GPUs
    (a worker thread per device, slices over both PCIe links, NCCL between the GPUs) and still produce the oracle's
    bytes -- including a frame whose incompressible middle hands literals from one GPU's partitions to the other's."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    env = dict(os.environ, AOCL_GPU_DEVICES="0,1", AOCL_GPU_SHARD="1")
    env.pop("AOCL_GPU_DEVICE", None)
    r = subprocess.run([sys.executable, "-c", f"ROOT = {ROOT!r}\n" + _HOST_SPLIT], env=env, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0 and "HOST SPLIT OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
