"""Interoperability with what a real AOCL host produces (SURVEY 8(a5/a6), 8(f3)):

  * RAP streams with an UNSATURATED thread count (T = min(omp threads, P), threads/threads.c:55-88): a 16- or 32-core
    host never emits the saturated layout the GPU encoder writes, its partitions are tens of MiB;
  * LZ4HC streams (plain LZ4 blocks, decoded by aocl_lz4_decompress: api/codec.h:168);
  * the encoder flavours that keep their hash tables in L2 (selected by frame size: > 11 x SMs partitions), forced
    here on the small golden inputs.
The streams come from the unmodified reference compiled under oracle/_ref (it travels to the GPU box prebuilt)."""
import hashlib
import os
import subprocess
import sys
import time

import numpy as np
import pytest

import kat
import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def _interop_input(n):
    from llc_b200 import gen
    parts = [gen.text_like(n // 2, seed=41), gen.log_like(n // 4, seed=42), gen.mixed_entropy(n - n // 2 - n // 4)]
    return np.concatenate(parts)


@pytest.mark.gpu
@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
def test_gpu_decodes_reference_streams_of_unsaturated_hosts(gpu_lib, ref, codec):
    """64 MiB compressed by the reference with T = 2 .. 32 OpenMP threads (partitions of 2 .. 32 MiB) decodes bit-exact
    through the C ABI and from device memory; the device-resident rate is printed (few units -> tile decoder)."""
    import torch
    import llc_b200
    if ref is None:
        pytest.skip("oracle/_ref not built")
    n = 64 << 20
    data = _interop_input(n)
    ctx = llc_b200.GpuContext(0)
    try:
        d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
        d_want = torch.from_numpy(data).cuda()
        for T in (2, 3, 5, 8, 16, 32):
            ref.set_threads(T)
            r, stream = ref.compress(data, codec)
            assert r > 0 and stream[:8] == b"AOCL_LLC"
            assert int.from_bytes(stream[12:16], "little") == T
            r2, back = gpu_lib.decompress(stream, codec, n)                    # host buffers, C ABI
            assert r2 == n and hashlib.sha256(back).digest() == hashlib.sha256(data.tobytes()).digest(), (codec, T)
            d_comp = torch.from_numpy(np.frombuffer(stream, dtype=np.uint8).copy()).cuda()
            best = 1e9
            for _ in range(3):
                d_back.zero_()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                assert ctx.decompress(codec, d_comp, len(stream), d_back) == n
                best = min(best, time.perf_counter() - t0)
                assert torch.equal(d_back, d_want), (codec, T)
            print(f"\n[interop] codec {codec} reference stream T={T:2d}: device-resident decode {n / best / 1e9:7.1f} GB/s")
    finally:
        ref.set_threads(os.cpu_count() or 8)
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
def test_gpu_writes_the_frames_of_unsaturated_hosts(oracle, ref, codec):
    """aocl_gpu_set_partitions(K) / AOCL_GPU_PARTITIONS=K: the GPU writes, byte for byte, the frame a host with K OpenMP
    threads writes -- min(K, P(n)) partitions of n / T bytes (threads/threads.c:55-88); checked against the oracle at
    that thread count and against the compiled reference itself."""
    import torch
    import llc_b200
    n = (6 << 20) + 12345
    data = _interop_input(n)
    L = llc_b200.load()
    ctx = llc_b200.GpuContext(0)
    try:
        d_in = torch.from_numpy(data).cuda()
        d_comp = torch.zeros(L.aocl_gpu_compress_bound(codec, n), dtype=torch.uint8, device="cuda")
        d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
        P = oracle.partition_count(n, codec)
        for K in (1, 2, 3, 5, 8, 16, P + 7):
            want = oracle.compress(data, codec, max_threads=K)
            assert ctx.set_partitions(K) == 0
            assert ctx.partition_count(codec, n) == min(K, P) == oracle.partition_count(n, codec, K)
            got = ctx.compress(codec, d_in, d_comp)
            assert got == len(want) and d_comp[:got].cpu().numpy().tobytes() == want, (codec, K)
            assert ctx.decompress(codec, d_comp, got, d_back) == n and torch.equal(d_back, d_in), (codec, K)
            if ref is not None and K <= 16:
                ref.set_threads(K)
                r, stream = ref.compress(data, codec)
                assert r == len(want) and stream == want, (codec, K, "the oracle's layout is not the reference's")
        assert ctx.set_partitions(0) == 0 and ctx.set_partitions(-1) == -5
        got = ctx.compress(codec, d_in, d_comp)
        assert d_comp[:got].cpu().numpy().tobytes() == oracle.compress(data, codec)
    finally:
        if ref is not None:
            ref.set_threads(os.cpu_count() or 8)
        ctx.close()


_HOST_LAYOUT = r"""
import sys
import numpy as np, torch
sys.path[:0] = ["tests", "aocl-compression_b200/python"]
import llc_b200, oracle_lib as ol
lib = ol.LlcLib(llc_b200.LIB_PATH)
orc = ol.Oracle()
from llc_b200 import gen
n = (40 << 20) + 777                                   # above the pipelined-transfer threshold
data = np.concatenate([gen.log_like(n // 2, seed=5), gen.mixed_entropy(n - n // 2)])
pin = torch.from_numpy(data).pin_memory().numpy()      # pinned: the striped, watermarked H2D path
for codec in (ol.LZ4, ol.SNAPPY):
    want = orc.compress(data, codec, max_threads=3)
    r, stream = lib.compress(pin, codec)
    assert r == len(want) and stream == want, (codec, r, len(want))
    assert int.from_bytes(stream[12:16], "little") == 3
    r2, back = lib.decompress(stream, codec, n)
    assert r2 == n and back == data.tobytes()
print("HOST LAYOUT OK")
"""


@pytest.mark.gpu
def test_host_api_follows_aocl_gpu_partitions():
    """AOCL_GPU_PARTITIONS=3 in the environment: aocl_llc_compress (pinned input, transfers pipelined behind the
    watermark with the 3-partition stripe geometry) writes the 3-thread host's frame."""
    env = dict(os.environ, AOCL_GPU_PARTITIONS="3")
    r = subprocess.run([sys.executable, "-c", _HOST_LAYOUT], env=env, capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and "HOST LAYOUT OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


@pytest.mark.gpu
def test_gpu_decodes_lz4hc_streams(gpu_lib, ref):
    """setup(LZ4HC) succeeds, decompress(LZ4HC) is the LZ4 decoder, compress(LZ4HC) is refused."""
    import ctypes as C
    for name in ("text", "mixed", "pages"):
        data = open(os.path.join(GOLDEN_DIR, f"lz4hc_{name}.src"), "rb").read()
        stream = open(os.path.join(GOLDEN_DIR, f"lz4hc_{name}.lz4hc"), "rb").read()
        for codec in (ol.LZ4HC, ol.LZ4):
            r, back = gpu_lib.decompress(stream, codec, len(data))
            assert r == len(data) and back == data, (name, codec)
        assert gpu_lib.decompress(stream, ol.LZ4HC, len(data) - 1)[0] < 0
    d = gpu_lib.new_desc(ol.LZ4HC)
    src = np.zeros(1000, dtype=np.uint8)
    assert gpu_lib.compress(src, ol.LZ4HC, desc=d)[0] == -3                   # ERR_EXCLUDED_METHOD
    if ref is not None:                                                        # a fresh stream from the reference's HC encoder
        from llc_b200 import gen
        data = gen.text_like(3 << 20, seed=77).tobytes()
        dref = ref.new_desc(ol.LZ4HC)
        dref.level = 9
        r, stream = ref.compress(np.frombuffer(data, dtype=np.uint8), ol.LZ4HC, desc=dref)
        assert r > 0
        r2, back = gpu_lib.decompress(stream, ol.LZ4HC, len(data))
        assert r2 == len(data) and back == data


def test_lz4hc_fixtures_decode_with_the_oracle(oracle):
    """The committed LZ4HC fixtures (tests/golden/make_lz4hc.py) are valid LZ4 blocks of their sources."""
    for name in ("text", "mixed", "pages"):
        data = open(os.path.join(GOLDEN_DIR, f"lz4hc_{name}.src"), "rb").read()
        stream = open(os.path.join(GOLDEN_DIR, f"lz4hc_{name}.lz4hc"), "rb").read()
        assert oracle.decompress(stream, kat.LZ4, len(data)) == data


@pytest.mark.gpu
@pytest.mark.parametrize("flavour", ["lz4_gtab", "snappy_gtab"])
def test_l2_table_encoder_flavours_match_the_goldens(flavour):
    """The 1 GiB bench runs the encoders whose hash tables live in L2 (frames of more than 11 x SMs LZ4 partitions /
    6 x SMs Snappy fragments).  Forced here for every size (fresh process: read when the context is created), they
    must reproduce the golden hashes of the compiled reference and the oracle's streams."""
    env = dict(os.environ)
    if flavour == "lz4_gtab":
        env.update(AOCL_GPU_STAB_CTAS="0", AOCL_GPU_GTAB_CTAS="32")
    else:
        env.update(AOCL_GPU_SNAPPY_STAB_CTAS="0", AOCL_GPU_SNAPPY_GTAB_CTAS="24")
    codec = "0" if flavour == "lz4_gtab" else "4"
    kf, pf = os.path.join(ROOT, "tests", "test_gpu_kat.py"), os.path.join(ROOT, "tests", "test_gpu_parity.py")
    ids = [f"{kf}::test_gpu_matches_golden[{g}-{codec}]" for g in kat.GOLDEN_GENS]
    ids += [f"{pf}::test_compress_matches_oracle[{g}-{codec}]" for g in ("mixed", "text", "random", "zeros", "period7")]
    ids += [f"{pf}::test_round_trip_large[{codec}]", f"{pf}::test_pipelined_host_transfers[{codec}]"]
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu"] + ids,
                       env=env, capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert f"{len(ids)} passed" in r.stdout, r.stdout[-500:]
