"""The separately named `fastparse` compress mode: a valid LZ4 RAP stream that the reference's own
decoder, the oracle and the GPU decoders read bit-exact; deterministic; never selected implicitly; its ratio is
reported next to the exact mode's."""
import numpy as np
import pytest

import kat
import oracle_lib as ol

pytestmark = pytest.mark.gpu


def _inputs():
    from llc_b200 import gen
    rng = np.random.default_rng(3)
    return {
        "text": gen.text_like(5 << 20, seed=71),
        "log": gen.log_like(3 << 20, seed=72),
        "mixed": gen.mixed_entropy(4 << 20),
        "random": rng.integers(0, 256, size=(1 << 20) + 77, dtype=np.uint8),
        "zeros": np.zeros((2 << 20) + 5, dtype=np.uint8),
        "period7": np.resize(np.frombuffer(b"abcdefg", dtype=np.uint8), 3 << 20).copy(),
        "long_repeats": np.tile(gen.text_like(100_000, seed=73), 40),
        "pages": gen.pages(40).reshape(-1),
    }


def test_fastparse_streams_decode_everywhere(oracle, ref):
    import torch
    import llc_b200
    ctx = llc_b200.GpuContext(0)
    try:
        for name, data in _inputs().items():
            n = len(data)
            d_in = torch.from_numpy(data).cuda()
            d_comp = torch.zeros(ctx.L.aocl_gpu_compress_bound(kat.LZ4, n), dtype=torch.uint8, device="cuda")
            d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            assert ctx.set_mode("exact") == 0
            exact = ctx.compress(kat.LZ4, d_in, d_comp)
            assert exact > 0 and d_comp[:exact].cpu().numpy().tobytes() == oracle.compress(data, kat.LZ4), name   # default untouched
            assert ctx.set_mode("fastparse") == 0
            c1 = ctx.compress(kat.LZ4, d_in, d_comp)
            assert c1 > 0, name
            stream = d_comp[:c1].cpu().numpy().tobytes()
            assert stream[:8] == b"AOCL_LLC" and int.from_bytes(stream[12:16], "little") == ctx.L.aocl_gpu_partition_count(kat.LZ4, n)
            c2 = ctx.compress(kat.LZ4, d_in, d_comp)
            assert c2 == c1 and d_comp[:c2].cpu().numpy().tobytes() == stream, (name, "not deterministic")
            assert oracle.decompress(stream, kat.LZ4, n) == data.tobytes(), (name, "oracle decoder")
            if ref is not None:
                r, back = ref.decompress(stream, kat.LZ4, n)
                assert r == n and back == data.tobytes(), (name, "reference decoder")
            assert ctx.decompress(kat.LZ4, d_comp, c1, d_back) == n and torch.equal(d_back, d_in), (name, "GPU decoder")
            print(f"\n[fastparse] {name:>12}: exact {exact / n:.4f}  fastparse {c1 / n:.4f}  delta {100.0 * (c1 - exact) / exact:+.2f} %")
            assert c1 <= exact * 1.35 + 4096, (name, exact, c1)
        assert ctx.set_mode("exact") == 0 and ctx.set_mode("nonsense") == -4
    finally:
        ctx.close()


def test_fastparse_with_an_imitated_host_layout(oracle):
    """fastparse packs positions in 19 bits.  With aocl_gpu_set_partitions() the partitions can be larger than that:
    those frames take the exact encoder (and are then the K-thread host's bytes); smaller ones stay fastparse."""
    import torch
    import llc_b200
    from llc_b200 import gen
    data = gen.text_like(3 << 20, seed=75)
    n = len(data)
    ctx = llc_b200.GpuContext(0)
    try:
        d_in = torch.from_numpy(data).cuda()
        d_comp = torch.zeros(ctx.L.aocl_gpu_compress_bound(kat.LZ4, n), dtype=torch.uint8, device="cuda")
        d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
        assert ctx.set_mode("fastparse") == 0
        assert ctx.set_partitions(2) == 0                                     # 1.5 MiB partitions: exact encoder
        c = ctx.compress(kat.LZ4, d_in, d_comp)
        assert c > 0 and d_comp[:c].cpu().numpy().tobytes() == oracle.compress(data, kat.LZ4, max_threads=2)
        assert ctx.set_partitions(8) == 0                                     # 384 KiB partitions: fastparse
        c8 = ctx.compress(kat.LZ4, d_in, d_comp)
        stream = d_comp[:c8].cpu().numpy().tobytes()
        assert c8 > 0 and int.from_bytes(stream[12:16], "little") == 8 and stream != oracle.compress(data, kat.LZ4, max_threads=8)
        assert oracle.decompress(stream, kat.LZ4, n) == data.tobytes()
        assert ctx.decompress(kat.LZ4, d_comp, c8, d_back) == n and torch.equal(d_back, d_in)
    finally:
        ctx.close()
