"""Device-level C ABI (include/aocl_llc_gpu.h) with buffers resident in HBM: whole-frame calls, partition
ranges (the multi-GPU sharding primitive), batched independent pages (BASELINE config 5), and the
alternative decoder organisations.  Everything is checked bit for bit against the oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import kat
import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def ctx(torch_mod):
    import llc_b200
    if not os.path.exists(llc_b200.LIB_PATH):
        llc_b200.build()
    c = llc_b200.GpuContext(0)
    yield c
    c.close()


def dev(torch, a: np.ndarray):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
def test_device_resident_round_trip(torch_mod, ctx, oracle, corpus, codec):
    torch = torch_mod
    data = np.concatenate([corpus["text"], corpus["mixed"]])
    d_in = dev(torch, data)
    cap = ctx.L.aocl_gpu_compress_bound(codec, len(data))
    d_comp = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()                                 # torch fills on its stream, the library runs on its own
    csz = ctx.compress(codec, d_in, d_comp)
    want = oracle.compress(data, codec)
    assert csz == len(want)
    assert d_comp[:csz].cpu().numpy().tobytes() == want
    d_back = torch.zeros(len(data), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    assert ctx.decompress(codec, d_comp, csz, d_back) == len(data)
    assert torch.equal(d_back, d_in)
    # host API with device pointers: no staging, same bytes
    import llc_b200
    desc = llc_b200.AoclDesc()
    assert ctx.L.aocl_llc_setup(C.byref(desc), codec) == 0
    d_comp2 = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    desc.inBuf, desc.inSize, desc.outBuf, desc.outSize = d_in.data_ptr(), len(data), d_comp2.data_ptr(), cap
    assert ctx.L.aocl_llc_compress(C.byref(desc), codec) == csz
    assert torch.equal(d_comp2[:csz], d_comp[:csz])
    assert ctx.L.aocl_skip_rap_frame_mt(C.c_void_p(d_comp2.data_ptr()), csz) == 16 + 12 * oracle.partition_count(len(data), codec)


@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
def test_partition_range_decode(torch_mod, ctx, oracle, corpus, codec):
    """Each 'rank' decodes its own partition range of one frame into a buffer that starts at its origin."""
    torch = torch_mod
    from llc_b200 import shard
    data = np.concatenate([corpus["text"], corpus["log"]])[:5000001]
    stream = np.frombuffer(oracle.compress(data, codec), dtype=np.uint8)
    frame, entries = shard.parse_frame(stream)
    T = len(entries)
    origins = shard.output_origins(entries)
    d_stream = dev(torch, stream)
    world = 3
    pieces = []
    for r in range(world):
        lo, hi = shard.partition_range(T, r, world)
        nbytes = int(sum(int(entries[i, 2]) for i in range(lo, hi) if entries[i, 1] > 0))
        d_out = torch.zeros(max(nbytes, 1), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.decompress_range_async(codec, d_stream, len(stream), d_out, lo, hi - lo, int(origins[lo]))
        assert ctx.finish() == nbytes
        pieces.append(d_out[:nbytes].cpu().numpy())
    assert np.concatenate(pieces).tobytes() == data.tobytes()


@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
def test_batched_pages(torch_mod, ctx, oracle, codec):
    """Independent frame-less 64 KiB pages (plus ragged and empty ones), one launch for all of them."""
    torch = torch_mod
    from llc_b200 import gen
    pages = gen.pages(24)
    sizes = [65536] * 20 + [1, 13, 4097, 0]
    ins = [np.ascontiguousarray(pages[i][: sizes[i]]) for i in range(24)]
    want = [oracle.compress(p, codec) for p in ins]
    bound = max(int(oracle.bound(65536, codec)), 64)
    d_in = dev(torch, np.concatenate([np.resize(p, 65536) if len(p) else np.zeros(65536, np.uint8) for p in ins]))
    d_out = torch.zeros(24 * bound, dtype=torch.uint8, device="cuda")
    in_ptrs = torch.tensor([d_in.data_ptr() + 65536 * i for i in range(24)], dtype=torch.int64, device="cuda")
    out_ptrs = torch.tensor([d_out.data_ptr() + bound * i for i in range(24)], dtype=torch.int64, device="cuda")
    in_sizes = torch.tensor(sizes, dtype=torch.int32, device="cuda")
    out_caps = torch.full((24,), bound, dtype=torch.int32, device="cuda")
    status = torch.zeros(24, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.compress_batch_async(codec, in_ptrs, in_sizes, out_ptrs, out_caps, status, 24)
    assert ctx.finish() == 0
    st = status.cpu().numpy()
    host = d_out.cpu().numpy()
    for i in range(24):
        assert st[i] == len(want[i]), (i, st[i], len(want[i]))
        assert host[bound * i: bound * i + st[i]].tobytes() == want[i], i
    # decode the pages back in one launch
    d_back = torch.zeros(24 * 65536, dtype=torch.uint8, device="cuda")
    back_ptrs = torch.tensor([d_back.data_ptr() + 65536 * i for i in range(24)], dtype=torch.int64, device="cuda")
    csizes = torch.tensor([len(w) for w in want], dtype=torch.int32, device="cuda")
    caps = torch.tensor([max(s, 0) for s in sizes], dtype=torch.int32, device="cuda")
    status2 = torch.zeros(24, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.decompress_batch_async(codec, out_ptrs, csizes, back_ptrs, caps, status2, 24)
    assert ctx.finish() == 0
    st2 = status2.cpu().numpy()
    hb = d_back.cpu().numpy()
    for i in range(24):
        assert st2[i] == sizes[i], (i, st2[i])
        assert hb[65536 * i: 65536 * i + sizes[i]].tobytes() == ins[i].tobytes(), i
    # one damaged page is reported without disturbing the others
    bad = d_out.clone()
    bad[bound * 3 + 100: bound * 3 + 116] = 0xFF
    bad_ptrs = torch.tensor([bad.data_ptr() + bound * i for i in range(24)], dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.decompress_batch_async(codec, bad_ptrs, csizes, back_ptrs, caps, status2, 24)
    failed = -ctx.finish()
    st3 = status2.cpu().numpy()
    assert failed in (0, 1) and all(st3[i] == sizes[i] for i in range(24) if i != 3)


@pytest.mark.parametrize("mode", ["rowq", "tile", "warp"])
def test_every_decoder_organisation(mode):
    """The default picks the decoder by the number of units in a launch (row decoder for large frames, tile decoder
    below); every organisation, forced for ALL sizes, must pass the whole parity / known-answer / hostile-stream
    suite (run in a fresh process: the choice is read once when the context is created)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, AOCL_GPU_DECODER=mode)
    files = [os.path.join(root, "tests", f) for f in ("test_gpu_parity.py", "test_gpu_kat.py", "test_synth_streams.py")]
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu"] + files,
                       env=env, capture_output=True, text=True, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
def test_one_launch_many_partitions_per_cta(torch_mod, ctx, codec):
    """A frame with more partitions than the persistent decode grid has CTAs (2 per SM), decoded by ONE launch
    from device memory: every CTA takes several partitions in a row, each starting at an arbitrary output
    alignment (the host API decodes such a frame slab by slab, one partition per CTA).  Checked through the
    size-independent round-trip property; the compressed stream itself is checked against the oracle on the
    first partitions."""
    torch = torch_mod
    from llc_b200 import gen
    n = 112 << 20                                            # 448 LZ4 / Snappy partitions > 2 x 148
    data = np.concatenate([gen.text_like(64 << 20, seed=31), gen.log_like(32 << 20, seed=32), gen.mixed_entropy(16 << 20)])
    assert len(data) == n
    d_in = dev(torch, data)
    d_comp = torch.zeros(ctx.L.aocl_gpu_compress_bound(codec, n), dtype=torch.uint8, device="cuda")
    d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    csz = ctx.compress(codec, d_in, d_comp)
    assert csz > 0
    for _ in range(2):
        d_back.zero_()
        torch.cuda.synchronize()
        assert ctx.decompress(codec, d_comp, csz, d_back) == n
        assert torch.equal(d_back, d_in)


@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
def test_decode_is_deterministic_under_repetition(torch_mod, ctx, oracle, codec):
    """The tile decoder resolves match sources with racy look-throughs and pointer jumping whose intermediate
    states depend on warp timing; the bytes it produces must not.  Decode oracle-compressed streams of every
    generator twenty times each and compare every result with the input."""
    torch = torch_mod
    for name in kat.GOLDEN_GENS:
        data = kat.make_input(name, 3 * 262144 + 1234)
        comp = oracle.compress(data, codec)
        d_comp = dev(torch, np.frombuffer(comp, dtype=np.uint8))
        d_in = dev(torch, data)
        d_back = torch.zeros(len(data), dtype=torch.uint8, device="cuda")
        for it in range(20):
            d_back.fill_(0xA5)
            torch.cuda.synchronize()
            assert ctx.decompress(codec, d_comp, len(comp), d_back) == len(data), (name, it)
            assert torch.equal(d_back, d_in), (name, it)


@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
def test_ranged_decode_respects_the_callers_buffer(torch_mod, ctx, oracle, codec):
    """aocl_gpu_decompress_range_async (the sharding entry point) trusts neither the RAP entries nor the caller's
    arithmetic: a range that does not fit [d_out, d_out + out_cap) once shifted by out_origin fails BEFORE anything is
    decoded, and the bytes around the buffer stay untouched."""
    torch = torch_mod
    from llc_b200 import shard
    data = kat.make_input("text", 6 * 262272 + 999)
    comp = oracle.compress(data, codec)
    frame, entries = shard.parse_frame(np.frombuffer(comp, dtype=np.uint8))
    T = len(entries)
    assert T >= 4
    origins = shard.output_origins(entries)
    d_comp = dev(torch, np.frombuffer(comp, dtype=np.uint8))
    first, count = 2, 2
    lo = int(origins[first]); hi = int(origins[first + count]) if first + count < T else len(data)
    need = hi - lo
    guard = 4096
    buf = torch.full((guard + need + guard,), 0x5A, dtype=torch.uint8, device="cuda")
    out = buf[guard:guard + need]
    torch.cuda.synchronize()
    ctx.decompress_range_async(codec, d_comp, len(comp), out, first, count, lo)
    assert ctx.finish() == need
    assert out.cpu().numpy().tobytes() == data[lo:hi].tobytes()
    assert bool((buf[:guard] == 0x5A).all()) and bool((buf[guard + need:] == 0x5A).all())
    for what, o, origin in (("capacity one byte short", buf[guard:guard + need - 1], lo),
                            ("origin too late (first partition would start before the buffer)", out, lo + 1),
                            ("origin too early (range would end past the buffer)", out, lo - 1 if lo else 0)):
        buf.fill_(0x5A)
        torch.cuda.synchronize()
        ctx.decompress_range_async(codec, d_comp, len(comp), o, first, count, origin)
        r = ctx.finish()
        if what.startswith("origin too early") and lo == 0:
            continue
        assert r < 0, what
        assert bool((buf == 0x5A).all()), what + ": wrote although the range was refused"
    # a hostile entry table: partition `first` claims a huge decomp_len
    bad = bytearray(comp)
    bad[16 + 12 * first + 8: 16 + 12 * first + 12] = (0x7fffffff).to_bytes(4, "little")
    d_bad = dev(torch, np.frombuffer(bytes(bad), dtype=np.uint8))
    buf.fill_(0x5A)
    torch.cuda.synchronize()
    ctx.decompress_range_async(codec, d_bad, len(bad), out, first, count, lo)
    assert ctx.finish() < 0
    assert bool((buf == 0x5A).all())
