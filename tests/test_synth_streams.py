"""Decoders on hand-built streams (tests/synth_streams.py): the oracle against the compiled reference on the
CPU, the CUDA path against the oracle on the GPU -- at several output alignments, because the tile decoder
works in 16-byte units relative to the output pointer."""
import numpy as np
import pytest

import kat
import oracle_lib as ol
import synth_streams as ss

CASES = [(p, s) for p in ss.PROFILES for s in (1, 2)]
TARGET = 700_000


def _stream(codec, profile, seed):
    return ss.lz4_block(profile, TARGET, seed) if codec == kat.LZ4 else ss.snappy_stream(profile, TARGET, seed)


@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
@pytest.mark.parametrize("profile,seed", CASES)
def test_oracle_decodes_synthetic_streams_like_the_reference(oracle, codec, profile, seed):
    stream = _stream(codec, profile, seed)
    cap = TARGET + 200_000
    want = oracle.decompress(stream, codec, cap)
    assert want is not None and len(want) >= TARGET
    if codec == kat.SNAPPY:
        assert len(want) == oracle.snappy_uncompressed_length(stream)
    ref = ol.ref_lib()
    if ref is not None:                                      # the unmodified reference, when it was compiled here
        r, got = ref.decompress(stream, codec, cap)
        assert r == len(want) and got == want
    # a truncated stream must not decode to the full length
    cut = oracle.decompress(stream[:len(stream) // 2], codec, cap)
    assert cut is None or len(cut) < len(want)


@pytest.mark.gpu
@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
@pytest.mark.parametrize("profile,seed", CASES)
def test_gpu_decodes_synthetic_streams(oracle, codec, profile, seed):
    import torch
    import llc_b200
    stream = _stream(codec, profile, seed)
    cap = TARGET + 200_000
    want = oracle.decompress(stream, codec, cap)
    assert want is not None
    n = len(want)
    ctx = llc_b200.GpuContext(0)
    try:
        d_comp = torch.from_numpy(np.frombuffer(stream, dtype=np.uint8).copy()).cuda()
        d_want = torch.from_numpy(np.frombuffer(want, dtype=np.uint8).copy()).cuda()
        d_buf = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
        for shift in (0, 1, 7, 15, 16, 33):                  # output alignment relative to 16 bytes
            d_buf.fill_(0x5A)
            torch.cuda.synchronize()
            out = d_buf[shift:shift + n]
            assert ctx.decompress(codec, d_comp, len(stream), out) == n, (profile, seed, shift)
            assert torch.equal(out, d_want), (profile, seed, shift)
            assert bool((d_buf[:shift] == 0x5A).all()) and bool((d_buf[shift + n:] == 0x5A).all()), "wrote outside the output"
        # capacity one byte short: refused, like the reference
        torch.cuda.synchronize()
        assert ctx.decompress(codec, d_comp, len(stream), d_buf[:n - 1]) < 0
    finally:
        ctx.close()
