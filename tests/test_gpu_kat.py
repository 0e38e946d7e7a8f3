"""GPU decoders/encoders against the reference's own known-answer vectors and the golden hashes,
through the C ABI; plus hostile-input safety (no crash, negative return or intact-size mismatch)."""
import json
import os

import numpy as np
import pytest

import kat
import oracle_lib as ol

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
@pytest.mark.parametrize("name", kat.GOLDEN_GENS)
def test_gpu_matches_golden(gpu_lib, codec, name):
    for c in [c for c in GOLDEN["cases"] if c["codec"] == codec and c["gen"] == name]:
        data = kat.make_input(name, c["size"])
        r, got = gpu_lib.compress(data, codec)
        assert r == c["out_len"], (name, c["size"], r)
        assert kat.sha(got) == c["out_sha256"], (name, c["size"])
        r2, back = gpu_lib.decompress(got, codec, max(c["size"], 1))
        assert r2 == c["size"] and back == data.tobytes()


def test_snappy_kats(gpu_lib, oracle):
    comp = lambda b: oracle.compress(np.frombuffer(b, dtype=np.uint8), kat.SNAPPY)
    for bad in kat.snappy_fail_cases(comp):
        r, _ = gpu_lib.decompress(bad, kat.SNAPPY, 4 << 20)
        assert r < 0, bad[:8]
    for good in kat.snappy_pass_cases():
        r, s = gpu_lib.compress(np.frombuffer(good, dtype=np.uint8), kat.SNAPPY)
        assert s == comp(good)
        r2, back = gpu_lib.decompress(s, kat.SNAPPY, max(len(good), 1))
        assert r2 == len(good) and back == good
    c4, src = kat.four_byte_offset()
    r, back = gpu_lib.decompress(c4, kat.SNAPPY, len(src))
    assert r == len(src) and back == src
    r, back = gpu_lib.decompress(b"\x01\x00x", kat.SNAPPY, 1)
    assert (r, back) == (1, b"x")


def test_empty_and_argument_errors(gpu_lib):
    empty = np.zeros(0, dtype=np.uint8)
    for codec in (kat.LZ4, kat.SNAPPY):                        # api_gtest.cpp:638-673: inSize == 0 -> 1 byte
        r, s = gpu_lib.compress(empty, codec, cap=64)
        assert (r, s) == (1, b"\x00")
        r, _ = gpu_lib.decompress(b"", codec, 16)              # api_gtest.cpp:832-905
        assert r < 0
    r, _ = gpu_lib.compress(np.arange(100, dtype=np.uint8), kat.SNAPPY, cap=100)   # api/codec.cpp:262-265
    assert r < 0
    r, _ = gpu_lib.compress(np.random.default_rng(0).integers(0, 256, 4096, dtype=np.uint8), kat.LZ4, cap=100)
    assert r < 0                                               # limitedOutput refusal, lz4.c:2523-2540


def test_lz4_limited_output_matches_oracle(gpu_lib, oracle, corpus):
    """Frame-less LZ4 with outSize < LZ4_compressBound runs the reference's limitedOutput checks."""
    data = corpus["text"][:30000]
    full = oracle.compress(data, kat.LZ4)
    for cap in (len(full) + 20, len(full) + 1, len(full), len(full) - 1, len(full) // 2, 17):
        want = oracle.compress(data, kat.LZ4, cap=cap)
        r, got = gpu_lib.compress(data, kat.LZ4, cap=cap)
        if want is None:
            assert r < 0, cap
        else:
            assert got == want, cap


def test_short_destination_and_truncation(gpu_lib, oracle, corpus):
    data = corpus["text"][:800]
    for codec in (kat.LZ4, kat.SNAPPY):
        s = oracle.compress(data, codec)
        assert gpu_lib.decompress(s, codec, 800)[0] == 800
        assert gpu_lib.decompress(s, codec, 799)[0] < 0        # lz4_gtest.cpp:289-313
        assert gpu_lib.decompress(s[:-1], codec, 800)[0] < 0


@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
def test_corrupted_streams_are_safe(gpu_lib, oracle, corpus, codec):
    """api_gtest.cpp:907-932: damage in the middle of the stream gives an error or different data, never a crash;
    the bytes outside the destination range stay untouched."""
    data = corpus["mixed"][: 1 << 20]
    good = bytearray(oracle.compress(data, codec))
    rng = np.random.default_rng(5)
    for trial in range(24):
        s = bytearray(good)
        pos = int(rng.integers(0, len(s) - 16))
        if trial % 3 == 0:
            pos = int(rng.integers(0, 16 + 12 * 4))             # hit the RAP frame itself
        s[pos:pos + 16] = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
        r, back = gpu_lib.decompress(bytes(s), codec, len(data))
        assert r < 0 or r == len(data)
        want = oracle.decompress(bytes(s), codec, len(data))
        if want is not None and r == len(data):
            assert back == want
    r, back = gpu_lib.decompress(bytes(good), codec, len(data))    # library still healthy afterwards
    assert r == len(data) and back == data.tobytes()


def test_measure_stats_fields(gpu_lib, corpus):
    """api_gtest.cpp:533-538: cSize / cTime / cSpeed arithmetic (api/api.cpp:69-75)."""
    data = corpus["text"][: 1 << 20]
    d = gpu_lib.new_desc(kat.LZ4, stats=1)
    assert d.workBuf is None
    r, s = gpu_lib.compress(data, kat.LZ4, desc=d)
    assert d.cSize == r and d.cTime > 0
    assert abs(d.cSpeed - len(data) * 1000.0 / d.cTime) <= 1e-3 * d.cSpeed
    r2, _ = gpu_lib.decompress(s, kat.LZ4, len(data), desc=d)
    assert d.dSize == r2 and abs(d.dSpeed - r2 * 1000.0 / d.dTime) <= 1e-3 * d.dSpeed


def test_opt_off_gives_frameless_lz4(gpu_lib, oracle, corpus):
    """optOff=1 selects the reference's single-threaded LZ4 layout (lz4.c:4927-4932): no RAP frame."""
    import ctypes as C
    data = corpus["text"][:600000]
    d0 = gpu_lib.new_desc(kat.LZ4)
    gpu_lib.L.aocl_llc_destroy(C.byref(d0), kat.LZ4)
    d = gpu_lib.new_desc(kat.LZ4, opt_off=1)
    try:
        r, s = gpu_lib.compress(data, kat.LZ4, desc=d)
        assert s[:8] != b"AOCL_LLC"
        assert s == oracle.compress(data, kat.LZ4, max_threads=1)
        assert gpu_lib.decompress(s, kat.LZ4, len(data), desc=d)[1] == data.tobytes()
    finally:
        gpu_lib.L.aocl_llc_destroy(C.byref(d), kat.LZ4)
        gpu_lib.new_desc(kat.LZ4)
