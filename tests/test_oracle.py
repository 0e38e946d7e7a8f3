"""CPU tests of the checker itself: the oracle restatement against (1) the golden hashes produced by the
compiled reference (tests/golden/golden.json), (2) the compiled reference when it is present
(oracle/_ref), and (3) the known-answer vectors of the reference's own tests (tests/kat.py)."""
import json
import os
import struct

import numpy as np
import pytest

import kat
import oracle_lib as ol

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
@pytest.mark.parametrize("name", kat.GOLDEN_GENS)
def test_oracle_matches_golden(oracle, codec, name):
    cases = [c for c in GOLDEN["cases"] if c["codec"] == codec and c["gen"] == name]
    assert len(cases) == len(kat.GOLDEN_SIZES)
    for c in cases:
        data = kat.make_input(name, c["size"])
        assert kat.sha(data.tobytes()) == c["in_sha256"], "generator drifted"
        got = oracle.compress(data, codec)
        assert len(got) == c["out_len"], (name, c["size"])
        assert kat.sha(got) == c["out_sha256"], (name, c["size"])
        if "out_hex" in c:
            assert got.hex() == c["out_hex"]
        assert oracle.partition_count(c["size"], codec) == c["partitions"]
        assert oracle.decompress(got, codec, c["size"]) == data.tobytes()


@pytest.mark.parametrize("codec", [kat.LZ4, kat.SNAPPY])
def test_oracle_matches_reference_at_every_thread_count(oracle, ref, corpus, codec):
    """Output depends on T = min(threads, P(n)) (threads/threads.c:55-88): check unsaturated layouts too."""
    if ref is None:
        pytest.skip("compiled reference not available")
    data = corpus["mixed"][: 3 << 20]
    for threads in (1, 2, 3, 5, 8, 64):
        ref.set_threads(threads)
        r, want = ref.compress(data, codec)
        got = oracle.compress(data, codec, max_threads=threads)
        assert got == want, (codec, threads)
        assert oracle.decompress(want, codec, len(data)) == data.tobytes()


def test_reference_decodes_oracle_streams(oracle, ref, corpus):
    if ref is None:
        pytest.skip("compiled reference not available")
    for codec in (kat.LZ4, kat.SNAPPY):
        for name in ("text", "random", "zeros"):
            data = corpus[name][: 1500001 if name == "text" else 1 << 20]
            stream = oracle.compress(data, codec)
            ref.set_threads(64)
            r, back = ref.decompress(stream, codec, len(data))
            assert r == len(data) and back == data.tobytes()


def test_bound_kats(oracle):
    for n, want in kat.LZ4_BOUND_KATS.items():          # gtest/lz4/lz4_gtest.cpp:323-326
        assert oracle.bound(n, kat.LZ4) == want
    for n, want in kat.SNAPPY_BOUND_KATS.items():       # gtest/snappy/snappy_gtest.cpp:514-520
        assert oracle.bound(n, kat.SNAPPY) == want


def test_partition_arithmetic(oracle):
    # SURVEY section 8 table (threads/threads.c:57,74-88)
    assert oracle.partition_count(64 << 20, kat.LZ4) == 256
    assert oracle.partition_count(1 << 30, kat.LZ4) == 4094
    assert oracle.partition_count(1 << 30, kat.SNAPPY) == 4096
    assert oracle.partition_count(262267, kat.LZ4) == 1 and oracle.partition_count(262268, kat.LZ4) == 1
    assert oracle.partition_count(393401, kat.LZ4) == 1 and oracle.partition_count(393402, kat.LZ4) == 2
    assert oracle.partition_count(1 << 30, kat.LZ4, max_threads=8) == 8


def test_rap_header_layout(oracle, corpus):
    """threads_gtest.cpp:512-545: magic @0, frame_len @8, T as int16 @12, child count 0 @14."""
    data = corpus["text"][:1500001]
    for codec in (kat.LZ4, kat.SNAPPY):
        s = oracle.compress(data, codec)
        T = oracle.partition_count(len(data), codec)
        assert s[:8] == b"AOCL_LLC"
        frame, main, child = struct.unpack_from("<IHH", s, 8)
        assert (frame, main, child) == (16 + 12 * T, T, 0)
        entries = [struct.unpack_from("<III", s, 16 + 12 * i) for i in range(T)]
        assert sum(e[2] for e in entries) == len(data)
        pos = frame + (0 if codec == kat.LZ4 else len(kat._varint(len(data))))
        for off, clen, dlen in entries:                # offsets are contiguous and absolute
            assert off == pos
            pos += clen
        assert pos == len(s)


def test_all_literal_carry_chain(oracle):
    """Incompressible partitions forward their bytes to the next token (lz4.c:2808-2822)."""
    data = np.random.default_rng(3).integers(0, 256, size=1 << 20, dtype=np.uint8)
    s = oracle.compress(data, kat.LZ4)
    entries = [struct.unpack_from("<III", s, 16 + 12 * i) for i in range(4)]
    assert entries[:3] == [(64, 0, 0)] * 3 and entries[3][2] == 1 << 20
    assert oracle.decompress(s, kat.LZ4, 1 << 20) == data.tobytes()


def test_snappy_decoder_kats(oracle):
    comp = lambda b: oracle.compress(np.frombuffer(b, dtype=np.uint8), kat.SNAPPY)
    for bad in kat.snappy_fail_cases(comp):
        assert oracle.decompress(bad, kat.SNAPPY, 4 << 20) is None
    for good in kat.snappy_pass_cases():
        s = comp(good)
        assert oracle.snappy_uncompressed_length(s) == len(good)
        assert oracle.decompress(s, kat.SNAPPY, len(good)) == good
    c4, src = kat.four_byte_offset()
    assert oracle.decompress(c4, kat.SNAPPY, len(src)) == src
    assert oracle.decompress(b"\x01\x00x", kat.SNAPPY, 1) == b"x"


def test_lz4_decoder_rejects_damage(oracle, corpus):
    data = corpus["text"][:800]                         # api_gtest.cpp:570-594 uses 800-byte inputs
    s = bytearray(oracle.compress(data, kat.LZ4))
    assert oracle.decompress(bytes(s), kat.LZ4, 800) == data.tobytes()
    assert oracle.decompress(bytes(s), kat.LZ4, 799) is None          # short destination, lz4_gtest.cpp:289-313
    assert oracle.decompress(bytes(s[:-1]), kat.LZ4, 800) is None      # truncated input
    assert oracle.decompress(b"", kat.LZ4, 800) is None


def test_empty_input_streams(oracle):
    empty = np.zeros(0, dtype=np.uint8)
    assert oracle.compress(empty, kat.LZ4) == b"\x00"                  # api_gtest.cpp:638-673
    assert oracle.compress(empty, kat.SNAPPY) == b"\x00"
