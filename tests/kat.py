"""Known-answer vectors restated from the reference's own tests (paths under /root/reference).

  snappy_fail_cases / snappy_pass_cases / four_byte_offset   gtest/snappy/snappy_gtest.cpp:75-222, 225-267, 357-385
  LZ4_BOUND_KATS                                             gtest/lz4/lz4_gtest.cpp:323-326
  SNAPPY_BOUND_KATS                                          gtest/snappy/snappy_gtest.cpp:514-520
  CASES                                                      the (generator, size, seed) grid behind tests/golden/golden.json
"""
from __future__ import annotations

import hashlib

import numpy as np

LZ4, SNAPPY = 0, 4

LZ4_BOUND_KATS = {65025: 65296, 0: 16, 255: 272}
SNAPPY_BOUND_KATS = {0: 32, 393216: 458784, 2147483647: 2505397620}


def _varint(n: int) -> bytes:
    out = bytearray()
    while n >= 128:
        out.append((n & 127) | 128)
        n >>= 7
    out.append(n)
    return bytes(out)


def _literal(s: bytes) -> bytes:
    """snappy_gtest.cpp AppendLiteral: tag + optional length bytes + data."""
    n = len(s) - 1
    if n < 60:
        return bytes([n << 2]) + s
    nb = (n.bit_length() + 7) // 8
    return bytes([(59 + nb) << 2]) + n.to_bytes(nb, "little") + s


def _copy4(offset: int, length: int) -> bytes:
    """snappy_gtest.cpp AppendCopy: 4-byte-offset copies of at most 64 bytes."""
    out = bytearray()
    while length > 0:
        take = min(length, 64) if length < 68 or length >= 68 else 64
        if length > 64 and length < 68:
            take = 60
        take = min(take, 64)
        out += bytes([3 | ((take - 1) << 2)]) + offset.to_bytes(4, "little")
        length -= take
    return bytes(out)


def snappy_fail_cases(compress) -> list[bytes]:
    """`compress(bytes) -> bytes` must be a frame-less Snappy compressor (the oracle)."""
    cases = [b"\x40\x12\x00\x00", b"\x05\x12\x00\x00", b"\xfb\xff\xff\xff\x7f", b"\x80\x80\x80\x80\x80\x0a", b"\xf0"]
    dest = bytearray(compress(b"making sure we don't crash with corrupted input"))
    dest[1] = (dest[1] - 1) & 255
    dest[3] = (dest[3] + 1) & 255
    cases.append(bytes(dest))
    dest = bytearray(compress(b"A" * 100000))
    dest[0:4] = b"\x00\x00\x00\x00"
    cases.append(bytes(dest))
    dest[0:4] = b"\xff\xff\xff\xff"
    dest[4] = ord("k")
    cases.append(bytes(dest))
    dest[0:3] = b"\xff\xff\xff"
    dest[3] = 0
    cases.append(bytes(dest))
    return cases


def snappy_pass_cases() -> list[bytes]:
    return [b"", b"a", b"abc", b"abcaaaaaaa" + b"b" * 65536 + b"aaaaa" + b"abc"]


def four_byte_offset() -> tuple[bytes, bytes]:
    f1, f2 = b"012345689abcdefghijklmnopqrstuvwxyz", b"some other string"
    n2 = 100000 // len(f2)
    length = 2 * len(f1) + n2 * len(f2)
    comp = _varint(length) + _literal(f1)
    src = f1
    for _ in range(n2):
        comp += _literal(f2)
        src += f2
    comp += _copy4(len(src), len(f1))
    src += f1
    return comp, src


# (codec, generator, size, seed) grid used for the golden hashes; sizes straddle every layout switch
GOLDEN_SIZES = [0, 1, 12, 13, 100, 4096, 65535, 65536, 65546, 65547, 131072, 262143, 262144, 262267, 262268,
                393215, 393216, 393401, 393402, 1048576, 1500001, 2097229]
GOLDEN_GENS = ["mixed", "text", "log", "random", "zeros", "period7"]


def make_input(name: str, size: int) -> np.ndarray:
    from llc_b200 import gen
    if name == "mixed":
        return gen.mixed_entropy(max(size, 1), seed=1234)[:size]
    if name == "text":
        return gen.text_like(max(size, 1), seed=11)[:size]
    if name == "log":
        return gen.log_like(max(size, 1), seed=12)[:size]
    if name == "random":
        return np.random.default_rng(7).integers(0, 256, size=size, dtype=np.uint8)
    if name == "zeros":
        return np.zeros(size, dtype=np.uint8)
    if name == "period7":
        return np.resize(np.frombuffer(b"abcdefg", dtype=np.uint8), size).copy() if size else np.zeros(0, dtype=np.uint8)
    raise KeyError(name)


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()
