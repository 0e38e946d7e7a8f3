"""Hand-built LZ4 blocks and Snappy streams for decoder tests: instead of running a compressor, draw the
sequence structure directly (literal runs, match lengths, offsets) so that the decoders see shapes a greedy
encoder rarely emits -- dense runs of minimal sequences, self-overlapping matches of every period, offsets
up to 65535 (sources that have left the tile decoder's 32 KiB ring), length runs that need extension bytes.
The streams are valid by construction (block formats: algos/lz4/lz4.c:3806-4305 end-of-block rules,
algos/snappy/snappy.cc:1466-1570 element layout); the expected output is what the oracle decoder produces.
Test infrastructure only."""
from __future__ import annotations

import numpy as np

PROFILES = {
    #            ll choices (weights)            ml choices                     offset classes (near, mid, far)
    "dense":    dict(ll=[(0, 6), (1, 2), (3, 1)], ml=[(4, 6), (5, 2), (9, 1)], off=(0.6, 0.35, 0.05)),
    "far":      dict(ll=[(0, 3), (2, 3), (9, 2), (40, 1)], ml=[(4, 3), (8, 4), (20, 2), (70, 1)], off=(0.1, 0.2, 0.7)),
    "overlap":  dict(ll=[(0, 4), (1, 3), (6, 1)], ml=[(6, 3), (19, 3), (64, 2), (200, 1)], off=(0.9, 0.1, 0.0)),
    "long":     dict(ll=[(0, 3), (5, 3), (30, 2), (300, 1), (700, 1)], ml=[(4, 3), (18, 3), (280, 2), (600, 1), (70000, 0.02)],
                     off=(0.3, 0.4, 0.3)),
}


def _pick(rng, table):
    vals = np.array([v for v, _ in table])
    w = np.array([x for _, x in table], dtype=np.float64)
    return int(vals[rng.choice(len(vals), p=w / w.sum())])


def _offset(rng, cls, produced):
    near, mid, far = cls
    r = rng.random()
    if r < near:
        hi = 16
    elif r < near + mid:
        hi = 2048
    else:
        hi = 65535
    lo = 1 if hi == 16 else (17 if hi == 2048 else 2049)
    hi = min(hi, produced, 65535)
    lo = min(lo, hi)
    return int(rng.integers(lo, hi + 1))


def _lz4_len_ext(out: bytearray, v: int):
    while v >= 255:
        out.append(255)
        v -= 255
    out.append(v)


def lz4_block(profile: str, target: int, seed: int) -> bytes:
    """A valid LZ4 block that decodes to roughly `target` bytes."""
    rng = np.random.default_rng(seed)
    pr = PROFILES[profile]
    out = bytearray()
    produced = 0
    lits = rng.integers(97, 123, size=target + 4096, dtype=np.uint8).tobytes()
    lp = 0
    first = True
    while produced < target:
        ll = _pick(rng, pr["ll"])
        if first:
            ll = max(ll, 8)                                  # something to copy from
            first = False
        ml = _pick(rng, pr["ml"]) + int(rng.integers(0, 4))
        off = _offset(rng, pr["off"], produced + ll)
        tok = (min(ll, 15) << 4) | min(ml - 4, 15)
        out.append(tok)
        if ll >= 15:
            _lz4_len_ext(out, ll - 15)
        out += lits[lp:lp + ll]
        lp = (lp + ll) % target
        out += bytes((off & 255, off >> 8))
        if ml - 4 >= 15:
            _lz4_len_ext(out, ml - 4 - 15)
        produced += ll + ml
    ll = 12 + int(rng.integers(0, 30))                       # closing literals (>= 12: every end-of-block rule holds)
    out.append(min(ll, 15) << 4)
    if ll >= 15:
        _lz4_len_ext(out, ll - 15)
    out += lits[lp:lp + ll]
    return bytes(out)


def _varint(v: int) -> bytes:
    b = bytearray()
    while v >= 128:
        b.append((v & 127) | 128)
        v >>= 7
    b.append(v)
    return bytes(b)


def snappy_stream(profile: str, target: int, seed: int) -> bytes:
    """A valid raw Snappy stream (varint length + elements) that decodes to roughly `target` bytes."""
    rng = np.random.default_rng(seed)
    pr = PROFILES[profile]
    body = bytearray()
    produced = 0
    lits = rng.integers(97, 123, size=target + 4096, dtype=np.uint8).tobytes()
    lp = 0
    while produced < target:
        ll = _pick(rng, pr["ll"])
        if produced == 0:
            ll = max(ll, 8)
        if ll:
            n = ll - 1
            if n < 60:
                body.append(n << 2)
            elif n < 256:
                body += bytes((60 << 2, n))
            else:
                body += bytes((61 << 2, n & 255, n >> 8))
            body += lits[lp:lp + ll]
            lp = (lp + ll) % target
            produced += ll
        ml = min(_pick(rng, pr["ml"]) + int(rng.integers(0, 4)), 2000)
        off = _offset(rng, pr["off"], produced)
        while ml > 0:                                        # copies carry at most 64 bytes
            n = min(ml, 64)
            if ml - n in (1, 2, 3) and n > 8:
                n -= 4                                       # keep the remainder encodable (>= 4 for COPY_1 is not required, but > 0)
            r = rng.random()
            if 4 <= n <= 11 and off < 2048 and r < 0.7:
                body += bytes((1 | ((n - 4) << 2) | ((off >> 8) << 5), off & 255))
            elif r < 0.97:
                body += bytes((2 | ((n - 1) << 2), off & 255, off >> 8))
            else:
                body += bytes((3 | ((n - 1) << 2), off & 255, (off >> 8) & 255, 0, 0))
            ml -= n
            produced += n
    return _varint(produced) + bytes(body)
