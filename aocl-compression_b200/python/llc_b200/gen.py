"""Synthetic inputs for the LZ4/Snappy RAP hot path.

All generators are deterministic functions of (size, seed) built on
numpy.random.default_rng, so the CPU baseline, the oracle and the GPU path always see
the same bytes.  Nothing here touches the GPU or the oracle.

  mixed_entropy  C1: 64 KiB segments cycling random / zipf words / run-length data
                 (the generator printed in BASELINE.md section 3, byte for byte)
  text_like      C2: zipf-distributed words from a 65,536-word vocabulary with
                 punctuation and newlines
  log_like       C3: "<ISO timestamp> <LEVEL> <component>: <template with numeric fields>\n"
  pages          C5: independent 64 KiB columnar-like pages
"""
from __future__ import annotations

import numpy as np


def mixed_entropy(n: int = 64 << 20, seed: int = 1234) -> np.ndarray:
    rng = np.random.default_rng(seed)
    words = [bytes(rng.integers(97, 123, size=rng.integers(2, 10)).astype(np.uint8)) for _ in range(4096)]
    out = bytearray()
    seg = 0
    while len(out) < n:
        k = seg % 3
        if k == 0:
            out += rng.integers(0, 256, size=65536, dtype=np.uint8).tobytes()
        elif k == 1:
            idx = rng.zipf(1.3, size=12000) % 4096
            out += b" ".join(words[i] for i in idx)[:65536]
        else:
            v = rng.integers(0, 256, size=512, dtype=np.uint8)
            l = rng.integers(1, 256, size=512)
            out += np.repeat(v, l).tobytes()[:65536]
        seg += 1
    return np.frombuffer(bytes(out[:n]), dtype=np.uint8).copy()


def _gather_tokens(pool: np.ndarray, starts: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """Concatenate pool[starts[i]:starts[i]+lens[i]] for all i (vectorised)."""
    ends = np.cumsum(lens)
    total = int(ends[-1]) if len(ends) else 0
    src = np.repeat(starts - (ends - lens), lens) + np.arange(total, dtype=np.int64)
    return pool[src]


def text_like(n: int, seed: int = 2024, vocab: int = 65536, zipf_a: float = 1.2) -> np.ndarray:
    rng = np.random.default_rng(seed)
    wl = rng.integers(2, 10, size=vocab).astype(np.int32) + 1          # +1: trailing separator slot
    starts = np.concatenate([[0], np.cumsum(wl)[:-1]]).astype(np.int32)
    pool = rng.integers(97, 123, size=int(wl.sum())).astype(np.uint8)
    seps = np.frombuffer(b"        ,.;\n", dtype=np.uint8)  # mostly spaces, some punctuation
    cdf = np.cumsum(1.0 / np.arange(1, vocab + 1, dtype=np.float64) ** zipf_a)
    cdf /= cdf[-1]                                         # zipf over ranks 1..vocab (inverse-CDF sampling)
    out = np.empty(n, dtype=np.uint8)
    pos = 0
    while pos < n:
        k = min(1 << 20, (n - pos) // 3 + 16)
        idx = np.minimum(np.searchsorted(cdf, rng.random(k)), vocab - 1).astype(np.int32)
        L = wl[idx]
        ends = np.cumsum(L, dtype=np.int32)
        src = np.repeat(starts[idx] - (ends - L), L) + np.arange(int(ends[-1]), dtype=np.int32)
        chunk = pool[src]
        chunk[ends - 1] = seps[rng.integers(0, len(seps), size=k)]
        m = min(len(chunk), n - pos)
        out[pos:pos + m] = chunk[:m]
        pos += m
    return out


_LEVELS = [b"TRACE", b"DEBUG", b"INFO ", b"WARN ", b"ERROR"]


def log_like(n: int, seed: int = 2025) -> np.ndarray:
    rng = np.random.default_rng(seed)
    comps = [bytes(rng.integers(97, 123, size=rng.integers(4, 13)).astype(np.uint8)) for _ in range(64)]
    vocab = [bytes(rng.integers(97, 123, size=rng.integers(3, 9)).astype(np.uint8)) for _ in range(512)]
    heads, tails = [], []
    for _ in range(256):
        nw = rng.integers(3, 9)
        words = [vocab[i] for i in rng.integers(0, 512, size=nw)]
        cut = rng.integers(1, nw)
        heads.append(b" ".join(words[:cut]) + b" id=")
        tails.append(b" " + b" ".join(words[cut:]) + b"\n")
    prefix = [b" " + lv + b" " + c + b": " for lv in _LEVELS for c in comps]

    def mkpool(items):
        lens = np.array([len(x) for x in items], dtype=np.int64)
        st = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
        return np.frombuffer(b"".join(items), dtype=np.uint8), st, lens

    p_pool, p_st, p_len = mkpool(prefix)
    h_pool, h_st, h_len = mkpool(heads)
    t_pool, t_st, t_len = mkpool(tails)
    digits = np.frombuffer(b"0123456789abcdef", dtype=np.uint8)

    out = np.empty(n, dtype=np.uint8)
    pos = 0
    t_ms = 0
    while pos < n:
        k = min(1 << 18, (n - pos) // 60 + 16)
        dt = rng.integers(0, 50, size=k).astype(np.int64)
        ms = t_ms + np.cumsum(dt)
        t_ms = int(ms[-1])
        # "2024-03-DDTHH:MM:SS.mmmZ" (24 bytes), digits computed arithmetically
        ts = np.empty((k, 24), dtype=np.uint8)
        ts[:] = np.frombuffer(b"2024-03-00T00:00:00.000Z", dtype=np.uint8)
        sec = ms // 1000
        day = 1 + (sec // 86400) % 28
        hh, mm, ss, mmm = (sec // 3600) % 24, (sec // 60) % 60, sec % 60, ms % 1000
        for col, val in ((8, day // 10), (9, day % 10), (11, hh // 10), (12, hh % 10), (14, mm // 10),
                         (15, mm % 10), (17, ss // 10), (18, ss % 10), (20, mmm // 100),
                         (21, (mmm // 10) % 10), (22, mmm % 10)):
            ts[:, col] = 48 + val
        lvl = rng.choice(5, size=k, p=[0.05, 0.25, 0.55, 0.1, 0.05])
        pre = lvl * 64 + rng.integers(0, 64, size=k)
        tpl = (rng.zipf(1.5, size=k) % 256).astype(np.int64)
        hexv = rng.integers(0, 1 << 32, size=k, dtype=np.uint64)
        hx = np.empty((k, 8), dtype=np.uint8)
        for c in range(8):
            hx[:, c] = digits[((hexv >> np.uint64(4 * (7 - c))) & np.uint64(15)).astype(np.int64)]
        # assemble 5 tokens per line: ts | prefix | head | hex | tail
        dyn = np.concatenate([ts.reshape(-1), hx.reshape(-1)])
        pool = np.concatenate([dyn, p_pool, h_pool, t_pool])
        o_p = len(dyn); o_h = o_p + len(p_pool); o_t = o_h + len(h_pool)
        st = np.stack([np.arange(k) * 24, o_p + p_st[pre], o_h + h_st[tpl],
                       k * 24 + np.arange(k) * 8, o_t + t_st[tpl]], axis=1).reshape(-1).astype(np.int64)
        ln = np.stack([np.full(k, 24), p_len[pre], h_len[tpl], np.full(k, 8), t_len[tpl]],
                      axis=1).reshape(-1).astype(np.int64)
        chunk = _gather_tokens(pool, st, ln)
        m = min(len(chunk), n - pos)
        out[pos:pos + m] = chunk[:m]
        pos += m
    return out


def pages(count: int, page_size: int = 65536, seed: int = 4000) -> np.ndarray:
    """`count` independent columnar-like pages, returned as a (count, page_size) uint8 array."""
    rng = np.random.default_rng(seed)
    out = np.empty((count, page_size), dtype=np.uint8)
    strs = [bytes(rng.integers(97, 123, size=rng.integers(3, 12)).astype(np.uint8)) + b"\0" for _ in range(1024)]
    s_pool = np.frombuffer(b"".join(strs), dtype=np.uint8)
    s_len = np.array([len(x) for x in strs], dtype=np.int64)
    s_st = np.concatenate([[0], np.cumsum(s_len)[:-1]]).astype(np.int64)
    nint = page_size // 4
    for p in range(count):
        kind = p % 3
        if kind == 0:    # sorted int32 keys, small deltas
            keys = (rng.integers(0, 1 << 20) + np.cumsum(rng.integers(0, 12, size=nint))).astype("<i4")
            out[p, :nint * 4] = keys.view(np.uint8)
            out[p, nint * 4:] = 0
        elif kind == 1:  # low-cardinality dictionary codes (int32) in short runs
            codes = np.repeat(rng.integers(0, 48, size=nint // 3 + 1), rng.integers(1, 7, size=nint // 3 + 1))[:nint]
            if len(codes) < nint:
                codes = np.resize(codes, nint)
            out[p, :nint * 4] = codes.astype("<i4").view(np.uint8)
            out[p, nint * 4:] = 0
        else:            # short strings
            idx = (rng.zipf(1.4, size=page_size // 4) % 1024).astype(np.int64)
            body = _gather_tokens(s_pool, s_st[idx], s_len[idx])
            if len(body) < page_size:
                body = np.resize(body, page_size)
            out[p] = body[:page_size]
    return out
