"""Executable model of the tile decoder (aocl-compression_b200/csrc/decode_tile.cuh) in numpy: the same
phases on the same data structures -- per-position "next token" links, pointer doubling to 64-hop links,
the chase with binary descent, sequence starts by binary-decomposed hops, block scan, start bitmap + covering
sequence per 32-byte row, the per-byte source step with look-through into already written entries, pointer
jumping over the 16-bit pointer/value table, flush -- so that the algorithm (not the CUDA code) can be
checked against the oracle on the CPU, where this container has no GPU.  Two schedules of the source step
are modelled: every earlier row already written ("front") and no row written yet ("none"); the real kernel
is somewhere in between (schedule "16": waves of 16 rows that cannot see each other, one row per warp), and
the bytes must not depend on it.

LZ4 blocks and raw Snappy streams (one element per "sequence") share everything but the parse.  Only the regular
path is modelled; an irregular sequence (a length that needs a 255 extension byte, the
closing sequences of a block) ends the group and is executed by a plain sequential step, as in the kernel.
Test infrastructure only (tests/test_tile_model.py)."""
from __future__ import annotations

import numpy as np

CHUNK = 4096
MARGIN = 384
SPAN = 16384
NONE = 0xFFFF
PTR = 0x4000
KNOWN = 0xFF00
MAXSEQ = 1024
RING = 32768


def _links_lz4(bp: np.ndarray, cbase: int, iend: int):
    """n1 for every byte position of the chunk (TileLz4::parse + ends_inside)."""
    i = np.arange(CHUNK)
    tok = bp[i].astype(np.int64)
    e1 = bp[i + 1].astype(np.int64)
    nib_l, nib_m = tok >> 4, tok & 15
    ext_l, ext_m = nib_l == 15, nib_m == 15
    ll = nib_l + np.where(ext_l, e1, 0)
    lit = i + 1 + ext_l
    q = lit + ll
    e2 = bp[np.minimum(q + 2, len(bp) - 1)].astype(np.int64)
    special = (ext_l & (e1 == 255)) | (ext_m & (e2 == 255))
    nxt = q + 2 + ext_m
    ok = (cbase + i < iend) & ~special & (cbase + lit + ll + 8 <= iend)
    return np.where(ok, nxt, NONE).astype(np.int64)


def _fields_lz4(bp: np.ndarray, p: np.ndarray):
    tok = bp[p].astype(np.int64)
    e1 = bp[p + 1].astype(np.int64)
    nib_l, nib_m = tok >> 4, tok & 15
    ext_l, ext_m = nib_l == 15, nib_m == 15
    ll = nib_l + np.where(ext_l, e1, 0)
    lit = p + 1 + ext_l
    q = lit + ll
    off = bp[q].astype(np.int64) | (bp[q + 1].astype(np.int64) << 8)
    ml = 4 + nib_m + np.where(ext_m, bp[q + 2].astype(np.int64), 0)
    return ll, ml, off, lit


def _double(tbl: np.ndarray) -> np.ndarray:
    inside = tbl < CHUNK
    return np.where(inside, tbl[np.where(inside, tbl, 0)], NONE)


def _slow_sequence_lz4(src: np.ndarray, ip: int, out: bytearray, cap: int, iend: int):
    """One checked sequence (tile_slow_step_lz4 for a last/vanilla block).  Returns (ip, done)."""
    tok = int(src[ip]); ip += 1
    ll = tok >> 4
    if ll == 15:
        while True:
            b = int(src[ip]); ip += 1; ll += b
            if b != 255:
                break
    out += src[ip:ip + ll].tobytes(); ip += ll
    if ip >= iend:
        return ip, True
    off = int(src[ip]) | (int(src[ip + 1]) << 8); ip += 2
    ml = tok & 15
    if ml == 15:
        while True:
            b = int(src[ip]); ip += 1; ml += b
            if b != 255:
                break
    ml += 4
    assert 0 < off <= len(out)
    for _ in range(ml):
        out.append(out[-off])
    return ip, False


def _links_snappy(bp: np.ndarray, cbase: int, iend: int):
    """n1 for every byte position of the chunk (TileSnappy::parse + ends_inside): one element per hop."""
    i = np.arange(CHUNK)
    tag = bp[i].astype(np.int64)
    b1 = bp[i + 1].astype(np.int64)
    kind, v = tag & 3, tag >> 2
    nxt = np.full(CHUNK, NONE, dtype=np.int64)
    lit_s = (kind == 0) & (v < 60)
    lit_1 = (kind == 0) & (v == 60)
    nxt = np.where(lit_s, i + 1 + v + 1, nxt)
    nxt = np.where(lit_1, i + 2 + b1 + 1, nxt)
    nxt = np.where(kind == 1, i + 2, nxt)
    nxt = np.where(kind == 2, i + 3, nxt)
    ok = (cbase + i < iend) & (nxt != NONE) & (cbase + nxt <= iend)
    return np.where(ok, nxt, NONE).astype(np.int64)


def _fields_snappy(bp: np.ndarray, p: np.ndarray):
    tag = bp[p].astype(np.int64)
    b1 = bp[p + 1].astype(np.int64)
    b2 = bp[p + 2].astype(np.int64)
    kind, v = tag & 3, tag >> 2
    ll = np.where(kind == 0, np.where(v < 60, v + 1, b1 + 1), 0)
    lit = np.where((kind == 0) & (v == 60), p + 2, p + 1)
    ml = np.where(kind == 1, 4 + (v & 7), np.where(kind == 2, 1 + v, 0))
    off = np.where(kind == 1, ((tag >> 5) << 8) | b1, np.where(kind == 2, b1 | (b2 << 8), 0))
    return ll, ml, off, lit


def _slow_element_snappy(src: np.ndarray, ip: int, out: bytearray, cap: int, iend: int):
    """One checked element (tile_slow_step_snappy).  Returns (ip, done)."""
    if ip >= iend:
        assert len(out) == cap
        return ip, True
    tag = int(src[ip]); ip += 1
    kind = tag & 3
    if kind == 0:
        n = (tag >> 2) + 1
        if n > 60:
            nb = n - 60
            n = int.from_bytes(src[ip:ip + nb].tobytes(), "little") + 1
            ip += nb
        out += src[ip:ip + n].tobytes()
        return ip + n, False
    if kind == 1:
        n, off = 4 + ((tag >> 2) & 7), ((tag >> 5) << 8) | int(src[ip]); ip += 1
    elif kind == 2:
        n, off = 1 + (tag >> 2), int(src[ip]) | (int(src[ip + 1]) << 8); ip += 2
    else:
        n, off = 1 + (tag >> 2), int.from_bytes(src[ip:ip + 4].tobytes(), "little"); ip += 4
    assert 0 < off <= len(out)
    for _ in range(n):
        out.append(out[-off])
    return ip, False


LZ4_FMT = dict(links=_links_lz4, fields=_fields_lz4, slow=_slow_sequence_lz4, end_slack=12, max_seq=1024)
SNAPPY_FMT = dict(links=_links_snappy, fields=_fields_snappy, slow=_slow_element_snappy, end_slack=0, max_seq=1280)


def decode_lz4_block(stream: bytes, cap: int, a: int = 0, schedule: str = "front", stats: dict | None = None) -> bytes:
    """Decode one LZ4 block the way the tile decoder does; `a` is the output alignment (out & 15)."""
    return _decode(LZ4_FMT, stream, cap, a, schedule, stats)


def decode_snappy_stream(stream: bytes, a: int = 0, schedule: str = "front", stats: dict | None = None) -> bytes:
    """Decode a raw Snappy stream (varint length + elements) the way the tile decoder does."""
    total, shift, n = 0, 0, 0
    while True:
        b = stream[n]; n += 1
        total |= (b & 127) << shift
        shift += 7
        if b < 128:
            break
    return _decode(SNAPPY_FMT, stream[n:], total, a, schedule, stats)


def _decode(fmt: dict, stream: bytes, cap: int, a: int, schedule: str, stats: dict | None) -> bytes:
    src = np.frombuffer(stream, dtype=np.uint8)
    iend = len(src)
    padded = np.concatenate([src, np.zeros(CHUNK + MARGIN + 8, dtype=np.uint8)])
    out = bytearray()                                         # out[j] = output position a + j
    ip = 0
    groups = rounds = visits = 0
    while True:
        # ---------------- one group ----------------
        chunk = ip >> 12
        cbase = chunk << 12
        bp = padded[cbase:cbase + CHUNK + MARGIN + 8]
        slow = True
        if ip < iend:
            n1 = fmt["links"](bp, cbase, iend)
            n2 = _double(n1); n4 = _double(n2); n8 = _double(n4); n16 = _double(n8); n32 = _double(n16); n64 = _double(n32)
            # chase: 64-hop anchors, then binary descent
            p = ip - cbase
            anchors = []
            while p < CHUNK and len(anchors) < fmt["max_seq"] // 64 - 1 and n64[p] != NONE:
                anchors.append(p); p = int(n64[p])
            anchors.append(p)
            rem = 0
            for tbl, w in ((n32, 32), (n16, 16), (n8, 8), (n4, 4), (n2, 2), (n1, 1)):
                if p < CHUNK and tbl[p] != NONE:
                    rem += w; p = int(tbl[p])
            end_special = p < CHUNK and n1[p] == NONE
            end_ip = cbase + p
            nseq = (len(anchors) - 1) * 64 + rem
            if nseq:
                # sequence starts by binary-decomposed hops
                k = np.arange(nseq)
                sp = np.array(anchors, dtype=np.int64)[k >> 6]
                for tbl, w in ((n32, 32), (n16, 16), (n8, 8), (n4, 4), (n2, 2), (n1, 1)):
                    sp = np.where(k & w, tbl[sp], sp)
                ll, ml, off, lit = fmt["fields"](bp, sp)
                ln = ll + ml
                op0 = a + len(out)
                base = op0 & ~15
                a0 = op0 - base
                dl = op0 + np.concatenate([[0], np.cumsum(ln)[:-1]])
                dm = dl + ll
                end = dm + ml
                slack = fmt["end_slack"]
                lim_o = min(a + cap - slack, base + SPAN) if a + cap >= op0 + slack else 0
                bad = ((ml != 0) & ((off == 0) | (off > dm - a))) | (end > lim_o)
                nexec = int(np.argmax(bad)) if bad.any() else nseq
                if nexec:
                    rel = (dl - base)[:nexec]
                    gend = int((end - base)[nexec - 1])
                    nrows = (gend + 31) >> 5
                    # start bitmap + 1 + covering sequence per row
                    startbits = np.zeros(nrows + 1, dtype=np.uint64)
                    np.bitwise_or.at(startbits, rel >> 5, np.uint64(1) << (rel & 31).astype(np.uint64))
                    row2seq = np.zeros(nrows + 1, dtype=np.int64)
                    erel = (end - base)[:nexec]
                    for kk in range(nexec):
                        for R in range((int(rel[kk]) + 31) >> 5, nrows):
                            if (R << 5) >= erel[kk]:
                                break
                            row2seq[R] = kk + 1
                    # per-byte source step
                    P = np.zeros(SPAN, dtype=np.int64)
                    x = np.arange(nrows * 32)
                    lane = x & 31
                    le_mask = ((np.uint64(2) << lane.astype(np.uint64)) - np.uint64(2)) & np.uint64(0xFFFFFFFF)
                    bits = startbits[x >> 5] & le_mask
                    pop = np.array([bin(int(b)).count("1") for b in bits])
                    kx = row2seq[x >> 5] + pop - 1
                    live = (x >= a0) & (x < gend)
                    kx = np.where(live, kx, 0)
                    r = x - rel[kx]
                    is_lit = live & (r >= 0) & (r < ll[kx])
                    pa = base + x - off[kx]
                    in_group = live & ~is_lit & (pa >= op0)
                    pre = live & ~is_lit & ~in_group
                    vals = np.zeros(len(x), dtype=np.int64)
                    vals[is_lit] = bp[(lit[kx] + r)[is_lit]]
                    if pre.any():
                        hist = np.frombuffer(bytes(out), dtype=np.uint8)
                        vals[pre] = hist[(pa - a)[pre]]               # ring or, beyond 32 KiB back, L2: same bytes
                    if schedule == "none":                            # nobody has written anything yet
                        e = np.where(in_group, PTR | (pa - base), KNOWN | vals)
                        P[x[live]] = e[live]
                    else:                                             # every earlier WAVE of rows has been written
                        # "front": one row at a time; "<w>": waves of w adjacent rows run side by side (the kernel: 16, one
                        # row per warp); "block<w>": the span is cut into w contiguous blocks, one per warp, and the
                        # warps walk their blocks in step
                        if schedule.startswith("block"):
                            nb = int(schedule[5:])
                            per = -(-nrows // nb)
                            steps = [[b * per + i for b in range(nb) if b * per + i < min(nrows, (b + 1) * per)] for i in range(per)]
                        else:
                            wave = 1 if schedule == "front" else int(schedule)
                            steps = [list(range(r0, min(nrows, r0 + wave))) for r0 in range(0, nrows, wave)]
                        for rows_now in steps:
                            if not rows_now:
                                continue
                            sel = np.concatenate([np.arange(r * 32, r * 32 + 32) for r in rows_now])
                            pp = np.where(in_group[sel], (pa - base)[sel], a0)
                            look = P[pp]
                            e = np.where(look == 0, PTR | pp, look)
                            e = np.where(in_group[sel], e, KNOWN | vals[sel])
                            P[x[sel][live[sel]]] = e[live[sel]]
                    # pointer jumping, two jumps per round
                    idx = x[live]
                    while True:
                        cur = P[idx]
                        un = cur < KNOWN
                        if not un.any():
                            break
                        rounds += 1
                        t = idx[un]
                        visits += len(t)
                        qv = P[P[t] - PTR]
                        assert (qv != 0).all()
                        q2 = np.where(qv < KNOWN, P[np.where(qv < KNOWN, qv - PTR, 0)], qv)
                        P[t] = q2
                    out += (P[a0:gend] & 255).astype(np.uint8).tobytes()
                    groups += 1
                    if nexec == nseq:
                        ip = end_ip
                        slow = bool(end_special)
                    else:
                        ip = cbase + int(sp[nexec])
                        slow = True
                else:
                    slow = True
            else:
                ip = end_ip
        if slow:
            ip, done = fmt["slow"](padded, ip, out, cap, iend)
            if done:
                break
    if stats is not None:
        stats.update(groups=groups, rounds=rounds, byte_rounds=visits / max(1, len(out)))
    return bytes(out)
