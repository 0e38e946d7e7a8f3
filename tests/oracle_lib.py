"""ctypes bindings for the two CHECKERS (test infrastructure only):

  Oracle  oracle/liboracle.so          -- our CPU restatement (oracle/llc_oracle.c)
  Ref     oracle/_ref/libaocl_ref.so   -- the unmodified reference compiled by oracle/Makefile

Nothing in the product package imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libaocl_ref.so")

LZ4, LZ4HC, SNAPPY = 0, 1, 4


def _u8p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    return np.frombuffer(bytes(data), dtype=np.uint8).copy()


def build_oracle() -> None:
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])


def build_ref() -> bool:
    """Compile the reference from /root/reference when it is present; else rely on a prebuilt file."""
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref", "-j8"])
    return os.path.exists(REF_SO)


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = C.CDLL(ORACLE_SO)
        i64, p8, ci = C.c_int64, C.POINTER(C.c_uint8), C.c_int
        L.orc_partition_count.restype = ci
        L.orc_partition_count.argtypes = [i64, ci, ci, ci]
        for name, args in {
            "orc_lz4_bound": [i64], "orc_snappy_bound": [i64],
            "orc_lz4_encode_partition": [p8, i64, p8, i64, ci, C.POINTER(i64)],
            "orc_lz4_compress": [p8, i64, p8, i64, ci],
            "orc_lz4_decode_partition": [p8, i64, p8, i64, ci],
            "orc_lz4_decompress": [p8, i64, p8, i64],
            "orc_snappy_encode_fragment": [p8, i64, p8],
            "orc_snappy_compress": [p8, i64, p8, ci],
            "orc_snappy_decode_body": [p8, i64, p8, i64],
            "orc_snappy_decompress": [p8, i64, p8, i64],
            "orc_snappy_uncompressed_length": [p8, i64],
        }.items():
            f = getattr(L, name)
            f.restype, f.argtypes = i64, args
        self.L = L

    def partition_count(self, n, codec, max_threads=0):
        win = 65567 if codec == LZ4 else 65536
        return self.L.orc_partition_count(n, win, 4, max_threads)

    def bound(self, n, codec):
        return self.L.orc_lz4_bound(n) if codec == LZ4 else self.L.orc_snappy_bound(n)

    def out_capacity(self, n, codec):
        T = self.partition_count(n, codec)
        return int(self.bound(n, codec) + 16 + 12 * T + 64)

    def compress(self, data, codec, max_threads=0, cap=None) -> bytes | None:
        src = as_u8(data)
        n = len(src)
        full = self.out_capacity(n, codec)
        dst = np.zeros(full, dtype=np.uint8)
        if codec == LZ4:
            r = self.L.orc_lz4_compress(_u8p(src), n, _u8p(dst), full if cap is None else cap, max_threads)
            if r <= 0:
                return None
        else:
            r = self.L.orc_snappy_compress(_u8p(src), n, _u8p(dst), max_threads)
        return dst[:r].tobytes()

    def decompress(self, data, codec, cap) -> bytes | None:
        src = as_u8(data)
        dst = np.zeros(max(cap, 1), dtype=np.uint8)
        f = self.L.orc_lz4_decompress if codec == LZ4 else self.L.orc_snappy_decompress
        r = f(_u8p(src), len(src), _u8p(dst), cap)
        return None if r < 0 else dst[:r].tobytes()

    def snappy_uncompressed_length(self, data) -> int:
        src = as_u8(data)
        return self.L.orc_snappy_uncompressed_length(_u8p(src), len(src))


class AoclDesc(C.Structure):
    """api/aocl_compression.h:125-152 (128 bytes on LP64)."""
    _fields_ = [
        ("inBuf", C.c_void_p), ("outBuf", C.c_void_p), ("workBuf", C.c_void_p),
        ("inSize", C.c_size_t), ("outSize", C.c_size_t), ("level", C.c_size_t), ("optVar", C.c_size_t),
        ("numThreads", C.c_int), ("numMPIranks", C.c_int), ("memLimit", C.c_size_t),
        ("measureStats", C.c_int), ("cSize", C.c_uint64), ("dSize", C.c_uint64),
        ("cTime", C.c_uint64), ("dTime", C.c_uint64), ("cSpeed", C.c_float), ("dSpeed", C.c_float),
        ("optOff", C.c_int), ("optLevel", C.c_int),
    ]


assert C.sizeof(AoclDesc) == 128


class LlcLib:
    """Any library exporting the aocl_llc_* C ABI (the reference build or the GPU build)."""

    def __init__(self, path: str, omp: bool = False):
        self.L = C.CDLL(path, mode=C.RTLD_GLOBAL if omp else C.DEFAULT_MODE)
        L = self.L
        dp = C.POINTER(AoclDesc)
        L.aocl_llc_setup.restype, L.aocl_llc_setup.argtypes = C.c_int32, [dp, C.c_int]
        L.aocl_llc_compress.restype, L.aocl_llc_compress.argtypes = C.c_int64, [dp, C.c_int]
        L.aocl_llc_decompress.restype, L.aocl_llc_decompress.argtypes = C.c_int64, [dp, C.c_int]
        L.aocl_llc_destroy.restype, L.aocl_llc_destroy.argtypes = None, [dp, C.c_int]
        L.aocl_llc_version.restype = C.c_char_p
        L.aocl_get_rap_frame_bound_mt.restype = C.c_int32
        L.aocl_skip_rap_frame_mt.restype = C.c_int32
        L.aocl_skip_rap_frame_mt.argtypes = [C.c_void_p, C.c_int32]
        self._omp = None
        if omp:
            self._omp = C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL)
            self._omp.omp_set_num_threads.argtypes = [C.c_int]

    def set_threads(self, t: int):
        if self._omp is not None:
            self._omp.omp_set_num_threads(int(t))

    def new_desc(self, codec, opt_off=0, stats=0):
        d = AoclDesc()
        d.optOff, d.optLevel, d.measureStats = opt_off, -1, stats
        rc = self.L.aocl_llc_setup(C.byref(d), codec)
        assert rc == 0, rc
        return d

    def compress(self, data, codec, cap=None, desc=None) -> tuple[int, bytes]:
        src = as_u8(data)
        n = len(src)
        if cap is None:
            cap = n + n // 6 + 16384 + 16 + 12 * 8192
        dst = np.zeros(max(cap, 1), dtype=np.uint8)
        d = desc or self.new_desc(codec)
        d.inBuf, d.inSize = src.ctypes.data, n
        d.outBuf, d.outSize = dst.ctypes.data, cap
        r = self.L.aocl_llc_compress(C.byref(d), codec)
        return r, (dst[:r].tobytes() if r > 0 else b"")

    GUARD = 256

    def decompress(self, data, codec, cap, desc=None) -> tuple[int, bytes]:
        """The destination sits between two guard regions; a call that touches them fails the test."""
        src = as_u8(data)
        g = self.GUARD
        buf = np.full(g + max(cap, 1) + g, 0xA5, dtype=np.uint8)
        dst = buf[g:g + max(cap, 1)]
        dst[:] = 0
        d = desc or self.new_desc(codec)
        d.inBuf, d.inSize = src.ctypes.data, len(src)
        d.outBuf, d.outSize = dst.ctypes.data, cap
        r = self.L.aocl_llc_decompress(C.byref(d), codec)
        assert bool((buf[:g] == 0xA5).all()) and bool((buf[g + max(cap, 1):] == 0xA5).all()), "wrote outside the destination"
        return r, (dst[:r].tobytes() if r > 0 else b"")


_ref = None


def ref_lib() -> LlcLib | None:
    global _ref
    if _ref is None and build_ref():
        _ref = LlcLib(REF_SO, omp=True)
    return _ref
