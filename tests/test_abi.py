"""The drop-in boundary: the product library loads without a GPU, exports every symbol the headers
declare, keeps the reference's descriptor layout, and refuses to work without a CUDA device
instead of falling back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import llc_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aocl_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(llc_b200.LIB_PATH):
        llc_b200.build()
    return llc_b200.load()


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions("aocl_llc.h") + declared_functions("aocl_llc_gpu.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert set(llc_b200.HOST_API) <= set(names) and set(llc_b200.GPU_API) <= set(names)
    # the codecs' native entry points (include/aocl_llc_native.h)
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "aocl_llc_native.h")).read(), flags=re.S)
    native = sorted(set(re.findall(r"\b((?:LZ4|snappy)_[A-Za-z0-9_]+)\s*\(", text)))
    assert native == sorted(llc_b200.NATIVE_API)
    for n in native:
        assert hasattr(lib, n), n


def test_native_bounds_without_gpu(lib):
    assert lib.LZ4_compressBound(65025) == 65296               # lz4_gtest.cpp:323-326
    assert lib.LZ4_compressBound(0x7E000001) == 0
    for n, want in ((0, 32), (1, 33), (6, 39), (100, 148), (65536, 76490)):   # snappy.cc:160-182
        assert lib.snappy_max_compressed_length(n) == want


def test_descriptor_layout_matches_reference():
    # api/aocl_compression.h:125-152 on LP64
    d = llc_b200.AoclDesc
    assert C.sizeof(d) == 128
    want = {"inBuf": 0, "outBuf": 8, "workBuf": 16, "inSize": 24, "outSize": 32, "level": 40, "optVar": 48,
            "numThreads": 56, "numMPIranks": 60, "memLimit": 64, "measureStats": 72, "cSize": 80, "dSize": 88,
            "cTime": 96, "dTime": 104, "cSpeed": 112, "dSpeed": 116, "optOff": 120, "optLevel": 124}
    for k, off in want.items():
        assert getattr(d, k).offset == off, k


def test_setup_return_codes(lib):
    d = llc_b200.AoclDesc()
    assert lib.aocl_llc_setup(C.byref(d), -1) == -4            # ERR_UNSUPPORTED_METHOD, api/api.cpp:133-138
    assert lib.aocl_llc_setup(C.byref(d), 7) == -4
    for excluded in (2, 3, 5, 6):                               # LZMA, BZIP2, ZLIB, ZSTD: api/api.cpp:156-162
        assert lib.aocl_llc_setup(C.byref(d), excluded) == -3
    # LZ4HC is served for decompress (api/codec.h:168): setup behaves like LZ4's (here: no device -> failure, on a
    # GPU box 0, tests/test_gpu_interop.py), compress is refused as an excluded method
    assert lib.aocl_llc_setup(C.byref(d), 1) in (0, -2)
    assert lib.aocl_llc_compress(C.byref(d), 1) == -3
    assert d.workBuf is None


def test_frame_helpers_without_gpu(lib):
    assert lib.aocl_get_rap_frame_bound_mt() == 16 + 12 * 8192
    assert lib.aocl_skip_rap_frame_mt(None, 100) == -5          # ERR_INVALID_INPUT, threads/threads.c:322-323
    buf = (C.c_char * 64)()
    assert lib.aocl_skip_rap_frame_mt(C.cast(buf, C.c_void_p), 64) == 0
    hdr = b"AOCL_LLC" + (16 + 12 * 5).to_bytes(4, "little") + (5).to_bytes(4, "little")
    buf[: len(hdr)] = hdr
    assert lib.aocl_skip_rap_frame_mt(C.cast(buf, C.c_void_p), 64) == 76
    assert lib.aocl_skip_rap_frame_mt(C.cast(buf, C.c_void_p), 7) == 0
    assert b"4.2.0" in lib.aocl_llc_version()


def test_partition_arithmetic_matches_oracle(lib, oracle):
    for n in [0, 1, 262143, 262144, 262267, 262268, 393215, 393216, 393401, 393402, 1 << 20, 1500001, 64 << 20, 1 << 30,
              (1 << 31) - 1]:
        for codec in (0, 4):
            assert lib.aocl_gpu_partition_count(codec, n) == oracle.partition_count(n, codec), (codec, n)
            assert lib.aocl_gpu_compress_bound(codec, n) >= oracle.bound(n, codec)


def test_no_cpu_fallback_without_device(lib):
    """On a box without CUDA every data call must fail (ERR_COMPRESSION_FAILED), never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    src = np.frombuffer(b"hello hello hello hello hello hello", dtype=np.uint8).copy()
    dst = np.zeros(256, dtype=np.uint8)
    d = llc_b200.AoclDesc()
    d.inBuf, d.inSize, d.outBuf, d.outSize = src.ctypes.data, len(src), dst.ctypes.data, len(dst)
    assert lib.aocl_llc_compress(C.byref(d), 0) == -2
    assert lib.aocl_llc_decompress(C.byref(d), 4) == -2
    assert not dst.any()
    h = C.c_void_p()
    assert lib.aocl_gpu_ctx_create(C.byref(h), -1, None) != 0
