"""Python-side plumbing for the B200 LZ4/Snappy RAP library.

The product is the C-ABI shared library ``aocl-compression_b200/lib/libaocl_compression.so``
(sources in ``aocl-compression_b200/csrc``).  This package only *binds* it with ctypes so that
tests and bench.py can drive it; torch is used by callers for device memory and
torch.distributed, never on the data path.  There is no CPU fallback: if the library is
missing or no CUDA device is usable, loading / calling fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # aocl-compression_b200/
REPO_ROOT = os.path.dirname(PKG_DIR)
def shard_unique_id() -> bytes:
    """128 bytes (an ncclUniqueId) that rank 0 hands to the other ranks before shard_init()."""
    buf = C.create_string_buffer(128)
    rc = load().aocl_gpu_shard_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"aocl_gpu_shard_unique_id failed ({rc}): libnccl.so.2 not loadable")
    return buf.raw


def shard_range(codec: int, n: int, rank: int, nranks: int):
    """-> (first partition, partition count, byte offset, byte length) of a rank's share, or None if n has fewer
    partitions than ranks."""
    f, c, o, ln = C.c_uint32(0), C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
    rc = load().aocl_gpu_shard_range(codec, n, rank, nranks, C.byref(f), C.byref(c), C.byref(o), C.byref(ln))
    return None if rc != 0 else (f.value, c.value, o.value, ln.value)


# AOCL_LLC_LIB points experiments at an A/B build of the same sources (make ... LIB=lib_x); the product is lib/
LIB_PATH = os.environ.get("AOCL_LLC_LIB") or os.path.join(PKG_DIR, "lib", "libaocl_compression.so")

LZ4, SNAPPY = 0, 4
CODEC_NAMES = {LZ4: "lz4", SNAPPY: "snappy"}


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into the in-tree shared library."""
    subprocess.check_call(["make", "-C", PKG_DIR, "-j4"] + ([] if verbose else ["-s"]))
    return LIB_PATH


class AoclDesc(C.Structure):
    """aocl_compression_desc (include/aocl_llc.h; reference api/aocl_compression.h:125-152)."""
    _fields_ = [
        ("inBuf", C.c_void_p), ("outBuf", C.c_void_p), ("workBuf", C.c_void_p),
        ("inSize", C.c_size_t), ("outSize", C.c_size_t), ("level", C.c_size_t), ("optVar", C.c_size_t),
        ("numThreads", C.c_int), ("numMPIranks", C.c_int), ("memLimit", C.c_size_t),
        ("measureStats", C.c_int), ("cSize", C.c_uint64), ("dSize", C.c_uint64),
        ("cTime", C.c_uint64), ("dTime", C.c_uint64), ("cSpeed", C.c_float), ("dSpeed", C.c_float),
        ("optOff", C.c_int), ("optLevel", C.c_int),
    ]


HOST_API = ["aocl_llc_compress", "aocl_llc_decompress", "aocl_llc_setup", "aocl_llc_destroy", "aocl_llc_version",
            "aocl_get_rap_frame_bound_mt", "aocl_skip_rap_frame_mt"]
NATIVE_API = ["LZ4_compressBound", "LZ4_compress_default", "LZ4_decompress_safe", "snappy_compress", "snappy_uncompress",
              "snappy_max_compressed_length", "snappy_uncompressed_length"]
GPU_API = ["aocl_gpu_ctx_create", "aocl_gpu_ctx_destroy", "aocl_gpu_ctx_stream", "aocl_gpu_partition_count",
           "aocl_gpu_compress_bound", "aocl_gpu_compress_async", "aocl_gpu_decompress_async", "aocl_gpu_finish",
           "aocl_gpu_compress", "aocl_gpu_decompress", "aocl_gpu_set_lz4_frameless",
           "aocl_gpu_decompress_range_async", "aocl_gpu_decompress_batch_async", "aocl_gpu_compress_batch_async",
           "aocl_gpu_launch_count", "aocl_gpu_set_profiling", "aocl_gpu_profile_count", "aocl_gpu_profile_get",
           "aocl_gpu_debug_counters", "aocl_gpu_set_input_watermark", "aocl_gpu_decompress_open_async",
           "aocl_gpu_decompress_slab_async", "aocl_gpu_decompress_close_async",
           "aocl_gpu_shard_unique_id", "aocl_gpu_shard_init", "aocl_gpu_shard_destroy", "aocl_gpu_shard_range",
           "aocl_gpu_compress_sharded", "aocl_gpu_decompress_sharded", "aocl_gpu_set_mode", "aocl_gpu_sharded_host_calls",
           "aocl_gpu_set_partitions", "aocl_gpu_ctx_partition_count"]

_lib = None


def load() -> C.CDLL:
    """Load the product library and declare every exported signature.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    dp, vp, sz, i32, i64, u32, u64 = C.POINTER(AoclDesc), C.c_void_p, C.c_size_t, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
    sig = {
        "aocl_llc_compress": (i64, [dp, C.c_int]), "aocl_llc_decompress": (i64, [dp, C.c_int]),
        "aocl_llc_setup": (i32, [dp, C.c_int]), "aocl_llc_destroy": (None, [dp, C.c_int]),
        "aocl_llc_version": (C.c_char_p, []), "aocl_get_rap_frame_bound_mt": (i32, []),
        "aocl_skip_rap_frame_mt": (i32, [vp, i32]),
        "aocl_gpu_ctx_create": (i32, [C.POINTER(vp), C.c_int, vp]), "aocl_gpu_ctx_destroy": (None, [vp]),
        "aocl_gpu_ctx_stream": (vp, [vp]), "aocl_gpu_partition_count": (i32, [i32, sz]),
        "aocl_gpu_compress_bound": (sz, [i32, sz]),
        "aocl_gpu_compress_async": (i32, [vp, i32, vp, sz, vp, sz]),
        "aocl_gpu_decompress_async": (i32, [vp, i32, vp, sz, vp, sz]),
        "aocl_gpu_finish": (i64, [vp]),
        "aocl_gpu_compress": (i64, [vp, i32, vp, sz, vp, sz]),
        "aocl_gpu_decompress": (i64, [vp, i32, vp, sz, vp, sz]),
        "aocl_gpu_set_lz4_frameless": (None, [vp, i32]),
        "aocl_gpu_decompress_range_async": (i32, [vp, i32, vp, sz, vp, sz, u32, u32, u64]),
        "aocl_gpu_decompress_batch_async": (i32, [vp, i32, vp, vp, vp, vp, vp, sz]),
        "aocl_gpu_compress_batch_async": (i32, [vp, i32, vp, vp, vp, vp, vp, sz]),
        "aocl_gpu_launch_count": (u64, []),
        "aocl_gpu_set_profiling": (None, [vp, i32]), "aocl_gpu_profile_count": (i32, [vp]),
        "aocl_gpu_profile_get": (C.c_float, [vp, i32, C.c_char_p, i32]),
        "aocl_gpu_debug_counters": (i32, [vp, i32]),
        "LZ4_compressBound": (C.c_int, [C.c_int]),
        "LZ4_compress_default": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "LZ4_decompress_safe": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "snappy_compress": (C.c_int, [vp, sz, vp, C.POINTER(sz)]),
        "snappy_uncompress": (C.c_int, [vp, sz, vp, C.POINTER(sz)]),
        "snappy_max_compressed_length": (sz, [sz]),
        "snappy_uncompressed_length": (C.c_int, [vp, sz, C.POINTER(sz)]),
        "aocl_gpu_set_input_watermark": (None, [vp, vp]),
        "aocl_gpu_decompress_open_async": (i32, [vp, i32, vp, sz, sz]),
        "aocl_gpu_decompress_slab_async": (i32, [vp, i32, vp, vp, u32, u32]),
        "aocl_gpu_decompress_close_async": (i32, [vp]),
        "aocl_gpu_set_mode": (i32, [vp, C.c_char_p]),
        "aocl_gpu_set_partitions": (i32, [vp, i32]), "aocl_gpu_ctx_partition_count": (i32, [vp, i32, sz]),
        "aocl_gpu_sharded_host_calls": (u64, []),
        "aocl_gpu_shard_unique_id": (i32, [vp]),
        "aocl_gpu_shard_init": (i32, [vp, vp, i32, i32]),
        "aocl_gpu_shard_destroy": (None, [vp]),
        "aocl_gpu_shard_range": (i32, [i32, sz, i32, i32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u64), C.POINTER(u64)]),
        "aocl_gpu_compress_sharded": (i64, [vp, i32, vp, sz, vp, sz, C.POINTER(u64), C.POINTER(u64)]),
        "aocl_gpu_decompress_sharded": (i64, [vp, i32, vp, sz, vp, sz, C.POINTER(u64), C.POINTER(u64)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)     # AttributeError here == a declared symbol is not exported
        f.restype, f.argtypes = res, args
    _lib = L
    return L


class GpuContext:
    """Thin RAII wrapper around aocl_gpu_ctx_t for torch callers (device pointers in, sizes out)."""

    def __init__(self, device: int = -1, stream: int | None = None):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.aocl_gpu_ctx_create(C.byref(h), device, C.c_void_p(stream) if stream else None)
        if rc != 0:
            raise RuntimeError(f"aocl_gpu_ctx_create failed ({rc}): a CUDA device is required, there is no CPU fallback")
        self.h = h

    def close(self):
        if self.h:
            self.L.aocl_gpu_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # tensors are torch.uint8 CUDA tensors; only data_ptr()/numel() are used
    def compress_async(self, codec, src, dst):
        return self.L.aocl_gpu_compress_async(self.h, codec, src.data_ptr(), src.numel(), dst.data_ptr(), dst.numel())

    def decompress_async(self, codec, src, n, dst):
        return self.L.aocl_gpu_decompress_async(self.h, codec, src.data_ptr(), n, dst.data_ptr(), dst.numel())

    def decompress_range_async(self, codec, src, n, dst, first, count, origin):
        return self.L.aocl_gpu_decompress_range_async(self.h, codec, src.data_ptr(), n, dst.data_ptr(), dst.numel(),
                                                       first, count, origin)

    def finish(self) -> int:
        return self.L.aocl_gpu_finish(self.h)

    def compress(self, codec, src, dst) -> int:
        self.compress_async(codec, src, dst)
        return self.finish()

    def decompress(self, codec, src, n, dst) -> int:
        self.decompress_async(codec, src, n, dst)
        return self.finish()

    def decompress_batch_async(self, codec, in_ptrs, in_sizes, out_ptrs, out_caps, status, count):
        return self.L.aocl_gpu_decompress_batch_async(self.h, codec, in_ptrs.data_ptr(), in_sizes.data_ptr(),
                                                       out_ptrs.data_ptr(), out_caps.data_ptr(), status.data_ptr(), count)

    def compress_batch_async(self, codec, in_ptrs, in_sizes, out_ptrs, out_caps, status, count):
        return self.L.aocl_gpu_compress_batch_async(self.h, codec, in_ptrs.data_ptr(), in_sizes.data_ptr(),
                                                     out_ptrs.data_ptr(), out_caps.data_ptr(), status.data_ptr(), count)

    def set_lz4_frameless(self, on: bool):
        self.L.aocl_gpu_set_lz4_frameless(self.h, 1 if on else 0)

    def set_mode(self, mode: str) -> int:
        """"exact" (default) or "fastparse" (the named non-exact LZ4 RAP encoder)."""
        return self.L.aocl_gpu_set_mode(self.h, mode.encode())

    def set_partitions(self, max_threads: int) -> int:
        """Write the frames of a host with `max_threads` OpenMP threads (0: the saturated layout, default)."""
        return self.L.aocl_gpu_set_partitions(self.h, max_threads)

    def partition_count(self, codec, n) -> int:
        return self.L.aocl_gpu_ctx_partition_count(self.h, codec, n)

    def set_profiling(self, on: bool):
        self.L.aocl_gpu_set_profiling(self.h, 1 if on else 0)

    def profile(self) -> list[tuple[str, float]]:
        """(kernel name, ms) of every kernel launched by the most recent enqueue (after finish())."""
        out = []
        buf = C.create_string_buffer(64)
        for i in range(self.L.aocl_gpu_profile_count(self.h)):
            ms = self.L.aocl_gpu_profile_get(self.h, i, buf, 64)
            out.append((buf.value.decode(), float(ms)))
        return out

    # ---- one frame over several GPUs (include/aocl_llc_gpu.h, "ONE frame over several GPUs")
    def shard_init(self, id_bytes: bytes, rank: int, nranks: int) -> int:
        buf = C.create_string_buffer(bytes(id_bytes), 128)
        return self.L.aocl_gpu_shard_init(self.h, C.cast(buf, C.c_void_p), rank, nranks)

    def compress_sharded(self, codec, src_slice, n_total, dst_slice):
        """-> (stream length or < 0, offset of this rank's bytes in the stream, their length)"""
        off, ln = C.c_uint64(0), C.c_uint64(0)
        r = self.L.aocl_gpu_compress_sharded(self.h, codec, src_slice.data_ptr(), n_total, dst_slice.data_ptr(), dst_slice.numel(),
                                             C.byref(off), C.byref(ln))
        return r, off.value, ln.value

    def decompress_sharded(self, codec, stream_base_ptr, n, dst_slice):
        """stream_base_ptr: device address that the stream's offset 0 maps to.  -> (total or < 0, output offset, length)"""
        off, ln = C.c_uint64(0), C.c_uint64(0)
        r = self.L.aocl_gpu_decompress_sharded(self.h, codec, stream_base_ptr, n, dst_slice.data_ptr(), dst_slice.numel(),
                                               C.byref(off), C.byref(ln))
        return r, off.value, ln.value

    @property
    def stream(self) -> int:
        return self.L.aocl_gpu_ctx_stream(self.h)
