"""Host-side sharding of RAP partitions across ranks.

RAP partitions are self-contained, so rank r of W owns the contiguous range
[r*T//W, (r+1)*T//W) of a frame's partitions; its input and output slices are contiguous.  The only
exchange is an all-gather of the 12-byte RAP entries {offset, comp_len, decomp_len} so that every
rank knows the global layout (where its output slice starts, how big the whole thing is).
Works with any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import struct

import numpy as np

RAP_MAGIC = b"AOCL_LLC"


def partition_range(T: int, rank: int, world: int) -> tuple[int, int]:
    return rank * T // world, (rank + 1) * T // world


def parse_frame(stream: bytes | np.ndarray) -> tuple[int, np.ndarray]:
    """(frame_len, entries[T,3] uint32) of a RAP stream; (0, empty) for a frame-less stream."""
    b = bytes(stream[:16])
    if len(b) < 16 or b[:8] != RAP_MAGIC:
        return 0, np.zeros((0, 3), dtype=np.uint32)
    frame, T = struct.unpack_from("<II", b, 8)
    ent = np.frombuffer(bytes(stream[16:16 + 12 * T]), dtype="<u4").reshape(T, 3)
    return frame, ent


def output_origins(entries: np.ndarray) -> np.ndarray:
    """Exclusive scan of decomp_len over partitions that carry data (comp_len > 0)."""
    d = np.where(entries[:, 1] > 0, entries[:, 2], 0).astype(np.int64)
    return np.concatenate([[0], np.cumsum(d)[:-1]])


def all_gather_entries(local_entries, dist, world: int):
    """All-gather equally sized [k,3] int32 entry tables; returns the [world*k, 3] global table.
    `local_entries` is a torch tensor on the backend's device."""
    import torch
    flat = local_entries.contiguous().view(-1)
    out = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(out, flat)
    return out.view(-1, 3)
