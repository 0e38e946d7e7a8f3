"""N > 1 host logic on CPU (gloo, world size 2): partition-range sharding of one RAP frame, the
all-gather of RAP entries, and reassembly.  The per-partition decode is done by the oracle here
(the CUDA kernels are covered by the gpu tests); what is under test is the sharding arithmetic and
the collective plumbing that bench.py uses with NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, codec, tmpdir):
    sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import kat
    import oracle_lib as ol
    from llc_b200 import shard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = ol.Oracle()
    data = kat.make_input("text", 1500001)
    stream = np.frombuffer(orc.compress(data, codec), dtype=np.uint8)
    frame, entries = shard.parse_frame(stream)
    T = len(entries)
    lo, hi = shard.partition_range(T, rank, world)
    origins = shard.output_origins(entries)
    # this rank decodes only its partitions, into a buffer that starts at its own origin
    my_origin = int(origins[lo])
    my_bytes = int(sum(int(entries[i, 2]) for i in range(lo, hi) if entries[i, 1] > 0))
    out = np.zeros(my_bytes, dtype=np.uint8)
    for i in range(lo, hi):
        off, clen, dlen = (int(x) for x in entries[i])
        if clen == 0:
            continue
        src = np.ascontiguousarray(stream[off:off + clen])
        dst = np.zeros(dlen, dtype=np.uint8)
        if codec == kat.LZ4:
            got = orc.L.orc_lz4_decode_partition(src.ctypes.data_as(C.POINTER(C.c_uint8)), clen,
                                                 dst.ctypes.data_as(C.POINTER(C.c_uint8)), dlen, int(i == T - 1))
        else:
            got = orc.L.orc_snappy_decode_body(src.ctypes.data_as(C.POINTER(C.c_uint8)), clen,
                                               dst.ctypes.data_as(C.POINTER(C.c_uint8)), dlen)
        assert got == dlen
        o = int(origins[i]) - my_origin
        out[o:o + dlen] = dst
    assert out.tobytes() == data[my_origin:my_origin + my_bytes].tobytes()
    # the only exchange: all-gather of the RAP entries each rank holds (padded to equal length)
    k = (T + world - 1) // world
    mine = np.zeros((k, 3), dtype=np.int32)
    mine[: hi - lo] = entries[lo:hi].astype(np.int32)
    table = shard.all_gather_entries(torch.from_numpy(mine), dist, world).numpy()
    rebuilt = np.concatenate([table[r * k: r * k + (shard.partition_range(T, r, world)[1] - shard.partition_range(T, r, world)[0])]
                              for r in range(world)])
    assert np.array_equal(rebuilt.astype(np.uint32), entries)
    sizes = torch.tensor([my_bytes], dtype=torch.int64)
    dist.all_reduce(sizes)
    assert int(sizes.item()) == len(data)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("codec", [0, 4])
def test_partition_sharding_world2(tmp_path, codec):
    port = 29500 + (os.getpid() % 2000) + codec
    mp.spawn(_worker, args=(2, port, codec, str(tmp_path)), nprocs=2, join=True)


def test_partition_range_covers_everything():
    sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python"))
    from llc_b200 import shard
    for T in (1, 2, 7, 4094, 4096):
        for W in (1, 2, 4, 8):
            got = [shard.partition_range(T, r, W) for r in range(W)]
            assert got[0][0] == 0 and got[-1][1] == T
            assert all(got[i][1] == got[i + 1][0] for i in range(W - 1))
