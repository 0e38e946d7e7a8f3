"""Decoder bring-up probe: decode oracle-compressed inputs of several generators/sizes with the decoder
AOCL_GPU_DECODER selects (auto | rowq | tile | warp), compare with the input, print the tile phase counters.
Run under gpurun."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np, torch, llc_b200
import kat, oracle_lib as ol
from llc_b200 import gen

L = llc_b200.load()
ctx = llc_b200.GpuContext(0)
orc = ol.Oracle()
cnt = (C.c_uint64 * 32)()

def counters(reset=True):
    L.aocl_gpu_debug_counters(C.cast(cnt, C.c_void_p), 1 if reset else 0)
    return list(cnt)

def run(name, data, codec):
    comp = orc.compress(data, codec)
    d_comp = torch.from_numpy(np.frombuffer(comp, dtype=np.uint8).copy()).cuda()
    d_back = torch.zeros(max(len(data), 1), dtype=torch.uint8, device="cuda")
    counters()
    torch.cuda.synchronize(); t = time.perf_counter()
    r = ctx.decompress(codec, d_comp, len(comp), d_back)
    dt = time.perf_counter() - t
    ok = r == len(data) and d_back[:len(data)].cpu().numpy().tobytes() == data.tobytes()
    c = counters()
    print(f"{name:>10} codec {codec} n={len(data):>9} -> r={r:>10} ok={ok} {dt*1e3:8.2f} ms  watchdog={c[24:28]} cyc/group={sum(c[:10])//max(c[17],1)}", flush=True)
    if any(c[:16]):
        tot = sum(c[:10]) or 1
        names = ["wait", "links", "chase", "expand", "fields", "lits", "match", "flush", "slow", "fwd"]
        print("     phases % :", " ".join(f"{n}={100*v/tot:.1f}" for n, v in zip(names, c[:10])),
              f"| tables={c[16]} groups={c[17]} seqs={c[18]} exec={c[19]} rounds={c[20]} slow={c[21]}", flush=True)
    return ok

which = sys.argv[1] if len(sys.argv) > 1 else "all"
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1000, 70000, 300000, 1 << 20, (4 << 20) + 12345]
bad = 0
for name in (kat.GOLDEN_GENS if which == "all" else which.split(",")):
    for n in sizes:
        data = kat.make_input(name, n)
        for codec in (kat.LZ4, kat.SNAPPY):
            bad += not run(name, data, codec)
print("FAILURES:", bad)
