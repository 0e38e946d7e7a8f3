"""Tile-decoder debugging: replay the golden 'mixed' LZ4 sequence through the host API and the device API
with dirty output buffers; print return codes and watchdog counters.  Run under gpurun."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, llc_b200
import kat, oracle_lib as ol

L = llc_b200.load()
gpu = ol.LlcLib(llc_b200.LIB_PATH)
orc = ol.Oracle()
cnt = (C.c_uint64 * 32)()
def counters():
    L.aocl_gpu_debug_counters(C.cast(cnt, C.c_void_p), 1)
    return list(cnt)[24:32]

names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["mixed"]
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1048576, 1500001, 2097229]
ctx = llc_b200.GpuContext(0)
for name in names:
    for codec in (kat.LZ4, kat.SNAPPY):
        for rep in range(2):
            for n in sizes:
                data = kat.make_input(name, n)
                comp = orc.compress(data, codec)
                counters()
                r, back = gpu.decompress(comp, codec, max(n, 1))
                okh = (r == n and back == data.tobytes())
                wh = counters()
                # device API, dirty output
                d_comp = torch.from_numpy(np.frombuffer(comp, dtype=np.uint8).copy()).cuda()
                d_back = torch.full((max(n, 1),), 0xA5, dtype=torch.uint8, device="cuda")
                r2 = ctx.decompress(codec, d_comp, len(comp), d_back)
                okd = (r2 == n and d_back[:n].cpu().numpy().tobytes() == data.tobytes())
                wd = counters()
                # device API, misaligned input and output (offset 5 / 3)
                d_comp2 = torch.empty(len(comp) + 64, dtype=torch.uint8, device="cuda"); d_comp2[5:5 + len(comp)] = d_comp
                d_back2 = torch.full((max(n, 1) + 64,), 0x5A, dtype=torch.uint8, device="cuda")
                r3 = ctx.decompress(codec, d_comp2[5:], len(comp), d_back2[3:3 + max(n, 1)])
                okm = (r3 == n and d_back2[3:3 + n].cpu().numpy().tobytes() == data.tobytes())
                wm = counters()
                print(f"{name} codec {codec} rep {rep} n={n}: host r={r} ok={okh} wd={wh} | dev r={r2} ok={okd} wd={wd} | misaligned r={r3} ok={okm} wd={wm}", flush=True)
