"""Break down the end-to-end call: PCIe copy bandwidth vs aocl_llc_* wall time (measureStats)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python"))
sys.path.insert(0, ROOT)
import torch, llc_b200
from bench import make_data
L = llc_b200.load()
U = 1 << 30
data = make_data("text_like", U, 2024)
h_in = torch.from_numpy(data).pin_memory()
d = torch.empty(U, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(h_in, non_blocking=True); torch.cuda.synchronize(); h2d = time.perf_counter() - t
    t = time.perf_counter(); h_in.copy_(d, non_blocking=True); torch.cuda.synchronize(); d2h = time.perf_counter() - t
print(f"pinned H2D {U/h2d/1e9:.1f} GB/s ({h2d*1e3:.1f} ms)  D2H {U/d2h/1e9:.1f} GB/s ({d2h*1e3:.1f} ms)")
cap = L.aocl_gpu_compress_bound(0, U)
h_comp = torch.empty(cap, dtype=torch.uint8).pin_memory()
h_back = torch.empty(U, dtype=torch.uint8).pin_memory()
desc = llc_b200.AoclDesc(); desc.measureStats = 1
assert L.aocl_llc_setup(C.byref(desc), 0) == 0
for it in range(3):
    desc.inBuf, desc.inSize, desc.outBuf, desc.outSize = h_in.data_ptr(), U, h_comp.data_ptr(), cap
    t = time.perf_counter(); c = L.aocl_llc_compress(C.byref(desc), 0); tc = time.perf_counter() - t
    desc.inBuf, desc.inSize, desc.outBuf, desc.outSize = h_comp.data_ptr(), c, h_back.data_ptr(), U
    t = time.perf_counter(); r = L.aocl_llc_decompress(C.byref(desc), 0); td = time.perf_counter() - t
    print(f"iter {it}: compress {tc*1e3:.1f} ms (cTime {desc.cTime/1e6:.1f}) -> {c} B; decompress {td*1e3:.1f} ms (dTime {desc.dTime/1e6:.1f}) -> {r} B")
