import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, llc_b200, oracle_lib as ol
from llc_b200 import gen
P = int(sys.argv[1]); PS = 65536; codec = 4
ctx = llc_b200.GpuContext(0); orc = ol.Oracle()
pages = gen.pages(P); d_in = torch.from_numpy(pages.reshape(-1)).cuda()
bound = 76544
d_comp = torch.zeros(P * bound, dtype=torch.uint8, device="cuda")
in_ptrs = (torch.arange(P, dtype=torch.int64, device="cuda") * PS + d_in.data_ptr())
out_ptrs = (torch.arange(P, dtype=torch.int64, device="cuda") * bound + d_comp.data_ptr())
in_sizes = torch.full((P,), PS, dtype=torch.int32, device="cuda"); out_caps = torch.full((P,), bound, dtype=torch.int32, device="cuda")
status = torch.zeros(P, dtype=torch.int64, device="cuda")
ctx.compress_batch_async(codec, in_ptrs, in_sizes, out_ptrs, out_caps, status, P)
r = ctx.finish(); print("compress finish", r, flush=True)
csz = status.cpu().numpy(); print("sizes min/max", csz.min(), csz.max(), flush=True)
bad = 0
for i in range(P):
    want = orc.compress(pages[i], codec)
    if csz[i] != len(want) or d_comp[i*bound:i*bound+csz[i]].cpu().numpy().tobytes() != want:
        bad += 1
        if bad < 5: print("page", i, "kind", i % 3, "got", csz[i], "want", len(want), flush=True)
print("bad pages", bad)
R = int(sys.argv[2]) if len(sys.argv) > 2 else 2
N = P * R
d_out = torch.zeros(N * PS, dtype=torch.uint8, device="cuda")
big_in_ptrs = out_ptrs.repeat(R); big_in_sizes = torch.from_numpy(csz.astype(np.int32)).cuda().repeat(R)
big_out_ptrs = (torch.arange(N, dtype=torch.int64, device="cuda") * PS + d_out.data_ptr())
big_caps = torch.full((N,), PS, dtype=torch.int32, device="cuda"); big_status = torch.zeros(N, dtype=torch.int64, device="cuda")
ctx.decompress_batch_async(codec, big_in_ptrs, big_in_sizes, big_out_ptrs, big_caps, big_status, N)
r = ctx.finish(); print("decompress finish", r, flush=True)
st = big_status.cpu().numpy(); print("status != PS:", int((st != PS).sum()), st[st != PS][:8], np.nonzero(st != PS)[0][:8], flush=True)
for rr in range(R): print("replica", rr, bool(torch.equal(d_out[rr*P*PS:(rr+1)*P*PS], d_in)), flush=True)
