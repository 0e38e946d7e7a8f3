"""fastparse bring-up: compress on the GPU, walk partition 0's LZ4 sequences on the host and report the first sequence
whose bytes differ from the input."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, llc_b200
from llc_b200 import gen
ctx = llc_b200.GpuContext(0)
data = gen.text_like(2 * 262272 + 1000, seed=71)
n = len(data)
d_in = torch.from_numpy(data).cuda()
d_comp = torch.zeros(ctx.L.aocl_gpu_compress_bound(0, n) + (1 << 20), dtype=torch.uint8, device="cuda")
ctx.set_mode("fastparse")
r = ctx.compress(0, d_in, d_comp)
print("compress ->", r)
if r <= 0:
    sys.exit(0)
s = d_comp[:r].cpu().numpy()
T = int.from_bytes(bytes(s[12:16]), "little")
ent = np.frombuffer(bytes(s[16:16 + 12 * T]), dtype="<u4").reshape(T, 3)
print("entries", ent[:4])
src = data.tobytes()
ip = int(ent[0, 0]); end = ip + int(ent[0, 1]) ; out = bytearray(); k = 0
b = bytes(s)
while ip < end:
    tok = b[ip]; ip0 = ip; ip += 1
    ll = tok >> 4
    if ll == 15:
        while True:
            x = b[ip]; ip += 1; ll += x
            if x != 255: break
    lit = b[ip:ip + ll]; ip += ll
    pos = len(out)
    if lit != src[pos:pos + ll]:
        print(f"seq {k} at stream {ip0}: literals differ at out {pos}: ll={ll} got {lit[:24]!r} want {src[pos:pos+24]!r}"); break
    out += lit
    if ip >= end: break
    off = b[ip] | (b[ip + 1] << 8); ip += 2
    ml = tok & 15
    if ml == 15:
        while True:
            x = b[ip]; ip += 1; ml += x
            if x != 255: break
    ml += 4
    pos = len(out)
    if off == 0 or off > pos:
        print(f"seq {k} at stream {ip0}: bad offset {off} at out {pos} (ll={ll}, ml={ml})"); break
    for j in range(ml): out.append(out[pos - off + j])
    if bytes(out[pos:pos + ml]) != src[pos:pos + ml]:
        print(f"seq {k} at stream {ip0}: match differs at out {pos}: off={off} ml={ml} ll={ll} got {bytes(out[pos:pos+24])!r} want {src[pos:pos+24]!r}"); break
    k += 1
print("walked", k, "sequences,", len(out), "bytes of partition 0 (expects", int(ent[0, 2]), ")")
