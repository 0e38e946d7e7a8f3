"""BASELINE configs[4]: batched decompression (and compression) of independent 64 KiB LZ4 / Snappy pages,
device-resident, one launch per batch.  16,384 distinct columnar-like pages (1 GiB) are generated and compressed
by the GPU batch encoder (checked against the oracle on a sample); the compressed pages are then replicated
`--rep` times on the device (SURVEY 8(d) allows this shortcut) so that one decode launch covers up to 1M pages.
Prints one JSON line per codec.  Run under gpurun:  python tools/pages_bench.py --rep 64"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, llc_b200
import oracle_lib as ol
from llc_b200 import gen

ap = argparse.ArgumentParser()
ap.add_argument("--pages", type=int, default=16384)
ap.add_argument("--rep", type=int, default=64)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--codec", default="lz4,snappy")
args = ap.parse_args()
P, PS = args.pages, 65536
L = llc_b200.load()
ctx = llc_b200.GpuContext(0)
stream = torch.cuda.ExternalStream(ctx.stream)
orc = ol.Oracle()
t0 = time.time()
pages = gen.pages(P)
print(f"generated {P} pages in {time.time()-t0:.1f}s", file=sys.stderr)
d_in = torch.from_numpy(pages.reshape(-1)).cuda()
for codec, name in [(c, nm) for c, nm in ((0, "lz4"), (4, "snappy")) if nm in args.codec.split(",")]:
    bound = int(orc.bound(PS, codec))
    bound = (bound + 255) // 256 * 256
    d_comp = torch.zeros(P * bound, dtype=torch.uint8, device="cuda")
    in_ptrs = (torch.arange(P, dtype=torch.int64, device="cuda") * PS + d_in.data_ptr())
    out_ptrs = (torch.arange(P, dtype=torch.int64, device="cuda") * bound + d_comp.data_ptr())
    in_sizes = torch.full((P,), PS, dtype=torch.int32, device="cuda")
    out_caps = torch.full((P,), bound, dtype=torch.int32, device="cuda")
    status = torch.zeros(P, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()      # torch fills the descriptor arrays on ITS stream; the library runs on its own
    best_c = 1e9
    for it in range(args.iters):
        e0.record(stream)
        ctx.compress_batch_async(codec, in_ptrs, in_sizes, out_ptrs, out_caps, status, P)
        e1.record(stream)
        rc = ctx.finish()
        assert rc == 0, ("compress", it, rc)
        best_c = min(best_c, e0.elapsed_time(e1))
    csz = status.cpu().numpy()
    for i in range(0, P, max(1, P // 32)):                   # sample check against the oracle
        want = orc.compress(pages[i], codec)
        assert csz[i] == len(want) and d_comp[i * bound: i * bound + csz[i]].cpu().numpy().tobytes() == want, i
    # replicate the compressed pages: rep * P page descriptors pointing into the same compressed bytes,
    # every one with its own output page
    R = args.rep
    N = P * R
    d_out = torch.empty(N * PS, dtype=torch.uint8, device="cuda")
    big_in_ptrs = out_ptrs.repeat(R)
    big_in_sizes = torch.from_numpy(csz.astype(np.int32)).cuda().repeat(R)
    big_out_ptrs = (torch.arange(N, dtype=torch.int64, device="cuda") * PS + d_out.data_ptr())
    big_caps = torch.full((N,), PS, dtype=torch.int32, device="cuda")
    big_status = torch.zeros(N, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    best_d = 1e9
    for it in range(args.iters):
        e0.record(stream)
        ctx.decompress_batch_async(codec, big_in_ptrs, big_in_sizes, big_out_ptrs, big_caps, big_status, N)
        e1.record(stream)
        rc = ctx.finish()
        assert rc == 0, ("decompress", it, rc)
        best_d = min(best_d, e0.elapsed_time(e1))
    assert bool((big_status == PS).all())
    for r in (0, R // 2, R - 1):                             # every replica decodes to the original pages
        assert torch.equal(d_out[r * P * PS:(r + 1) * P * PS], d_in), r
    C = int(csz.sum())
    print(json.dumps({"workload": f"{N} independent 64 KiB {name} pages (BASELINE configs[4]); {P} distinct pages x{R} replicas of the compressed bytes",
                      "decoder": os.environ.get("AOCL_GPU_DECODER", "tile (default)"),
                      "decompress_ms": best_d, "decompress_GBps": N * PS / best_d / 1e6, "pages_per_s": N / best_d * 1e3,
                      "compress_ms_16384_pages": best_c, "compress_GBps": P * PS / best_c / 1e6, "ratio": C / (P * PS),
                      "hbm_algorithmic_GBps": (N * PS + C * R) / best_d / 1e6}))
    del d_out, d_comp
    torch.cuda.empty_cache()
