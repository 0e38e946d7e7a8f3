"""Replay test_gpu_matches_golden[mixed-0] exactly (compress + decompress through the host API for every golden size)
and print per-size results + tile watchdog counters."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, llc_b200
import kat, oracle_lib as ol
L = llc_b200.load()
gpu = ol.LlcLib(llc_b200.LIB_PATH)
cnt = (C.c_uint64 * 32)()
def counters():
    L.aocl_gpu_debug_counters(C.cast(cnt, C.c_void_p), 1)
    return list(cnt)[24:32]
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["mixed"]
do_compress = (sys.argv[2] != "nocomp") if len(sys.argv) > 2 else True
orc = ol.Oracle()
for name in names:
  for codec in (kat.LZ4, kat.SNAPPY):
    for c in [c for c in GOLDEN["cases"] if c["codec"] == codec and c["gen"] == name]:
        data = kat.make_input(name, c["size"])
        if do_compress:
            r, got = gpu.compress(data, codec)
            okc = (r == c["out_len"] and kat.sha(got) == c["out_sha256"])
        else:
            got = orc.compress(data, codec); okc = None
        counters()
        r2, back = gpu.decompress(got, codec, max(c["size"], 1))
        ok = (r2 == c["size"] and back == data.tobytes())
        w = counters()
        flag = "" if ok else "   <<<<<<<< FAIL"
        print(f"{name} codec {codec} size {c['size']}: compress ok={okc} decompress r={r2} ok={ok} wd={w}{flag}", flush=True)
        if not ok:
            for rep in range(3):
                r3, back3 = gpu.decompress(got, codec, max(c["size"], 1))
                print("    retry", rep, r3, r3 == c["size"] and back3 == data.tobytes(), counters(), flush=True)
