#!/bin/bash
# encoder iteration: parity tests of the compress path, then the bench line without the CPU legs
tag=${1:-enc}
mkdir -p gpurun_out
timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_kat.py tests/test_gpu_parity.py tests/test_gpu_interop.py -k "not unsaturated_hosts or writes" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${tag}_pytest.log
timeout 400 python bench.py --steps 5 --no-cpu-baseline --configs 2 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - "$tag" <<'PY'
import json,sys
j=json.loads(open(f"gpurun_out/{sys.argv[1]}_bench.json").read().strip().splitlines()[-1]); d=j["detail"]
print("value", round(j["value"],2), "compress_ms", round(d["compress_ms"],2), "decompress_ms", round(d["decompress_ms"],2), "fastparse", round(d["fastparse"]["compress_ms"],2), "identical", d.get("bytes_identical_to_reference"))
for k,v in d.get("configs",{}).items(): print(k, v.get("value"), {a:round(b,2) for a,b in v.get("detail",{}).items() if a in ("compress_ms","decompress_ms")})
PY
if [ -n "$2" ]; then
timeout 500 ncu --set full --clock-control none --import-source on -k regex:lz4_encode_parts -c 1 -f -o gpurun_out/${tag}_enc \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/${tag}_ncu_enc.log 2>&1; echo "ncu enc rc=$?"
fi
