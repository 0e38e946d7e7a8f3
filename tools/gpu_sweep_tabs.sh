#!/bin/bash
# LZ4 exact encoder: shared-memory-table warps + L2-table warps per SM side by side (AOCL_GPU_STAB_CTAS / AOCL_GPU_GTAB_CTAS)
mkdir -p gpurun_out
for cfg in "0 32" "3 25" "5 23" "7 21" "9 19" "11 17" "11 21"; do
  set -- $cfg
  AOCL_GPU_STAB_CTAS=$1 AOCL_GPU_GTAB_CTAS=$2 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --configs none > gpurun_out/sweep_$1_$2.json 2> gpurun_out/sweep_$1_$2.err
  python - "$1" "$2" <<'PY'
import json,sys
try:
    j=json.loads(open(f"gpurun_out/sweep_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1]); d=j["detail"]
    print("stab", sys.argv[1], "gtab", sys.argv[2], "compress_ms", round(d["compress_ms"],2), "identical", d.get("bytes_identical_to_reference"))
except Exception as e: print("stab", sys.argv[1], "gtab", sys.argv[2], "failed", e)
PY
done
