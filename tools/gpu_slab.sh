#!/bin/bash
# host-buffer decompress: partitions per slab of the H2D / decode / D2H pipeline
mkdir -p gpurun_out
for per in 0 148 296 444 592 1184; do
  if [ $per = 0 ]; then unset AOCL_GPU_SLAB_PARTS; else export AOCL_GPU_SLAB_PARTS=$per; fi
  timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --configs none > gpurun_out/slab_$per.json 2> gpurun_out/slab_$per.err
  python - "$per" <<'PY'
import json,sys
j=json.loads(open(f"gpurun_out/slab_{sys.argv[1]}.json").read().strip().splitlines()[-1]); e=j["e2e"]
print("slab parts", sys.argv[1], "e2e", round(e["value"],2), "compress_ms", round(e["compress_ms"],2), "decompress_ms", round(e["decompress_ms"],2))
PY
done
