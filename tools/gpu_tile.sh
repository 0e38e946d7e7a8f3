#!/bin/bash
# Tile decoder bring-up: parity tests with the tile decoder selected, then a short bench.
tag=${1:-tile}
mkdir -p gpurun_out
AOCL_GPU_DECODER=tile timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 25 gpurun_out/${tag}_pytest.log
AOCL_GPU_DECODER=tile timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_lz4.json 2> gpurun_out/${tag}_bench_lz4.err; echo "bench lz4 rc=$?"
tail -c 1500 gpurun_out/${tag}_bench_lz4.err
python - <<PY
import json
for w in ("lz4",):
    try:
        j=json.loads(open("gpurun_out/${tag}_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, "compress_ms", round(j["detail"]["compress_ms"],2), "decompress_ms", round(j["detail"]["decompress_ms"],3), j["detail"]["kernels_ms"])
    except Exception as e: print("no bench", w, e)
PY
