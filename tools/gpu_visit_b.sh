#!/bin/bash
# GPU visit B: parity tests under both decoders, then bench lines for both workloads.
# Usage (under gpurun): bash tools/gpu_visit_b.sh <tag>
tag=${1:-visB}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest(default decoder) rc=$?"
tail -n 15 gpurun_out/${tag}_pytest.log
AOCL_GPU_DECODER=warp timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_warp.log 2>&1; echo "pytest(warp decoder) rc=$?"
tail -n 3 gpurun_out/${tag}_pytest_warp.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_lz4.json 2> gpurun_out/${tag}_bench_lz4.err; echo "bench lz4 rc=$?"
tail -n 5 gpurun_out/${tag}_bench_lz4.err
timeout 400 python bench.py --workload snappy_log --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_snappy.json 2> gpurun_out/${tag}_bench_snappy.err; echo "bench snappy rc=$?"
tail -n 5 gpurun_out/${tag}_bench_snappy.err
for f in lz4 snappy; do python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/${tag}_bench_${f}.json").read().strip().splitlines()[-1])
    print("${f}", "value", round(j["value"],2), "e2e", round(j["e2e"]["value"],2), j["e2e"].get("detail"), "c_ms", round(j["detail"]["compress_ms"],2), "d_ms", round(j["detail"]["decompress_ms"],3), j["detail"]["kernels_ms"])
except Exception as e: print("no bench ${f}", e)
PY
done
