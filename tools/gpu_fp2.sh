#!/bin/bash
tag=${1:-fp2}
mkdir -p gpurun_out
timeout 600 python -m pytest -x -q -s -m gpu tests/test_gpu_fastparse.py tests/test_gpu_consumers.py > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; grep "fastparse\]\|passed\|failed\|Error\|assert\|speed" gpurun_out/${tag}_pytest.log | cut -c1-220 | head -40
timeout 400 python bench.py --steps 2 --no-cpu-baseline --configs none > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python -c "
import json,sys
j=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1]); print(j['detail'].get('fastparse'), 'exact ms', j['detail']['compress_ms'])"
AOCL_GPU_MODE=fastparse timeout 500 ncu --set full --clock-control none --import-source on -k regex:lz4_fastparse -c 1 -f -o gpurun_out/${tag}_fastparse \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
