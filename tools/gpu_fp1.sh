#!/bin/bash
tag=${1:-fp1}
mkdir -p gpurun_out
timeout 600 python -m pytest -x -q -s -m gpu tests/test_gpu_fastparse.py > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; grep "fastparse\]\|passed\|failed\|Error\|assert" gpurun_out/${tag}_pytest.log | head -30
for cfg in "default" "stab" ; do
  if [ $cfg = stab ]; then export AOCL_GPU_STAB_CTAS=11 AOCL_GPU_GTAB_CTAS=0; fi
  timeout 400 python bench.py --steps 2 --no-cpu-baseline --configs none > gpurun_out/${tag}_bench_${cfg}.json 2> gpurun_out/${tag}_bench_${cfg}.err; echo "bench $cfg rc=$?"
  python -c "
import json,sys
j=json.loads(open('gpurun_out/${tag}_bench_${cfg}.json').read().strip().splitlines()[-1]); print('$cfg', j['detail'].get('fastparse'), 'exact ms', j['detail']['compress_ms'])"
  unset AOCL_GPU_STAB_CTAS AOCL_GPU_GTAB_CTAS
done
tail -n 3 gpurun_out/${tag}_bench_default.err
