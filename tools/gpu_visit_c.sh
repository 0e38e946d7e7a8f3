#!/bin/bash
tag=${1:-visC}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8 9 10; do AOCL_GPU_VERBOSE=1 timeout 300 python -m pytest tests/test_gpu_device_api.py tests/test_gpu_kat.py -m gpu -x -q -s -k "device_resident or golden" 2>&1 | grep -E "passed|failed|FAILED|assert \(|call failed" | head -8; done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 500 ncu --set full --clock-control none --import-source on -k regex:lz4_encode_parts -c 1 -f -o gpurun_out/${tag}_enc \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_enc.log 2>&1; echo "ncu enc rc=$?"
