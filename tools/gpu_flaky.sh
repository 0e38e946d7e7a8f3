#!/bin/bash
for i in 1 2 3 4 5 6 7 8; do AOCL_GPU_VERBOSE=1 timeout 300 python -m pytest tests/test_gpu_device_api.py tests/test_gpu_kat.py -m gpu -x -q -s -k "device_resident or golden" 2>&1 | grep -E "passed|failed|FAILED|assert \(|call failed" | head -8; echo "--- run $i"; done
