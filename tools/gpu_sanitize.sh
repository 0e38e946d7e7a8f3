#!/bin/bash
# compute-sanitizer memcheck over the smoke round trip and a slice of the GPU parity tests (small inputs).
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python __graft_entry__.py smoke > gpurun_out/san_smoke.log 2>&1; echo "memcheck smoke rc=$?"
tail -n 6 gpurun_out/san_smoke.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kat.py tests/test_gpu_device_api.py -m gpu -x -q -k "not many_partitions and not repetition" > gpurun_out/san_tests.log 2>&1; echo "memcheck tests rc=$?"
tail -n 8 gpurun_out/san_tests.log
