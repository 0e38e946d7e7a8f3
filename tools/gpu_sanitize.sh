#!/bin/bash
# compute-sanitizer over the code added in round 2 (small inputs): memcheck on the smoke round trip, the known-answer /
# device-API / fastparse tests in the default configuration and with the row decoder forced; racecheck (shared-memory
# hazards) on the row decoder and the fastparse encoder through the probe / fastparse test.
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python __graft_entry__.py smoke > gpurun_out/san_smoke.log 2>&1; echo "memcheck smoke rc=$?"
tail -n 3 gpurun_out/san_smoke.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kat.py tests/test_gpu_device_api.py tests/test_gpu_fastparse.py -m gpu -x -q -k "not many_partitions and not repetition and not every_decoder" > gpurun_out/san_tests.log 2>&1; echo "memcheck tests rc=$?"
tail -n 4 gpurun_out/san_tests.log
AOCL_GPU_DECODER=rowq timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kat.py tests/test_synth_streams.py -m gpu -x -q > gpurun_out/san_tests_rowq.log 2>&1; echo "memcheck tests (row decoder) rc=$?"
tail -n 4 gpurun_out/san_tests_rowq.log
AOCL_GPU_DECODER=rowq timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python tools/decode_probe.py text,mixed,period7 70000,300000 > gpurun_out/san_race_rowq.log 2>&1; echo "racecheck row decoder rc=$?"
grep -E "RACECHECK SUMMARY|FAILURES|hazard" gpurun_out/san_race_rowq.log | head -5
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python tools/fp_debug.py > gpurun_out/san_race_fastparse.log 2>&1; echo "racecheck fastparse rc=$?"
grep -E "RACECHECK SUMMARY|walked|hazard" gpurun_out/san_race_fastparse.log | head -5
