#!/bin/bash
# Snappy encoder: shared-memory-table and global-table warps side by side.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/k_pytest.log
{
timeout 200 python tools/enc_sweep.py snappy_log 3
for cfg in "24 0" "24 4" "22 5" "20 6" "26 4" "0 6"; do set -- $cfg; AOCL_GPU_SNAPPY_GTAB_CTAS=$1 AOCL_GPU_SNAPPY_STAB_CTAS=$2 timeout 200 python tools/enc_sweep.py snappy_log 3; done
} 2>&1 | grep -v Warning | tee gpurun_out/k_sweep.txt
