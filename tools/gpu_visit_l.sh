#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/l_pytest.log
{
timeout 200 python tools/enc_sweep.py lz4_text 3 2>&1 | grep -v Warning
timeout 300 python tools/pages_bench.py --rep 16 --codec lz4 2>/dev/null
} | tee gpurun_out/l_sweep.txt | cut -c1-420
