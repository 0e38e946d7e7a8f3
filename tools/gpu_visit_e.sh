#!/bin/bash
# Tile decoder (thread-per-byte copies): parity tests, A/B against lib_base, phase profile.
mkdir -p gpurun_out
D=$PWD/aocl-compression_b200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/e_pytest.log
{
for w in lz4_text snappy_log; do
AOCL_LLC_LIB=$D/lib_base/libaocl_compression.so timeout 200 python tools/enc_sweep.py $w 3
timeout 200 python tools/enc_sweep.py $w 3
AOCL_LLC_LIB=$D/lib_tprof/libaocl_compression.so timeout 200 python tools/enc_sweep.py $w 2
done
} 2>&1 | grep -v Warning | tee gpurun_out/e_sweep.txt
