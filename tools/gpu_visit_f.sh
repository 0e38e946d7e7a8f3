#!/bin/bash
# Tile decoder iteration: parity tests, then timing + phase profile of the default build and any lib_v* variants.
mkdir -p gpurun_out
D=$PWD/aocl-compression_b200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/f_pytest.log
{
for w in lz4_text snappy_log; do
timeout 200 python tools/enc_sweep.py $w 3
AOCL_LLC_LIB=$D/lib_tprof/libaocl_compression.so timeout 200 python tools/enc_sweep.py $w 2
for v in $D/lib_v*; do [ -f $v/libaocl_compression.so ] && AOCL_LLC_LIB=$v/libaocl_compression.so timeout 200 python tools/enc_sweep.py $w 3; done
done
} 2>&1 | grep -v Warning | tee gpurun_out/f_sweep.txt
