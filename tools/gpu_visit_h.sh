#!/bin/bash
mkdir -p gpurun_out
D=$PWD/aocl-compression_b200
{
for v in $D/lib_v*; do echo "== $v"; AOCL_LLC_LIB=$v/libaocl_compression.so timeout 200 python tools/enc_sweep.py lz4_text 2 2>&1 | tail -2; done
} 2>&1 | grep -v Warning | tee gpurun_out/h_sweep.txt
