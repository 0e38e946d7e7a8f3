#!/bin/bash
mkdir -p gpurun_out
D=$PWD/aocl-compression_b200
{ timeout 100 python tools/enc_sweep.py lz4_text 2; AOCL_LLC_LIB=$D/lib_vpf/libaocl_compression.so timeout 100 python tools/enc_sweep.py lz4_text 2; } 2>&1 | grep -v Warning | tee gpurun_out/o_sweep.txt | cut -c1-260
