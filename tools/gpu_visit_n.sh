#!/bin/bash
mkdir -p gpurun_out
D=$PWD/aocl-compression_b200
{ for w in lz4_text snappy_log; do timeout 200 python tools/enc_sweep.py $w 3; for v in $D/lib_v*; do AOCL_LLC_LIB=$v/libaocl_compression.so timeout 200 python tools/enc_sweep.py $w 3; done; done; } 2>&1 | grep -v Warning | tee gpurun_out/n_sweep.txt | cut -c1-260
