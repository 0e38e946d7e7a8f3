#!/bin/bash
# Encoder L1 / shared-memory carve-out experiments.
mkdir -p gpurun_out
D=$PWD/aocl-compression_b200
{
timeout 200 python tools/enc_sweep.py lz4_text 3
for c in 10 25 50 100; do AOCL_GPU_ENC_CARVEOUT=$c timeout 200 python tools/enc_sweep.py lz4_text 3; done
AOCL_LLC_LIB=$D/lib_vclaim/libaocl_compression.so timeout 200 python tools/enc_sweep.py lz4_text 3
for c in 5 25; do AOCL_GPU_ENC_CARVEOUT=$c AOCL_LLC_LIB=$D/lib_vclaim/libaocl_compression.so timeout 200 python tools/enc_sweep.py lz4_text 3; done
timeout 200 python tools/enc_sweep.py snappy_log 3
for c in 25 100; do AOCL_GPU_ENC_CARVEOUT=$c timeout 200 python tools/enc_sweep.py snappy_log 3; done
} 2>&1 | grep -v Warning | tee gpurun_out/i_sweep.txt
