#!/bin/bash
# Session-5 visit 1: parity of the Snappy check-nibble encoder, CTAs-per-SM sweep, source prefetch A/B, tile phase profile.
mkdir -p gpurun_out
D=$PWD/aocl-compression_b200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/d_pytest.log
{
AOCL_LLC_LIB=$D/lib_base/libaocl_compression.so timeout 200 python tools/enc_sweep.py snappy_log
for g in 14 16 18 20; do AOCL_GPU_SNAPPY_GTAB_CTAS=$g timeout 200 python tools/enc_sweep.py snappy_log; done
for g in 16 20; do AOCL_LLC_LIB=$D/lib_pf/libaocl_compression.so AOCL_GPU_SNAPPY_GTAB_CTAS=$g timeout 200 python tools/enc_sweep.py snappy_log; done
AOCL_LLC_LIB=$D/lib_base/libaocl_compression.so timeout 200 python tools/enc_sweep.py lz4_text
timeout 200 python tools/enc_sweep.py lz4_text
AOCL_LLC_LIB=$D/lib_pf/libaocl_compression.so timeout 200 python tools/enc_sweep.py lz4_text
AOCL_LLC_LIB=$D/lib_pf2/libaocl_compression.so timeout 200 python tools/enc_sweep.py lz4_text
AOCL_LLC_LIB=$D/lib_tprof/libaocl_compression.so timeout 200 python tools/enc_sweep.py lz4_text 2
AOCL_LLC_LIB=$D/lib_tprof/libaocl_compression.so timeout 200 python tools/enc_sweep.py snappy_log 2
} 2>&1 | grep -v Warning | tee gpurun_out/d_sweep.txt
