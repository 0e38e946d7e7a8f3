#!/bin/bash
# Tile decoder: variants + one full ncu capture (source-level) of the LZ4 tile decoder.
mkdir -p gpurun_out
D=$PWD/aocl-compression_b200
{
for w in lz4_text snappy_log; do
timeout 200 python tools/enc_sweep.py $w 3
for v in $D/lib_v*; do [ -f $v/libaocl_compression.so ] && AOCL_LLC_LIB=$v/libaocl_compression.so timeout 200 python tools/enc_sweep.py $w 3; done
done
} 2>&1 | grep -v Warning | tee gpurun_out/g_sweep.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_parts_tile -c 1 -f -o gpurun_out/g_dec \
    python tools/enc_sweep.py lz4_text 1 > gpurun_out/g_ncu.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
