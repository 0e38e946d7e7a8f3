#!/bin/bash
mkdir -p gpurun_out
{ for w in lz4_mixed snappy_mixed lz4_log snappy_text; do timeout 300 python tools/enc_sweep.py $w 2 2>&1 | grep -v Warning | tail -3; done; } | tee gpurun_out/m_sweep.txt | cut -c1-300
