#!/bin/bash
# Pages (BASELINE configs[4]) with both decoder organisations; smaller replica count to bound time.
mkdir -p gpurun_out
{
timeout 300 python tools/pages_bench.py --rep 16 2>/dev/null
AOCL_GPU_DECODER=tile timeout 300 python tools/pages_bench.py --rep 16 2>/dev/null
} | tee gpurun_out/j_pages.jsonl | cut -c1-400
