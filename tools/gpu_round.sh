#!/bin/bash
# One GPU-box visit: the whole -m gpu suite, the default bench line, the reference arm, the ncu launch list and full
# ncu captures of the hot kernels.  Usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/${tag}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -n 4 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "bench reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:lz4_encode_parts -c 1 -f -o gpurun_out/${tag}_enc \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/${tag}_ncu_enc.log 2>&1; echo "ncu enc rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:decode_parts_tile -c 1 -f -o gpurun_out/${tag}_dec \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/${tag}_ncu_dec.log 2>&1; echo "ncu dec rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:snappy_encode_frags -c 1 -f -o gpurun_out/${tag}_senc \
    python bench.py --workload snappy_log --steps 1 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/${tag}_ncu_senc.log 2>&1; echo "ncu snappy enc rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:decode_parts_tile -c 1 -f -o gpurun_out/${tag}_sdec \
    python bench.py --workload snappy_log --steps 1 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/${tag}_ncu_sdec.log 2>&1; echo "ncu snappy dec rc=$?"
cut -c1-700 gpurun_out/${tag}_bench.json; cut -c1-400 gpurun_out/${tag}_bench_reference.json
