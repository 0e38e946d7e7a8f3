#!/bin/bash
# One GPU-box visit: parity tests, both bench workloads, the launch list and a full ncu capture.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_lz4.json 2> gpurun_out/${tag}_bench_lz4.err; echo "bench lz4 rc=$?"
timeout 600 python bench.py --workload snappy_log --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_snappy.json 2> gpurun_out/${tag}_bench_snappy.err; echo "bench snappy rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --size 268435456 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
tail -n 3 gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_bench_lz4.json gpurun_out/${tag}_bench_snappy.json
