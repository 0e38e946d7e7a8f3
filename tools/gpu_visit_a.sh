#!/bin/bash
# GPU visit A: parity tests, bench lines for both workloads and both decoders, full ncu captures of
# the LZ4 encoder and the two LZ4 decoders (1 GiB workload).  Usage (under gpurun): bash tools/gpu_visit_a.sh <tag>
tag=${1:-visA}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/${tag}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_lz4.json 2> gpurun_out/${tag}_bench_lz4.err; echo "bench lz4 rc=$?"
AOCL_GPU_DECODER=tile timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_lz4_tile.json 2> gpurun_out/${tag}_bench_lz4_tile.err; echo "bench lz4 tile rc=$?"
timeout 400 python bench.py --workload snappy_log --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_snappy.json 2> gpurun_out/${tag}_bench_snappy.err; echo "bench snappy rc=$?"
AOCL_GPU_DECODER=tile timeout 400 python bench.py --workload snappy_log --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_snappy_tile.json 2> gpurun_out/${tag}_bench_snappy_tile.err; echo "bench snappy tile rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:lz4_encode_parts -c 1 -f -o gpurun_out/${tag}_enc \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_enc.log 2>&1; echo "ncu enc rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:decode_parts -c 1 -f -o gpurun_out/${tag}_dec_warp \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_dec_warp.log 2>&1; echo "ncu dec warp rc=$?"
AOCL_GPU_DECODER=tile timeout 500 ncu --set full --clock-control none --import-source on -k regex:decode_parts -c 1 -f -o gpurun_out/${tag}_dec_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_dec_tile.log 2>&1; echo "ncu dec tile rc=$?"
tail -n 3 gpurun_out/${tag}_pytest.log
for f in lz4 lz4_tile snappy snappy_tile; do python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/${tag}_bench_${f}.json").read().strip().splitlines()[-1])
    print("${f}", "value", round(j["value"],2), "e2e", round(j["e2e"]["value"],2), "c_ms", round(j["detail"]["compress_ms"],2), "d_ms", round(j["detail"]["decompress_ms"],3), j["detail"]["kernels_ms"], j.get("cpu_baseline"))
except Exception as e: print("no bench ${f}", e)
PY
done
ls -la gpurun_out/
