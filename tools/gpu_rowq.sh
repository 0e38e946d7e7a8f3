#!/bin/bash
# Row-decoder iteration visit: probe -> forced-decoder parity suite -> bench -> one full ncu capture.
tag=${1:-rq}
mkdir -p gpurun_out
AOCL_GPU_DECODER=rowq AOCL_GPU_VERBOSE=1 timeout 300 python tools/decode_probe.py all > gpurun_out/${tag}_probe.log 2>&1; echo "probe rc=$?"
tail -n 1 gpurun_out/${tag}_probe.log
AOCL_GPU_DECODER=rowq timeout 600 python -m pytest -x -q -m gpu tests/test_gpu_parity.py tests/test_gpu_kat.py tests/test_synth_streams.py > gpurun_out/${tag}_pytest_rowq.log 2>&1; echo "pytest(rowq) rc=$?"
tail -n 3 gpurun_out/${tag}_pytest_rowq.log
for mode in rowq; do
  AOCL_GPU_DECODER=$mode timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_lz4_${mode}.json 2> gpurun_out/${tag}_bench_lz4_${mode}.err; echo "bench lz4 $mode rc=$?"
  AOCL_GPU_DECODER=$mode timeout 400 python bench.py --workload snappy_log --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_snappy_${mode}.json 2> gpurun_out/${tag}_bench_snappy_${mode}.err; echo "bench snappy $mode rc=$?"
done
python - "$tag" <<'PY'
import json,glob,sys
for f in sorted(glob.glob("gpurun_out/"+sys.argv[1]+"_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); d=j["detail"]
        print(f.split("/")[-1], "dec_ms", round(d["decompress_ms"],3), "dec GB/s", round(d["decompress_GBps"],1), "comp_ms", round(d["compress_ms"],2), {k:v for k,v in d["kernels_ms"].items() if "decode" in k})
    except Exception as e: print(f, "unreadable", e)
PY
AOCL_GPU_DECODER=rowq timeout 500 ncu --set full --clock-control none --import-source on -k regex:decode_parts_rowq -c 1 -f -o gpurun_out/${tag}_dec_rowq \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_dec.log 2>&1; echo "ncu dec rc=$?"
