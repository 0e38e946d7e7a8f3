#!/bin/bash
# N-GPU bench under torchrun exactly as the driver launches it (gpurun --gpus N -- bash tools/gpu_n8.sh <tag> <N>)
tag=${1:-n8}; N=${2:-8}
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 5 --warmup 3 \
   > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2>&1 | grep real; echo "bench n$N rc=$?"; grep -v "^\[W\|Warning\|^W1\|^\*\*\*\|OMP_NUM" gpurun_out/${tag}_bench.err | tail -n 5
python - "$tag" <<'PY'
import json,sys
j=json.loads(open(f"gpurun_out/{sys.argv[1]}_bench.json").read().strip().splitlines()[-1])
d=j["detail"]
print("value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], j["e2e"].get("pcie"))
print("exchange_ms", d["exchange_ms"], [ (r["rank"], round(r["step_ms"],2), round(r["exchange_ms"],3)) for r in d.get("per_rank",[])])
for k,v in d["configs"].items(): print(k, json.dumps(v)[:500])
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err ) 2>&1 | grep real; cut -c1-300 gpurun_out/${tag}_ref.json
