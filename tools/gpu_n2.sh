#!/bin/bash
# two-GPU visit (gpurun --gpus 2): the sharding test, the bench line under torchrun, the reference arm
tag=${1:-n2}
mkdir -p gpurun_out
timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_shard.py > gpurun_out/${tag}_pytest_shard.log 2>&1; echo "pytest shard rc=$?"; tail -n 3 gpurun_out/${tag}_pytest_shard.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 \
   > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench n2 rc=$?"; grep -v "^\[W\|Warning\|^W1\|^\*\*\*\|OMP_NUM" gpurun_out/${tag}_bench.err | tail -n 5
python - "$tag" <<'PY'
import json,sys
j=json.loads(open(f"gpurun_out/{sys.argv[1]}_bench.json").read().strip().splitlines()[-1])
d=j["detail"]
print("value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], j["e2e"].get("pcie"))
print("exchange_ms", d["exchange_ms"], [ (r["rank"], round(r["step_ms"],2), round(r["exchange_ms"],3)) for r in d.get("per_rank",[])])
for k,v in d["configs"].items(): print(k, json.dumps(v)[:600])
PY
