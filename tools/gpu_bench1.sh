#!/bin/bash
# bench.py bring-up: small-size run of every config, then decoder A/B on the many-unit configs, then the full default line
tag=${1:-b1}
mkdir -p gpurun_out
timeout 600 python bench.py --size $((256<<20)) --steps 2 > gpurun_out/${tag}_small.json 2> gpurun_out/${tag}_small.err; echo "small rc=$?"
tail -n 5 gpurun_out/${tag}_small.err
python - "$tag" <<'PY'
import json,sys
j=json.loads(open(f"gpurun_out/{sys.argv[1]}_small.json").read().strip().splitlines()[-1])
print("value", j["value"], "e2e", j["e2e"]["value"], "identical", j["detail"].get("bytes_identical_to_reference"), "cpu", j["cpu_baseline"] and j["cpu_baseline"]["value"])
for k,v in j["detail"]["configs"].items(): print(k, json.dumps(v)[:400])
PY
for mode in tile rowq; do
  AOCL_GPU_DECODER=$mode timeout 600 python bench.py --steps 2 --no-cpu-baseline --configs 3,4 > gpurun_out/${tag}_cfg34_${mode}.json 2> gpurun_out/${tag}_cfg34_${mode}.err; echo "cfg34 $mode rc=$?"
  python - "$tag" "$mode" <<'PY'
import json,sys
j=json.loads(open(f"gpurun_out/{sys.argv[1]}_cfg34_{sys.argv[2]}.json").read().strip().splitlines()[-1])
c=j["detail"]["configs"]
print(sys.argv[2], "1GiB dec ms", j["detail"]["decompress_ms"], "| frames16", c.get("3_lz4_16_frames_decode",{}).get("decompress_ms"), c.get("3_lz4_16_frames_decode",{}).get("error"), "| pages", {k:(v.get("decompress_ms"), round(v.get("decompress_GBps",0),1)) for k,v in c.get("4_pages_1M_decode",{}).items() if isinstance(v,dict)}, c.get("4_pages_1M_decode",{}).get("error"))
PY
done
(time timeout 900 python bench.py) > gpurun_out/${tag}_full.json 2> gpurun_out/${tag}_full.err; echo "full rc=$?"; tail -n 4 gpurun_out/${tag}_full.err
cut -c1-600 gpurun_out/${tag}_full.json
(time timeout 600 python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err; echo "ref rc=$?"; tail -n 4 gpurun_out/${tag}_ref.err; cut -c1-500 gpurun_out/${tag}_ref.json
