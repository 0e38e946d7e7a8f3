#!/bin/bash
# encoder placement / L2 policy sweep (compress_ms only)
run() { echo "== $*"; env "$@" AOCL_GPU_VERBOSE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> /tmp/err.txt | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('compress_ms', round(j['detail']['compress_ms'],2), 'decompress_ms', round(j['detail']['decompress_ms'],2), 'e2e', round(j['e2e']['value'],2))"; grep -m2 "L2 persisting" /tmp/err.txt; }
run X=1
run AOCL_GPU_NO_L2_PERSIST=1
run AOCL_GPU_STAB_CTAS=11 AOCL_GPU_GTAB_CTAS=17
run AOCL_GPU_STAB_CTAS=8 AOCL_GPU_GTAB_CTAS=20
