#!/bin/bash
run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload snappy_log --steps 3 --warmup 3 --no-cpu-baseline 2> /tmp/err.txt | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('compress_ms', round(j['detail']['compress_ms'],2), 'decompress_ms', round(j['detail']['decompress_ms'],2), 'e2e', round(j['e2e']['value'],2))"; tail -n 2 /tmp/err.txt; }
for g in 12 20 24 28; do run AOCL_GPU_SNAPPY_GTAB_CTAS=$g; done
