#!/bin/bash
# A/B: default lib vs variants in lib_*; prints compress ms of the LZ4 bench; then parity tests with the default lib
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> /tmp/err.txt | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('compress_ms', round(j['detail']['compress_ms'],2), 'decompress_ms', round(j['detail']['decompress_ms'],2), 'e2e', round(j['e2e']['value'],2))"; tail -n 3 /tmp/err.txt; }
run X=1
for d in aocl-compression_b200/lib_*; do [ -f $d/libaocl_compression.so ] && run AOCL_LLC_LIB=$PWD/$d/libaocl_compression.so; done
[ -z "$SKIP_TESTS" ] && timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
