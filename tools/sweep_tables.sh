for g in 0 16 24 32 48; do AOCL_GPU_SNAPPY_GTAB_CTAS=$g timeout 300 python bench.py --workload snappy_log --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('snappy gtab $g', round(j['detail']['compress_ms'],2), round(j['detail']['decompress_ms'],2), j['detail']['ratio'])
"; done
