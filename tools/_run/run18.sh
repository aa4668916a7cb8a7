set -x
for g in 0 148 74 37; do HD_EDIT_FUSED_GRID=$g timeout 600 python tools/bench_brush.py --edits 60 --cpu-sample 0 --radii 2,8,32,64,128,256 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print([(r['radius'], r['gpu_ms_median']) for r in d['rows']])"; done
