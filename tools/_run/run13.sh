set -x
timeout 1200 python -m pytest tests/test_gpu_trace.py tests/test_gpu_edit.py -m gpu -x -q 2>&1 | tail -3
for v in "" 0 "" 0; do HD_TRACE_VARIANT=$v timeout 300 python tools/trace_probe.py --frames 20 2>&1 | tail -1; done
for w in 0 1; do
HD_EDIT_SCAN_WIDE=$w HD_EDIT_FAST_TRACE=1 timeout 900 python tools/edit_probe.py --reps 1 --color 24 --mid 100 > gpurun_out/r2j_w$w.log 2>&1; tail -1 gpurun_out/r2j_w$w.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['color']['median_all'], d['color']['median_fill'], d['color']['median_paint']); print(d['mid'])"
grep "fused edit: 1 editors" -A1 gpurun_out/r2j_w$w.log | tail -1
grep "fused edit: 100 editors" -A1 gpurun_out/r2j_w$w.log | tail -1
HD_EDIT_SCAN_WIDE=$w timeout 600 python tools/edit_probe.py --reps 1 --mid 33 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['mid'])"
HD_EDIT_SCAN_WIDE=$w timeout 600 python tools/bench_brush.py --edits 60 --cpu-sample 0 --radii 2,32,128,256 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print([(r['radius'], r['gpu_ms_median']) for r in d['rows']])"
done
