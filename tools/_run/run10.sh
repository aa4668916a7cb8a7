set -x
timeout 1200 python -m pytest tests/test_gpu_edit.py tests/test_gpu_color_edit.py -m gpu -x -q 2>&1 | tail -3
HD_EDIT_FAST_TRACE=1 timeout 900 python tools/edit_probe.py --reps 1 --color 24 --mid 100 > gpurun_out/r2i_color_cfg3.log 2>&1; tail -1 gpurun_out/r2i_color_cfg3.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['color']['median_all'], d['color']['median_fill'], d['color']['median_paint']); print(d['mid'])"
grep "fused edit: 1 editors" -A1 gpurun_out/r2i_color_cfg3.log | tail -2
grep "fused edit: 100 editors" -A1 gpurun_out/r2i_color_cfg3.log | tail -2
timeout 600 python tools/edit_probe.py --reps 1 --mid 33 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['mid'])"
timeout 600 python tools/bench_brush.py --edits 60 --cpu-sample 20 --radii 2,32,128,256 2>&1 | tail -1
timeout 600 python tools/bench_color_edit.py --edits 60 --cpu-edits 10 2>&1 | tail -1
