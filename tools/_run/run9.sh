set -x
timeout 1200 python -m pytest tests/test_gpu_color_edit.py tests/test_gpu_edit.py tests/test_gpu_gc_io.py tests/test_cpp_host.py tests/test_gpu_more.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/stress_edit.py 2>&1 | tail -3
HD_EDIT_FAST_TRACE=1 timeout 900 python tools/edit_probe.py --reps 1 --color 24 --mid 100 > gpurun_out/r2h_color_cfg3.log 2>&1; tail -1 gpurun_out/r2h_color_cfg3.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['color']); print(d['mid'])"
grep "fused edit: 1 editors" -A1 gpurun_out/r2h_color_cfg3.log | tail -4
timeout 600 python tools/bench_brush.py --edits 60 --cpu-sample 20 --radii 2,32,128,256 2>&1 | tail -1
timeout 600 python tools/bench_color_edit.py --edits 60 --cpu-edits 10 2>&1 | tail -1
