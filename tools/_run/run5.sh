set -x
timeout 1200 python -m pytest tests/test_gpu_color_edit.py tests/test_gpu_edit.py tests/test_gpu_gc_io.py tests/test_cpp_host.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/bench_color_edit.py --edits 60 --cpu-edits 10 2>&1 | tail -1
HD_COLOR_OVERLAP=0 timeout 600 python tools/bench_color_edit.py --edits 60 --cpu-edits 0 2>&1 | tail -1
timeout 600 python tools/bench_color_edit.py --edits 60 --cpu-edits 10 --radius 32 2>&1 | tail -1
HD_EDIT_FAST_TRACE=1 timeout 600 python tools/edit_probe.py --reps 1 --mid 100 > gpurun_out/r2d_mid100.log 2>&1; tail -1 gpurun_out/r2d_mid100.log
HD_EDIT_FAST_TRACE=1 timeout 600 python tools/edit_probe.py --reps 1 --mid 33 > gpurun_out/r2d_mid33.log 2>&1; tail -1 gpurun_out/r2d_mid33.log
timeout 600 python tools/bench_brush.py --edits 60 --cpu-sample 0 --radii 2,32,128,256 2>&1 | tail -1
