set -x
timeout 1200 python -m pytest tests/test_gpu_trace.py tests/test_gpu_edit.py tests/test_cpp_host.py tests/test_gpu_color_edit.py -m gpu -x -q 2>&1 | tail -8
HD_TRACE_TABLE=2 HD_TRACE_TABLE_VERBOSE=1 timeout 300 python tools/trace_probe.py --frames 20 --lod 2>&1 | tail -3
timeout 300 python tools/trace_probe.py --frames 20 --lod 2>&1 | tail -1
HD_EDIT_FAST_TRACE=1 timeout 600 python tools/edit_probe.py --reps 1 --mid 100 > gpurun_out/r2c_mid100.log 2>&1; tail -1 gpurun_out/r2c_mid100.log
HD_EDIT_FAST_TRACE=1 timeout 600 python tools/edit_probe.py --reps 1 --mid 33 > gpurun_out/r2c_mid33.log 2>&1; tail -1 gpurun_out/r2c_mid33.log
HD_EDIT_FAST_TRACE=1 timeout 600 python tools/edit_probe.py --reps 1 --mid 500 > gpurun_out/r2c_mid500.log 2>&1; tail -1 gpurun_out/r2c_mid500.log
timeout 600 python tools/bench_brush.py --edits 60 --cpu-sample 0 --radii 2,32,128 > gpurun_out/r2c_brush.log 2>&1; tail -1 gpurun_out/r2c_brush.log
