set -x
timeout 1200 python -m pytest tests/test_gpu_color_edit.py tests/test_gpu_edit.py tests/test_gpu_gc_io.py tests/test_cpp_host.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/bench_color_edit.py --edits 60 --cpu-edits 10 2>&1 | tail -1
HD_COLOR_OVERLAP=0 timeout 600 python tools/bench_color_edit.py --edits 60 --cpu-edits 0 2>&1 | tail -1
timeout 600 python tools/bench_color_edit.py --edits 60 --cpu-edits 10 --radius 32 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2e_color_launches.csv python tools/bench_color_edit.py --edits 6 --cpu-edits 0 > /dev/null 2>&1
