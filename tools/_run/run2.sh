set -x
timeout 900 python -m pytest tests/test_gpu_trace.py tests/test_gpu_more.py -m gpu -x -q 2>&1 | tail -5
for t in 0 1; do HD_TRACE_TABLE=$t timeout 300 python tools/trace_probe.py --frames 20 --lod 2>&1 | tail -1; done
HD_TRACE_VARIANT=4 timeout 300 python tools/trace_probe.py --frames 20 --lod 2>&1 | tail -1
HD_TRACE_TABLE=1 timeout 300 python tools/trace_probe.py --frames 20 --lod 2>&1 | tail -1
HD_TRACE_TABLE=1 timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:trace_kernel -s 4 -c 2 --csv --log-file gpurun_out/r2b_trace_table.csv python tools/trace_probe.py --frames 3 > gpurun_out/r2b_trace_table.log 2>&1
HD_TRACE_TABLE=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -s 4 -c 1 -f -o gpurun_out/r2b_trace_table_full python tools/trace_probe.py --frames 3 > gpurun_out/r2b_trace_table_full.log 2>&1
HD_EDIT_FAST_TRACE=1 timeout 600 python tools/edit_probe.py --reps 1 --mid 100 > gpurun_out/r2b_mid100.log 2>&1; tail -1 gpurun_out/r2b_mid100.log
HD_EDIT_FAST_TRACE=1 timeout 600 python tools/edit_probe.py --reps 1 --mid 33 > gpurun_out/r2b_mid33.log 2>&1; tail -1 gpurun_out/r2b_mid33.log
HD_EDIT_FAST_TRACE=1 timeout 600 python tools/bench_brush.py --edits 40 --cpu-sample 0 --radii 2,32,128 > gpurun_out/r2b_brush.log 2>&1; tail -1 gpurun_out/r2b_brush.log
