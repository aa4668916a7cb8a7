set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_trace.py tests/test_gpu_more.py -m gpu -x -q 2>&1 | tail -5
for v in 0 3 2 0 3 2; do HD_TRACE_VARIANT=$v timeout 300 python tools/trace_probe.py --frames 20 --lod 2>&1 | tail -1; done
for v in 0 3 2; do HD_TRACE_VARIANT=$v timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:trace_kernel -s 4 -c 2 --csv --log-file gpurun_out/r2b_trace_v$v.csv python tools/trace_probe.py --frames 3 > gpurun_out/r2b_trace_v$v.log 2>&1; done
