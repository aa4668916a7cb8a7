#!/bin/bash
# Round 2, second pass (run on the GPU box): the full bench line and reference arm, the launch list of the same bench command,
# one ncu --set full capture of the product trace kernel, launch lists of the cfg3 edit batch and of the brush loop.
# Outputs under gpurun_out/; the summaries that are judged are copied into profiles/ by hand.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
tail -c 400 gpurun_out/r2b_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_reference.json 2>> gpurun_out/r2b_bench_n1.err
# launch list of the headline command (no CPU legs: they launch nothing)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launch_list_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_ncu_bench.log 2>&1
# the product trace kernel, full capture with source (cameras 1 and 2 of the probe; lean instantiation, fetch at the loop head)
ncu --set full --import-source on --clock-control none -k regex:trace_kernel -s 4 -c 2 -f -o gpurun_out/r2b_trace_product \
    python tools/trace_probe.py --frames 3 > gpurun_out/r2b_ncu_trace.log 2>&1
ncu -i gpurun_out/r2b_trace_product.ncu-rep --page raw --csv > gpurun_out/r2b_trace_product_raw.csv 2>/dev/null
# launch list of ONE warm cfg3 batch (bracketed by cudaProfilerStart/Stop) with its DRAM bytes
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2b_launch_list_edit_cfg3.csv python tools/edit_probe.py --reps 1 > gpurun_out/r2b_ncu_edit.log 2>&1
ls -la gpurun_out | tail -12
