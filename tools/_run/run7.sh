set -x
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -c 6000 gpurun_out/r2f_bench_n1.json; tail -5 gpurun_out/r2f_bench_n1.err
