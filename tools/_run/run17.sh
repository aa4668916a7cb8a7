set -x
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n8.json 2> gpurun_out/r2b_bench_n8.err; tail -c 300 gpurun_out/r2b_bench_n8.err; tail -c 1500 gpurun_out/r2b_bench_n8.json
