set -x
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_more.py tests/test_gpu_gc_io.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; tail -c 300 gpurun_out/r2b_bench_n2.err; tail -c 3000 gpurun_out/r2b_bench_n2.json
