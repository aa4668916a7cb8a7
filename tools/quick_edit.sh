#!/bin/bash
# quick parity + speed check of the edit kernels on the GPU box (used while optimising)
python -m pytest tests/test_gpu_edit.py tests/test_cpp_host.py -m gpu -x -q 2>&1 | tail -2
BB=10,10,10,10,10,10,10,10,10,16,16,16,16,18,18,18
python tools/bench_edit.py --verify --cpu-sample 50 --bucket-bits $BB 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch s', d['batch_seconds'], 'terrain s', d['terrain_build_seconds'], 'value', d['value'], 'parity', d['parity'], 'x cpu', d['speedup_vs_cpu_per_edit'])"
