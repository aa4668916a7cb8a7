#!/bin/bash
# quick parity + speed check of the trace kernel on the GPU box (used while optimising)
python -m pytest tests/test_gpu_trace.py tests/test_gpu_more.py -m gpu -x -q -k "trace or colour" 2>&1 | tail -2
python bench.py --no-cpu-baseline --steps 50 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','value_lod','ms_per_step')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])"
