set -x
timeout 1200 python -m pytest tests/test_gpu_edit.py tests/test_gpu_gc_io.py tests/test_cpp_host.py tests/test_gpu_more.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/stress_edit.py 2>&1 | tail -2
for m in 1 0; do HD_EDIT_BUCKET_MERGE=$m timeout 600 python tools/edit_probe.py --reps 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['batch_s'], d['terrain_s'], d['stats'])"; done
HD_EDIT_VERIFY=1 timeout 600 python tools/edit_probe.py --reps 1 2>&1 | tail -3 | cut -c1-600
timeout 900 python tools/bench_edit.py --verify --cpu-sample 20 --bucket-bits 10,10,10,10,10,10,10,10,10,16,16,16,16,18,18,18 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch s', d['batch_seconds'], 'terrain s', d['terrain_build_seconds'], 'parity', d['parity'])"
