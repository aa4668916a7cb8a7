set -x
timeout 1200 python -m pytest tests/test_gpu_edit.py tests/test_gpu_color_edit.py tests/test_cpp_host.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/bench_brush.py --edits 60 --cpu-sample 20 --radii 2,32,128,256 2>&1 | tail -1
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; tail -c 300 gpurun_out/r2b_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['value_lod'], d['e2e']['value'], d['roofline']['frac'], d['parity']['cfg2'], d['parity']['cfg3_batch'], d['parity']['trace_rows'])
print(d['edit']['batch']['seconds'], d['edit']['color_brush'])
print(d['interactive'])
PY
