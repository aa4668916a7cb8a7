set -x
HD_EDIT_FAST_TRACE=1 timeout 900 python tools/edit_probe.py --reps 1 --color 24 > gpurun_out/r2g_color_cfg3.log 2>&1; tail -1 gpurun_out/r2g_color_cfg3.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['color'])"
grep "fused edit" -A1 gpurun_out/r2g_color_cfg3.log | tail -4
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_color_cfg3_launches.csv python tools/edit_probe.py --reps 1 --color 9 > /dev/null 2>&1
