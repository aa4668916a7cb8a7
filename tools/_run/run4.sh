set -x
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_edit_fused -s 8 -c 1 -f -o gpurun_out/r2c_fused_mid100 python tools/edit_probe.py --reps 1 --mid 100 > gpurun_out/r2c_fused_mid100.log 2>&1
tail -2 gpurun_out/r2c_fused_mid100.log
