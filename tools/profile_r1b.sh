#!/bin/bash
# second profiling pass of round 1 (run on the GPU box): full bench line, reference arm, launch lists and ncu --set full
# captures of the new edit kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
tail -c 600 gpurun_out/bench_r1b.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1b_reference.json 2>> gpurun_out/bench_r1b.err
BB=10,10,10,10,10,10,10,10,10,16,16,16,16,18,18,18
# launch list of the cfg3 batch (terrain build + one 10k-sphere batch)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b_edit.csv \
    python tools/bench_edit.py --cpu-sample 1 --bucket-bits $BB > gpurun_out/ncu_edit_r1b.log 2>&1
# launch list of the brush loop (fused kernel, one launch per edit)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b_brush.csv \
    python tools/bench_brush.py --radii 2,32,128,256 --edits 12 --cpu-sample 0 > gpurun_out/ncu_brush_r1b.log 2>&1
# full captures: fused brush kernel (r = 128), grouped upsert + leaf + down kernels of the batch
ncu --set full --clock-control none --import-source on -k regex:k_edit_fused -s 8 -c 2 -o gpurun_out/brush_r1b \
    python tools/bench_brush.py --radii 128 --edits 12 --cpu-sample 0 >> gpurun_out/ncu_brush_r1b.log 2>&1
# the big launches of the batch, one capture each (kernel base names; -s skips the terrain build's launches of the same kernel)
E="python tools/bench_edit.py --cpu-sample 1 --bucket-bits $BB"
ncu --set full --clock-control none --import-source on -k k_leaf_half -c 1 -o gpurun_out/edit_r1b_leaf $E >> gpurun_out/ncu_edit_r1b.log 2>&1
ncu --set full --clock-control none --import-source on -k k_leaf -c 1 -o gpurun_out/edit_r1b_leaf_terrain $E >> gpurun_out/ncu_edit_r1b.log 2>&1
ncu --set full --clock-control none --import-source on -k k_upsert_grouped -s 7 -c 3 -o gpurun_out/edit_r1b_grouped $E >> gpurun_out/ncu_edit_r1b.log 2>&1
ncu --set full --clock-control none --import-source on -k k_down -s 28 -c 2 -o gpurun_out/edit_r1b_down $E >> gpurun_out/ncu_edit_r1b.log 2>&1
for f in brush_r1b edit_r1b_leaf edit_r1b_leaf_terrain edit_r1b_grouped edit_r1b_down; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -20
