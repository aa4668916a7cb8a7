#!/usr/bin/env python
"""tools/edit_probe.py — the cfg3 edit batch of bench.py in isolation (GPU box): wall time per hd_edit_batch call and, under
ncu, the launch list of ONE warm batch (the batch is bracketed by cudaProfilerStart/Stop):

  python tools/edit_probe.py [--reps 3] [--brush 0]
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/edit_launches.csv \
      python tools/edit_probe.py --reps 1
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vkhashdag_b200 as v  # noqa: E402
from vkhashdag_b200 import abi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--mid", type=int, default=0, help="also time batches of this many editors (mid-size batches)")
    ap.add_argument("--l14", type=int, default=0, help="override the bucket bits of node level 14")
    ap.add_argument("--color", type=int, default=0, help="also time this many coloured r=128 brushes like bench.py's color_brush leg")
    a = ap.parse_args()
    if a.l14:
        bench.CFG3_BUCKET_BITS[14] = a.l14
    cfg = bench.cfg3_config()
    vl = cfg.voxel_level
    pool = v.DAGNodePool(cfg, device=0)
    terrain = v.TerrainEditor(vl, extent_bits=bench.CFG3_PATCH_BITS)
    spheres = abi.random_spheres(bench.EDIT_BATCH, vl, seed=1234, rmin=16, rmax=256, extent_bits=bench.CFG3_PATCH_BITS)
    arr = abi.edit_array(spheres)
    pool.EditBatch(pool.Edit(abi.NULL, terrain), arr)   # rehearsal (allocator pool growth)
    times, tt = [], []
    for r in range(a.reps):
        pool.Clear()
        t0 = time.perf_counter()
        root = pool.Edit(abi.NULL, terrain)
        tt.append(time.perf_counter() - t0)
        pool.Sync()
        last = r == a.reps - 1
        if last:
            torch.cuda.profiler.start()
        t0 = time.perf_counter()
        root_b = pool.EditBatch(root, arr)
        times.append(time.perf_counter() - t0)
        if last:
            torch.cuda.profiler.stop()
        assert pool.last_stats["overflow_count"] == 0
    bw = pool.ReadBucketWords()
    bases = cfg.level_bases() + [len(bw)]
    out = {"max_bucket_words_per_level": [int(bw[bases[l]:bases[l + 1]].max()) for l in range(cfg.node_levels)],
           "batch_s": [round(t, 4) for t in times], "terrain_s": [round(t, 4) for t in tt], "stats": pool.last_stats}
    if a.mid:
        ms = []
        rb = root_b
        for k in range(12):
            sub = abi.edit_array(abi.random_spheres(a.mid, vl, seed=900 + k, rmin=16, rmax=128, extent_bits=bench.CFG3_PATCH_BITS))
            t0 = time.perf_counter()
            rb = pool.EditBatch(rb, sub)
            ms.append((time.perf_counter() - t0) * 1e3)
        out["mid"] = {"editors": a.mid, "ms_median": round(float(np.median(ms[2:])), 4), "path": pool.last_stats["path"]}
    if a.color:
        pool.SetRoot(root_b)
        pool.ColorConfig(bench.COLOR_LEAF_LEVEL)
        res3 = 1 << vl
        pool.EditColor(root_b, abi.sphere((res3 // 2,) * 3, 3 * res3 * res3), 0x60C0E0, paint=True)   # base coat
        scale = (1 << bench.CFG3_PATCH_BITS) / float(res3)
        g = pool.Trace(bench.camera(cfg, root_b, 5, 96, 54, False, scale=scale), want=("hits",))["hits"].reshape(-1)
        g = g[(g["packed"] >> 31) != 0]
        picks = g[np.linspace(0, len(g) - 1, a.color).astype(int)]
        palette = [0xE04020, 0x20A040, 0x3060E0, 0xE0C020, 0xA040C0]
        broot, ms, kinds = root_b, [], []
        torch.cuda.profiler.start()
        for i, h in enumerate(picks):
            d = abi.sphere(tuple(int(c) for c in h["vox"]), 128 * 128)
            t0 = time.perf_counter()
            broot, _ = pool.EditColor(broot, d, palette[i % 5], i % 3 == 2)
            ms.append(round((time.perf_counter() - t0) * 1e3, 4))
            kinds.append("paint" if i % 3 == 2 else "fill")
        torch.cuda.profiler.stop()
        out["color"] = {"ms": ms, "kinds": kinds, "median_fill": float(np.median([m for m, k in zip(ms, kinds) if k == "fill"][2:])),
                        "median_paint": float(np.median([m for m, k in zip(ms, kinds) if k == "paint"][1:])),
                        "median_all": float(np.median(ms[4:]))}
    print(json.dumps(out))
    pool.close()


if __name__ == "__main__":
    main()
