#!/usr/bin/env python
"""tools/bench_color_edit.py — coloured brush edits (vbr_edit of src/main.cpp:224-230, the reference's interactive
right-mouse path) on a 2^17 world: GPU hd_edit_color vs the reference's ThreadedEdit(max_task_level = colour leaf level).
One JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edits", type=int, default=40)
    ap.add_argument("--cpu-edits", type=int, default=10)
    ap.add_argument("--radius", type=int, default=128)
    a = ap.parse_args()
    import vkhashdag_b200 as v
    from oracle import bindings as B
    from vkhashdag_b200 import abi
    cfg = abi.custom_config([10] * 9 + [16] * 4 + [18] * 3)
    vl, ext, LL = cfg.voxel_level, 15, 10
    t = abi.terrain(vl, extent_bits=ext)
    O = B.Oracle()
    rng = np.random.default_rng(5)
    brushes = []
    for i in range(a.edits):
        x, z = (int(q) for q in rng.integers(2000, (1 << ext) - 2000, 2))
        y = O.terrain_height(t, x, z)
        brushes.append((abi.sphere((x, y, z), a.radius ** 2), int(rng.integers(0, 1 << 24)), bool(i % 3 == 2)))
    pool = v.DAGNodePool(cfg)
    pool.ColorConfig(LL)
    root = pool.Edit(abi.NULL, t)
    root, _ = pool.EditColor(root, abi.aabb((0, 0, 0), (1 << ext, 1 << 14, 1 << ext)), 0x60A060)   # base coat (untimed)
    times = []
    for d, rgb, paint in brushes:
        t0 = time.perf_counter()
        root, _ = pool.EditColor(root, d, rgb, paint)
        times.append(time.perf_counter() - t0)
        assert pool.last_stats["overflow_count"] == 0
    cn, cl = pool.ReadColor()
    cores = os.cpu_count() or 1
    cpu = None
    if B.Ref.available() and a.cpu_edits:
        R = B.Ref()
        rp, cp = R.pool(cfg), R.color_pool(LL, node_capacity=1 << 22, leaf_word_capacity=1 << 28)
        rr = rp.edit(B.NULL, t, threads=cores, max_task_level=10)
        rr = rp.edit_color(cp, rr, abi.aabb((0, 0, 0), (1 << ext, 1 << 14, 1 << ext)), 0x60A060, threads=cores)
        ct = []
        for d, rgb, paint in brushes[:a.cpu_edits]:
            t0 = time.perf_counter()
            rr = rp.edit_color(cp, rr, d, rgb, paint, threads=cores)
            ct.append(time.perf_counter() - t0)
        cpu = {"kind": "reference", "cores": cores, "edits": len(ct), "ms_per_edit_median": round(float(np.median(ct)) * 1e3, 3),
               "sample": f"first {len(ct)} brushes, ThreadedEdit(busy_pool({cores}), max_task_level={LL}) through VBREditorWrapper"}
    print(json.dumps({"metric": "coloured brush edit latency (vbr_edit)", "unit": "ms", "radius": a.radius, "edits": a.edits,
                      "gpu_ms_per_edit_median": round(float(np.median(times)) * 1e3, 3),
                      "gpu_ms_per_edit_p90": round(float(np.percentile(times, 90)) * 1e3, 3),
                      "color_pool_words": {"nodes": int(cn.size), "leaves": int(cl.size)}, "cpu_baseline": cpu,
                      "config": {"workload": f"2^{vl} world, terrain patch 2^{ext}, colour leaf level {LL} (2^21 voxels per colour leaf), "
                                             f"r={a.radius} sphere fills (2 of 3) and paints (1 of 3) with random colours at the terrain surface"}}))
    pool.close()


if __name__ == "__main__":
    main()
