#!/usr/bin/env python
"""tools/bench_brush.py — latency of ONE brush edit per call (the reference's interactive use: src/main.cpp:214-238,321).

  python tools/bench_brush.py [--level 17] [--patch 15] [--radii 2,32,128,256] [--edits 200] [--cpu-sample 30]

Scene: the cfg3 terrain patch in a 2^level world.  For every radius, `edits` sphere brushes (fill/dig alternating, centres
on the terrain surface) are applied one hd_edit_batch call each; the median wall time per call is reported next to the
reference's ThreadedEdit on the same brushes (all host cores, max_task_level 10) and the path that served the calls.
HD_EDIT_FAST=0 forces the general path (for the comparison in DESIGN.md).  One JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=17)
    ap.add_argument("--patch", type=int, default=15)
    ap.add_argument("--radii", default="2,32,128,256")
    ap.add_argument("--edits", type=int, default=200)
    ap.add_argument("--cpu-sample", type=int, default=30)
    a = ap.parse_args()
    import vkhashdag_b200 as v
    from oracle import bindings as B
    from vkhashdag_b200 import abi

    bits17 = [10] * 9 + [16] * 4 + [18] * 3
    cfg = abi.custom_config(bits17) if a.level == 17 else abi.default_config(level_count=a.level, bucket_bits_per_bottom_level=17)
    vl = cfg.voxel_level
    ext = a.patch if a.patch < vl else 0
    terrain = abi.terrain(vl, extent_bits=ext)
    O = B.Oracle()
    pool = v.DAGNodePool(cfg)
    root = pool.Edit(abi.NULL, terrain)
    assert pool.last_stats["overflow_count"] == 0
    span = 1 << (ext or vl)
    rng = np.random.default_rng(5)
    cores = os.cpu_count() or 1
    kind = "reference" if B.Ref.available() else "port"
    host = hroot = None
    if a.cpu_sample:
        host = (B.Ref() if kind == "reference" else O).pool(cfg)
        hroot = host.edit(B.NULL, terrain, threads=cores, max_task_level=10) if kind == "reference" else host.edit(B.NULL, terrain)
    out = []
    for r in (int(x) for x in a.radii.split(",")):
        brushes = []
        for i in range(a.edits):
            x, z = (int(t) for t in rng.integers(span // 8, 7 * span // 8, 2))
            brushes.append(abi.sphere((x, O.terrain_height(terrain, x, z), z), r * r, dig=bool(i & 1)))
        arrs = [abi.edit_array([b]) for b in brushes]
        ms, paths, leaves = [], {}, 0
        for arr in arrs:
            t = time.perf_counter()
            root = pool.EditBatch(root, arr)
            ms.append((time.perf_counter() - t) * 1e3)
            st = pool.last_stats
            assert st["overflow_count"] == 0
            paths[st["path"]] = paths.get(st["path"], 0) + 1
            leaves += st["visited_leaves"]
        row = {"radius": r, "gpu_ms_median": round(float(np.median(ms)), 4), "gpu_ms_p90": round(float(np.percentile(ms, 90)), 4),
               "paths": paths, "leaves_per_edit": leaves // a.edits}
        if host is not None:
            n = min(a.cpu_sample, a.edits)
            cms = []
            for b in brushes[:n]:
                t = time.perf_counter()
                hroot = host.edit(hroot, b, threads=cores, max_task_level=10) if kind == "reference" else host.edit(hroot, b)
                cms.append((time.perf_counter() - t) * 1e3)
            row["cpu_ms_median"] = round(float(np.median(cms)), 4)
            row["speedup"] = round(row["cpu_ms_median"] / row["gpu_ms_median"], 2)
        out.append(row)
    print(json.dumps({"metric": "brush edit latency (one editor per call)", "unit": "ms", "world": f"2^{vl}", "rows": out,
                      "cpu": {"kind": kind, "cores": cores, "what": "ThreadedEdit(busy_pool(cores), max_task_level=10)" if kind == "reference" else "serial Edit port"},
                      "fast_path_enabled": os.environ.get("HD_EDIT_FAST", "1") != "0"}))
    pool.close()


if __name__ == "__main__":
    main()
