#!/usr/bin/env python
"""tools/bench_gc.py — garbage collection (row N1) on the cfg3 scene: 2^17 world, 2^15 terrain patch, 10 000 sphere
edits applied in `--batches` batches (every batch leaves the previous upper-level nodes behind as garbage).

GPU hd_gc vs the reference's NodePoolThreadedGC::ThreadedGC (all host cores) on a mirror of the same pool.  Reports
time, words before/after and the canonical-DAG check (voxel count and per-level census unchanged by either GC).
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
    ap.add_argument("--edits", type=int, default=10000)
    ap.add_argument("--batches", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import vkhashdag_b200 as v
    from oracle import bindings as B
    from vkhashdag_b200 import abi
    cfg = abi.custom_config([10] * 9 + [16] * 4 + [18] * 3)
    vl, ext = cfg.voxel_level, 15
    pool = v.DAGNodePool(cfg)
    root = pool.Edit(abi.NULL, abi.terrain(vl, extent_bits=ext))
    spheres = abi.random_spheres(a.edits, vl, seed=1234, rmin=16, rmax=256, extent_bits=ext)
    per = max(1, a.edits // a.batches)
    for i in range(0, a.edits, per):
        root = pool.EditBatch(root, abi.edit_array(spheres[i:i + per]))
        assert pool.last_stats["overflow_count"] == 0
    used0 = pool.UsedWords()
    cpu = None
    if not a.no_cpu:
        cores = os.cpu_count() or 1
        kind = "reference" if B.Ref.available() else None
        if kind:
            host = B.Ref().pool(cfg)
            ranges, bw = pool.Download()
            for off, words in ranges.items():
                host.words_np(off, len(words))[:] = words
            host.bucket_words_np()[:] = bw
            O = B.Oracle()
            before = O.canonical(host.words_ptr, cfg.node_levels, root)
            t = time.perf_counter()
            hroot = host.gc(root, threads=cores)
            dt = time.perf_counter() - t
            after = O.canonical(host.words_ptr, cfg.node_levels, hroot)
            cpu = {"kind": kind, "cores": cores, "seconds": round(dt, 4), "words_after": int(host.bucket_words_np().sum()),
                   "dag_unchanged": before["hash"] == after["hash"], "nodes": after["by_ptr"]}
    pool.Sync()
    t = time.perf_counter()
    new_root = pool.ThreadedGC(root)
    dt = time.perf_counter() - t
    used1 = pool.UsedWords()
    out = {"metric": "garbage collection (hd_gc)", "unit": "s", "gpu_seconds": round(dt, 4), "words_before": used0, "words_after": used1,
           "kept_nodes": pool.last_gc_nodes, "cpu_baseline": cpu,
           "config": {"workload": f"cfg3 scene: 2^{vl} world, 2^{ext} terrain patch, {a.edits} sphere edits in {a.batches} batches"}}
    if cpu:
        out["speedup"] = round(cpu["seconds"] / dt, 1)
        # the GPU-compacted pool holds the same canonical DAG as before (and as the reference's compacted pool)
        ranges, bw = pool.Download()
        host.words_np(0, host.total_words)[:] = 0
        for off, words in ranges.items():
            host.words_np(off, len(words))[:] = words
        got = O.canonical(host.words_ptr, cfg.node_levels, new_root)
        out["gpu_dag_unchanged"] = got["hash"] == before["hash"] and got["by_ptr"] == got["by_content"] == before["by_ptr"]
    # the DAG after the GPU GC still holds the same scene: the next edit continues from it
    r2 = pool.EditBatch(new_root, abi.edit_array(spheres[:8]))
    out["edit_after_gc_path"] = pool.last_stats["path"]
    print(json.dumps(out))
    pool.close()


if __name__ == "__main__":
    main()
