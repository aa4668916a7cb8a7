#!/usr/bin/env python
"""tools/bench_edit.py — secondary metric of BASELINE.json: edited voxels/s of the batched GPU edit rebuild
(config 3: 2^17 world, noise-terrain patch + 10 000 random sphere fill/dig edits, alternating, applied in index order).

  python tools/bench_edit.py [--edits 10000] [--level 17] [--patch 15] [--cpu-sample 300] [--verify]

Prints one JSON line: GPU batch time, in-range voxels/s (implementation-independent unit), the reference's
CPU ThreadedEdit timed on a bounded sample of the same edit list (sequential, all host cores), and a parity check of
sampled voxels against the analytic definition of the scene (terrain height + ordered sphere fold).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def analytic_voxels(O, terrain, spheres, pts):
    """Ground truth by definition: terrain column test, then the ordered fill/dig fold."""
    out = np.zeros(len(pts), bool)
    ext = terrain.p1[1]
    c = np.array([[s.p0[0], s.p0[1], s.p0[2]] for s in spheres], np.int64)
    r2 = np.array([s.r2 for s in spheres], np.int64)
    dig = np.array([s.kind == 2 for s in spheres])
    for i, (x, y, z) in enumerate(pts):
        inside = ext == 0 or ((x >> ext) == 0 and (z >> ext) == 0)
        v = bool(inside and y < O.terrain_height(terrain, int(x), int(z)))
        d = c - np.array([x, y, z], np.int64)
        hit = np.nonzero((d * d).sum(1) <= r2)[0]
        for j in hit:
            v = not dig[j]
        out[i] = v
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edits", type=int, default=10000)
    ap.add_argument("--level", type=int, default=17)
    ap.add_argument("--patch", type=int, default=15)
    ap.add_argument("--cpu-sample", type=int, default=300)
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--bucket-bits", default="", help="comma-separated bucket bits per node level (default: DefaultConfig, bottom 17)")
    ap.add_argument("--usage", action="store_true", help="print per-level pool usage")
    ap.add_argument("--sequential", action="store_true", help="also time one hd_edit_batch call per edit")
    a = ap.parse_args()
    import vkhashdag_b200 as v
    from oracle import bindings as B
    from vkhashdag_b200 import abi

    bits = [int(b) for b in a.bucket_bits.split(",")] if a.bucket_bits else None
    if bits:
        assert len(bits) == a.level - 1
        cfg = abi.custom_config(bits)
    else:
        cfg = abi.default_config(level_count=a.level, top_level_count=9, bucket_bits_per_bottom_level=17)
    vl = cfg.voxel_level
    ext = a.patch if a.patch < vl else 0
    terrain = abi.terrain(vl, extent_bits=ext)
    spheres = abi.random_spheres(a.edits, vl, seed=1234, rmin=16, rmax=256, extent_bits=ext)
    O = B.Oracle()
    pool = v.DAGNodePool(cfg)
    t = time.perf_counter()
    root0 = pool.Edit(abi.NULL, terrain)
    t_terrain = time.perf_counter() - t
    st_terrain = dict(pool.last_stats)
    assert st_terrain["overflow_count"] == 0

    sphere_arr = abi.edit_array(spheres)   # descriptor marshalling is host-side set-up, not part of the batch
    t = time.perf_counter()
    root = pool.EditBatch(root0, sphere_arr)
    t_batch = time.perf_counter() - t
    st = dict(pool.last_stats)
    if a.usage:
        bwv = pool.ReadBucketWords()
        bases = cfg.level_bases() + [len(bwv)]
        for l in range(cfg.node_levels):
            seg = bwv[bases[l]:bases[l + 1]]
            print(f"level {l}: buckets 2^{cfg.bucket_bits_each_level[l]} used {int(seg.sum()) / 1e6:.1f} M words, max bucket {int(seg.max())}", file=sys.stderr)
    assert st["overflow_count"] == 0, st
    in_range = sum(O.in_range_voxels(s, vl) for s in spheres)

    seq = None
    if a.sequential:
        pool2 = v.DAGNodePool(cfg)
        r2 = pool2.Edit(abi.NULL, terrain)
        n = min(a.edits, 2000)
        t = time.perf_counter()
        for s in spheres[:n]:
            r2 = pool2.Edit(r2, s)
        seq = {"edits": n, "seconds": round(time.perf_counter() - t, 4)}
        pool2.close()

    # CPU baseline: the reference's ThreadedEdit (or the port) on the first `cpu_sample` edits, sequentially
    cores = os.cpu_count() or 1
    kind = "reference" if B.Ref.available() else "port"
    host = (B.Ref() if kind == "reference" else O).pool(cfg)
    ranges, bw = pool.Download() if False else (None, None)
    # scene for the CPU: mirror the GPU-built terrain (state before the spheres) is not available any more ->
    # rebuild the terrain with the CPU editor itself (untimed set-up)
    t = time.perf_counter()
    hr = host.edit(B.NULL, terrain, threads=cores, max_task_level=10) if kind == "reference" else host.edit(B.NULL, terrain)
    t_cpu_terrain = time.perf_counter() - t
    n_cpu = min(a.cpu_sample, a.edits)
    t = time.perf_counter()
    for s in spheres[:n_cpu]:
        hr = host.edit(hr, s, threads=cores, max_task_level=10) if kind == "reference" else host.edit(hr, s)
    t_cpu = time.perf_counter() - t
    in_range_cpu = sum(O.in_range_voxels(s, vl) for s in spheres[:n_cpu])

    parity = None
    if a.verify:
        # sampled voxels near sphere surfaces and terrain surface vs the analytic definition
        rng = np.random.default_rng(7)
        pts = []
        for s in rng.choice(len(spheres), 400, replace=False):
            c, r = np.array(spheres[s].p0[:], np.int64), int(np.sqrt(spheres[s].r2))
            for _ in range(6):
                d = rng.normal(size=3)
                d /= np.linalg.norm(d)
                p = c + np.round(d * (r + rng.integers(-2, 3))).astype(np.int64)
                if (p >= 0).all() and (p < (1 << vl)).all():
                    pts.append(p)
        pts = np.array(pts)
        ranges, bw = pool.Download()
        m = O.pool(cfg)
        for off, words in ranges.items():
            m.words_np(off, len(words))[:] = words
        exp = analytic_voxels(O, terrain, spheres, pts)
        got = np.array([O.voxel_get(m.words_ptr, cfg.node_levels, root, int(x), int(y), int(z)) for x, y, z in pts])
        parity = {"sampled_voxels": int(len(pts)), "mismatches": int((exp != got).sum()), "set_fraction": float(exp.mean())}

    print(json.dumps({
        "metric": "edited voxels/s per batch (in-range voxels)", "unit": "voxels/s",
        "value": round(in_range / t_batch), "batch_seconds": round(t_batch, 4), "edits": a.edits,
        "config": {"workload": f"cfg3: 2^{vl} world, terrain patch 2^{ext or vl}, {a.edits} random spheres r in [16,256] alternating fill/dig, one batch",
                   "pool": f"DefaultConfig(level_count={a.level}, bottom bucket bits 17)"},
        "in_range_voxels": int(in_range), "leaf_voxels_per_s": round(st["visited_leaves"] * 64 / t_batch),
        "stats": st, "terrain_build_seconds": round(t_terrain, 4), "terrain_stats": st_terrain,
        "gpu_sequential": seq,
        "cpu_baseline": {"kind": kind, "cores": cores, "edits": n_cpu, "seconds": round(t_cpu, 3),
                         "value": round(in_range_cpu / t_cpu), "unit": "voxels/s", "ms_per_edit": round(t_cpu / n_cpu * 1e3, 3),
                         "sample": f"first {n_cpu} edits of the same list applied sequentially by "
                                   f"{'ThreadedEdit(busy_pool(%d), max_task_level=10)' % cores if kind == 'reference' else 'the serial port'}; "
                                   f"terrain built by the same CPU editor in {t_cpu_terrain:.1f} s (untimed)"},
        "speedup_vs_cpu_per_edit": round((t_cpu / n_cpu) / (t_batch / a.edits), 1),
        "parity": parity}))
    pool.close()


if __name__ == "__main__":
    main()
