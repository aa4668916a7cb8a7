#!/usr/bin/env python
"""tools/simt_model.py — drive tools/simt_model.cpp: estimate warp instructions per ray of different loop organisations of
the trace kernel on the cfg2 scene, from real rays replayed on the CPU (design tool; needs no GPU).

  python tools/simt_model.py [--level 15] [--frames 3] [--cta-step 97]

The scene is built with the compiled reference when available (oracle/_ref), else the oracle port, and cached under /tmp.
"""
import argparse
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402  (tools/ are not product code)
from vkhashdag_b200 import abi  # noqa: E402
import bench  # noqa: E402


def scene(level):
    cfg = abi.default_config(level_count=level, top_level_count=9, bucket_bits_per_bottom_level=17 if level >= 15 else 16)
    cache = f"/tmp/simt_scene_{level}.npz"
    host = (B.Ref() if B.Ref.available() else B.Oracle()).pool(cfg)
    if os.path.exists(cache):
        z = np.load(cache)
        offs, cnts, data = z["offs"], z["cnts"], z["data"]
        p = 0
        for o, c in zip(offs, cnts):
            host.words_np(int(o), int(c))[:] = data[p:p + c]
            p += c
        return cfg, host, int(z["root"])
    if B.Ref.available():
        root = host.edit(B.NULL, B.terrain(cfg.voxel_level), threads=os.cpu_count(), max_task_level=10)
    else:
        root = host.edit(B.NULL, B.terrain(cfg.voxel_level))
    r = host.used_ranges()
    np.savez(cache, offs=np.array([o for o, _ in r], np.int64), cnts=np.array([c for _, c in r], np.int64),
             data=np.concatenate([host.words_np(o, c) for o, c in r]), root=root)
    return cfg, host, root


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=15)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--cta-step", type=int, default=97)
    ap.add_argument("--breakdown", action="store_true")
    ap.add_argument("--hist", action="store_true", help="only print the per-level census of PUSH / ADVANCE / POP transitions")
    a = ap.parse_args()
    so = os.path.join(ROOT, "tools", "_simt_model.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "simt_model.cpp")])
    L = C.CDLL(so)
    L.simt_model.argtypes = [C.POINTER(C.c_uint32), C.POINTER(abi.HdTraceParams), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                             C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    cfg, host, root = scene(a.level)
    bench.LEVEL_COUNT = a.level
    if a.hist:
        for f in range(a.frames):
            P = bench.camera(cfg, root, f * 7, 3840, 2160, False)
            out = (C.c_uint64 * 32)()
            L.simt_model(host.words_ptr, C.byref(P), 0, 0, 1, a.cta_step, None, out)
        h = (C.c_uint64 * (24 * 3))()
        L.simt_hist(h, 1)
        h = np.array(h[:], np.uint64).reshape(24, 3).astype(np.float64)
        tot = h.sum()
        print("level (0 = root's children)  push%  advance%  pop-to%   cumulative share of all transitions")
        cum = 0.0
        for scale in range(22, 22 - a.level - 1, -1):
            cum += h[scale].sum()
            print(f"  level {22 - scale:2d}   {100 * h[scale, 0] / tot:6.2f} {100 * h[scale, 1] / tot:6.2f} {100 * h[scale, 2] / tot:6.2f}    {100 * cum / tot:6.2f}")
        return
    rows = []
    for name, policy, refill_min, patches in (("baseline (one transition per trip)", 0, 0, 1),
                                              ("two-phase (advance until PUSH/POP)", 1, 0, 1),
                                              ("baseline + refill>=16 idle, 8 patches/warp", 2, 16, 8),
                                              ("two-phase + refill>=16 idle, 8 patches/warp", 3, 16, 8),
                                              ("baseline + deferred POP (>= 4 lanes)", 8, 4, 1),
                                              ("baseline + deferred POP (>= 8 lanes)", 8, 8, 1),
                                              ("baseline + deferred POP (>= 12 lanes)", 8, 12, 1),
                                              ("baseline + deferred POP (>= 16 lanes)", 8, 16, 1),
                                              ("greedy region scheduling, pop_min 1", 4, 1, 1),
                                              ("greedy region scheduling, pop_min 6", 4, 6, 1),
                                              ("greedy region scheduling, pop_min 12", 4, 12, 1)):
        tot = np.zeros(32, np.uint64)
        for f in range(a.frames):
            P = bench.camera(cfg, root, f * 7, 3840, 2160, False)
            out = (C.c_uint64 * 32)()
            L.simt_model(host.words_ptr, C.byref(P), policy, refill_min, patches, a.cta_step, None, out)
            tot += np.array(out[:], np.uint64)
        wi, ti, rays, trips, hits = (int(x) for x in tot[:5])
        rows.append((name, wi / rays * 32, ti / wi, trips / (rays / 32), hits / rays))
        print(f"{name:48s} warp-instr/warp-of-32-rays {wi / rays * 32:8.0f}   lanes/instr {ti / wi:5.2f}   trips {trips / (rays / 32):6.1f}"
              f"   hit rate {hits / rays:.3f}", flush=True)
        if a.breakdown:
            for i, rn in enumerate(("loop", "fetch_inner", "fetch_leaf", "fetch_sub", "test", "push", "advance", "pop", "setup")):
                ex, ls, cs = (int(x) for x in tot[5 + 3 * i:8 + 3 * i])
                if ex:
                    print(f"      {rn:12s} executions/warp {ex / (rays / 32):6.1f}  lanes {ls / ex:5.1f}  share of warp instr {cs / wi:6.1%}")
    base = rows[0][1]
    for r in rows:
        print(f"{r[0]:48s} {base / r[1]:.3f}x")


if __name__ == "__main__":
    main()
