#!/usr/bin/env python
"""tools/trace_probe.py — quick A/B of trace-kernel variants on the cfg2 scene (GPU box; also the ncu target).

  HD_TRACE_VARIANT=1 python tools/trace_probe.py --frames 20          device-timed Mrays/s, full detail and LOD
  ncu --set full --import-source on -k regex:trace_kernel -s 4 -c 1 -o gpurun_out/x python tools/trace_probe.py --frames 2

Checks every variant's frames against variant-independent checksums printed with the numbers (same scene, same cameras),
so two runs can be compared by eye: equal checksums = identical pixels.
"""
import argparse
import json
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vkhashdag_b200 as v  # noqa: E402
from vkhashdag_b200 import abi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--level", type=int, default=15)
    ap.add_argument("--lod", action="store_true")
    ap.add_argument("--ab", default="", help="comma list of variant:cta pairs measured round-robin in ONE process, "
                                             "e.g. 0:128,0:64,0:256,5:128 (HD_TRACE_VARIANT / HD_TRACE_CTA are read per call)")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--scene", default="cfg2", help="cfg2 (2^15 terrain) or cfg3 (2^17 world, terrain patch + 10 000 sphere edits)")
    ap.add_argument("--tiled", action="store_true", help="trace through the tile-shard kernels (64x64 tiles, world 1)")
    ap.add_argument("--size", default="4k", help="4k or 8k frame")
    a = ap.parse_args()
    bench.LEVEL_COUNT = a.level
    scale = 1.0
    if a.scene == "cfg3":
        cfg = bench.cfg3_config()
        scale = (1 << bench.CFG3_PATCH_BITS) / (1 << cfg.voxel_level)
        pool = v.DAGNodePool(cfg, device=0)
        root = pool.Edit(abi.NULL, v.TerrainEditor(cfg.voxel_level, extent_bits=bench.CFG3_PATCH_BITS))
        root = pool.EditBatch(root, abi.edit_array(abi.random_spheres(bench.EDIT_BATCH, cfg.voxel_level, seed=1234, rmin=16, rmax=256,
                                                                     extent_bits=bench.CFG3_PATCH_BITS)))
    else:
        cfg = bench.scene_config()
        pool = v.DAGNodePool(cfg, device=0)
        root = pool.Edit(abi.NULL, v.TerrainEditor(cfg.voxel_level))
    assert pool.last_stats["overflow_count"] == 0
    W, H = (bench.W4K, bench.H4K) if a.size == "4k" else (2 * bench.W4K, 2 * bench.H4K)
    a.cam_scale, a.shard = scale, ((64, 64, 0, 1) if a.tiled else None)
    stream = torch.cuda.ExternalStream(pool.stream, device=0)
    rgba = torch.zeros(W * H, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if a.ab:
        ab(a, pool, cfg, root, W, H, stream, rgba, flush)
        pool.close()
        return
    out = {}
    for lod in ((False, True) if a.lod else (False,)):
        ev, crc = [], 0
        with torch.cuda.stream(stream):
            for s in range(-3, a.frames):
                P = bench.camera(cfg, root, s + 1000 * lod, W, H, lod)
                flush.fill_(s & 0xFF)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pool.TraceDev(P, rgba8=rgba.data_ptr())
                e1.record()
                if s >= 0:
                    ev.append((e0, e1))
                if s in (0, a.frames - 1):
                    pool.Sync()
                    crc = zlib.crc32(rgba.cpu().numpy().tobytes(), crc)
            torch.cuda.synchronize()
        ms = [x.elapsed_time(y) for x, y in ev]
        out["lod" if lod else "full"] = {"Mrays_s": round(W * H * len(ms) / sum(ms) / 1e3, 1), "ms_min": round(min(ms), 4),
                                         "ms_max": round(max(ms), 4), "crc": crc}
    print(json.dumps({"variant": os.environ.get("HD_TRACE_VARIANT", "0"), "level": a.level, **out}))
    pool.close()


def ab(a, pool, cfg, root, W, H, stream, rgba, flush):
    """Round-robin A/B of (variant, CTA size) pairs inside one process: same box, same clocks, same scene."""
    pairs = [tuple((x.split(":") + ["0"])[:3]) for x in a.ab.split(",")]  # variant : CTA threads [: persistent CTA threads]
    res = {p: {"full": [], "lod": [], "crc": {}} for p in pairs}
    with torch.cuda.stream(stream):
        for rnd in range(a.rounds):
            for p in pairs:
                os.environ["HD_TRACE_VARIANT"], os.environ["HD_TRACE_CTA"], os.environ["HD_TRACE_PERSIST"] = p
                for lod in ((False, True) if a.lod else (False,)):
                    if lod:  # the LOD default is the hoisted loop: variant 0 -> 4, 5 -> 6
                        os.environ["HD_TRACE_VARIANT"] = {"0": "4", "5": "6"}.get(p[0], p[0])
                    ev, crc = [], 0
                    for s in range(-2, a.frames):
                        P = bench.camera(cfg, root, s + 1000 * lod, W, H, lod, scale=a.cam_scale)
                        flush.fill_(s & 0xFF)
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        pool.TraceDev(P, rgba8=rgba.data_ptr(), shard=a.shard)
                        e1.record()
                        if s >= 0:
                            ev.append((e0, e1))
                        if s in (0, a.frames - 1):
                            pool.Sync()
                            crc = zlib.crc32(rgba.cpu().numpy().tobytes(), crc)
                    torch.cuda.synchronize()
                    ms = [x.elapsed_time(y) for x, y in ev]
                    res[p]["lod" if lod else "full"].append(round(W * H * len(ms) / sum(ms) / 1e3, 1))
                    res[p]["crc"]["lod" if lod else "full"] = crc
    for p in pairs:
        print(json.dumps({"variant": p[0], "cta": p[1], "persist": p[2], **res[p]}))


if __name__ == "__main__":
    main()
