#!/usr/bin/env python
"""tools/strong_probe.py — BASELINE config 4 in isolation (run under torchrun): ONE 7680x4320 frame on the cfg3 scene, tile-
sharded over the ranks, for several tile sizes; device-timed, max over ranks.  Every rank builds its own replica (no sync)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vkhashdag_b200 as v  # noqa: E402
from vkhashdag_b200 import abi  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = bench.cfg3_config()
    vl = cfg.voxel_level
    scale = (1 << bench.CFG3_PATCH_BITS) / (1 << vl)
    pool = v.DAGNodePool(cfg, device=local)
    root = pool.Edit(abi.NULL, v.TerrainEditor(vl, extent_bits=bench.CFG3_PATCH_BITS))
    root = pool.EditBatch(root, abi.edit_array(abi.random_spheres(bench.EDIT_BATCH, vl, seed=1234, rmin=16, rmax=256,
                                                                   extent_bits=bench.CFG3_PATCH_BITS)))
    stream = torch.cuda.ExternalStream(pool.stream, device=local)
    W, H = bench.W8K, bench.H8K
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    tiles = [int(t) for t in (sys.argv[1:] or ["32", "64", "128", "256"])]
    for T in tiles:
        th = T if T < 256 else 256
        shard = (T, th, rank, world) if world > 1 else None
        P0 = bench.camera(cfg, root, 0, W, H, False, scale=scale)
        n_local = pool.ShardPixels(P0, shard) if shard else W * H
        buf = torch.zeros(n_local, dtype=torch.int32, device="cuda")
        for lod in (False, True):
            ev = []
            with torch.cuda.stream(stream):
                for s in range(-3, 10):
                    P = bench.camera(cfg, root, s + 100, W, H, lod, scale=scale)
                    flush.fill_(s & 0xFF)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    pool.TraceDev(P, rgba8=buf.data_ptr(), shard=shard)
                    e1.record()
                    if s >= 0:
                        ev.append((e0, e1))
                torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in ev) / 10
            t = torch.tensor([ms, -ms], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[f"tile{T}_{'lod' if lod else 'full'}"] = {"ms_max": round(float(t[0]), 4), "ms_min": round(-float(t[1]), 4),
                                                          "mrays_s": round(W * H / float(t[0]) / 1e3, 1)}
    if rank == 0:
        print(json.dumps({"n_gpus": world, **out}))
    pool.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
