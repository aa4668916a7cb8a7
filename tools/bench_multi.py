#!/usr/bin/env python
"""tools/bench_multi.py — BASELINE.json configs 4 and 5 on 1..8 GPUs of one node (run under torchrun for N > 1).

  --mode cfg4   2^17 edited DAG (terrain patch 2^15 + 10 000 sphere edits), ONE 7680x4320 frame tile-sharded (64x64,
                round-robin) over the ranks with a replicated pool: STRONG scaling, Mrays/s = frame rays / max-rank time.
  --mode cfg5   interactive loop on the same scene: per frame one r=128 sphere brush (fill/dig alternating) at the
                centre-pixel hit on rank 0, ONE NCCL broadcast of the packed dirty ranges, 3840x2160 trace sharded
                over all ranks; reports median edit / sync / trace / total milliseconds.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 \
      tools/bench_multi.py --mode cfg4 --steps 20
Every rank builds its own replica of the scene (the canonical DAG is identical, pointers are rank-local), so no
scene transfer is needed before the first frame.  One JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vkhashdag_b200 as v  # noqa: E402
from vkhashdag_b200 import abi, replica  # noqa: E402

BITS17 = [10] * 9 + [16] * 4 + [18] * 3   # bucket bits per node level for the 2^17 scene


def camera(cfg, root, step, W, H, scale, lod=False):
    yaw = 0.6 + 0.37 * step
    pos = ((0.5 + 0.12 * np.sin(0.9 * step)) * scale, (0.62 + 0.03 * np.cos(1.3 * step)) * scale,
           (0.5 + 0.12 * np.cos(0.7 * step)) * scale)
    return abi.camera_params(cfg, root, pos, yaw, -0.5236, W, H, color_root=(1 << 30) | 0x60C0E0, lod=lod)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="cfg4", choices=["cfg4", "cfg5"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--edits", type=int, default=10000)
    ap.add_argument("--level", type=int, default=17)
    ap.add_argument("--patch", type=int, default=15)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    cfg = abi.custom_config(BITS17) if a.level == 17 else abi.default_config(level_count=a.level, bucket_bits_per_bottom_level=17)
    vl = cfg.voxel_level
    ext = a.patch if a.patch < vl else 0
    scale = (1 << (ext or vl)) / (1 << vl)
    pool = v.DAGNodePool(cfg, device=local)
    t0 = time.perf_counter()
    # cfg4 only reads: every rank builds its own replica (same canonical DAG, rank-local pointers).
    # cfg5 edits on rank 0 and ships pointer-valued ranges: the replicas must be COPIES of rank 0's pool, so the
    # scene is built once on rank 0 and published (one broadcast of the whole used pool) before the loop.
    builds_here = a.mode == "cfg4" or rank == 0
    if builds_here:
        root = pool.Edit(abi.NULL, v.TerrainEditor(vl, extent_bits=ext))
        root = pool.EditBatch(root, abi.random_spheres(a.edits, vl, seed=1234, rmin=16, rmax=256, extent_bits=ext))
        assert pool.last_stats["overflow_count"] == 0
        pool.SetRoot(root)
    if a.mode == "cfg5" and dist:
        first_sync_bytes = replica.ReplicaSync(pool, dist, device=f"cuda:{local}").publish(src=0)
        root = pool.GetRoot()
    else:
        pool.DirtyReset()
        first_sync_bytes = 0
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    stream = torch.cuda.ExternalStream(pool.stream, device=local)
    T = 64
    shard = (T, T, rank, world) if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    if a.mode == "cfg4":
        W, H = 7680, 4320
        P0 = camera(cfg, root, 0, W, H, scale)
        n_local = pool.ShardPixels(P0, shard) if shard else W * H
        rgba = torch.zeros(n_local, dtype=torch.int32, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        res = {}
        for lod in (False, True):
            ev = []
            barrier()
            with torch.cuda.stream(stream):
                for s in range(-a.warmup, a.steps):
                    P = camera(cfg, root, s + 100, W, H, scale, lod)
                    flush.fill_(s & 0xFF)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    pool.TraceDev(P, rgba8=rgba.data_ptr(), shard=shard)
                    e1.record()
                    if s >= 0:
                        ev.append((e0, e1))
                torch.cuda.synchronize()
            barrier()
            ms = max_over_ranks(sum(x.elapsed_time(y) for x, y in ev)) / a.steps
            res["lod" if lod else "full"] = (ms, W * H / ms / 1e3)
        if rank == 0:
            print(json.dumps({"metric": "Mrays/s primary-ray traversal @8K (cfg4, strong scaling)", "unit": "Mrays/s",
                              "value": round(res["full"][1], 1), "value_lod": round(res["lod"][1], 1),
                              "ms_per_frame": round(res["full"][0], 4), "ms_per_frame_lod": round(res["lod"][0], 4),
                              "n_gpus": world, "steps": a.steps, "scaling": "strong",
                              "config": {"workload": f"cfg4: 2^{vl} DAG (terrain patch 2^{ext or vl} + {a.edits} sphere edits), 7680x4320 frame, "
                                                     f"64x64 tiles round-robin over {world} GPU(s), replicated pool", "scene_build_s": round(build_s, 3),
                                         "pool_used_MB": round(pool.UsedWords() * 4 / 1e6, 1)}}))
    else:
        W, H = 3840, 2160
        sync = replica.ReplicaSync(pool, dist, device=f"cuda:{local}") if dist else None
        P0 = camera(cfg, root, 0, W, H, scale)
        n_local = pool.ShardPixels(P0, shard) if shard else W * H
        host = torch.zeros(n_local, dtype=torch.int32).pin_memory()
        out = {"rgba8": host.numpy().view(np.uint32)}
        r = 128
        rows = []
        paths = {}
        res_vox = 1 << vl
        for f in range(-a.warmup, a.steps):
            barrier()
            t0 = time.perf_counter()
            if rank == 0:   # pick ray at the centre pixel (main.cpp:320-324), brush at the hit voxel
                cur = pool.GetRoot()
                P1 = camera(cfg, cur, f, 1, 1, scale)
                hp = pool.Traversal(cur, tuple(P1.pos), tuple(P1.look))   # the reference's own pick: Traversal<float>
                c = tuple(int(x * res_vox) for x in hp) if hp is not None else (res_vox // 8, res_vox // 12, res_vox // 8)
                t_pick = time.perf_counter()
                new_root = pool.Edit(cur, v.SphereEditor(c, r * r, "dig" if f & 1 else "fill"))
                assert pool.last_stats["overflow_count"] == 0
                paths[pool.last_stats["path"]] = paths.get(pool.last_stats["path"], 0) + 1
                pool.SetRoot(new_root)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            nbytes = sync.publish(src=0) if sync else 0
            torch.cuda.synchronize()
            if not sync:
                pool.DirtyReset()
            t2 = time.perf_counter()
            P = camera(cfg, pool.GetRoot(), f, W, H, scale, lod=True)
            pool.Trace(P, want=("rgba8",), shard=shard, out=out)
            barrier()
            t3 = time.perf_counter()
            if f >= 0:
                rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, nbytes,
                             (t_pick - t0) * 1e3 if rank == 0 else 0.0))
        rows = np.array(rows)
        # edit and sync are rank 0's own clocks (the replicas sit in the broadcast while rank 0 edits, so their
        # "sync" interval would contain the edit); trace and total are the max over ranks
        med = [max_over_ranks(float(np.median(rows[:, i])) if (rank == 0 or i >= 2) else 0.0) for i in range(4)]
        if rank == 0:
            print(json.dumps({"metric": "interactive loop latency (cfg5)", "unit": "ms", "n_gpus": world, "frames": a.steps,
                              "edit_ms": round(med[0], 3), "of_which_pick_ray_ms": round(float(np.median(rows[:, 5])), 3), "sync_ms": round(med[1], 3), "trace_ms": round(med[2], 3),
                              "total_ms": round(med[3], 3), "edit_paths": paths, "sync_KB_median": round(float(np.median(rows[:, 4])) / 1e3, 1),
                              "config": {"workload": f"cfg5: 2^{vl} DAG, per frame one r={r} sphere brush at the centre-pixel hit on GPU0, "
                                                     f"one NCCL broadcast of the dirty ranges, 3840x2160 LOD trace + host read-back sharded "
                                                     f"over {world} GPU(s)", "scene_build_s": round(build_s, 3),
                                         "initial_replica_sync_MB": round(first_sync_bytes / 1e6, 1),
                                         "per_frame_sync_bytes": [int(x) for x in rows[:8, 4]]}}))
    pool.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
