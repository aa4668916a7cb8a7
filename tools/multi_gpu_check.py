#!/usr/bin/env python
"""tools/multi_gpu_check.py — run under torchrun (one rank per GPU): replicated pool, tile-sharded trace, edit on rank 0
with ONE NCCL broadcast of the packed dirty ranges per edit batch (BASELINE config 5 in miniature).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/multi_gpu_check.py [--level 13] [--frames 8]

Checks (rank 0 gathers the tile shards): after every edit+sync the stitched N-GPU frame is pixel-identical to a
full-frame trace on rank 0, and every replica holds the same bucket cursors.  Prints per-frame edit/sync/trace ms.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vkhashdag_b200 as v  # noqa: E402
from vkhashdag_b200 import abi, replica  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=13)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = abi.default_config(level_count=a.level, top_level_count=9, bucket_bits_per_bottom_level=15)
    pool = v.DAGNodePool(cfg, device=local)
    sync = replica.ReplicaSync(pool, dist, device=f"cuda:{local}")
    vl = cfg.voxel_level
    res = 1 << vl
    W, H, T = a.width, a.height, 64
    shard = (T, T, rank, world)

    # initial scene on rank 0 only, then published to the replicas
    if rank == 0:
        root = pool.Edit(abi.NULL, v.TerrainEditor(vl))
        pool.SetRoot(root)
    t0 = time.perf_counter()
    nbytes = sync.publish(src=0)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"initial sync: {nbytes / 1e6:.1f} MB in {(time.perf_counter() - t0) * 1e3:.1f} ms")
    ok = True
    for f in range(a.frames):
        te = time.perf_counter()
        if rank == 0:   # brush edit at the frame centre hit (main.cpp:320-343)
            P1 = abi.camera_params(cfg, pool.GetRoot(), (0.5, 0.7, 0.5), 0.6 + 0.2 * f, -0.6, 1, 1, lod=False)
            h = pool.Trace(P1, want=("hits",))["hits"][0, 0]
            c = tuple(int(x) for x in h["vox"]) if h["packed"] >> 31 else (res // 2, res // 3, res // 2)
            r = max(4, res // 64)
            new_root = pool.Edit(pool.GetRoot(), v.SphereEditor(c, r * r, "dig" if f & 1 else "fill"))
            assert pool.last_stats["overflow_count"] == 0
            pool.SetRoot(new_root)
        t_edit = time.perf_counter() - te
        ts = time.perf_counter()
        nbytes = sync.publish(src=0)
        torch.cuda.synchronize()
        t_sync = time.perf_counter() - ts
        root = pool.GetRoot()
        P = abi.camera_params(cfg, root, (0.5, 0.7, 0.5), 0.6 + 0.2 * f, -0.6, W, H, color_root=(1 << 30) | 0x80C0FF)
        tt = time.perf_counter()
        part = pool.Trace(P, want=("rgba8",), shard=shard)["rgba8"]
        dist.barrier()
        t_trace = time.perf_counter() - tt
        # gather shards on rank 0 and compare with a full-frame trace there
        n_max = max(pool.ShardPixels(P, (T, T, r_, world)) for r_ in range(world))
        buf = torch.zeros(n_max, dtype=torch.int32, device="cuda")
        buf[:part.size] = torch.from_numpy(part.view(np.int32)).cuda()
        gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, gathered, dst=0)
        bw = torch.from_numpy(pool.ReadBucketWords().astype(np.int64)).cuda()
        bw_sum = torch.tensor([int(bw.sum()), root], device="cuda", dtype=torch.int64)
        all_bw = [torch.zeros_like(bw_sum) for _ in range(world)]
        dist.all_gather(all_bw, bw_sum)
        same = all(bool((x == all_bw[0]).all()) for x in all_bw)
        if rank == 0:
            parts = [g.cpu().numpy().view(np.uint32)[:pool.ShardPixels(P, (T, T, r_, world))] for r_, g in enumerate(gathered)]
            frame = replica.assemble_frame(parts, W, H, T, T, world)
            full = pool.Trace(P, want=("rgba8",))["rgba8"]
            eq = bool(np.array_equal(frame, full))
            ok &= eq and same
            print(f"frame {f}: edit {t_edit * 1e3:6.2f} ms  sync {t_sync * 1e3:6.2f} ms ({nbytes / 1e3:.1f} KB)  "
                  f"trace+D2H {t_trace * 1e3:6.2f} ms  stitched==full {eq}  replicas identical {same}")
    if rank == 0:
        print("MULTI-GPU CHECK", "OK" if ok else "FAILED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
