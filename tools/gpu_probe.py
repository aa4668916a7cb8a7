"""Scratch probe: build the cfg2 terrain on the GPU and time the 4K trace (device-resident outputs)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vkhashdag_b200 as v
from vkhashdag_b200 import abi

lc = int(sys.argv[1]) if len(sys.argv) > 1 else 15
W, H = (3840, 2160)
cfg = abi.default_config(level_count=lc, top_level_count=9)
print("total words", cfg.total_words(), "GB", cfg.total_words()*4/1e9)
t = time.time(); pool = v.DAGNodePool(cfg); print("create", time.time()-t)
t = time.time(); root = pool.Edit(abi.NULL, v.TerrainEditor(cfg.voxel_level, amp_div=int(sys.argv[2]) if len(sys.argv) > 2 else 8)); dt = time.time()-t
print("terrain build s", dt, "root", root, pool.last_stats, "used words", pool.UsedWords())
bw = pool.ReadBucketWords(); bases = cfg.level_bases()
for l in range(cfg.node_levels):
    lo = bases[l]; hi = bases[l+1] if l+1 < cfg.node_levels else len(bw)
    print("level", l, "max bucket", int(bw[lo:hi].max()), "mean", float(bw[lo:hi].mean()), "sum", int(bw[lo:hi].sum()))
stream = torch.cuda.ExternalStream(pool.stream)
rgba = torch.zeros(H*W, dtype=torch.int32, device="cuda")
iters = torch.zeros(H*W, dtype=torch.int32, device="cuda")
for (pos, yaw, pitch) in [((0.5, 0.75, 0.5), 0.6, -0.5236), ((0.5, 0.62, 0.1), 0.0, -0.3), ((0.3, 0.9, 0.3), 0.8, -1.0)]:
    for lod in (True, False):
        P = abi.camera_params(cfg, root, pos, yaw, pitch, W, H, color_root=(1 << 30) | 0x80C0FF, lod=lod)
        with torch.cuda.stream(stream):
            for _ in range(3): pool.TraceDev(P, rgba8=rgba.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): pool.TraceDev(P, rgba8=rgba.data_ptr())
            e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1)/10
            pool.TraceDev(P, rgba8=rgba.data_ptr(), iters=iters.data_ptr()); pool.Sync()
        it = iters.cpu().numpy().astype(np.int64)
        hitfrac = float(((rgba.cpu().numpy().view(np.uint32) & 0xFFFFFF) != 0).mean())
        print(f"cam {pos} yaw {yaw} pitch {pitch} lod {lod}: {ms:.3f} ms  {W*H/ms/1e3:.1f} Mrays/s  mean iters {it.mean():.1f} max {it.max()} hit {hitfrac:.3f}")
