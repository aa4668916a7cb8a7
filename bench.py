#!/usr/bin/env python
"""bench.py — headline benchmark: primary-ray traversal throughput (Mrays/s) on the cfg2 scene of BASELINE.json
("2^15 procedural noise-terrain DAG, 3840x2160 primary-ray traversal on 1xB200").

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm: CUDA kernels behind the C ABI
  python bench.py --impl reference [...]                       reference arm: the reference's own CPU code

A "step" is one frame: every primary ray of a 3840x2160 frame (per GPU: for N > 1 the frame grows to N x 4K pixels and
is sharded over the ranks by 64x64 screen tiles, pool replicated, no data-path collective -> weak scaling).
The camera moves every step.  Headline workload is FULL DETAIL (proj_factor = +inf, no LOD cut-off) because that is
what the reference's CPU tracer (NodePoolTraversal::Traversal<float>) computes; the LOD-on figure (the reference's
GPU shader behaviour, trace.frag:148) is reported beside it as `value_lod`.

Prints ONE JSON line (rank 0).  Uses oracle/ only for the `cpu_baseline` leg and the `--impl reference` arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W4K, H4K = 3840, 2160
EDIT_BATCH = 10000        # sphere edits of the secondary metric (BASELINE.json config 3)
CFG3_BUCKET_BITS = [10] * 9 + [16] * 4 + [18] * 3   # bucket bits per node level of the 2^17 pool
CFG3_PATCH_BITS = 15
LEVEL_COUNT = 15          # 2^15 voxels per axis
BOTTOM_BUCKET_BITS = 17   # DefaultConfig with 2^17 buckets per bottom level: the default 2^16 overflows level 12
TILE = 64


def scene_config():
    from vkhashdag_b200 import abi
    return abi.default_config(level_count=LEVEL_COUNT, top_level_count=9, bucket_bits_per_bottom_level=BOTTOM_BUCKET_BITS)


def camera(cfg, root, step, width, height, lod):
    """Orbiting camera above the terrain, pitched about -30 degrees."""
    from vkhashdag_b200 import abi
    yaw = 0.6 + 0.37 * step
    pos = (0.5 + 0.12 * np.sin(0.9 * step), 0.62 + 0.03 * np.cos(1.3 * step), 0.5 + 0.12 * np.cos(0.7 * step))
    return abi.camera_params(cfg, root, pos, yaw, -0.5236, width, height, color_root=(1 << 30) | 0x60C0E0,
                             color_leaf_level=10, type_=0, lod=lod)


def frame_dims(n):
    """Global frame for n GPUs: n x 4K pixels (weak scaling), as square as the factorisation allows."""
    a = 1
    while a * a * 2 <= n:
        a *= 2
    b = n // a
    return W4K * b, H4K * a


class ClockSampler:
    """nvidia-smi clocks/throttle sampling DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic():
    """Per-launch DRAM bytes of the trace kernel from the committed ncu capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "trace_ncu_summary.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import vkhashdag_b200 as v
    from vkhashdag_b200 import abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    cfg = scene_config()
    pool = v.DAGNodePool(cfg, device=local)
    t0 = time.time()
    root = pool.Edit(abi.NULL, v.TerrainEditor(cfg.voxel_level))   # every rank builds its replica of the pool
    build_s = time.time() - t0
    stats = pool.last_stats
    assert stats["overflow_count"] == 0, f"bucket overflow ({stats['overflow_count']}): parity void, enlarge the pool"
    pool.SetRoot(root)

    n = max(world, 1)
    GW, GH = frame_dims(n)
    shard = (TILE, TILE, rank, n) if n > 1 else None
    probe = camera(cfg, root, 0, GW, GH, False)
    local_px = pool.ShardPixels(probe, shard) if shard else GW * GH
    rays_per_step = GW * GH  # whole job

    stream = torch.cuda.ExternalStream(pool.stream, device=local)
    rgba = torch.zeros(local_px, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def timed(lod, steps, warmup, first_step=0):
        """Device-timed steps (CUDA events on the pool's stream), L2 flushed between steps (outside the events)."""
        per = []
        with torch.cuda.stream(stream):
            for s in range(-warmup, steps):
                P = camera(cfg, root, first_step + s + 1000 * lod, GW, GH, lod)
                flush.fill_(s & 0xFF)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pool.TraceDev(P, rgba8=rgba.data_ptr(), shard=shard)
                e1.record()
                if s >= 0:
                    per.append((e0, e1))
            torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in per]

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = v.kernel_launches()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    ms_steps = timed(False, args.steps, args.warmup)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = v.kernel_launches() - launches0 - args.warmup
    ms_total = float(sum(ms_steps))
    ms_lod = float(sum(timed(True, args.steps, 1)))
    if dist:
        t = torch.tensor([ms_total, ms_lod], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_lod = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    value = rays_per_step / ms_per_step / 1e3            # Mrays/s, whole job
    value_lod = rays_per_step / (ms_lod / args.steps) / 1e3
    # LOD + beam pre-pass (the reference's optional BEAM_OPTIMIZATION path, row N3); single-GPU frames only
    value_lod_beam = None
    if n == 1:
        beam_buf = torch.zeros(((GW + 7) // 8) * ((GH + 7) // 8), dtype=torch.float32, device="cuda")
        ev = []
        with torch.cuda.stream(stream):
            for s in range(-1, args.steps):
                P = camera(cfg, root, s + 1000, GW, GH, True)
                B = abi.beam_params(P)
                flush.fill_(s & 0xFF)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pool.BeamDev(B, beam_buf.data_ptr())
                pool.TraceBeamDev(P, beam_buf.data_ptr(), B.width, B.height, rgba8=rgba.data_ptr())
                e1.record()
                if s >= 0:
                    ev.append((e0, e1))
            torch.cuda.synchronize()
        value_lod_beam = rays_per_step / (sum(a_.elapsed_time(b_) for a_, b_ in ev) / args.steps) / 1e3

    # ---- e2e: through the C ABI with HOST buffers, wall clock incl. the copies ----
    # Every step passes the 84-byte parameter block in and reads the shaded frame (rgba8) back into pinned host
    # memory.  hd_trace_submit/collect keeps two frames in flight (the reference keeps kFrameCount = 3,
    # src/main.cpp:20), so the read-back of frame k overlaps the trace of frame k+1; every frame is collected and
    # folded into a checksum inside the timed region.  `e2e_sync` is the same loop with the blocking hd_trace call.
    hosts = [torch.zeros(local_px, dtype=torch.int32).pin_memory() for _ in range(2)]
    views = [h.numpy().view(np.uint32) for h in hosts]

    def e2e_loop(steps, first):
        acc = 0
        for s in range(steps):
            slot = s & 1
            if s >= 2:
                pool.TraceCollect(slot)
                acc ^= int(views[slot][::4099].sum())
            pool.TraceSubmit(camera(cfg, root, first + s, GW, GH, False), views[slot], slot, shard=shard)
        for s in range(max(steps - 2, 0), steps):
            pool.TraceCollect(s & 1)
            acc ^= int(views[s & 1][::4099].sum())
        return acc

    e2e_loop(3, 500)
    barrier()
    te = time.perf_counter()
    checksum = e2e_loop(args.steps, 0) & 0xFFFFFFFF
    barrier()
    e2e_s = time.perf_counter() - te
    out = {"rgba8": views[0]}
    barrier()
    ts = time.perf_counter()
    for s in range(args.steps):
        pool.Trace(camera(cfg, root, s, GW, GH, False), want=("rgba8",), shard=shard, out=out)
    barrier()
    e2e_sync_s = time.perf_counter() - ts
    if dist:
        t = torch.tensor([e2e_s, e2e_sync_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, e2e_sync_s = float(t[0]), float(t[1])
    e2e_value = rays_per_step * args.steps / e2e_s / 1e6
    e2e_sync_value = rays_per_step * args.steps / e2e_sync_s / 1e6

    # ---- roofline of the dominant kernel (trace_kernel): algorithmic bytes = 4F + 16 per ray ----
    fet = torch.zeros(local_px, dtype=torch.int32, device="cuda")
    f_sum = 0
    with torch.cuda.stream(stream):
        for s in range(args.steps):
            pool.TraceDev(camera(cfg, root, s, GW, GH, False), fetches=fet.data_ptr(), shard=shard)
            pool.Sync()
            f_sum += int(fet.sum(dtype=torch.int64))
    if dist:
        t = torch.tensor([f_sum], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        f_sum = int(t[0])
    F = f_sum / (rays_per_step * args.steps)
    bytes_per_ray = 4.0 * F + 16.0
    peak, peak_src = measured_peak()
    achieved = value * 1e6 * bytes_per_ray / 1e9 / n   # GB/s per GPU
    roofline = {"bound": "hbm", "kernel": "trace_kernel", "achieved": round(achieved, 2), "peak": peak,
                "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)" if peak_src == "measured" else "fallback",
                "unit": "GB/s", "frac": round(achieved / peak, 5), "traffic": ncu_traffic(),
                "words_per_ray_F": round(F, 3), "bytes_per_ray": round(bytes_per_ray, 2),
                "sector_granular_GBps": round(value * 1e6 * 32.0 * F / 1e9 / n, 1),
                "note": "gather/latency-bound: dependent 4-byte loads; see profiles/ for L2 sectors and stall reasons"}

    # ---- edit throughput (secondary metric: edited voxels/s) measured on the scene build ----
    edit = {"workload": f"terrain fill 2^{LEVEL_COUNT} in one batched pass (cold: the first GPU work of the process)", "seconds": round(build_s, 4),
            "visited_leaves": stats["visited_leaves"], "leaf_voxels_per_s": round(stats["visited_leaves"] * 64 / build_s),
            "appended_nodes": stats["appended_nodes"], "overflow_count": stats["overflow_count"]}

    cpu_baseline = None
    if rank == 0 and n == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_leg(pool, cfg, root, args)

    # ---- secondary metric of BASELINE.json: "edited voxels/s per batch" on config 3 (2^17 world, 2^15 terrain patch,
    # 10 000 random sphere fill/dig edits in index order) and the latency of one brush edit per call; N = 1 only ----
    if rank == 0 and n == 1:
        cfg3 = abi.custom_config(CFG3_BUCKET_BITS)
        vl3 = cfg3.voxel_level
        pool3 = v.DAGNodePool(cfg3, device=local)
        spheres = abi.random_spheres(EDIT_BATCH, vl3, seed=1234, rmin=16, rmax=256, extent_bits=CFG3_PATCH_BITS)
        arr = abi.edit_array(spheres)
        # One untimed rehearsal of the whole sequence first (the warm-up step of this one-shot workload): the GPU idled
        # while the host ran the CPU tracer leg, and the first pass through the general edit path grows the stream-
        # ordered allocator's pool by several GB.  Then the pool is cleared and the sequence is run again, timed.
        pool3.EditBatch(pool3.Edit(abi.NULL, v.TerrainEditor(vl3, extent_bits=CFG3_PATCH_BITS)), arr)
        pool3.Clear()
        t0 = time.perf_counter()
        root3 = pool3.Edit(abi.NULL, v.TerrainEditor(vl3, extent_bits=CFG3_PATCH_BITS))
        edit["terrain_patch_warm_seconds"] = round(time.perf_counter() - t0, 4)
        assert pool3.last_stats["overflow_count"] == 0
        mirror = pool3.Download() if cpu_baseline is not None else None   # un-edited scene for the CPU editor (timed later)
        P3 = abi.camera_params(cfg3, root3, (0.06, 0.09, 0.06), 0.8, -0.5236, W4K, H4K, lod=False)
        with torch.cuda.stream(torch.cuda.ExternalStream(pool3.stream, device=local)):
            for _ in range(40):   # the download above let the clocks drop again
                pool3.TraceDev(P3, rgba8=rgba.data_ptr())
            pool3.Sync()
        t0 = time.perf_counter()
        root_b = pool3.EditBatch(root3, arr)
        dt = time.perf_counter() - t0
        st = pool3.last_stats
        assert st["overflow_count"] == 0, "bucket overflow in the edit batch: parity void"
        in_range = abi.spheres_in_range_voxels(spheres, vl3)
        edit["batch"] = {"workload": f"cfg3: 2^{vl3} world, 2^{CFG3_PATCH_BITS} terrain patch, {EDIT_BATCH} random sphere fill/dig edits "
                                     f"(r 16..256, xorshift32 seed 1234), one hd_edit_batch call", "seconds": round(dt, 4),
                         "edited_voxels_per_s": round(in_range / dt), "in_range_voxels": in_range,
                         "visited_leaves": st["visited_leaves"], "appended_nodes": st["appended_nodes"], "path": st["path"]}
        ms = []
        for e in abi.random_spheres(60, vl3, seed=77, rmin=128, rmax=128, extent_bits=CFG3_PATCH_BITS):
            a1 = abi.edit_array([e])
            t0 = time.perf_counter()
            root_b = pool3.EditBatch(root_b, a1)
            ms.append((time.perf_counter() - t0) * 1e3)
            assert pool3.last_stats["overflow_count"] == 0
        edit["brush"] = {"workload": "one r=128 sphere brush per hd_edit_batch call on the edited cfg3 scene (60 calls, fill/dig alternating)",
                         "ms_per_edit_median": round(float(np.median(ms[5:])), 4), "path": pool3.last_stats["path"]}
        if mirror is not None:
            cpu_baseline["edit"] = cpu_edit_leg(mirror, cfg3, root3, spheres[:40])
        pool3.close()

    if rank == 0:
        line = {
            "metric": "Mrays/s primary-ray traversal @4K", "value": round(value, 2), "unit": "Mrays/s", "n_gpus": n,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
            "config": {"workload": f"cfg2: 2^{LEVEL_COUNT} noise-terrain DAG (seed 0x5EED, GPU-built), {W4K}x{H4K} primary rays "
                                   f"per GPU, full detail (no LOD cut-off)", "frame": [GW, GH], "tile": TILE if n > 1 else None,
                       "pool": f"DefaultConfig(level_count={LEVEL_COUNT}, bottom bucket bits {BOTTOM_BUCKET_BITS})",
                       "pool_used_MB": round(pool.UsedWords() * 4 / 1e6, 1),
                       "l2": "flushed between steps (256 MB fill outside the timed events); pool 800 MB > 126 MB L2",
                       "parallelism": f"screen-tile shard x{n}, replicated pool" if n > 1 else "single GPU"},
            "value_lod": round(value_lod, 2), "value_lod_beam": round(value_lod_beam, 2) if value_lod_beam else None,
            "wall_s_timed_region": round(wall, 4),
            "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": 84,
                    "d2h_bytes_per_step": int(local_px * 4 * n), "frame_checksum": checksum,
                    "value_blocking_call": round(e2e_sync_value, 2),
                    "how": "hd_trace_submit/collect per step (2 frames in flight), wall clock incl. the D2H of every shaded "
                           "frame into pinned host memory; value_blocking_call = same loop through the blocking hd_trace"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "edit": edit,
        }
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    pool.close()
    if dist:
        dist.destroy_process_group()


def cpu_baseline_leg(pool, cfg, root, args):
    """The reference's own CPU tracer (oracle/_ref, else the oracle port) on a bounded sample of the same frames:
    the GPU-built pool is mirrored into the CPU pool's memory, every `row_step`-th row is traced on all host cores."""
    from oracle import bindings as B
    cores = os.cpu_count() or 1
    kind = "reference" if B.Ref.available() else "port"
    ranges, bw = pool.Download()
    if kind == "reference":
        host = B.Ref().pool(cfg)
    else:
        host = B.Oracle().pool(cfg)
    for off, words in ranges.items():
        host.words_np(off, len(words))[:] = words
    host.bucket_words_np()[:] = bw
    row_step = 24   # 90 of 2160 rows -> 345,600 rays per frame
    P = camera(cfg, root, 0, W4K, H4K, False)
    rays = len(range(0, H4K, row_step)) * W4K
    t = time.perf_counter()
    frames = 0
    while frames < 2 or (time.perf_counter() - t < 10 and frames < 40):
        P = camera(cfg, root, frames, W4K, H4K, False)
        if kind == "reference":
            host.trace_frame_host(P, row_step=row_step, threads=cores, want_pos=False)
        else:
            B.Oracle().trace_frame_host(host.words_ptr, cfg.node_levels, P, row_step=row_step, threads=cores, want_pos=False)
        frames += 1
    dt = time.perf_counter() - t
    # sector-exact F of the same sample from the shader restatement (port), as a cross-check of the GPU counter
    fr = B.Oracle().trace_frame(host.words_ptr, camera(cfg, root, 0, W4K, H4K, False), rows=(0, H4K), row_step=row_step,
                                threads=cores, want=())
    return {"value": round(rays * frames / dt / 1e6, 3), "unit": "Mrays/s", "cores": cores, "kind": kind,
            "sample": f"every {row_step}th row of {frames} 4K frames ({rays} rays/frame), full detail, "
                      f"{'NodePoolTraversal::Traversal<float>' if kind == 'reference' else 'oracle port'} on {cores} threads",
            "words_per_ray_F_sample": round(fr["fetches"] / rays, 3)}


def cpu_edit_leg(mirror, cfg3, root3, sample):
    """The reference's CPU editor on the first spheres of the cfg3 edit list: the GPU-built terrain pool is mirrored
    into the CPU pool, then one call per edit — ThreadedEdit on all host cores with max_task_level = 10 as in
    src/main.cpp:216 (oracle/_ref), or the serial Edit port when the reference could not be compiled."""
    from oracle import bindings as B
    from vkhashdag_b200 import abi
    cores = os.cpu_count() or 1
    kind = "reference" if B.Ref.available() else "port"
    host = (B.Ref() if kind == "reference" else B.Oracle()).pool(cfg3)
    ranges, bw = mirror
    for off, words in ranges.items():
        host.words_np(off, len(words))[:] = words
    host.bucket_words_np()[:] = bw
    hroot = root3
    t = time.perf_counter()
    for e in sample:
        hroot = host.edit(hroot, e, threads=cores, max_task_level=10) if kind == "reference" else host.edit(hroot, e)
    dt = time.perf_counter() - t
    return {"edits": len(sample), "seconds": round(dt, 4), "ms_per_edit": round(dt / len(sample) * 1e3, 4),
            "edited_voxels_per_s": round(abi.spheres_in_range_voxels(sample, cfg3.voxel_level) / dt), "kind": kind, "cores": cores,
            "what": "ThreadedEdit(busy_pool(cores), max_task_level=10), one call per edit, first 40 edits of the batch"
                    if kind == "reference" else "serial Edit port, first 40 edits of the batch"}


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation end to end: ThreadedEdit builds the scene (untimed set-up), then
    Traversal<float> traces a bounded sample of each 4K frame on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import bindings as B
    cores = os.cpu_count() or 1
    cfg = scene_config()
    kind = "reference" if B.Ref.available() else "port"
    t0 = time.perf_counter()
    if kind == "reference":
        host = B.Ref().pool(cfg)
        root = host.edit(B.NULL, B.terrain(cfg.voxel_level), threads=cores, max_task_level=10)
    else:
        host = B.Oracle().pool(cfg)
        root = host.edit(B.NULL, B.terrain(cfg.voxel_level))
    build_s = time.perf_counter() - t0
    row_step = 24
    rays = len(range(0, H4K, row_step)) * W4K

    def step(s):
        P = camera(cfg, root, s, W4K, H4K, False)
        if kind == "reference":
            host.trace_frame_host(P, row_step=row_step, threads=cores, want_pos=False)
        else:
            B.Oracle().trace_frame_host(host.words_ptr, cfg.node_levels, P, row_step=row_step, threads=cores, want_pos=False)

    for s in range(args.warmup):
        step(-1 - s)
    t = time.perf_counter()
    for s in range(args.steps):
        step(s)
    dt = time.perf_counter() - t
    value = rays * args.steps / dt / 1e6
    sample = (f"every {row_step}th row of each 4K frame ({rays} rays/step), full detail, "
              f"{'NodePoolTraversal::Traversal<float> (oracle/_ref)' if kind == 'reference' else 'oracle port'} on {cores} threads; "
              f"scene built by {'ThreadedEdit' if kind == 'reference' else 'serial Edit port'} in {build_s:.1f} s (untimed)")
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s primary-ray traversal @4K", "value": round(value, 3), "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
        "config": {"workload": f"cfg2: 2^{LEVEL_COUNT} noise-terrain DAG (seed 0x5EED), {W4K}x{H4K} primary rays, full detail; "
                               f"bounded sample per step", "pool": f"DefaultConfig(level_count={LEVEL_COUNT}, bottom bucket bits {BOTTOM_BUCKET_BITS})"},
        "cpu_baseline": {"value": round(value, 3), "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
