#!/usr/bin/env python
"""bench.py — headline benchmark: primary-ray traversal throughput (Mrays/s) on the cfg2 scene of BASELINE.json
("2^15 procedural noise-terrain DAG, 3840x2160 primary-ray traversal on 1xB200"), plus the other BASELINE configs as
extra keys of the same JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm: CUDA kernels behind the C ABI
  python bench.py --impl reference [...]                       reference arm: the reference's own CPU code

A "step" is one frame: every primary ray of a 3840x2160 frame (per GPU: for N > 1 the frame grows to N x 4K pixels and
is sharded over the ranks by 64x64 screen tiles, pool replicated, no data-path collective -> weak scaling).  The camera
moves every step; both arms use the same K cameras.  Headline workload is FULL DETAIL (proj_factor = +inf, no LOD cut-off)
because that is what the reference's CPU tracer (NodePoolTraversal::Traversal<float>) computes, shaded through a REAL
colour pool (painted on the GPU, VBR leaves decoded per hit, trace.frag:272-364); the LOD-on figure (the reference's GPU
shader behaviour, trace.frag:148) and the constant-colour figure are reported beside it.

Extra keys (every N): `strong_8k` = BASELINE config 4 (ONE 7680x4320 frame on the 2^17 edited DAG, tile-sharded, strong
scaling), `interactive` = config 5 (brush edit on rank 0 -> ONE NCCL broadcast of the packed dirty ranges -> sharded 4K
trace, with a per-frame check that every replica renders what rank 0 renders), `value_2p17_edited_4k` = the north-star
target config.  N = 1 only: `parity` (bench-scale canonical-DAG comparison of the GPU-built cfg2 scene and of the cfg3
10 000-edit batch against the reference's ThreadedEdit, and sampled frame rows against the restated shader),
`cpu_baseline`, `edit` (config 3) with its own roofline.

Prints ONE JSON line (rank 0).  Uses oracle/ only as the checker (`parity`), for the `cpu_baseline` legs and in the
`--impl reference` arm.
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
W8K, H8K = 7680, 4320
LEVEL_COUNT = 15          # cfg2: 2^15 voxels per axis
BOTTOM_BUCKET_BITS = 17   # DefaultConfig with 2^17 buckets per bottom level: the default 2^16 overflows level 12
TILE = 64
ROW_STEP = 24             # CPU legs trace every 24th row of a 4K frame (90 rows = 345 600 rays per frame)
COLOR_LEAF_LEVEL = 10     # DAGColorPool::Config::leaf_level of src/main.cpp:206
COLOR_SPHERES = 300       # paint spheres of the headline scene's colour pool (each re-encodes up to ~2 500 colour leaves: ~1.3 M
                          # chunk words per sphere, and the 30-bit leaf index of DAGColorPool.hpp:23-34 caps the pool at 2^30 words)
EDIT_BATCH = 10000        # sphere edits of BASELINE.json config 3
# cfg3 = "2^17 terrain with 10k edits".  A FULL 2^17 x 2^17 terrain of this noise function does not fit the reference's
# address space: the 2^16 x 2^16 patch below stores 528.6 M words of 8^3-voxel nodes (level 14, measured with the
# reference's own editor), the full extent has 4x the surface = 2.11 G words at that level alone, a level's bucket count is
# a power of two (Config.hpp:21) and 2^21 buckets x 2048 words = 2^32 words is already the whole pointer range
# (Config.hpp:48-56: total words <= 2^32 - 1), while 2^20 buckets (2.15 G words) would have to be 98 % full and overflow.
# The largest square power-of-two patch that fits is therefore 2^16 (1/4 of the world's area): 3.64 G words of address
# space with the per-level bucket bits below (level 14 at 2^20 buckets is 25 % full after the terrain).
CFG3_PATCH_BITS = 16
CFG3_BUCKET_BITS = [10] * 9 + [16, 16, 16, 17, 18, 19, 17]


def scene_config():
    from vkhashdag_b200 import abi
    return abi.default_config(level_count=LEVEL_COUNT, top_level_count=9, bucket_bits_per_bottom_level=BOTTOM_BUCKET_BITS)


def cfg3_config():
    from vkhashdag_b200 import abi
    return abi.custom_config(CFG3_BUCKET_BITS)


def camera(cfg, root, step, width, height, lod, color_root=(1 << 30) | 0x60C0E0, scale=1.0):
    """Orbiting camera above the terrain, pitched about -30 degrees.  `scale` shrinks the orbit onto a
    terrain patch that covers only part of the world (cfg3)."""
    from vkhashdag_b200 import abi
    yaw = 0.6 + 0.37 * step
    pos = ((0.5 + 0.12 * np.sin(0.9 * step)) * scale, (0.62 + 0.03 * np.cos(1.3 * step)) * scale,
           (0.5 + 0.12 * np.cos(0.7 * step)) * scale)
    return abi.camera_params(cfg, root, pos, yaw, -0.5236, width, height, color_root=color_root,
                             color_leaf_level=COLOR_LEAF_LEVEL, type_=0, lod=lod)


def frame_dims(n):
    """Global frame for n GPUs: n x 4K pixels (weak scaling), as square as the factorisation allows."""
    a = 1
    while a * a * 2 <= n:
        a *= 2
    b = n // a
    return W4K * b, H4K * a


def make_config(n):
    """The `config` object of the JSON line — identical in both arms (the reference arm times a bounded sample of it)."""
    GW, GH = frame_dims(n)
    return {"workload": f"cfg2: 2^{LEVEL_COUNT} noise-terrain DAG (seed 0x5EED), {W4K}x{H4K} primary rays per GPU, full detail "
                        f"(no LOD cut-off), camera orbit step s = 0..steps-1",
            "frame": [GW, GH], "tile": TILE if n > 1 else None,
            "pool": f"DefaultConfig(level_count={LEVEL_COUNT}, bottom bucket bits {BOTTOM_BUCKET_BITS})",
            "parallelism": f"screen-tile shard x{n}, replicated pool" if n > 1 else "single GPU"}


def paint_sequence(cfg):
    """The headline scene's colour edits: one base colour over everything, then COLOR_SPHERES random paint spheres
    (SphereEditor<kPaint>, src/main.cpp:107-149: colour only, geometry untouched)."""
    from vkhashdag_b200 import abi
    res = 1 << cfg.voxel_level
    seq = [(abi.sphere((res // 2,) * 3, 3 * res * res), 0x60C0E0)]
    palette = [0xE04020, 0x20A040, 0x3060E0, 0xE0C020, 0xA040C0, 0x20C0C0, 0xF08030, 0x808080]
    for i, s in enumerate(abi.random_spheres(COLOR_SPHERES, cfg.voxel_level, seed=4242, rmin=48, rmax=640)):
        s.kind = abi.EDIT_SPHERE_FILL
        seq.append((s, palette[i % len(palette)]))
    return seq


def pin_to_gpu_numa(local):
    """Pin this process to the CPUs that are local to its GPU before any pinned host memory is allocated, so that the
    frame read-back buffers are first-touched on the GPU's own NUMA node (the e2e path is one 33 MB D2H per frame per
    rank).  Returns a short description for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"
        node = open(path + "/numa_node").read().strip()
        cpus = open(path + "/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return {"numa_node": node, "local_cpulist": cpus, "pinned_cpus": len(ids)}
    except Exception as e:  # noqa: BLE001 — topology files are optional
        return {"error": str(e)[:80]}


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
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_static(key):
    """A per-launch figure from the committed ncu capture (profiles/trace_ncu_summary.json), or None.  STATIC: measured
    once under ncu on the same command, not in this run."""
    path = os.path.join(ROOT, "profiles", "trace_ncu_summary.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(key)
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import vkhashdag_b200 as v
    from vkhashdag_b200 import abi, replica

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback")
    torch.cuda.set_device(local)
    numa = pin_to_gpu_numa(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    n = max(world, 1)
    dev = f"cuda:{local}"

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*xs):
        if not dist:
            return list(xs)
        t = torch.tensor(xs, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    # ---- cfg2 scene: every rank builds its replica (no scene transfer), then paints the colour pool ----
    cfg = scene_config()
    pool = v.DAGNodePool(cfg, device=local)
    t0 = time.time()
    root = pool.Edit(abi.NULL, v.TerrainEditor(cfg.voxel_level))
    build_s = time.time() - t0
    stats = pool.last_stats
    assert stats["overflow_count"] == 0, f"bucket overflow ({stats['overflow_count']}): parity void, enlarge the pool"
    pool.SetRoot(root)
    pool.ColorConfig(COLOR_LEAF_LEVEL)
    t0 = time.time()
    for desc, rgb in paint_sequence(cfg):
        r2, croot = pool.EditColor(root, desc, rgb, paint=True)
        assert r2 == root, "a paint edit must not change the geometry"
    paint_s = time.time() - t0
    croot = pool.ColorRoot()

    GW, GH = frame_dims(n)
    shard = (TILE, TILE, rank, n) if n > 1 else None
    probe = camera(cfg, root, 0, GW, GH, False)
    local_px = pool.ShardPixels(probe, shard) if shard else GW * GH
    rays_per_step = GW * GH  # whole job

    stream = torch.cuda.ExternalStream(pool.stream, device=local)
    rgba = torch.zeros(local_px, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def timed(lod, steps, warmup, color_root, first_step=0):
        """Device-timed steps (CUDA events on the pool's stream), L2 flushed between steps (outside the events)."""
        per = []
        with torch.cuda.stream(stream):
            for s in range(-warmup, steps):
                P = camera(cfg, root, first_step + s + 1000 * lod, GW, GH, lod, color_root)
                flush.fill_(s & 0xFF)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pool.TraceDev(P, rgba8=rgba.data_ptr(), shard=shard)
                e1.record()
                if s >= 0:
                    per.append((e0, e1))
            torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in per]

    CONST = (1 << 30) | 0x60C0E0
    launches0 = v.kernel_launches()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    ms_steps = timed(False, args.steps, args.warmup, croot)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = v.kernel_launches() - launches0 - args.warmup
    ms_total = float(sum(ms_steps))
    ms_const = float(sum(timed(False, args.steps, 1, CONST)))
    ms_lod = float(sum(timed(True, args.steps, 1, croot)))
    ms_total, ms_const, ms_lod = max_over_ranks(ms_total, ms_const, ms_lod)
    ms_per_step = ms_total / args.steps
    value = rays_per_step / ms_per_step / 1e3            # Mrays/s, whole job
    value_const = rays_per_step / (ms_const / args.steps) / 1e3
    value_lod = rays_per_step / (ms_lod / args.steps) / 1e3
    # LOD + beam pre-pass (the reference's optional BEAM_OPTIMIZATION path, row N3); single-GPU frames only
    value_lod_beam = None
    if n == 1:
        beam_buf = torch.zeros(((GW + 7) // 8) * ((GH + 7) // 8), dtype=torch.float32, device="cuda")
        ev = []
        with torch.cuda.stream(stream):
            for s in range(-1, args.steps):
                P = camera(cfg, root, s + 1000, GW, GH, True, croot)
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
    # The frame buffers are page-locked (hd_host_alloc).  With N GPUs writing frames into one socket's memory the host's DMA
    # write rate is what bounds e2e: `d2h_copy_only_GBps_slowest_rank` below times the same copies with no kernel at all
    # (measured on this pool's 8-GPU box: 54.6 GB/s alone, 11.85 GB/s per GPU with eight ranks copying = ~95 GB/s aggregate,
    # which the pipelined frames reach).  HD_BENCH_WC=1 allocates write-combined buffers instead (43 GB/s alone: slower here).
    use_wc = os.environ.get("HD_BENCH_WC", "0") != "0"
    hosts = [v.HostBuffer(local_px, write_combined=use_wc) for _ in range(2)]
    views = [h.array for h in hosts]

    def e2e_loop(steps, first):
        acc = 0
        for s in range(steps):
            slot = s & 1
            if s >= 2:
                pool.TraceCollect(slot)
                acc ^= int(views[slot][::65537].sum())
            pool.TraceSubmit(camera(cfg, root, first + s, GW, GH, False, croot), views[slot], slot, shard=shard)
        for s in range(max(steps - 2, 0), steps):
            pool.TraceCollect(s & 1)
            acc ^= int(views[s & 1][::65537].sum())
        return acc

    e2e_loop(3, 500)
    # ceiling of the read-back alone: every rank copies a frame-sized device buffer to its host buffer back to back
    copy_stream = torch.cuda.Stream(device=local)
    host_t = torch.from_numpy(views[0].view(np.int32))   # a tensor view of the page-locked buffer: copy_ is a plain async D2H
    barrier()
    tc = time.perf_counter()
    with torch.cuda.stream(copy_stream):
        for _ in range(args.steps):
            host_t.copy_(rgba, non_blocking=True)
        copy_stream.synchronize()
    my_copy_s = time.perf_counter() - tc
    barrier()
    copy_rate = -max_over_ranks(-(local_px * 4 * args.steps / my_copy_s / 1e9))[0]
    del host_t
    barrier()
    te = time.perf_counter()
    checksum = e2e_loop(args.steps, 0) & 0xFFFFFFFF
    my_e2e_s = time.perf_counter() - te
    barrier()
    e2e_s = time.perf_counter() - te
    out = {"rgba8": views[0]}
    barrier()
    ts = time.perf_counter()
    for s in range(args.steps):
        pool.Trace(camera(cfg, root, s, GW, GH, False, croot), want=("rgba8",), shard=shard, out=out)
    barrier()
    e2e_sync_s = time.perf_counter() - ts
    e2e_s, e2e_sync_s = max_over_ranks(e2e_s, e2e_sync_s)
    e2e_value = rays_per_step * args.steps / e2e_s / 1e6
    e2e_sync_value = rays_per_step * args.steps / e2e_sync_s / 1e6
    # slowest rank's own read-back rate (its loop time includes its traces, so this is a lower bound on the copy rate)
    d2h_rate = -max_over_ranks(-(local_px * 4 * args.steps / my_e2e_s / 1e9))[0]

    # ---- roofline of the dominant kernel (trace_kernel): algorithmic bytes = 4F + 16 per ray ----
    fet = torch.zeros(local_px, dtype=torch.int32, device="cuda")
    f_sum = 0
    with torch.cuda.stream(stream):
        for s in range(args.steps):
            pool.TraceDev(camera(cfg, root, s, GW, GH, False, croot), fetches=fet.data_ptr(), shard=shard)
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
                "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / peak, 5),
                "traffic": ncu_static("dram_bytes_per_launch"),
                "traffic_source": "profiles/trace_ncu_summary.json (static: one ncu --set full capture of this command, "
                                  "not measured in this run)",
                "words_per_ray_F": round(F, 3), "bytes_per_ray": round(bytes_per_ray, 2),
                "sector_granular_GBps": round(value * 1e6 * 32.0 * F / 1e9 / n, 1),
                # lts__t_sectors, l1tex sectors per request, L2 / DRAM GB/s of the same static capture
                "ncu_memory_system": ncu_static("memory_system"),
                # the ceiling that binds: warp instructions issued against 148 SMs x 4 schedulers x clock (same static capture)
                "ncu_issue_roofline": ncu_static("issue_roofline"),
                "limiter": "instruction issue, not HBM: ncu issue-active / lanes per instruction in profiles/ (static)",
                "note": "gather-bound: dependent 4-byte loads; the HBM fraction is reported because the contract asks "
                        "for it, the kernel's real ceiling is the SM issue rate (DESIGN.md §3.1)"}

    edit = {"workload": f"terrain fill 2^{LEVEL_COUNT} in one batched pass (cold: the first GPU work of the process)",
            "seconds": round(build_s, 4), "visited_leaves": stats["visited_leaves"],
            "leaf_voxels_per_s": round(stats["visited_leaves"] * 64 / build_s), "appended_nodes": stats["appended_nodes"],
            "overflow_count": stats["overflow_count"],
            "colour_pool": {"paint_edits": COLOR_SPHERES + 1, "seconds": round(paint_s, 3),
                            "ms_per_paint_edit": round(paint_s / (COLOR_SPHERES + 1) * 1e3, 3)}}

    parity, cpu_baseline, cpu_ctx = None, None, None
    if rank == 0 and n == 1 and not args.no_cpu_baseline:
        parity = {}
        cpu_baseline, cpu_ctx = cpu_trace_leg(pool, cfg, root, croot, args, parity)

    cpu_ctx = None

    # ---- BASELINE configs 3, 4, 5 on the 2^17 scene ----
    scene3 = cfg3_section(args, v, abi, replica, torch, dist, rank, n, local, dev, barrier, max_over_ranks, flush,
                          edit, parity, cpu_baseline)

    if rank == 0:
        line = {
            "metric": "Mrays/s primary-ray traversal @4K", "value": round(value, 2), "unit": "Mrays/s", "n_gpus": n,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
            "config": make_config(n),
            "details": {"pool_used_MB": round(pool.UsedWords() * 4 / 1e6, 1),
                        "colour": f"colour pool painted on the GPU: base colour + {COLOR_SPHERES} paint spheres, colour leaf level "
                                  f"{COLOR_LEAF_LEVEL}, VBR chunks decoded per hit",
                        "l2": "flushed between steps (256 MB fill outside the timed events); pool 800 MB > 126 MB L2",
                        "host_numa": numa},
            "value_constant_colour": round(value_const, 2),
            "value_lod": round(value_lod, 2), "value_lod_beam": round(value_lod_beam, 2) if value_lod_beam else None,
            "value_2p17_edited_4k": scene3.get("value_2p17_edited_4k"),
            "wall_s_timed_region": round(wall, 4),
            "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": 84,
                    "d2h_bytes_per_step": int(local_px * 4 * n), "frame_checksum": checksum,
                    "value_blocking_call": round(e2e_sync_value, 2), "d2h_GBps_slowest_rank": round(d2h_rate, 2),
                    "d2h_copy_only_GBps_slowest_rank": round(copy_rate, 2), "host_buffers": "write-combined pinned" if use_wc else "pinned",
                    "how": "hd_trace_submit/collect per step (2 frames in flight), wall clock incl. the D2H of every shaded "
                           "frame into pinned host memory (allocated after pinning the process to the GPU's NUMA node); "
                           "value_blocking_call = same loop through the blocking hd_trace"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "edit": edit,
            "strong_8k": scene3.get("strong_8k"), "interactive": scene3.get("interactive"),
        }
        if parity is not None:
            line["parity"] = parity
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
        if parity is not None and not all(val is True for key, val in parity.items() if not key.startswith("_")):
            raise SystemExit(f"bench.py: PARITY FAILURE against the reference: {parity}")
    pool.close()
    if dist:
        dist.destroy_process_group()


def cfg3_section(args, v, abi, replica, torch, dist, rank, n, local, dev, barrier, max_over_ranks, flush, edit, parity,
                 cpu_baseline):
    """BASELINE configs 3 (edit batch), 4 (strong-scaling 8K frame) and 5 (interactive loop) on the 2^17 scene.  Rank 0
    builds and edits; replicas are COPIES of rank 0's pool made by the NCCL dirty-range broadcast (replica.ReplicaSync),
    which is also what carries every later edit."""
    out = {}
    cfg3 = cfg3_config()
    vl3 = cfg3.voxel_level
    scale = (1 << CFG3_PATCH_BITS) / (1 << vl3)
    pool3 = v.DAGNodePool(cfg3, device=local)
    stream3 = torch.cuda.ExternalStream(pool3.stream, device=local)
    spheres = abi.random_spheres(EDIT_BATCH, vl3, seed=1234, rmin=16, rmax=256, extent_bits=CFG3_PATCH_BITS)
    terrain = v.TerrainEditor(vl3, extent_bits=CFG3_PATCH_BITS)
    warm = torch.zeros(W4K * H4K, dtype=torch.int32, device="cuda")
    mirror3 = None
    if rank == 0:
        arr = abi.edit_array(spheres)
        # One untimed rehearsal of the whole sequence first (the warm-up step of this one-shot workload): the first pass
        # through the general edit path grows the stream-ordered allocator's pool by several GB.  Then the pool is
        # cleared and the sequence is run again, timed.
        pool3.EditBatch(pool3.Edit(abi.NULL, terrain), arr)
        pool3.Clear()
        t0 = time.perf_counter()
        root3 = pool3.Edit(abi.NULL, terrain)
        edit["terrain_patch_warm_seconds"] = round(time.perf_counter() - t0, 4)
        edit["terrain_patch_visited_leaves"] = pool3.last_stats["visited_leaves"]
        assert pool3.last_stats["overflow_count"] == 0
        if parity is not None:   # the un-edited scene for the oracle's S sample (timed later, on the CPU)
            from oracle import bindings as B
            mirror3 = B.Oracle().pool(cfg3)
            pool3.DownloadInto(mirror3)
        P3 = camera(cfg3, root3, 0, W4K, H4K, False, scale=scale)
        with torch.cuda.stream(stream3):
            for _ in range(30):   # the download above let the clocks drop again
                pool3.TraceDev(P3, rgba8=warm.data_ptr())
            pool3.Sync()
        t0 = time.perf_counter()
        root_b = pool3.EditBatch(root3, arr)
        dt = time.perf_counter() - t0
        st = pool3.last_stats
        assert st["overflow_count"] == 0, "bucket overflow in the edit batch: parity void"
        in_range = abi.spheres_in_range_voxels(spheres, vl3)
        edit["batch"] = {"workload": f"cfg3: 2^{vl3} world, 2^{CFG3_PATCH_BITS} x 2^{CFG3_PATCH_BITS} terrain patch (the largest that "
                                     f"fits the 2^32-word pointer range, see bench.py), {EDIT_BATCH} random sphere fill/dig edits "
                                     f"(r 16..256, xorshift32 seed 1234), one hd_edit_batch call",
                         "seconds": round(dt, 4), "edited_voxels_per_s": round(in_range / dt), "in_range_voxels": in_range,
                         "visited_leaves": st["visited_leaves"], "upserts": st["upserts"], "appended_nodes": st["appended_nodes"],
                         "appended_words": st["appended_words"], "scan_words": st["scan_words"], "path": st["path"],
                         "pool_used_MB": round(pool3.UsedWords() * 4 / 1e6, 1)}
        pool3.SetRoot(root_b)
        # ---- coloured brush edits (vbr_edit of src/main.cpp:224-230, the reference's interactive right-mouse path) in the
        # reference's own configuration: 2^17 world, colour leaf level 10 (128^3-voxel colour leaves, main.cpp:195-211) ----
        pool3.ColorConfig(COLOR_LEAF_LEVEL)
        res3 = 1 << vl3
        _, _ = pool3.EditColor(root_b, abi.sphere((res3 // 2,) * 3, 3 * res3 * res3), 0x60C0E0, paint=True)   # base coat
        g = pool3.Trace(camera(cfg3, root_b, 5, 96, 54, False, scale=scale), want=("hits",))["hits"].reshape(-1)
        g = g[(g["packed"] >> 31) != 0]
        picks = g[np.linspace(0, len(g) - 1, 36).astype(int)] if len(g) else []
        palette = [0xE04020, 0x20A040, 0x3060E0, 0xE0C020, 0xA040C0]
        brushes = [(abi.sphere(tuple(int(c) for c in h["vox"]), 128 * 128), palette[i % 5], i % 3 == 2) for i, h in enumerate(picks)]
        broot, ms = root_b, []
        for d, rgb, paint in brushes:
            t0 = time.perf_counter()
            broot, _ = pool3.EditColor(broot, d, rgb, paint)
            ms.append((time.perf_counter() - t0) * 1e3)
        if ms:
            edit["color_brush"] = {"workload": "r=128 coloured sphere brushes at visible surface points of the edited cfg3 scene, colour leaf "
                                               "level 10 (2 of 3 fill + colour, 1 of 3 paint), one hd_edit_color call each",
                                   "edits": len(ms), "ms_per_edit_median": round(float(np.median(ms[4:])), 4),
                                   "ms_per_edit_p90": round(float(np.percentile(ms[4:], 90)), 4)}
        if parity is not None:
            cfg3_parity_and_cpu_leg(pool3, cfg3, root_b, spheres, mirror3, root3, edit, parity, cpu_baseline, brushes[:12])
            mirror3 = None
    # ---- replicas: ONE packed broadcast of everything rank 0 built (the same path every later edit takes) ----
    sync = None
    if dist:
        sync = replica.ReplicaSync(pool3, dist, device=dev)
        barrier()
        t0 = time.perf_counter()
        nbytes = sync.publish(src=0)
        barrier()
        out["initial_replica_sync"] = {"MB": round(nbytes / 1e6, 1), "seconds": round(time.perf_counter() - t0, 4),
                                       "collectives": sync.collectives}
    else:
        pool3.DirtyReset()
    root3e = pool3.GetRoot()

    def frames_agree(tag):
        """Every rank traces the same small UNSHARDED probe frame from its own replica; the frames must be identical."""
        if not dist:
            return True
        P = camera(cfg3, pool3.GetRoot(), 3, 480, 270, True, scale=scale)
        buf = torch.zeros(480 * 270, dtype=torch.int32, device="cuda")
        with torch.cuda.stream(stream3):
            pool3.TraceDev(P, rgba8=buf.data_ptr())
            pool3.Sync()
        w = torch.arange(1, buf.numel() + 1, device="cuda", dtype=torch.int64)
        h = torch.stack([(buf.to(torch.int64) * w).sum(), (buf != -16777216).sum().to(torch.int64)])
        hs = [torch.zeros_like(h) for _ in range(n)]
        dist.all_gather(hs, h)
        ok = all(bool((x == hs[0]).all()) for x in hs) and int(hs[0][1]) > 0
        if not ok and rank == 0:
            sys.stderr.write(f"[bench] replica frames differ after {tag}: {[x.tolist() for x in hs]}\n")
        return ok

    replicas_ok = frames_agree("initial sync")

    # ---- config 4: ONE 8K frame, tile-sharded, strong scaling; also the north-star 4K figure on this scene ----
    shard8 = (TILE, TILE, rank, n) if n > 1 else None
    P0 = camera(cfg3, root3e, 0, W8K, H8K, False, scale=scale)
    n_local = pool3.ShardPixels(P0, shard8) if shard8 else W8K * H8K
    buf8 = torch.zeros(max(n_local, W8K * H8K if rank == 0 else 0), dtype=torch.int32, device="cuda")
    steps8 = max(4, min(args.steps, 10))

    def timed8(W, H, sh, lod, only_rank0=False):
        ev = []
        if only_rank0 and rank != 0:
            return 0.0
        with torch.cuda.stream(stream3):
            for s in range(-3, steps8):
                P = camera(cfg3, root3e, s + 100, W, H, lod, scale=scale)
                flush.fill_(s & 0xFF)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pool3.TraceDev(P, rgba8=buf8.data_ptr(), shard=sh)
                e1.record()
                if s >= 0:
                    ev.append((e0, e1))
            torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev) / steps8

    barrier()
    ms8, ms8_lod = max_over_ranks(timed8(W8K, H8K, shard8, False), timed8(W8K, H8K, shard8, True))
    barrier()
    ms8_one = max_over_ranks(timed8(W8K, H8K, None, False, only_rank0=True))[0] if n > 1 else ms8
    barrier()
    ms4 = max_over_ranks(timed8(W4K, H4K, None, False, only_rank0=True))[0]
    out["value_2p17_edited_4k"] = round(W4K * H4K / ms4 / 1e3, 1)
    out["strong_8k"] = {"workload": f"cfg4: 2^{vl3} edited DAG (2^{CFG3_PATCH_BITS} terrain patch + {EDIT_BATCH} sphere edits), ONE "
                                    f"{W8K}x{H8K} frame, {TILE}x{TILE} tiles over {n} GPU(s), replicated pool, full detail",
                        "mrays_s": round(W8K * H8K / ms8 / 1e3, 1), "ms_per_frame": round(ms8, 4),
                        "mrays_s_lod": round(W8K * H8K / ms8_lod / 1e3, 1),
                        "ms_per_frame_one_gpu_same_run": round(ms8_one, 4), "speedup_vs_one_gpu": round(ms8_one / ms8, 3),
                        "efficiency": round(ms8_one / ms8 / n, 4), "frames": steps8, "scaling": "strong",
                        "timing": "CUDA events on each rank's stream, max over ranks"}

    # ---- config 5: interactive loop ----
    frames = 30
    P0 = camera(cfg3, root3e, 0, W4K, H4K, True, scale=scale)
    shard4 = (TILE, TILE, rank, n) if n > 1 else None
    n_local = pool3.ShardPixels(P0, shard4) if shard4 else W4K * H4K
    host = torch.zeros(n_local, dtype=torch.int32).pin_memory()
    hout = {"rgba8": host.numpy().view(np.uint32)}
    res_vox = 1 << vl3
    rows, paths, agree = [], {}, True
    for f in range(-3, frames):
        barrier()
        t0 = time.perf_counter()
        t_pick = t0
        if rank == 0:   # pick ray at the centre pixel (main.cpp:320-324), brush at the hit voxel
            cur = pool3.GetRoot()
            P1 = camera(cfg3, cur, f, 1, 1, True, scale=scale)
            hp = pool3.Traversal(cur, tuple(P1.pos), tuple(P1.look))   # the reference's own pick: Traversal<float>
            c = tuple(int(x * res_vox) for x in hp) if hp is not None else (res_vox // 8, res_vox // 12, res_vox // 8)
            t_pick = time.perf_counter()
            new_root = pool3.Edit(cur, v.SphereEditor(c, 128 * 128, "dig" if f & 1 else "fill"))
            assert pool3.last_stats["overflow_count"] == 0
            paths[pool3.last_stats["path"]] = paths.get(pool3.last_stats["path"], 0) + 1
            pool3.SetRoot(new_root)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        nbytes = sync.publish(src=0) if sync else 0
        torch.cuda.synchronize()
        if not sync:
            pool3.DirtyReset()
        t2 = time.perf_counter()
        P = camera(cfg3, pool3.GetRoot(), f, W4K, H4K, True, scale=scale)
        pool3.Trace(P, want=("rgba8",), shard=shard4, out=hout)
        barrier()
        t3 = time.perf_counter()
        if f >= 0:
            rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, nbytes, (t_pick - t0) * 1e3))
        if f < 0 or f % 5 == 4:   # outside the timed part of the frame
            agree = frames_agree(f"interactive frame {f}") and agree
    rows = np.array(rows)
    # edit and sync are rank 0's own clocks (the replicas sit in the broadcast while rank 0 edits, so their "sync"
    # interval would contain the edit); trace and total are the max over ranks
    med = [max_over_ranks(float(np.median(rows[:, i])) if (rank == 0 or i >= 2) else 0.0)[0] for i in range(4)]
    out["interactive"] = {"workload": f"cfg5: 2^{vl3} edited DAG, per frame one r=128 sphere brush at the centre-pixel hit on GPU0, ONE "
                                      f"NCCL broadcast of the packed dirty ranges, {W4K}x{H4K} LOD trace + host read-back sharded "
                                      f"over {n} GPU(s)", "frames": frames,
                          "edit_ms": round(med[0], 3), "of_which_pick_ray_ms": round(float(np.median(rows[:, 5])), 3),
                          "sync_ms": round(med[1], 3), "trace_ms": round(med[2], 3), "total_ms": round(med[3], 3),
                          "edit_paths": paths, "sync_KB_median": round(float(np.median(rows[:, 4])) / 1e3, 1),
                          "collective": "ncclBroadcast via torch.distributed (1 per frame)" if sync else "none (single GPU)",
                          "collectives_issued": sync.collectives if sync else 0,
                          "replica_frames_identical": bool(agree and replicas_ok) if dist else None,
                          "initial_replica_sync": out.get("initial_replica_sync")}
    if dist and not (agree and replicas_ok):
        raise SystemExit("bench.py: a replica rendered a different frame than rank 0 after a publish")
    pool3.close()
    return out


def cpu_trace_leg(pool, cfg, root, croot, args, parity):
    """N = 1: (a) canonical-DAG parity of the GPU-built cfg2 scene against the reference's own build, (b) per-pixel parity
    of sampled rows of the benched frames against the restated shader, (c) the reference's CPU tracer timed on the same
    cameras and the same rows."""
    from oracle import bindings as B
    cores = os.cpu_count() or 1
    O = B.Oracle()
    kind = "reference" if B.Ref.available() else "port"
    mirror = O.pool(cfg)
    t0 = time.perf_counter()
    pool.DownloadInto(mirror)
    download_s = time.perf_counter() - t0
    if kind == "reference":   # the reference builds the same scene with its own editor: the canonical DAGs must agree
        t0 = time.perf_counter()
        host = B.Ref().pool(cfg)
        hroot = host.edit(B.NULL, B.terrain(cfg.voxel_level), threads=cores, max_task_level=10)
        ref_build_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        a, b = O.canonical_fast(mirror.words_ptr, cfg, root), O.canonical_fast(host.words_ptr, cfg, hroot)
        parity["cfg2"] = bool(a == b and a["by_ptr"] == a["by_content"] and a["by_ptr"] > 0)
        parity["_cfg2"] = {"unique_nodes": a["by_ptr"], "voxels": a["voxels"], "hash": f"{a['hash']:016x}",
                           "reference_hash": f"{b['hash']:016x}", "reference_build_s": round(ref_build_s, 2),
                           "canonical_s": round(time.perf_counter() - t0, 2), "download_s": round(download_s, 2),
                           "checker": "oracle canonicaliser over the GPU pool's mirror vs over the pool the reference's "
                                      "ThreadedEdit built (NodePoolThreadedEdit.hpp:104-126)"}
    else:
        host, hroot = mirror, root
        parity["_cfg2"] = "oracle/_ref missing: canonical comparison against the reference skipped"
    # (b) sampled rows of the benched frames, GPU vs the restated shader on the GPU pool's mirror + colour pool
    cn, cl = pool.ReadColor()
    ok, f_words, rays = True, 0, len(range(0, H4K, ROW_STEP)) * W4K
    for s in (0, max(args.steps // 2, 1), max(args.steps - 1, 2)):
        P = camera(cfg, root, s, W4K, H4K, False, croot)
        g = pool.Trace(P, want=("rgba8", "hits", "fetches"))
        e = O.trace_frame(mirror.words_ptr, P, cn, cl, rows=(0, H4K), row_step=ROW_STEP, threads=cores, want=("rgba8", "hits"))
        sel = slice(0, H4K, ROW_STEP)
        ok = ok and np.array_equal(g["rgba8"][sel], e["rgba8"][sel]) and np.array_equal(g["hits"][sel], e["hits"][sel])
        ok = ok and int(g["fetches"][sel].sum()) == int(e["fetches"])
        f_words += int(e["fetches"])
    parity["trace_rows"] = bool(ok)
    parity["_trace_rows"] = f"3 of the benched cameras, every {ROW_STEP}th row: rgba8, hit records and fetched-word count F, GPU == restated shader"
    # (c) the reference's CPU tracer on the SAME cameras and rows
    t = time.perf_counter()
    frames = 0
    while frames < args.steps and (frames < 2 or time.perf_counter() - t < 15):
        P = camera(cfg, hroot, frames, W4K, H4K, False)
        if kind == "reference":
            host.trace_frame_host(P, row_step=ROW_STEP, threads=cores, want_pos=False)
        else:
            O.trace_frame_host(host.words_ptr, cfg.node_levels, P, row_step=ROW_STEP, threads=cores, want_pos=False)
        frames += 1
    dt = time.perf_counter() - t
    return ({"value": round(rays * frames / dt / 1e6, 3), "unit": "Mrays/s", "cores": cores, "kind": kind,
             "sample": f"every {ROW_STEP}th row of the first {frames} benched 4K frames ({rays} rays/frame), full detail, "
                       f"{'NodePoolTraversal::Traversal<float>' if kind == 'reference' else 'oracle port'} on {cores} threads "
                       f"(geometry only: the host tracer has no colour path)",
             "words_per_ray_F_sample": round(f_words / (3 * rays), 3)},
            {"host": host, "hroot": hroot, "kind": kind})


def cpu_color_brush_leg(ctx, cfg, brushes):
    """The reference's coloured brush (VBREditorWrapper + VBRChunkWriter, VBREditor.hpp:26-107) on all host cores."""
    from oracle import bindings as B
    from vkhashdag_b200 import abi
    cores = os.cpu_count() or 1
    host, hroot = ctx["host"], ctx["hroot"]
    cp = B.Ref().color_pool(COLOR_LEAF_LEVEL, node_capacity=1 << 22, leaf_word_capacity=1 << 28)
    res = 1 << cfg.voxel_level
    hroot = host.edit_color(cp, hroot, abi.sphere((res // 2,) * 3, 3 * res * res), 0x60C0E0, True, threads=cores)   # base coat
    ms = []
    for d, rgb, paint in brushes:
        t0 = time.perf_counter()
        hroot = host.edit_color(cp, hroot, d, rgb, paint, threads=cores)
        ms.append((time.perf_counter() - t0) * 1e3)
    return round(float(np.median(ms)), 4)


def cfg3_parity_and_cpu_leg(pool3, cfg3, root_b, spheres, mirror_terrain, root3, edit, parity, cpu_baseline, brushes=()):
    """N = 1: the reference applies the SAME cfg3 sequence (terrain patch, then the 10 000 spheres one ThreadedEdit call
    each, NodePool.hpp:405-417 / NodePoolThreadedEdit.hpp:104-126) — its wall time is the CPU baseline of the edit
    metric and its canonical DAG must equal the GPU batch's.  Also the edit roofline of SURVEY §8d."""
    from oracle import bindings as B
    from vkhashdag_b200 import abi
    cores = os.cpu_count() or 1
    O = B.Oracle()
    vl3 = cfg3.voxel_level
    # S = words the reference's linear bucket scan reads per upsert (NodePool.hpp:79-132), from the instrumented oracle
    # port on the first edits of the same batch applied to a mirror of the same terrain
    sample = spheres[:64]
    mirror_terrain.reset_stats()
    t0 = time.perf_counter()
    r = root3
    for e in sample:
        r = mirror_terrain.edit(r, e)
    port_s = time.perf_counter() - t0
    ps = mirror_terrain.stats()
    ups = max(ps["upserts"], 1)
    S, k_read, k_write = ps["scan_words"] / ups, ps["read_words"] / ups, ps["appended_words"] / ups
    b = edit["batch"]
    alg_bytes = 4.0 * (k_read + k_write + S) * b["upserts"] + 16.0 * b["visited_leaves"]
    peak, peak_src = measured_peak()
    ach = alg_bytes / b["seconds"] / 1e9
    edit["roofline"] = {"bound": "hbm", "kernel": "hd_edit_batch (all kernels of the batch; dominant: k_upsert_grouped, k_down_leaf, "
                                                  "k_dedup — launch list in profiles/r2_launch_list_edit_cfg3.csv)",
                        "achieved": round(ach, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": round(ach / peak, 5),
                        "traffic": ncu_static("edit_batch_dram_bytes"),
                        "traffic_source": "profiles/trace_ncu_summary.json (static ncu capture of the batch), or null",
                        "bytes_per_upsert": round(4.0 * (k_read + k_write + S), 1), "S_scan_words_per_upsert": round(S, 1),
                        "k_read": round(k_read, 2), "k_write": round(k_write, 2), "gpu_upserts": b["upserts"],
                        "gpu_scan_words": b["scan_words"], "leaf_bytes": 16 * b["visited_leaves"],
                        "S_source": f"instrumented oracle port, first {len(sample)} edits of the batch on the same terrain "
                                    f"({ps['upserts']} upserts, {port_s:.2f} s serial)",
                        "note": "unit = one upsert, bytes/upsert = 4(k_read + k_write + S) + 16 B per evaluated leaf (SURVEY §8d); "
                                "the GPU batch merges the 10 000 edits into one pass, so it performs far fewer upserts than "
                                "10 000 sequential reference edits do"}
    mirror_terrain.close()
    if not B.Ref.available():
        parity["_cfg3_batch"] = "oracle/_ref missing: comparison against the reference skipped"
        return
    host = B.Ref().pool(cfg3)
    t0 = time.perf_counter()
    hroot = host.edit(B.NULL, B.terrain(vl3, extent_bits=CFG3_PATCH_BITS), threads=cores, max_task_level=10)
    terrain_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    hroot = host.edit_batch(hroot, spheres, threads=cores, max_task_level=10)
    batch_s = time.perf_counter() - t0
    mirror = O.pool(cfg3)
    t0 = time.perf_counter()
    pool3.DownloadInto(mirror)
    a, c = O.canonical_fast(mirror.words_ptr, cfg3, root_b), O.canonical_fast(host.words_ptr, cfg3, hroot)
    parity["cfg3_batch"] = bool(a == c and a["by_ptr"] == a["by_content"] and a["by_ptr"] > 0)
    parity["_cfg3_batch"] = {"unique_nodes": a["by_ptr"], "voxels": a["voxels"], "hash": f"{a['hash']:016x}",
                             "reference_hash": f"{c['hash']:016x}", "reference_terrain_s": round(terrain_s, 2),
                             "download_and_canonical_s": round(time.perf_counter() - t0, 2),
                             "checker": "canonical DAG of the GPU's ONE batched pass vs the reference's 10 000 sequential "
                                        "ThreadedEdit calls on its own terrain build"}
    in_range = b["in_range_voxels"]
    cpu_baseline["edit"] = {"edits": len(spheres), "seconds": round(batch_s, 3), "ms_per_edit": round(batch_s / len(spheres) * 1e3, 4),
                            "edited_voxels_per_s": round(in_range / batch_s), "kind": "reference", "cores": cores,
                            "what": "the WHOLE cfg3 batch: ThreadedEdit(busy_pool(cores), max_task_level=10), one call per edit, "
                                    "10 000 edits in index order"}
    edit["batch"]["speedup_vs_reference_cpu"] = round(batch_s / b["seconds"], 1)
    if brushes and "color_brush" in edit:   # the same coloured brushes through the reference's VBREditorWrapper on its own scene
        ref_ms = cpu_color_brush_leg({"host": host, "hroot": hroot}, cfg3, brushes)
        edit["color_brush"]["reference_ms_per_edit_median"] = ref_ms
        edit["color_brush"]["speedup_vs_reference_cpu"] = round(ref_ms / edit["color_brush"]["ms_per_edit_median"], 2)
        cpu_baseline["color_brush"] = {"ms_per_edit_median": ref_ms, "kind": "reference", "cores": cores,
                                       "what": f"ThreadedEdit(max_task_level = colour leaf level) through VBREditorWrapper, first {len(brushes)} "
                                               f"brushes of the same list on the reference-built cfg3 scene (base coat, then the brushes)"}
    host.close(), mirror.close()


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation end to end: ThreadedEdit builds the scene (untimed set-up), then
    Traversal<float> traces a bounded sample (every ROW_STEP-th row) of the same K cameras on all host cores."""
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
    rays = len(range(0, H4K, ROW_STEP)) * W4K

    def step(s):
        P = camera(cfg, root, s, W4K, H4K, False)
        if kind == "reference":
            host.trace_frame_host(P, row_step=ROW_STEP, threads=cores, want_pos=False)
        else:
            B.Oracle().trace_frame_host(host.words_ptr, cfg.node_levels, P, row_step=ROW_STEP, threads=cores, want_pos=False)

    for s in range(args.warmup):
        step(-1 - s)
    t = time.perf_counter()
    for s in range(args.steps):
        step(s)
    dt = time.perf_counter() - t
    value = rays * args.steps / dt / 1e6
    sample = (f"every {ROW_STEP}th row of each of the same {args.steps} 4K camera frames ({rays} rays/step), full detail, "
              f"{'NodePoolTraversal::Traversal<float> (oracle/_ref)' if kind == 'reference' else 'oracle port'} on {cores} threads; "
              f"scene built by {'ThreadedEdit' if kind == 'reference' else 'serial Edit port'} in {build_s:.1f} s (untimed)")
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s primary-ray traversal @4K", "value": round(value, 3), "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
        "config": make_config(max(args.gpus, 1)),
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
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU legs and the bench-scale parity checks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
