"""GPU parity of kernel #1 (trace) against the CPU oracle, through the C ABI with host buffers."""
import numpy as np
import pytest

from vkhashdag_b200 import abi

pytestmark = pytest.mark.gpu

NULL = abi.NULL


def build_scene(oracle, level_count=10, edits=None):
    cfg = abi.default_config(level_count=level_count, top_level_count=9)
    pool = oracle.pool(cfg)
    res = 1 << cfg.voxel_level
    if edits is None:
        edits = [abi.sphere((res // 2,) * 3, (res // 3) ** 2)]
    root = pool.edit_batch(NULL, edits)
    return cfg, pool, root


def compare_frames(oracle, hd, cfg, opool, root, cams, W, H, color=None, types=(0,)):
    dev = hd.DAGNodePool(cfg)
    dev.UploadFrom(opool)
    color_root, cn, cl, cleaf = abi.COLOR_NULL, None, None, 10
    if color is not None:
        color_root, cn, cl, cleaf = color
        dev.UploadColor(cn, cl)
    for (pos, yaw, pitch, lod) in cams:
        for t in types:
            P = abi.camera_params(cfg, root, pos, yaw, pitch, W, H, color_root=color_root, color_leaf_level=cleaf,
                                  type_=t, lod=lod)
            exp = oracle.trace_frame(opool.words_ptr, P, cn, cl)
            got = dev.Trace(P)
            assert np.array_equal(got["iters"], exp["iters"]), f"iters differ cam={pos, yaw, pitch, lod}"
            assert np.array_equal(got["hits"], exp["hits"]), f"hit records differ cam={pos, yaw, pitch, lod}"
            assert np.array_equal(got["rgba8"], exp["rgba8"]), f"frame differs cam={pos, yaw, pitch, lod} type={t}"
    dev.close()
    return exp


def test_trace_sphere_cfg1(oracle, hd):
    """BASELINE config 1: 2^10 sphere, 1280x720, camera (0.5,0.5,1.5) looking -z."""
    cfg, opool, root = build_scene(oracle)
    exp = compare_frames(oracle, hd, cfg, opool, root, [((0.5, 0.5, 1.5), np.pi, 0.0, True)], 1280, 720, types=(0, 1, 2))
    assert (exp["hits"]["packed"] >> 31).sum() > 10000


def test_trace_cameras_lod_and_full_detail(oracle, hd):
    res = 1 << 10
    edits = [abi.sphere((512, 512, 512), 341 ** 2), abi.sphere((512, 512, 300), 150 ** 2, dig=True),
             abi.sphere((700, 600, 512), 200 ** 2), abi.aabb((100, 50, 100), (400, 90, 900))]
    cfg, opool, root = build_scene(oracle, edits=edits)
    cams = [((0.5, 0.5, 1.5), np.pi, 0.0, False), ((0.1, 0.9, 0.1), 0.7, -0.6, True), ((0.5, 0.5, 0.5), 2.0, 0.3, True),
            ((-0.2, 0.3, 0.4), 1.4, 0.1, False), ((0.52, 0.97, 0.51), 0.0, -1.5, True)]
    compare_frames(oracle, hd, cfg, opool, root, cams, 320, 200)


def test_trace_odd_sizes_and_empty(oracle, hd):
    cfg, opool, root = build_scene(oracle, level_count=8)
    compare_frames(oracle, hd, cfg, opool, root, [((0.5, 0.5, 1.4), np.pi, 0.0, True)], 333, 77)
    compare_frames(oracle, hd, cfg, opool, NULL, [((0.5, 0.5, 1.4), np.pi, 0.0, True)], 64, 32)
    compare_frames(oracle, hd, cfg, opool, root, [((0.5, 0.5, 1.4), np.pi, 0.0, True)], 1, 1)


def test_trace_tile_shards_cover_frame(oracle, hd):
    cfg, opool, root = build_scene(oracle, level_count=9)
    dev = hd.DAGNodePool(cfg)
    dev.UploadFrom(opool)
    from vkhashdag_b200 import replica
    tw, th = 64, 64
    # 400 px: 7 tiles per row (row-major round robin); 512 px: 8 per row (the (tx + ty) % world map for world 2, 4, 8)
    for W, H, worlds in ((400, 250, (1, 2, 3, 8)), (512, 200, (2, 3, 4, 8))):
        P = abi.camera_params(cfg, root, (0.5, 0.6, 1.3), np.pi, -0.2, W, H)
        full = dev.Trace(P)
        for world in worlds:
            seen = np.zeros((H, W), bool)
            for rank in range(world):
                part = dev.Trace(P, shard=(tw, th, rank, world))
                tiles = replica.local_tiles(W, H, tw, th, rank, world)
                assert part["rgba8"].size == len(tiles) * tw * th
                for lt, tx, ty in tiles:
                    assert not seen[ty * th, tx * tw]   # no tile is owned twice
                    x0, y0 = tx * tw, ty * th
                    w, h = min(tw, W - x0), min(th, H - y0)
                    for name in ("rgba8", "iters", "hits"):
                        blk = part[name][lt * tw * th:(lt + 1) * tw * th].reshape(th, tw)[:h, :w]
                        assert np.array_equal(blk, full[name][y0:y0 + h, x0:x0 + w])
                    seen[y0:y0 + h, x0:x0 + w] = True
            assert seen.all()
    dev.close()


def test_pick_ray_matches_host_traversal(oracle, hd):
    cfg, opool, root = build_scene(oracle)
    dev = hd.DAGNodePool(cfg)
    dev.UploadFrom(opool)
    rng = np.random.default_rng(7)
    rays = [((0.5, 0.5, 1.5), (0, 0, -1)), ((0.5, 0.5, 1.5), (0.1, -0.2, -1)), ((-0.2, 0.3, 0.4), (1, 0.3, 0.2)),
            ((0.5, 0.5, 1.5), (0.6, 0, -1))]
    for _ in range(60):
        rays.append((tuple(rng.uniform(-0.5, 1.5, 3)), tuple(rng.normal(size=3))))
    for o, d in rays:
        d = np.array(d, np.float32)
        d = d / np.float32(np.sqrt(np.float32(np.dot(d, d))))
        exp = oracle.traverse(opool.words_ptr, cfg.node_levels, root, o, d)
        got = dev.Traversal(root, o, d)
        assert (exp is None) == (got is None)
        if exp is not None:
            assert np.array_equal(exp.view(np.uint32), got.view(np.uint32))
    assert dev.Traversal(NULL, (0.5, 0.5, 1.5), (0, 0, -1)) is None
    dev.close()


def test_pipelined_frames_equal_blocking_frames(oracle, hd):
    """hd_trace_submit/collect (two frames in flight) returns the same frames as the blocking hd_trace."""
    import torch
    cfg, opool, root = build_scene(oracle, level_count=9)
    dev = hd.DAGNodePool(cfg)
    dev.UploadFrom(opool)
    W, H = 640, 360
    cams = [abi.camera_params(cfg, root, (0.5, 0.5 + 0.02 * i, 1.4), np.pi + 0.05 * i, -0.1, W, H) for i in range(7)]
    want = [dev.Trace(P, want=("rgba8",))["rgba8"].copy() for P in cams]
    for shard in (None, (64, 64, 1, 3)):
        n = dev.ShardPixels(cams[0], shard) if shard else W * H
        bufs = [torch.zeros(n, dtype=torch.int32).pin_memory() for _ in range(2)]
        views = [b.numpy().view(np.uint32) for b in bufs]
        got = []
        for i, P in enumerate(cams):
            slot = i & 1
            if i >= 2:
                dev.TraceCollect(slot)
                got.append(views[slot].copy())
            dev.TraceSubmit(P, views[slot], slot, shard=shard)
        for i in range(len(cams) - 2, len(cams)):
            dev.TraceCollect(i & 1)
            got.append(views[i & 1].copy())
        for i, P in enumerate(cams):
            exp = want[i].reshape(-1) if shard is None else dev.Trace(P, want=("rgba8",), shard=shard)["rgba8"]
            assert np.array_equal(got[i], exp), (i, shard)
    dev.close()


def test_beam_prepass_and_beam_trace(oracle, hd):
    """Row N3: beam.frag pre-pass and trace.frag with BEAM_OPTIMIZATION, bit-exact against their restatement;
    the beam-optimised frame equals the plain frame wherever both hit (the beam is conservative)."""
    cfg = abi.default_config(level_count=10, top_level_count=9)
    opool = oracle.pool(cfg)
    root = opool.edit_batch(NULL, [abi.terrain(cfg.voxel_level)] + abi.random_spheres(60, cfg.voxel_level, seed=9, rmin=8, rmax=60))
    dev = hd.DAGNodePool(cfg)
    dev.UploadFrom(opool)
    W, H = 645, 363   # not multiples of 8: partial beam texels
    for cam in (((0.5, 0.75, 0.5), 0.6, -0.5236), ((0.1, 0.6, 0.9), 2.3, -0.2), ((0.5, 1.05, 0.5), 0.0, -1.4)):
        P = abi.camera_params(cfg, root, *cam, W, H, color_root=(1 << 30) | 0x3377AA)
        B = abi.beam_params(P)
        assert (B.width, B.height) == ((W + 7) // 8, (H + 7) // 8)
        exp_beam = oracle.beam_frame(opool.words_ptr, B)
        exp = oracle.trace_frame(opool.words_ptr, P, beam=exp_beam)
        got = dev.TraceBeam(P, B)
        assert np.array_equal(got["beam"].view(np.uint32), exp_beam.view(np.uint32))
        for k in ("hits", "iters", "rgba8"):
            assert np.array_equal(got[k], exp[k]), (k, cam)
        plain = dev.Trace(P)
        assert plain["iters"].sum() > got["iters"].sum()          # fewer iterations: that is the optimisation
        both = (plain["hits"]["packed"] >> 31 == 1) & (got["hits"]["packed"] >> 31 == 1)
        same = (plain["hits"]["packed"] == got["hits"]["packed"]) & (plain["hits"]["vox"] == got["hits"]["vox"]).all(-1)
        assert both.sum() > 1000 and same[both].mean() > 0.98      # LOD bias differs by the beam term on a few pixels
    dev.close()


def test_exact_arithmetic_helpers(hd):
    """The trace kernel's shortcuts (unchecked reciprocal / square root, x / D with a folded reciprocal, the centre planes
    as one fma) return the bits of the IEEE operations the shader writes, for every input they can see."""
    assert hd.api.selftest_exact_arith(0) == 0


def test_lean_frames_equal_the_oracle_on_a_frame_size_change(oracle, hd):
    """The lean instantiation (rgba8 only: per-column / per-row ray tables, no iteration plane) against the oracle, with
    the frame size changing between calls (the ray tables are rebuilt) and on odd sizes."""
    cfg = abi.default_config(level_count=10, top_level_count=9)
    opool = oracle.pool(cfg)
    root = opool.edit_batch(NULL, [abi.terrain(cfg.voxel_level)] + abi.random_spheres(40, cfg.voxel_level, seed=4, rmin=8, rmax=70))
    dev = hd.DAGNodePool(cfg)
    dev.UploadFrom(opool)
    for (W, H), cam in (((640, 360), ((0.5, 0.8, 0.5), 0.6, -0.6)), ((333, 517), ((0.2, 0.7, 0.9), 2.3, -0.3)),
                        ((1280, 720), ((0.5, 1.05, 0.5), 0.1, -1.3)), ((640, 360), ((0.9, 0.6, 0.1), 4.0, -0.2))):
        for lod in (False, True):
            P = abi.camera_params(cfg, root, *cam, W, H, color_root=(1 << 30) | 0x80C040, lod=lod)
            exp = oracle.trace_frame(opool.words_ptr, P)
            got = dev.Trace(P, want=("rgba8",))
            assert np.array_equal(got["rgba8"], exp["rgba8"]), (W, H, cam, lod)
    dev.close()


def test_cta_shapes_and_experimental_kernels_are_bit_exact(oracle, hd, monkeypatch):
    """HD_TRACE_CTA, HD_TRACE_PERSIST and HD_TRACE_VARIANT are read per call: every CTA shape of the grid-per-patch kernel,
    the persistent chunk-queue kernel and the no-allocate leaf loads must shade exactly the frame the oracle shades, on
    odd frame sizes (partial patches, partial chunks) with and without LOD; the tile-shard kernels with both CTA shapes."""
    cfg = abi.default_config(level_count=10, top_level_count=9)
    opool = oracle.pool(cfg)
    root = opool.edit_batch(NULL, [abi.terrain(cfg.voxel_level)] + abi.random_spheres(40, cfg.voxel_level, seed=4, rmin=8, rmax=70))
    dev = hd.DAGNodePool(cfg)
    dev.UploadFrom(opool)
    knobs = [{}] + [{"HD_TRACE_CTA": c} for c in ("64", "128", "256", "2560", "512", "1024")] \
        + [{"HD_TRACE_PERSIST": c} for c in ("256", "512", "1024")] + [{"HD_TRACE_VARIANT": x} for x in ("5", "6")]
    for (W, H), cam in (((333, 517), ((0.2, 0.7, 0.9), 2.3, -0.3)), ((640, 360), ((0.5, 0.8, 0.5), 0.6, -0.6))):
        for lod in (False, True):
            P = abi.camera_params(cfg, root, *cam, W, H, color_root=(1 << 30) | 0x80C040, lod=lod)
            exp = oracle.trace_frame(opool.words_ptr, P)["rgba8"]
            for env in knobs:
                with monkeypatch.context() as m:
                    for k, val in env.items():
                        m.setenv(k, val)
                    got = dev.Trace(P, want=("rgba8",))["rgba8"]
                assert np.array_equal(got, exp), (W, H, lod, env)
            shards = []
            for cta in ("128", "256"):
                with monkeypatch.context() as m:
                    m.setenv("HD_TRACE_CTA", cta)
                    shards.append(dev.Trace(P, want=("rgba8",), shard=(64, 64, 1, 3))["rgba8"])
            assert np.array_equal(shards[0], shards[1]), (W, H, lod)
    dev.close()


def test_staged_top_levels_equal_the_pool_path(oracle, hd):
    """The staged copy of the top node levels (built once a root is traced for the second time) must not change a single
    output: frames, hit records, iteration counts and F before and after the table exists, across edits (new roots),
    a clear + re-upload, shards, and a DAG too shallow to stage anything."""
    import os
    os.environ["HD_TRACE_TABLE"] = "1"     # the library reads it per call; off by default (it loses on cfg2)
    try:
        _staged_top_levels(oracle, hd)
    finally:
        os.environ.pop("HD_TRACE_TABLE", None)


def _staged_top_levels(oracle, hd):
    import os
    cfg = abi.default_config(level_count=10, top_level_count=9)
    opool = oracle.pool(cfg)
    roots = [opool.edit_batch(NULL, [abi.terrain(cfg.voxel_level)] + abi.random_spheres(30, cfg.voxel_level, seed=2, rmin=8, rmax=80))]
    roots.append(opool.edit_batch(roots[0], abi.random_spheres(10, cfg.voxel_level, seed=3, rmin=20, rmax=90)))
    dev = hd.DAGNodePool(cfg)
    dev.UploadFrom(opool)
    W, H = 480, 270
    cams = [((0.5, 0.8, 0.5), 0.6, -0.6, False), ((0.2, 0.7, 0.9), 2.3, -0.3, True), ((0.5, 1.05, 0.5), 0.1, -1.3, False)]
    for rnd in range(2):
        for root in roots:
            for k in range(3):            # the 2nd and 3rd frame of a root read the table
                for cam in cams:
                    P = abi.camera_params(cfg, root, *cam[:3], W, H, color_root=(1 << 30) | 0x4080C0, lod=cam[3])
                    exp = oracle.trace_frame(opool.words_ptr, P)
                    got = dev.Trace(P, want=("rgba8", "hits", "iters", "fetches"))
                    for key in ("rgba8", "hits", "iters"):
                        assert np.array_equal(got[key], exp[key]), (rnd, root, k, cam, key)
                    assert int(got["fetches"].sum(dtype=np.uint64)) == exp["fetches"], (rnd, root, k, cam, "F")
                    lean = dev.Trace(P, want=("rgba8",))
                    assert np.array_equal(lean["rgba8"], exp["rgba8"]), (rnd, root, k, cam, "lean")
                    sh = dev.Trace(P, want=("rgba8",), shard=(64, 64, 1, 2))
                    ref_sh = oracle_shard(exp["rgba8"], P, (64, 64, 1, 2), dev)
                    assert np.array_equal(sh["rgba8"], ref_sh), (rnd, root, k, cam, "shard")
            if os.environ.get("HD_TRACE_VARIANT", "0") == "0":   # a forced A/B variant never stages anything
                troot, tlevels, tnodes = dev.TraceTableInfo()   # the last frames of this root really read the table
                assert troot == root and 1 <= tlevels <= cfg.node_levels - 2 and tnodes > tlevels, (troot, tlevels, tnodes)
        dev.Clear()                        # invalidates the table; the same pointers come back with the re-upload
        dev.UploadFrom(opool)
    dev.close()
    # 2 node levels: nothing can be staged, the pool path serves every frame
    cfg2 = abi.default_config(level_count=3, top_level_count=9)
    op2 = oracle.pool(cfg2)
    r2 = op2.edit_batch(NULL, [abi.sphere((4, 4, 4), 9)])
    d2 = hd.DAGNodePool(cfg2)
    d2.UploadFrom(op2)
    for k in range(3):
        P = abi.camera_params(cfg2, r2, (0.5, 0.5, 1.6), np.pi, 0.0, 160, 90, color_root=(1 << 30) | 0x112233)
        exp = oracle.trace_frame(op2.words_ptr, P)
        got = d2.Trace(P)
        for key in ("rgba8", "hits", "iters"):
            assert np.array_equal(got[key], exp[key]), (k, key)
    d2.close()


def oracle_shard(frame, P, shard, dev):
    """Rearrange a full frame into the tile-major layout hd_trace_tiles returns for `shard` = (tile_w, tile_h, rank, world)."""
    from vkhashdag_b200 import replica
    tw, th, rank, world = shard
    H, W = frame.shape
    tiles = replica.local_tiles(W, H, tw, th, rank, world)
    out = np.zeros(len(tiles) * tw * th, np.uint32)
    for lt, tx, ty in tiles:
        tile = np.zeros((th, tw), np.uint32)
        part = frame[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
        tile[:part.shape[0], :part.shape[1]] = part
        out[lt * tw * th:(lt + 1) * tw * th] = tile.reshape(-1)
    return out
