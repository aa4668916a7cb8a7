"""More GPU parity: colour decode in the kernel, the fetch counter, replica sync on one device, BASELINE-size
properties (cfg2: 2^15 terrain, 3840x2160)."""
import os

import numpy as np
import pytest
import torch

from vkhashdag_b200 import abi

pytestmark = pytest.mark.gpu
NULL = abi.NULL
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def mirror_of(oracle, dev, cfg):
    m = oracle.pool(cfg)
    ranges, bw = dev.Download()
    for off, words in ranges.items():
        m.words_np(off, len(words))[:] = words
    m.bucket_words_np()[:] = bw
    return m


def test_colour_scene_from_reference_fixture(oracle, hd):
    """Node pool + colour pool written by the REFERENCE (VBREditorWrapper, VBRChunkWriter) -> kernel colour per pixel."""
    z = np.load(os.path.join(GOLD, "color_scene.npz"))
    cfg = abi.HdConfig()
    cfg.word_bits_per_page, cfg.page_bits_per_bucket = 9, 2
    cfg.node_levels = len(z["bucket_bits"])
    for i, b in enumerate(z["bucket_bits"]):
        cfg.bucket_bits_each_level[i] = int(b)
    host = oracle.pool(cfg)
    pos = 0
    for off, n in zip(z["range_offsets"].tolist(), z["range_lengths"].tolist()):
        host.words_np(off, n)[:] = z["words"][pos:pos + n]
        pos += n
    host.bucket_words_np()[:] = z["bucket_words"]
    dev = hd.DAGNodePool(cfg)
    dev.UploadFrom(host)
    cn, cl = z["color_nodes"], z["color_leaves"]
    dev.UploadColor(cn, cl)
    root, croot, cleaf = int(z["node_root"]), int(z["color_root"]), int(z["leaf_level"])
    n_colors = set()
    for cam in (((0.45, 0.6, 1.4), np.pi, -0.25), ((1.3, 0.7, 0.4), -1.9, -0.3), ((0.4, 0.95, 0.45), 0.3, -1.3)):
        for lod in (True, False):
            P = abi.camera_params(cfg, root, *cam, 320, 180, color_root=croot, color_leaf_level=cleaf, lod=lod)
            exp = oracle.trace_frame(host.words_ptr, P, cn, cl)
            got = dev.Trace(P, want=("rgba8", "hits", "iters", "fetches"))
            assert np.array_equal(got["hits"], exp["hits"])
            assert np.array_equal(got["rgba8"], exp["rgba8"])
            assert np.array_equal(got["iters"], exp["iters"])
            assert int(got["fetches"].sum(dtype=np.uint64)) == exp["fetches"]     # F of SURVEY §8d, word for word
            n_colors |= set(np.unique(got["hits"]["packed"][got["hits"]["packed"] >> 31 == 1] & 0xFFFFFF).tolist())
    assert len(n_colors) >= 4
    dev.close()


def test_replica_sync_pack_apply_same_device(oracle, hd):
    """hd_dirty_pack_dev -> (the broadcast would go here) -> hd_dirty_apply_dev reproduces the pool on a replica."""
    cfg = abi.default_config(level_count=9, top_level_count=9)
    a, b = hd.DAGNodePool(cfg), hd.DAGNodePool(cfg)
    stage = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    root = NULL
    batches = [[abi.terrain(cfg.voxel_level)], abi.random_spheres(20, cfg.voxel_level, seed=4, rmin=4, rmax=30),
               [abi.sphere((256, 200, 256), 50 ** 2, dig=True)]]
    for edits in batches:
        root = a.EditBatch(root, edits)
        a.SetRoot(root)
        n_ranges, need = a.DirtyCount()
        assert n_ranges > 0 and sum(c for _, c in a.DirtyRanges()) * 4 + 32 + 12 * n_ranges == need
        n = a.DirtyPack(stage.data_ptr(), stage.numel())
        assert n == need
        a.DirtyReset()
        assert a.DirtyCount()[0] == 0
        b.DirtyApply(stage.data_ptr(), n)
        assert b.GetRoot() == root
        assert np.array_equal(a.ReadBucketWords(), b.ReadBucketWords())
    ma, mb = mirror_of(oracle, a, cfg), mirror_of(oracle, b, cfg)
    for off, cnt in ma.used_ranges():
        assert np.array_equal(ma.words_np(off, cnt), mb.words_np(off, cnt))
    P = abi.camera_params(cfg, root, (0.5, 0.8, 0.5), 0.6, -0.6, 256, 144)
    fa, fb = a.Trace(P), b.Trace(P)
    assert np.array_equal(fa["hits"], fb["hits"]) and (fa["hits"]["packed"] >> 31).sum() > 1000
    # a replica can keep editing after a sync (its bucket cursors moved with the data)
    e = [abi.sphere((100, 180, 100), 25 ** 2)]
    ra, rb = a.EditBatch(root, e), b.EditBatch(root, e)
    ma, mb = mirror_of(oracle, a, cfg), mirror_of(oracle, b, cfg)   # keep the mirrors alive while they are walked
    ca = oracle.canonical(ma.words_ptr, cfg.node_levels, ra)
    cb = oracle.canonical(mb.words_ptr, cfg.node_levels, rb)
    assert ca == cb
    a.close(), b.close()


@pytest.fixture(scope="module")
def cfg2_scene(hd):
    import bench
    cfg = bench.scene_config()
    dev = hd.DAGNodePool(cfg)
    root = dev.Edit(NULL, hd.TerrainEditor(cfg.voxel_level))
    stats = dict(dev.last_stats)
    yield cfg, dev, root, stats
    dev.close()


def test_cfg2_build_properties(oracle, hd, cfg2_scene):
    """BASELINE config 2 at full size: no overflow, idempotent, voxel columns equal the height function."""
    cfg, dev, root, stats = cfg2_scene
    assert stats["overflow_count"] == 0 and root != NULL
    assert stats["visited_leaves"] > 100_000_000
    again = dev.Edit(root, hd.TerrainEditor(cfg.voxel_level))
    assert again == root and dev.last_stats["appended_nodes"] == 0          # idempotence at full size
    m = mirror_of(oracle, dev, cfg)
    t = abi.terrain(cfg.voxel_level)
    rng = np.random.default_rng(15)
    res = 1 << cfg.voxel_level
    for x, z in rng.integers(0, res, (1500, 2)).tolist():
        h = oracle.terrain_height(t, x, z)
        for y in (h - 1, h, 0, res - 1, int(rng.integers(0, res))):
            assert oracle.voxel_get(m.words_ptr, cfg.node_levels, root, x, y, z) == (y < h)
    cfg2_scene[1].mirror = m


def test_cfg2_4k_trace_sample_rows_vs_oracle(oracle, hd, cfg2_scene):
    """3840x2160 on the 2^15 terrain: every 45th row of the frame is checked ray by ray against the oracle;
    the whole frame is deterministic across launches and independent of tile sharding."""
    import bench
    cfg, dev, root, _ = cfg2_scene
    m = getattr(dev, "mirror", None) or mirror_of(oracle, dev, cfg)
    for step, lod in ((0, False), (3, True)):
        P = bench.camera(cfg, root, step, 3840, 2160, lod)
        got = dev.Trace(P, want=("rgba8", "hits", "iters"))
        exp = oracle.trace_frame(m.words_ptr, P, rows=(7, 2160), row_step=45)
        rows = np.arange(7, 2160, 45)
        for k in ("rgba8", "hits", "iters"):
            assert np.array_equal(got[k][rows], exp[k][rows]), (k, step, lod)
        assert (got["hits"]["packed"][rows] >> 31).mean() > 0.2
        again = dev.Trace(P, want=("rgba8",))
        assert np.array_equal(again["rgba8"], got["rgba8"])
        from vkhashdag_b200 import replica
        parts = [dev.Trace(P, want=("rgba8",), shard=(64, 64, r, 4))["rgba8"] for r in range(4)]
        assert np.array_equal(replica.assemble_frame(parts, 3840, 2160, 64, 64, 4), got["rgba8"])


def test_two_gpu_replica_if_available(oracle, hd):
    """With >= 2 devices: edit on cuda:0, copy the packed dirty ranges peer-to-peer, apply on cuda:1, same frame."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cfg = abi.default_config(level_count=9, top_level_count=9)
    a, b = hd.DAGNodePool(cfg, device=0), hd.DAGNodePool(cfg, device=1)
    root = a.EditBatch(NULL, [abi.terrain(cfg.voxel_level)] + abi.random_spheres(30, cfg.voxel_level, seed=2, rmin=4, rmax=30))
    a.SetRoot(root)
    s0 = torch.empty(64 << 20, dtype=torch.uint8, device="cuda:0")
    n = a.DirtyPack(s0.data_ptr(), s0.numel())
    s1 = s0[:n].to("cuda:1")
    torch.cuda.synchronize()
    b.DirtyApply(s1.data_ptr(), n)
    P = abi.camera_params(cfg, root, (0.5, 0.8, 0.5), 0.6, -0.6, 256, 144)
    assert np.array_equal(a.Trace(P)["hits"], b.Trace(P)["hits"])
    a.close(), b.close()


def test_torchrun_two_gpu_nccl_replica_sync():
    """One process per GPU under torchrun (2 ranks): scene built on rank 0 and published with the NCCL dirty-range
    broadcast (replaces DAGNodePool::Flush, src/DAGNodePool.cpp:56-85, for the replicated pool), then per frame a brush
    edit on rank 0 -> ONE broadcast -> tile-sharded trace; the stitched frame must equal rank 0's full-frame trace and
    every replica must hold the same bucket cursors and root (ordering contract of src/main.cpp:281-296)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(root, "tools", "multi_gpu_check.py"), "--level", "11", "--frames", "4",
           "--width", "1280", "--height", "720"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI-GPU CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_read_subtree_matches_word_reads(oracle, hd):
    """hd_pool_read_subtree: breadth-first records of a subtree == the same nodes read word by word (what a host-side
    visitor such as NodePoolBase::Iterate needs, test/test.cpp:36-52); depth and capacity limits are honoured."""
    cfg = abi.default_config(level_count=7, top_level_count=9)
    dev = hd.DAGNodePool(cfg)
    root = dev.EditBatch(NULL, [abi.sphere((60, 60, 60), 40 ** 2), abi.aabb((5, 5, 5), (30, 20, 90))])
    recs, truncated = dev.ReadSubtree(root)
    assert not truncated and recs[0][0] == root and recs[0][1] == 0
    seen = {}
    for ptr, level, words in recs:
        leaf = level == cfg.node_levels - 1
        n = 2 if leaf else 1 + bin(int(dev.ReadWords(ptr, 1)[0]) & 0xFF).count("1")
        assert words == [int(w) for w in dev.ReadWords(ptr, n)]
        seen.setdefault(level, set()).add(ptr)
    m = mirror_of(oracle, dev, cfg)
    canon = oracle.canonical(m.words_ptr, cfg.node_levels, root)
    assert [len(seen[l]) for l in range(cfg.node_levels)] == canon["per_level"]     # every reachable node, nothing else
    top, _ = dev.ReadSubtree(root, depth=2)
    assert {lvl for _, lvl, _ in top} == {0, 1} and len(top) == 1 + len(recs[0][2]) - 1
    few, truncated = dev.ReadSubtree(root, capacity=10)
    assert truncated and len(few) == 10
    assert dev.ReadSubtree(NULL)[0] == []
    dev.close()


def test_host_buffers_for_frame_readback(hd):
    """hd_host_alloc: page-locked (optionally write-combined) frame buffers for hd_trace_submit / collect."""
    cfg = abi.default_config(level_count=7, top_level_count=9)
    dev = hd.DAGNodePool(cfg)
    root = dev.Edit(NULL, hd.SphereEditor((64, 64, 64), 40 ** 2))
    P = abi.camera_params(cfg, root, (0.5, 0.5, 1.6), np.pi, 0.0, 128, 72, color_root=(1 << 30) | 0x4080C0)
    ref = dev.Trace(P, want=("rgba8",))["rgba8"].reshape(-1)
    for wc in (False, True):
        buf = hd.HostBuffer(128 * 72, write_combined=wc)
        dev.TraceSubmit(P, buf.array, 0)
        dev.TraceCollect(0)
        assert np.array_equal(np.array(buf.array), ref)
        buf.free()
    dev.close()
