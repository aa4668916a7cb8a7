"""GPU GC/compaction (row N1) and pool serialisation (row N4): canonical DAG preserved, nothing unreachable left."""
import os

import numpy as np
import pytest
import torch

from vkhashdag_b200 import abi

pytestmark = pytest.mark.gpu
NULL = abi.NULL


def mirror_of(oracle, dev, cfg):
    m = oracle.pool(cfg)
    ranges, bw = dev.Download()
    for off, words in ranges.items():
        m.words_np(off, len(words))[:] = words
    m.bucket_words_np()[:] = bw
    return m


def scene_edits(cfg):
    vl = cfg.voxel_level
    return [abi.terrain(vl)] + abi.random_spheres(40, vl, seed=3, rmin=4, rmax=40)


def test_gc_keeps_dag_drops_garbage(oracle, hd):
    cfg = abi.default_config(level_count=9, top_level_count=9)
    dev = hd.DAGNodePool(cfg)
    root = NULL
    for e in scene_edits(cfg):           # one call per edit: every intermediate version leaves garbage behind
        root = dev.Edit(root, e)
    dev.SetRoot(root)
    m0 = mirror_of(oracle, dev, cfg)
    before = oracle.canonical(m0.words_ptr, cfg.node_levels, root)
    stored0 = oracle.count_stored_nodes(m0)
    P = abi.camera_params(cfg, root, (0.5, 0.8, 0.5), 0.6, -0.6, 256, 144)
    frame0 = dev.Trace(P)

    new_root = dev.ThreadedGC(root)
    assert dev.GetRoot() == new_root
    m1 = mirror_of(oracle, dev, cfg)
    after = oracle.canonical(m1.words_ptr, cfg.node_levels, new_root)
    assert after == before                                     # same DAG: hash, unique counts, voxels, per level
    stored1 = oracle.count_stored_nodes(m1)
    assert sum(stored1) < sum(stored0)
    # exactly the reachable nodes + the filled chain survive (NodePool.hpp:54)
    filled_chain = oracle.canonical(m1.words_ptr, cfg.node_levels, dev.FilledNodes()[0])
    assert sum(stored1) == dev.last_gc_nodes
    assert all(s >= a for s, a in zip(stored1, after["per_level"]))
    assert sum(stored1) <= after["by_ptr"] + filled_chain["by_ptr"]
    # the frame is unchanged
    P2 = abi.camera_params(cfg, new_root, (0.5, 0.8, 0.5), 0.6, -0.6, 256, 144)
    frame1 = dev.Trace(P2)
    assert np.array_equal(frame0["hits"], frame1["hits"]) and np.array_equal(frame0["rgba8"], frame1["rgba8"])
    # editing continues to work on the compacted pool and stays canonical
    more = abi.random_spheres(10, cfg.voxel_level, seed=77, rmin=4, rmax=30)
    r2 = dev.EditBatch(new_root, more)
    opool = oracle.pool(cfg)
    oroot = opool.edit_batch(NULL, scene_edits(cfg) + more)
    m2 = mirror_of(oracle, dev, cfg)
    assert oracle.canonical(m2.words_ptr, cfg.node_levels, r2) == opool.canonical(oroot)
    dev.close()


def test_gc_matches_reference_gc(oracle, ref, hd):
    """Same scene, reference ThreadedGC vs hd_gc: identical stored-node census per level, identical DAG."""
    cfg = abi.default_config(level_count=9, top_level_count=9)
    rp = ref.pool(cfg)
    dev = hd.DAGNodePool(cfg)
    rr = gr = NULL
    for e in scene_edits(cfg):
        rr, gr = rp.edit(rr, e), dev.Edit(gr, e)
    rr2 = rp.gc(rr, threads=4)
    gr2 = dev.ThreadedGC(gr)
    m = mirror_of(oracle, dev, cfg)
    assert oracle.canonical(m.words_ptr, cfg.node_levels, gr2) == oracle.canonical(rp.words_ptr, cfg.node_levels, rr2)
    assert oracle.count_stored_nodes(m) == oracle.count_stored_nodes(rp)
    assert int(m.bucket_words_np().sum()) <= int(rp.bucket_words_np().sum()) * 1.01 + 64   # same words up to page padding
    dev.close()


def test_gc_null_root_and_multiple_roots(oracle, hd):
    cfg = abi.default_config(level_count=7, top_level_count=9)
    dev = hd.DAGNodePool(cfg)
    a = dev.Edit(NULL, hd.SphereEditor((60, 60, 60), 30 ** 2))
    b = dev.Edit(a, hd.SphereEditor((70, 60, 60), 20 ** 2, "dig"))
    c = dev.Edit(NULL, hd.AABBEditor((10, 10, 10), (90, 20, 90)))
    m0 = mirror_of(oracle, dev, cfg)
    ca, cb = (oracle.canonical(m0.words_ptr, cfg.node_levels, r) for r in (a, b))
    na, nb, nn = dev.ThreadedGC([a, b, NULL])          # c becomes garbage, a and b (sharing structure) survive
    assert nn == NULL
    m1 = mirror_of(oracle, dev, cfg)
    assert oracle.canonical(m1.words_ptr, cfg.node_levels, na) == ca
    assert oracle.canonical(m1.words_ptr, cfg.node_levels, nb) == cb
    assert dev.ThreadedGC(NULL) == NULL                # only the filled chain remains
    m2 = mirror_of(oracle, dev, cfg)
    assert oracle.count_stored_nodes(m2) == [1] * cfg.node_levels
    dev.close()


def test_replica_resync_after_gc(oracle, hd):
    cfg = abi.default_config(level_count=8, top_level_count=9)
    a, b = hd.DAGNodePool(cfg), hd.DAGNodePool(cfg)
    stage = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
    root = NULL
    for e in scene_edits(cfg)[:12]:
        root = a.Edit(root, e)
    a.SetRoot(root)
    n = a.DirtyPack(stage.data_ptr(), stage.numel())
    a.DirtyReset()
    b.DirtyApply(stage.data_ptr(), n)
    new_root = a.ThreadedGC(root)
    n = a.DirtyPack(stage.data_ptr(), stage.numel())      # carries the clear-first flag
    a.DirtyReset()
    b.DirtyApply(stage.data_ptr(), n)
    assert b.GetRoot() == new_root
    assert np.array_equal(a.ReadBucketWords(), b.ReadBucketWords())
    ma, mb = mirror_of(oracle, a, cfg), mirror_of(oracle, b, cfg)
    assert oracle.count_stored_nodes(ma) == oracle.count_stored_nodes(mb)
    assert oracle.canonical(ma.words_ptr, cfg.node_levels, new_root) == oracle.canonical(mb.words_ptr, cfg.node_levels, new_root)
    a.close(), b.close()


def test_save_load_roundtrip(oracle, hd, tmp_path):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "color_scene.npz"))
    cfg = abi.custom_config([int(b) for b in z["bucket_bits"]])
    host = oracle.pool(cfg)
    pos = 0
    for off, n in zip(z["range_offsets"].tolist(), z["range_lengths"].tolist()):
        host.words_np(off, n)[:] = z["words"][pos:pos + n]
        pos += n
    host.bucket_words_np()[:] = z["bucket_words"]
    dev = hd.DAGNodePool(cfg)
    dev.UploadFrom(host)
    dev.UploadColor(z["color_nodes"], z["color_leaves"])
    root, croot, cleaf = int(z["node_root"]), int(z["color_root"]), int(z["leaf_level"])
    dev.SetRoot(root)
    path = str(tmp_path / "scene.hdag")
    dev.Save(path)
    assert os.path.getsize(path) > 4 * int(z["bucket_words"].sum())
    back = hd.DAGNodePool.Load(path)
    assert back.GetRoot() == root and back.GetConfig().bucket_bits() == cfg.bucket_bits()
    assert np.array_equal(back.ReadBucketWords(), dev.ReadBucketWords())
    P = abi.camera_params(cfg, root, (0.45, 0.6, 1.4), np.pi, -0.25, 320, 180, color_root=croot, color_leaf_level=cleaf)
    f0, f1 = dev.Trace(P), back.Trace(P)
    for k in ("rgba8", "hits", "iters"):
        assert np.array_equal(f0[k], f1[k])
    assert len(np.unique(f1["hits"]["packed"] & 0xFFFFFF)) > 4
    # a loaded pool is editable
    r2 = back.Edit(root, hd.SphereEditor((128, 100, 128), 20 ** 2, "dig"))
    assert r2 != root
    with pytest.raises(hd.HashDagError):
        hd.DAGNodePool.Load(str(tmp_path / "missing.hdag"))
    dev.close(), back.close()
