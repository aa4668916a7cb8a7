"""GPU GC/compaction (row N1) and pool serialisation (row N4): canonical DAG preserved, nothing unreachable left."""
import os

import numpy as np
import pytest
import torch

from vkhashdag_b200 import abi, replica

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


def test_gc_overflow_leaves_pool_untouched(oracle, hd, monkeypatch):
    """A bucket that fills up during the rebuild (remapped children re-randomise the inner-node hashes) must not corrupt
    the pool: hd_gc builds the compacted pool beside the old one and returns HD_ERR_OVERFLOW with everything unchanged.
    The overflow is injected at an inner level through the library's test hook."""
    cfg = abi.default_config(level_count=8, top_level_count=9)
    dev = hd.DAGNodePool(cfg)
    root = NULL
    for e in scene_edits(cfg)[:15]:
        root = dev.Edit(root, e)
    dev.SetRoot(root)
    m0 = mirror_of(oracle, dev, cfg)
    before = oracle.canonical(m0.words_ptr, cfg.node_levels, root)
    bw0 = dev.ReadBucketWords()
    monkeypatch.setenv("HD_GC_INJECT_OVERFLOW_LEVEL", "3")
    with pytest.raises(hd.HashDagError) as ei:
        dev.ThreadedGC(root)
    assert ei.value.status == 4          # HD_ERR_OVERFLOW
    monkeypatch.delenv("HD_GC_INJECT_OVERFLOW_LEVEL")
    assert dev.GetRoot() == root and np.array_equal(dev.ReadBucketWords(), bw0)
    m1 = mirror_of(oracle, dev, cfg)
    assert oracle.canonical(m1.words_ptr, cfg.node_levels, root) == before
    r2 = dev.Edit(root, hd.SphereEditor((100, 90, 100), 15 ** 2))     # still editable, then compactable
    r3 = dev.ThreadedGC(r2)
    m = mirror_of(oracle, dev, cfg)
    opool = oracle.pool(cfg)
    oroot = opool.edit_batch(NULL, scene_edits(cfg)[:15] + [abi.sphere((100, 90, 100), 15 ** 2)])
    assert oracle.canonical(m.words_ptr, cfg.node_levels, r3) == opool.canonical(oroot)
    dev.close()


def test_load_rejects_corrupt_files(hd, tmp_path):
    """hd_pool_load trusts nothing in the file: sizes are checked against the file length before anything is allocated, and
    every packed range is bounds-checked on the device before the first write."""
    import struct
    cfg = abi.default_config(level_count=7, top_level_count=9)
    dev = hd.DAGNodePool(cfg)
    root = dev.EditBatch(NULL, [abi.terrain(cfg.voxel_level), abi.sphere((60, 50, 60), 20 ** 2)])
    dev.SetRoot(root)
    good = str(tmp_path / "good.hdag")
    dev.Save(good)
    raw = bytearray(open(good, "rb").read())
    hd.DAGNodePool.Load(good).close()
    head = 8 + 4 + 4 * (3 + abi.MAX_LEVELS)          # magic, version, hd_config
    head += (-head) % 8                              # the u64 fields are 8-byte aligned
    blob_bytes, = struct.unpack_from("<Q", raw, head)
    blob0 = head + 24 + 8                            # three u64 sizes, colour root + leaf level

    def bad(name, data):
        path = str(tmp_path / name)
        open(path, "wb").write(bytes(data))
        with pytest.raises(hd.HashDagError) as ei:
            hd.DAGNodePool.Load(path)
        assert ei.value.status == 1, name          # HD_ERR_INVALID, not a crash / CUDA error

    bad("truncated.hdag", raw[:len(raw) - 40])
    huge = bytearray(raw)
    struct.pack_into("<Q", huge, head, 1 << 60)      # blob size far beyond the file
    bad("huge_blob.hdag", huge)
    cwords = bytearray(raw)
    struct.pack_into("<Q", cwords, head + 8, 1 << 61)  # colour node words: would throw length_error if trusted
    bad("huge_colour.hdag", cwords)
    assert blob_bytes == len(raw) - blob0
    n_ranges, = struct.unpack_from("<I", raw, blob0)
    assert n_ranges > 4
    tri = blob0 + 32
    oob = bytearray(raw)
    struct.pack_into("<I", oob, tri, 0xFFFFFF00)     # first range: word offset outside the pool
    bad("range_outside_pool.hdag", oob)
    straddle = bytearray(raw)
    struct.pack_into("<I", straddle, tri + 4, 5000)  # first range: longer than a bucket (and than the payload)
    bad("range_past_bucket.hdag", straddle)
    many = bytearray(raw)
    struct.pack_into("<I", many, blob0, 0x7FFFFFFF)  # more ranges than the blob holds
    bad("too_many_ranges.hdag", many)
    dev.close()


def test_replica_sync_carries_colour_edits(oracle, hd):
    """Coloured edits on the editing pool reach a replica through the SAME packed blob: appended colour nodes, appended leaf
    chunks, and chunks rewritten in place (keep_history = false); the replica's colour buffers equal the editor's word for
    word and both render the same frame (replaces DAGColorPool::Flush for the replicated pool, src/main.cpp:245-251)."""
    cfg = abi.default_config(level_count=8, top_level_count=9)
    a, b = hd.DAGNodePool(cfg), hd.DAGNodePool(cfg)
    a.ColorConfig(4), b.ColorConfig(4)
    stage = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    root = a.Edit(NULL, abi.terrain(cfg.voxel_level))
    steps = [(abi.aabb((0, 0, 0), (256, 70, 256)), 0x40A040, False), (abi.sphere((128, 90, 128), 40 ** 2), 0xC08040, False),
             (abi.sphere((100, 80, 120), 30 ** 2), 0x2040F0, True), (abi.sphere((100, 80, 120), 12 ** 2), 0xF0F020, True),
             (abi.sphere((100, 80, 120), 12 ** 2), 0x00FF00, True),      # same shape, new colour: fits the old chunks
             (abi.sphere((105, 82, 118), 9 ** 2), 0x10F0F0, True), (abi.sphere((60, 75, 60), 22 ** 2), 0xF0F020, False)]
    in_place = 0
    for i, (desc, rgb, paint) in enumerate(steps):
        lw0 = a.ReadColor()[1].size
        root, croot = a.EditColor(root, desc, rgb, paint)
        a.SetRoot(root)
        n = a.DirtyPack(stage.data_ptr(), stage.numel())
        blob = stage[:n].view(torch.int32).cpu().numpy().view(np.uint32)
        head = blob[:8]
        assert head[3] & 2 and n == replica.packed_bytes(head)
        sec = 8 + 3 * int(head[0]) + int(head[1])
        tri = blob[sec + 2:sec + 2 + 3 * int(head[4])].reshape(-1, 3)
        # leaf-tagged ranges that start below the previously synchronised leaf words are chunks rewritten in place
        in_place += int(((tri[:, 0] >> 31 == 1) & ((tri[:, 0] & 0x7FFFFFFF) < lw0)).sum()) if i else 0
        a.DirtyReset()
        b.DirtyApply(stage.data_ptr(), n)
        assert b.GetRoot() == root and b.ColorRoot() == croot
        (an, al), (bn, bl) = a.ReadColor(), b.ReadColor()
        assert np.array_equal(an, bn) and np.array_equal(al, bl), i
        P = abi.camera_params(cfg, root, (0.5, 0.7, 1.3), np.pi, -0.35, 320, 180, color_root=croot, color_leaf_level=4)
        fa, fb = a.Trace(P), b.Trace(P)
        assert np.array_equal(fa["rgba8"], fb["rgba8"]) and np.array_equal(fa["hits"], fb["hits"])
    assert in_place >= 1
    # a geometry-only edit afterwards carries no colour section
    root = a.Edit(root, hd.SphereEditor((150, 95, 150), 10 ** 2, "dig"))
    a.SetRoot(root)
    n = a.DirtyPack(stage.data_ptr(), stage.numel())
    assert not stage[:32].view(torch.int32).cpu().numpy().view(np.uint32)[3] & 2
    b.DirtyApply(stage.data_ptr(), n)
    # a colour root that does not exist on a pool is refused instead of read out of bounds
    c = hd.DAGNodePool(cfg)
    c.ColorConfig(4)
    c.Edit(NULL, abi.terrain(cfg.voxel_level))
    with pytest.raises(hd.HashDagError):
        c.Trace(abi.camera_params(cfg, c.GetRoot(), (0.5, 0.7, 1.3), np.pi, -0.35, 64, 36, color_root=croot, color_leaf_level=4))
    a.close(), b.close(), c.close()
