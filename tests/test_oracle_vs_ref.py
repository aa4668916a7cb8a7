"""Live differential tests: the CPU restatement against the reference's own headers compiled in place (oracle/_ref).
Skipped where neither the reference tree nor a prebuilt oracle/_ref exists."""
import numpy as np
import pytest

from oracle import bindings as B


def canon(oracle, pool, root):
    return oracle.canonical(pool.words_ptr, pool.cfg.node_levels, root)


def test_random_edit_sequences_pointer_exact(oracle, ref):
    """Serial Edit: identical pointers, identical pool words, for random AABB/sphere sequences."""
    rng = np.random.default_rng(11)
    for trial in range(6):
        cfg = B.default_config(level_count=int(rng.integers(5, 9)), top_level_count=9)
        res = 1 << cfg.voxel_level
        edits = []
        for _ in range(12):
            if rng.random() < 0.3:
                lo = rng.integers(0, res, 3)
                hi = np.minimum(lo + rng.integers(1, res // 2, 3), res)
                edits.append(B.aabb([int(v) for v in lo], [int(v) for v in hi]))
            else:
                r = int(rng.integers(1, res // 3))
                edits.append(B.sphere([int(v) for v in rng.integers(0, res, 3)], r * r, dig=bool(rng.random() < 0.4)))
        op, rp = oracle.pool(cfg), ref.pool(cfg)
        ro = rr = B.NULL
        for e in edits:
            ro, rr = op.edit(ro, e), rp.edit(rr, e)
            assert ro == rr
        assert np.array_equal(op.bucket_words_np(), rp.bucket_words_np())
        for off, cnt in op.used_ranges():
            assert np.array_equal(op.words_np(off, cnt), rp.words_np(off, cnt))
        assert canon(oracle, op, ro) == canon(oracle, rp, rr)


def test_threaded_edit_canonical_equal(oracle, ref):
    """ThreadedEdit returns different pointers but the same canonical DAG == the serial restatement."""
    cfg = B.default_config(level_count=9, top_level_count=9)
    edits = [B.terrain(cfg.voxel_level)] + B.random_spheres(40, cfg.voxel_level, seed=5, rmin=4, rmax=40)
    op, rp = oracle.pool(cfg), ref.pool(cfg)
    ro = op.edit_batch(B.NULL, edits)
    rr = rp.edit_batch(B.NULL, edits, threads=4, max_task_level=5)
    a, b = canon(oracle, op, ro), canon(oracle, rp, rr)
    assert a == b and a["by_ptr"] == a["by_content"]


def test_terrain_editor_pointer_exact(oracle, ref):
    cfg = B.default_config(level_count=8, top_level_count=9)
    for t in (B.terrain(cfg.voxel_level), B.terrain(cfg.voxel_level, seed=77, octaves=3, amp_div=4),
              B.terrain(cfg.voxel_level, extent_bits=cfg.voxel_level - 1)):
        op, rp = oracle.pool(cfg), ref.pool(cfg)
        assert op.edit(B.NULL, t) == rp.edit(B.NULL, t)
        assert np.array_equal(op.bucket_words_np(), rp.bucket_words_np())


def test_traversal_random_rays(oracle, ref):
    cfg = B.default_config(level_count=9, top_level_count=9)
    rp = ref.pool(cfg)
    root = rp.edit_batch(B.NULL, [B.terrain(cfg.voxel_level)] + B.random_spheres(30, cfg.voxel_level, seed=8, rmin=4, rmax=50))
    rng = np.random.default_rng(3)
    hits = 0
    for _ in range(1500):
        o = rng.uniform(-0.3, 1.3, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        d = d / np.float32(np.sqrt(np.float32(np.dot(d, d))))
        a, b = oracle.traverse(rp.words_ptr, cfg.node_levels, root, o, d), rp.traverse(root, o, d)
        assert (a is None) == (b is None)
        if a is not None:
            hits += 1
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert hits > 200


def test_frame_host_tracer_equal(oracle, ref):
    cfg = B.default_config(level_count=9, top_level_count=9)
    rp = ref.pool(cfg)
    root = rp.edit(B.NULL, B.terrain(cfg.voxel_level))
    for cam in (((0.5, 0.8, 0.5), 0.6, -0.6), ((0.1, 0.5, 0.9), 2.2, -0.1)):
        P = B.camera_params(cfg, root, *cam, 128, 72)
        a = oracle.trace_frame_host(rp.words_ptr, cfg.node_levels, P, threads=2)
        b = rp.trace_frame_host(P, threads=2)
        assert np.array_equal(a["hit"], b["hit"]) and a["n_hits"] > 500
        assert np.array_equal(a["pos"].view(np.uint32) * a["hit"][..., None], b["pos"].view(np.uint32) * b["hit"][..., None])


def test_colour_pool_decode(oracle, ref):
    """Colour octree + VBR leaves written by the reference's VBREditorWrapper/VBRChunkWriter, decoded by the
    trace.frag restatement, compared with the reference's own VBRChunkIterator."""
    cfg = B.default_config(level_count=7, top_level_count=9)
    vl = cfg.voxel_level
    pool, cp = ref.pool(cfg), ref.color_pool(leaf_level=3)
    root = B.NULL
    root = pool.edit_color(cp, root, B.aabb((5, 5, 5), (100, 40, 90)), 0x8040C0)
    root = pool.edit_color(cp, root, B.sphere((60, 50, 60), 30 ** 2), 0x10F0A0)
    root = pool.edit_color(cp, root, B.sphere((40, 30, 50), 25 ** 2), 0xFF2010, paint=True)
    root = pool.edit(root, B.sphere((70, 45, 70), 18 ** 2, dig=True))
    cn, cl = cp.arrays()
    rng = np.random.default_rng(1)
    checked = 0
    for x, y, z in rng.integers(0, 1 << vl, (8000, 3)).tolist():
        if not oracle.voxel_get(pool.words_ptr, cfg.node_levels, root, x, y, z):
            continue
        exp = cp.color_at(vl, x, y, z)
        if exp is None:
            continue
        got = oracle.color_fetch(cn, cl, cp.root, vl, 3, x, y, z)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (x, y, z)
        checked += 1
    assert checked > 300


def test_degenerate_edits_pointer_exact(oracle, ref):
    """Edge cases of the editor contract, restatement vs the reference's headers, pointer for pointer: single-voxel and
    zero-volume edits, edits entirely outside the world, whole-world fill and clear, digging in empty space, the same
    edit twice (idempotence), and a centre on the world's far corner."""
    cfg = B.default_config(level_count=6, top_level_count=9)
    res = 1 << cfg.voxel_level
    edits = [
        B.sphere((5, 5, 5), 9, dig=True),                       # dig in an empty world -> Null stays Null
        B.sphere((10, 11, 12), 0),                              # r2 = 0: exactly one voxel
        B.aabb((3, 3, 3), (3, 9, 9)),                           # empty box (lo == hi on one axis)
        B.aabb((res - 1, res - 1, res - 1), (res, res, res)),   # the last voxel of the world
        B.sphere((res - 1, res - 1, res - 1), 50),              # ball clipped by three faces
        B.sphere((10, 11, 12), 0),                              # again: nothing changes
        B.aabb((0, 0, 0), (res, res, res)),                     # whole world -> the filled root
        B.sphere((res // 2, res // 2, res // 2), 7 ** 2, dig=True),
        B.sphere((res // 2, res // 2, res // 2), (3 * res) ** 2, dig=True),  # everything away again -> Null
        B.aabb((1, 2, 3), (4, 5, 6)),
    ]
    op, rp = oracle.pool(cfg), ref.pool(cfg)
    ro = rr = B.NULL
    seen = []
    for e in edits:
        ro, rr = op.edit(ro, e), rp.edit(rr, e)
        assert ro == rr
        seen.append(ro)
    assert seen[0] == B.NULL and seen[2] == seen[1] and seen[5] == seen[4]
    assert seen[6] == op.filled_nodes()[0] == rp.filled_nodes()[0] and seen[8] == B.NULL
    assert np.array_equal(op.bucket_words_np(), rp.bucket_words_np())
    for off, cnt in op.used_ranges():
        assert np.array_equal(op.words_np(off, cnt), rp.words_np(off, cnt))


def test_bucket_overflow_fallback_pointer_exact(oracle, ref):
    """Collisions: with one 32-word bucket per level the pool overflows almost at once; a full bucket keeps the OLD
    node (NodePool.hpp:137-139,195) and both sides must agree on every pointer, every bucket cursor and every word —
    including the zeroed page tails a node would have straddled."""
    cfg = B.HdConfig()
    cfg.word_bits_per_page, cfg.page_bits_per_bucket, cfg.node_levels = 4, 1, 4
    for l in range(4):
        cfg.bucket_bits_each_level[l] = 1 if l < 3 else 2
    res = 1 << (cfg.node_levels + 1)
    rng = np.random.default_rng(23)
    op, rp = oracle.pool(cfg), ref.pool(cfg)
    ro = rr = B.NULL
    for _ in range(40):
        c = [int(v) for v in rng.integers(0, res, 3)]
        r = int(rng.integers(1, res // 2))
        e = B.sphere(c, r * r, dig=bool(rng.random() < 0.3)) if rng.random() < 0.7 else \
            B.aabb(c, [min(res, v + int(rng.integers(1, res // 2))) for v in c])
        ro, rr = op.edit(ro, e), rp.edit(rr, e)
        assert ro == rr
    assert op.stats()["overflow"] > 0                        # the case was actually exercised
    assert np.array_equal(op.bucket_words_np(), rp.bucket_words_np())
    for off, cnt in op.used_ranges():
        assert np.array_equal(op.words_np(off, cnt), rp.words_np(off, cnt))
