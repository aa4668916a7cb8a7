"""GPU parity of kernel family #2 (batched edit rebuild) against the CPU oracle: identical canonical DAG."""
import ctypes as C

import numpy as np
import pytest

from vkhashdag_b200 import abi

pytestmark = pytest.mark.gpu
NULL = abi.NULL


def device_canonical(oracle, dev, cfg, root):
    """Read the device pool back into a host mirror and canonicalise it with the oracle's walker."""
    mirror = oracle.pool(cfg)
    ranges, bw = dev.Download()
    for off, words in ranges.items():
        mirror.words_np(off, len(words))[:] = words
    mirror.bucket_words_np()[:] = bw
    c = oracle.canonical(mirror.words_ptr, cfg.node_levels, root)
    return c, mirror


def check_equal(oracle, hd, cfg, edits, batches=None, root0=NULL):
    """Apply `edits` sequentially on the oracle and in `batches` on the GPU; compare canonical DAGs."""
    opool = oracle.pool(cfg)
    oroot = opool.edit_batch(root0, edits)
    dev = hd.DAGNodePool(cfg)
    groot = root0
    batches = batches or [len(edits)]
    i = 0
    overflow = 0
    for b in batches:
        groot = dev.EditBatch(groot, edits[i:i + b])
        overflow += dev.last_stats["overflow_count"]
        i += b
    assert i == len(edits)
    assert overflow == 0
    exp = opool.canonical(oroot)
    got, mirror = device_canonical(oracle, dev, cfg, groot)
    assert (oroot == NULL) == (groot == NULL)
    if got["hash"] != exp["hash"]:  # keep the full evidence: this must never happen
        print("CANONICAL MISMATCH got", got, "exp", exp, "stats", dev.last_stats, "filled", dev.FilledNodes(), flush=True)
    assert got["hash"] == exp["hash"]
    assert got["by_ptr"] == got["by_content"] == exp["by_ptr"], "duplicate or missing nodes"
    assert got["voxels"] == exp["voxels"]
    assert got["per_level"] == exp["per_level"]
    return dev, groot, opool, oroot, mirror


def test_single_sphere_cfg1(oracle, hd):
    cfg = abi.default_config(level_count=10, top_level_count=9)
    dev, groot, opool, oroot, _ = check_equal(oracle, hd, cfg, [abi.sphere((512, 512, 512), 341 ** 2)])
    # no garbage: the GPU pool holds exactly the words the reference-order build holds (filled nodes included)
    assert dev.UsedWords() == int(opool.bucket_words_np().sum())
    # idempotence (test/test.cpp:161-172 intent): same edit again -> same root, nothing appended
    again = dev.EditBatch(groot, [abi.sphere((512, 512, 512), 341 ** 2)])
    assert again == groot and dev.last_stats["appended_nodes"] == 0
    dev.close()


def test_sequence_fill_dig_aabb(oracle, hd):
    cfg = abi.default_config(level_count=10, top_level_count=9)
    edits = [abi.sphere((512, 512, 512), 341 ** 2), abi.sphere((512, 512, 300), 150 ** 2, dig=True),
             abi.sphere((700, 600, 512), 200 ** 2), abi.aabb((100, 50, 100), (400, 90, 900)),
             abi.sphere((250, 70, 500), 60 ** 2, dig=True)]
    check_equal(oracle, hd, cfg, edits)[0].close()                      # one batch
    check_equal(oracle, hd, cfg, edits, batches=[1, 1, 1, 1, 1])[0].close()  # one call per edit
    check_equal(oracle, hd, cfg, edits, batches=[2, 3])[0].close()


def test_batch_order_matters(oracle, hd):
    """fill then dig != dig then fill: the batch must honour index order per voxel."""
    cfg = abi.default_config(level_count=8, top_level_count=9)
    a, b = abi.sphere((128, 128, 128), 60 ** 2), abi.sphere((150, 128, 128), 50 ** 2, dig=True)
    d1 = check_equal(oracle, hd, cfg, [a, b])
    d2 = check_equal(oracle, hd, cfg, [b, a])
    assert d1[3] != NULL and d2[3] != NULL
    assert d1[2].canonical(d1[3])["hash"] != d2[2].canonical(d2[3])["hash"]
    d1[0].close(), d2[0].close()


def test_random_sphere_batch(oracle, hd):
    cfg = abi.default_config(level_count=9, top_level_count=9)
    edits = abi.random_spheres(200, cfg.voxel_level, seed=99, rmin=3, rmax=40, y_lo=100, y_hi=400)
    check_equal(oracle, hd, cfg, edits)[0].close()


def test_random_sphere_batch_general_path(oracle, hd):
    """More than 1024 editors: the host-driven level-synchronous path with the fused last-level + leaf kernel
    (k_down_leaf), batch dedup and the bucket-grouped find-or-insert."""
    cfg = abi.default_config(level_count=9, top_level_count=9)
    edits = abi.random_spheres(1100, cfg.voxel_level, seed=7, rmin=3, rmax=24, y_lo=60, y_hi=450)
    dev = check_equal(oracle, hd, cfg, edits)[0]
    assert dev.last_stats["path"] == "general" and dev.last_stats["visited_leaves"] > 100000
    dev.close()
    # the same list as two batches on top of each other (the second one edits an existing scene)
    check_equal(oracle, hd, cfg, edits, batches=[1050, 50])[0].close()


def test_terrain_and_spheres(oracle, hd):
    cfg = abi.default_config(level_count=9, top_level_count=9)
    vl = cfg.voxel_level
    edits = [abi.terrain(vl)] + abi.random_spheres(64, vl, seed=1234, rmin=4, rmax=32)
    dev, groot, opool, oroot, mirror = check_equal(oracle, hd, cfg, edits)
    # spot-check voxels against the height function
    t = abi.terrain(vl)
    rng = np.random.default_rng(3)
    dev2, g2, op2, or2, m2 = check_equal(oracle, hd, cfg, [t])
    for _ in range(200):
        x, z = (int(v) for v in rng.integers(0, 1 << vl, 2))
        h = oracle.terrain_height(t, x, z)
        for y in (h - 1, h):
            if 0 <= y < (1 << vl):
                assert oracle.voxel_get(m2.words_ptr, cfg.node_levels, g2, x, y, z) == (y < h)
    assert oracle.canonical(m2.words_ptr, cfg.node_levels, g2)["voxels"] == oracle.in_range_voxels(t, vl)
    dev.close(), dev2.close()


def test_whole_world_fill_and_clear(oracle, hd):
    cfg = abi.default_config(level_count=6, top_level_count=9)
    res = 1 << cfg.voxel_level
    dev, groot, *_ = check_equal(oracle, hd, cfg, [abi.aabb((0, 0, 0), (res, res, res))])
    assert groot == dev.FilledNodes()[0]
    dev.close()
    # dig everything away again -> Null root
    big = abi.sphere((res // 2,) * 3, (2 * res) ** 2, dig=True)
    dev, groot, *_ = check_equal(oracle, hd, cfg, [abi.aabb((0, 0, 0), (res, res, res)), big])
    assert groot == NULL
    dev.close()
    # edit that touches nothing
    dev, groot, *_ = check_equal(oracle, hd, cfg, [abi.sphere((5, 5, 5), 9, dig=True)])
    assert groot == NULL
    dev.close()


def test_subset_and_superset_edits(oracle, hd):
    """test/test.cpp:161-199 intent: subset edit keeps the root, superset edit changes it."""
    cfg = abi.default_config(level_count=5, top_level_count=9)
    dev = hd.DAGNodePool(cfg)
    r1 = dev.Edit(NULL, hd.AABBEditor((0, 0, 0), (4, 4, 4)))
    assert r1 != NULL and dev.Edit(r1, hd.AABBEditor((0, 0, 0), (4, 4, 4))) == r1
    r3 = dev.Edit(r1, hd.AABBEditor((1, 1, 1), (5, 5, 5)))
    assert r3 != r1
    assert dev.Edit(r3, hd.AABBEditor((1, 2, 3), (3, 5, 5))) == r3
    c, mirror = device_canonical(oracle, dev, cfg, r3)
    assert oracle.voxel_get(mirror.words_ptr, cfg.node_levels, r3, 4, 3, 3)
    assert oracle.voxel_get(mirror.words_ptr, cfg.node_levels, r1, 3, 3, 3)
    assert not oracle.voxel_get(mirror.words_ptr, cfg.node_levels, r1, 4, 3, 3)
    dev.close()


def test_upsert_dedup_known_answers(oracle, hd):
    """test/test.cpp:118-147 intent."""
    cfg = abi.default_config(level_count=5, top_level_count=9)
    dev = hd.DAGNodePool(cfg)
    shift = cfg.word_bits_per_page + cfg.page_bits_per_bucket
    n0 = [0b11, 0x2300, 0x4500]
    p = dev.Upsert(0, [n0, n0, n0], 3)
    assert p[0] != NULL and p[0] == p[1] == p[2]
    assert dev.Upsert(0, [n0], 3)[0] == p[0]
    bw = dev.ReadBucketWords()
    assert bw[p[0] >> shift] == 3
    n1 = [0b110, 0x2300, 0x4400]
    p3 = dev.Upsert(0, [n1], 3)[0]
    p4 = dev.Upsert(1, [n1], 3)[0]
    assert p3 != p[0] and p4 != p3
    n2 = [0xFF, 0x2300, 0x4400, 0x5500, 0x6600, 0x7700, 0x8800, 0x9900, 0x100]
    p5 = dev.Upsert(2, [n2], 9)[0]
    assert dev.ReadBucketWords()[p5 >> shift] == 9
    p6 = dev.Upsert(3, [[0x23, 0x55]], 2)[0]
    assert dev.ReadBucketWords()[p6 >> shift] == 2
    # bucket placement equals the reference hash (Hasher.hpp:21-49)
    opool = oracle.pool(cfg)
    for lvl, node in ((0, n0), (0, n1), (1, n1), (2, n2), (3, [0x23, 0x55])):
        assert (opool.upsert(lvl, node) >> shift) == (dev.Upsert(lvl, [node], len(node))[0] >> shift)
    dev.close()


@pytest.mark.parametrize("top_bits", [4, 0])  # 0: tiny pool -> the sequential bucket walk is used
def test_bucket_full_and_no_page_straddle(oracle, hd, top_bits):
    """test/test.cpp:148-160 intent, with a one-bucket level instead of a zero hasher."""
    cfg = abi.HdConfig()
    cfg.word_bits_per_page, cfg.page_bits_per_bucket, cfg.node_levels = 4, 1, 3
    cfg.bucket_bits_each_level[0], cfg.bucket_bits_each_level[1], cfg.bucket_bits_each_level[2] = top_bits, 0, 4
    dev = hd.DAGNodePool(cfg)
    wpp, pages = 16, 2
    nodes = [[0b11, 0x1000 + i, 0x4500] for i in range(40)]
    ptrs = dev.Upsert(1, nodes, 3)
    ok = ptrs[ptrs != NULL]
    assert len(ok) == (wpp // 3) * pages
    assert all((int(p) % wpp) % 3 == 0 for p in ok)
    # again: the stored ones are found, the others still do not fit
    again = dev.Upsert(1, nodes, 3)
    assert np.array_equal(again, ptrs)
    dev.close()


def test_brush_sequence_low_latency_path(oracle, hd):
    """The interactive loop (src/main.cpp:214-238): one brush editor per call, fill/dig alternating, on a terrain.
    Every call is served by the low-latency path (one cooperative kernel — or, with HD_EDIT_FAST=2, one CUDA graph — and
    device-resident queue counts) and the DAG after the whole sequence is canonically equal to sequential CPU edits."""
    cfg = abi.default_config(level_count=9, top_level_count=9)
    vl = cfg.voxel_level
    res = 1 << vl
    rng = np.random.default_rng(17)
    edits = [abi.terrain(vl)]
    for i in range(24):
        c = tuple(int(v) for v in rng.integers(res // 4, 3 * res // 4, 3))
        edits.append(abi.sphere(c, int(rng.integers(3, 40)) ** 2, dig=bool(i & 1)))
    edits.append(abi.aabb((10, 10, 10), (60, 30, 90)))
    opool = oracle.pool(cfg)
    oroot = opool.edit_batch(NULL, edits)
    dev = hd.DAGNodePool(cfg)
    groot = dev.EditBatch(NULL, edits[:1])
    assert dev.last_stats["path"] == "general"          # terrain fills always take the general path
    for e in edits[1:]:
        groot = dev.EditBatch(groot, [e])
        assert dev.last_stats["path"] in ("fused", "graph") and dev.last_stats["overflow_count"] == 0
    exp = opool.canonical(oroot)
    got, _ = device_canonical(oracle, dev, cfg, groot)
    assert got["hash"] == exp["hash"] and got["by_ptr"] == got["by_content"] == exp["by_ptr"]
    assert got["voxels"] == exp["voxels"] and got["per_level"] == exp["per_level"]
    # a small multi-editor batch (<= 32) is one graph launch too, and order inside it is honoured
    g2 = dev.EditBatch(groot, edits[1:9])
    assert dev.last_stats["path"] in ("fused", "graph")
    o2 = opool.edit_batch(oroot, edits[1:9])
    assert device_canonical(oracle, dev, cfg, g2)[0]["hash"] == opool.canonical(o2)["hash"]
    dev.close()


_OVERFLOW_SCRIPT = r"""
import sys
sys.path.insert(0, {root!r})
import vkhashdag_b200 as v
from oracle import bindings
from vkhashdag_b200 import abi
O = bindings.Oracle()
cfg = abi.default_config(level_count=9, top_level_count=9)
edits = [abi.sphere((256, 256, 256), 150 ** 2), abi.sphere((256, 256, 300), 3 ** 2, dig=True)]
opool = O.pool(cfg)
oroot = opool.edit_batch(abi.NULL, edits)
dev = v.DAGNodePool(cfg)
r = dev.EditBatch(abi.NULL, edits[:1])
assert dev.last_stats["path"] == "general", dev.last_stats     # work queues of 64 items cannot hold this sphere
used = dev.UsedWords()
r = dev.EditBatch(r, edits[1:])
assert dev.last_stats["path"] in ("fused", "graph"), dev.last_stats       # the small brush fits
mirror = O.pool(cfg)
ranges, bw = dev.Download()
for off, words in ranges.items():
    mirror.words_np(off, len(words))[:] = words
got, exp = O.canonical(mirror.words_ptr, cfg.node_levels, r), opool.canonical(oroot)
assert got == exp, (got, exp)
# the aborted low-latency attempt wrote nothing: the pool holds exactly what the reference-order build holds
assert dev.UsedWords() == int(opool.bucket_words_np().sum()), (dev.UsedWords(), int(opool.bucket_words_np().sum()))
print("OK")
"""


def test_low_latency_queue_overflow_falls_back(oracle, hd):
    """With tiny fixed work queues (HD_EDIT_FAST_CAP=64) a big sphere overflows them: the call must be redone by the
    general path with nothing written in between, and give the same canonical DAG."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HD_EDIT_FAST_CAP="64")
    out = subprocess.run([sys.executable, "-c", _OVERFLOW_SCRIPT.format(root=root)], env=env, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def test_degenerate_edits(oracle, hd):
    """Edge cases of the editor contract on the GPU paths (one call per edit -> low-latency path; one batch; a batch padded
    past 32 editors -> long lists on the low-latency path; and one padded past 1024 editors -> general path): single-voxel and zero-volume edits, edits clipped by or outside the world,
    whole-world fill and clear, digging in empty space, repeated edits."""
    cfg = abi.default_config(level_count=6, top_level_count=9)
    res = 1 << cfg.voxel_level
    edits = [
        abi.sphere((5, 5, 5), 9, dig=True), abi.sphere((10, 11, 12), 0), abi.aabb((3, 3, 3), (3, 9, 9)),
        abi.aabb((res - 1, res - 1, res - 1), (res, res, res)), abi.sphere((res - 1, res - 1, res - 1), 50),
        abi.sphere((10, 11, 12), 0), abi.aabb((1, 2, 3), (40, 50, 60)),
        abi.sphere((res // 2, res // 2, res // 2), 7 ** 2, dig=True), abi.aabb((1, 2, 3), (4, 5, 6)),
    ]
    check_equal(oracle, hd, cfg, edits, batches=[1] * len(edits))[0].close()
    check_equal(oracle, hd, cfg, edits)[0].close()
    padded = edits + [abi.sphere((20, 20, 20), 4, dig=True)] * 30      # 39 editors: lists longer than 32 on the one-launch path
    dev = check_equal(oracle, hd, cfg, padded)[0]
    assert dev.last_stats["path"] in ("fused", "graph")
    dev.close()
    padded = edits + [abi.sphere((20, 20, 20), 4, dig=True)] * 1030    # past the one-launch path's 1024 editors: general path
    dev = check_equal(oracle, hd, cfg, padded)[0]
    assert dev.last_stats["path"] == "general"
    dev.close()
    # whole world, then everything away again: filled root, then Null, on the low-latency path
    dev = hd.DAGNodePool(cfg)
    r = dev.EditBatch(NULL, [abi.aabb((0, 0, 0), (res, res, res))])
    assert r == dev.FilledNodes()[0]
    r = dev.EditBatch(r, [abi.sphere((res // 2,) * 3, (3 * res) ** 2, dig=True)])
    assert r == NULL
    assert dev.EditBatch(NULL, []) == NULL                                # empty batch: root unchanged
    dev.close()


def test_mid_size_batches_one_launch(oracle, hd):
    """Batches of 33 .. 1000 editors stay on the one-launch path: their lists are longer than 32 entries in the levels
    next to the root (a warp per (item, child) pair filters them, phase_down_long), shorter below.  Mixed fill / dig / AABB
    editors, overlapping on purpose so that the order inside a list matters, applied on top of an existing scene."""
    cfg = abi.default_config(level_count=9, top_level_count=9)
    vl = cfg.voxel_level
    base = [abi.terrain(vl)]
    for n, seed in ((33, 1), (100, 2), (400, 3), (1000, 4)):
        rng = np.random.default_rng(seed)
        edits = []
        for i in range(n):
            c = tuple(int(v) for v in rng.integers(120, 390, 3))
            if i % 7 == 3:
                edits.append(abi.aabb(c, tuple(v + int(rng.integers(4, 50)) for v in c)))
            else:
                edits.append(abi.sphere(c, int(rng.integers(4, 60)) ** 2, dig=bool(rng.integers(0, 2))))
        dev = check_equal(oracle, hd, cfg, base + edits, batches=[1, n])[0]
        assert dev.last_stats["path"] in ("fused", "graph"), (n, dev.last_stats)
        dev.close()


def test_edit_node8_equals_eight_edit_node_calls(hd):
    """The shared-plane classification of a node's eight children (edit_node8) returns what EditNode returns for each child,
    on 64 M pseudo-random editor / node pairs placed around the decision boundaries."""
    assert hd.api.selftest_edit_node8(1 << 26, 0) == 0
