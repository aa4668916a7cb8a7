"""Row N2: colour-aware edits on the GPU against the REFERENCE itself (oracle/_ref: VBREditorWrapper + VBRChunkWriter +
the DAGColorPool layout).  The two colour pools must be structurally identical: same tags and colours at every octree
position, and every leaf chunk word-for-word equal (macro blocks, block headers, weight bits) — only node ids and chunk
placement may differ."""
import numpy as np
import pytest

from vkhashdag_b200 import abi

pytestmark = pytest.mark.gpu
NULL = abi.NULL


def chunk_payload(leaves, idx):
    macro, blocks, ww = (int(v) for v in leaves[idx + 1:idx + 4])
    n = 3 + 2 * macro + 2 * blocks + ww
    return leaves[idx + 1:idx + 1 + n]


def compare_color_trees(a, b, leaf_level):
    """a, b = (nodes, leaves, root).  Returns the number of (octree nodes, leaf chunks) compared."""
    (na, la, ra), (nb, lb, rb) = a, b
    stack, n_nodes, n_leaves = [(ra, rb, 0)], 0, 0
    while stack:
        pa, pb, lvl = stack.pop()
        ta, tb = pa >> 30, pb >> 30
        assert ta == tb, f"tag mismatch at level {lvl}: {ta} vs {tb}"
        if ta == 1:
            assert (pa & 0xFFFFFF) == (pb & 0xFFFFFF), f"colour mismatch at level {lvl}"
        elif ta == 2:
            assert lvl == leaf_level
            ca, cb = chunk_payload(la, pa & 0x3FFFFFFF), chunk_payload(lb, pb & 0x3FFFFFFF)
            assert len(ca) == len(cb) and np.array_equal(ca, cb), f"leaf chunk differs ({len(ca)} vs {len(cb)} words)"
            n_leaves += 1
        elif ta == 0:
            assert lvl < leaf_level
            n_nodes += 1
            ia, ib = (pa & 0x3FFFFFFF) << 3, (pb & 0x3FFFFFFF) << 3
            for c in range(8):
                stack.append((int(na[ia + c]), int(nb[ib + c]), lvl + 1))
    return n_nodes, n_leaves


def mirror_of(oracle, dev, cfg):
    m = oracle.pool(cfg)
    ranges, bw = dev.Download()
    for off, words in ranges.items():
        m.words_np(off, len(words))[:] = words
    m.bucket_words_np()[:] = bw
    return m


def run_sequence(oracle, ref, hd, cfg, leaf_level, steps):
    """steps: list of (desc, rgb8 | None, paint).  rgb8 None = stateless edit (e.g. dig).  Checks after EVERY step."""
    rp, cp = ref.pool(cfg), ref.color_pool(leaf_level=leaf_level)
    dev = hd.DAGNodePool(cfg)
    dev.ColorConfig(leaf_level)
    rr = gr = NULL
    stats = []
    for i, (desc, rgb8, paint) in enumerate(steps):
        if rgb8 is None:
            rr, gr = rp.edit(rr, desc), dev.Edit(gr, desc)
        else:
            rr = rp.edit_color(cp, rr, desc, rgb8, paint)
            gr, gcol = dev.EditColor(gr, desc, rgb8, paint)
            assert gcol == dev.ColorRoot()
        m = mirror_of(oracle, dev, cfg)
        assert oracle.canonical(m.words_ptr, cfg.node_levels, gr) == oracle.canonical(rp.words_ptr, cfg.node_levels, rr), i
        rn, rl = cp.arrays()
        gn, gl = dev.ReadColor()
        stats.append(compare_color_trees((gn, gl, dev.ColorRoot()), (rn, rl, cp.root), leaf_level))
    return dev, gr, rp, cp, rr, stats


def test_color_edits_match_reference_main_scene(oracle, ref, hd):
    """The shape of the reference's initial scene (src/main.cpp:252-279) at 2^8: two coloured AABBs, a paint sphere,
    a dig, then brush fills and paints."""
    cfg = abi.default_config(level_count=8, top_level_count=9)
    steps = [(abi.aabb((20, 10, 20), (200, 60, 220)), 0xFFFFFF, False),
             (abi.aabb((0, 0, 0), (90, 90, 90)), 0x00FFFF, False),
             (abi.sphere((128, 100, 128), 50 ** 2), 0x3060C0, False),
             (abi.sphere((100, 80, 100), 40 ** 2), 0x007FFF, True),
             (abi.sphere((150, 100, 150), 35 ** 2, dig=True), None, False),
             (abi.sphere((60, 60, 160), 30 ** 2), 0x20C040, False),
             (abi.sphere((70, 65, 150), 25 ** 2), 0x20C040, False),        # same colour again: merges runs
             (abi.sphere((150, 100, 150), 20 ** 2), 0xFF0000, True),       # paint into the dug hole: mostly empty space
             (abi.aabb((0, 0, 0), (256, 256, 256)), 0x7F7F7F, False)]      # whole world: collapses to one colour
    dev, gr, rp, cp, rr, stats = run_sequence(oracle, ref, hd, cfg, 4, steps)
    assert max(s[1] for s in stats) > 20 and max(s[0] for s in stats) > 10     # real leaf chunks and octree nodes
    assert stats[-1] == (0, 0) and dev.ColorRoot() >> 30 == 1                   # solid colour at the root
    dev.close()


def test_color_frames_match_reference_pool(oracle, ref, hd):
    cfg = abi.default_config(level_count=8, top_level_count=9)
    steps = [(abi.terrain(cfg.voxel_level), None, False),
             (abi.aabb((0, 0, 0), (256, 70, 256)), 0x40A040, False),
             (abi.sphere((128, 90, 128), 40 ** 2), 0xC08040, False),
             (abi.sphere((100, 80, 120), 30 ** 2), 0x2040F0, True),
             (abi.sphere((140, 95, 140), 18 ** 2, dig=True), None, False),
             (abi.sphere((60, 75, 60), 22 ** 2), 0xF0F020, False)]
    dev, gr, rp, cp, rr, _ = run_sequence(oracle, ref, hd, cfg, 4, steps)
    rn, rl = cp.arrays()
    for cam in (((0.5, 0.7, 1.3), np.pi, -0.35), ((0.9, 0.8, 0.2), -0.9, -0.5)):
        Pg = abi.camera_params(cfg, gr, *cam, 320, 180, color_root=dev.ColorRoot(), color_leaf_level=4)
        Pr = abi.camera_params(cfg, rr, *cam, 320, 180, color_root=cp.root, color_leaf_level=4)
        got = dev.Trace(Pg)
        exp = oracle.trace_frame(rp.words_ptr, Pr, rn, rl)
        assert np.array_equal(got["hits"], exp["hits"]) and np.array_equal(got["rgba8"], exp["rgba8"])
        assert len(np.unique(got["hits"]["packed"] & 0xFFFFFF)) >= 4
    dev.close()


def test_color_random_brushes(oracle, ref, hd):
    cfg = abi.default_config(level_count=7, top_level_count=9)
    rng = np.random.default_rng(21)
    res = 1 << cfg.voxel_level
    steps = [(abi.aabb((0, 0, 0), (res, res // 3, res)), 0x808080, False)]
    palette = [0xFF0000, 0x00FF00, 0x0000FF, 0xFFFF00, 0x123456]
    for i in range(24):
        c = [int(v) for v in rng.integers(10, res - 10, 3)]
        c[1] = int(rng.integers(res // 4, res // 2))
        r = int(rng.integers(4, 26))
        kind = rng.integers(0, 4)
        if kind == 0:
            steps.append((abi.sphere(c, r * r, dig=True), None, False))
        elif kind == 1:
            steps.append((abi.sphere(c, r * r), int(rng.choice(palette)), True))
        elif kind == 2:
            lo = [max(0, v - r) for v in c]
            steps.append((abi.aabb(lo, [v + r for v in c]), int(rng.choice(palette)), False))
        else:
            steps.append((abi.sphere(c, r * r), int(rng.choice(palette)), False))
    run_sequence(oracle, ref, hd, cfg, 3, steps)[0].close()


def test_color_large_leaves_many_macro_blocks(oracle, ref, hd):
    """64^3-voxel colour leaves (16 macro blocks, 512 8^3 blocks each): the block path of color.cu — one-colour blocks from
    the editor, from empty space and from single old runs next to per-voxel blocks, repeated rewrites in place and appends."""
    cfg = abi.default_config(level_count=9, top_level_count=9)
    steps = [(abi.terrain(cfg.voxel_level), None, False),
             (abi.aabb((0, 0, 0), (512, 150, 512)), 0x40A040, False),
             (abi.sphere((256, 160, 256), 90 ** 2), 0xC08040, False),
             (abi.sphere((200, 150, 240), 70 ** 2), 0x2040F0, True),
             (abi.sphere((200, 150, 240), 70 ** 2), 0x2040F0, True),       # same paint again: nothing changes
             (abi.sphere((230, 170, 250), 33 ** 2), 0xF02020, True),
             (abi.sphere((280, 190, 280), 40 ** 2, dig=True), None, False),
             (abi.sphere((120, 140, 120), 50 ** 2), 0xF0F020, False),
             (abi.aabb((100, 100, 100), (140, 190, 130)), 0x10D0D0, False),
             (abi.sphere((128, 150, 128), 9 ** 2), 0x123456, True)]
    dev, gr, rp, cp, rr, stats = run_sequence(oracle, ref, hd, cfg, 3, steps)
    assert max(s[1] for s in stats) >= 8
    rn, rl = cp.arrays()
    Pg = abi.camera_params(cfg, gr, (0.5, 0.75, 1.2), np.pi, -0.45, 320, 180, color_root=dev.ColorRoot(), color_leaf_level=3)
    Pr = abi.camera_params(cfg, rr, (0.5, 0.75, 1.2), np.pi, -0.45, 320, 180, color_root=cp.root, color_leaf_level=3)
    got, exp = dev.Trace(Pg), oracle.trace_frame(rp.words_ptr, Pr, rn, rl)
    assert np.array_equal(got["hits"], exp["hits"]) and np.array_equal(got["rgba8"], exp["rgba8"])
    dev.close()
