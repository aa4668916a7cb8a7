#!/usr/bin/env python
"""tools/make_golden.py — regenerate tests/golden/* from the COMPILED REFERENCE (oracle/_ref).

Run in the build container (where /root/reference is mounted):  python tools/make_golden.py
Everything written here is produced by the reference's own headers (include/hashdag/*.hpp, compiled in place by
oracle/Makefile) — never by the oracle restatement or the CUDA product, which are what the vectors check.
The only non-reference ingredient is the CANONICAL hash (oracle.canonical): a pure read-only walk of a word array,
applied to the reference's pool memory.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cfg_dict(cfg):
    return {"word_bits_per_page": cfg.word_bits_per_page, "page_bits_per_bucket": cfg.page_bits_per_bucket,
            "bucket_bits": cfg.bucket_bits()}


def desc_dict(d):
    return {"kind": d.kind, "p0": list(d.p0), "p1": list(d.p1), "aux": d.aux, "r2": int(d.r2)}


def main(out_dir=None):
    global OUT
    if out_dir:
        OUT = out_dir
    B.build(ref=True)
    R, O = B.Ref(), B.Oracle()
    os.makedirs(OUT, exist_ok=True)
    g = {"generator": "tools/make_golden.py over oracle/_ref (reference headers compiled in place)"}

    # ---- Hasher.hpp:21-49 ----
    rng = np.random.default_rng(20240117)
    inner = [[0x3, 0x23, 0x45], [0xFF, 1, 2, 3, 4, 5, 6, 7, 8], [0x80, 0xFFFFFFFE]]
    for _ in range(40):
        mask = int(rng.integers(1, 256))
        inner.append([mask] + [int(v) for v in rng.integers(256, 2 ** 32 - 1, bin(mask).count("1"))])
    leaves = [[0x23, 0x55], [0xFFFFFFFF, 0xFFFFFFFF], [1, 0]] + [[int(v) for v in rng.integers(0, 2 ** 32, 2)] for _ in range(40)]
    g["hash_inner"] = [{"words": w, "hash": R.hash_inner(w)} for w in inner]
    g["hash_leaf"] = [{"words": w, "hash": R.hash_leaf(w)} for w in leaves]
    g["hash_leaf_words_through_inner_overload"] = {"words": [0x23, 0x55], "hash": R.hash_inner([0x23, 0x55])}

    # ---- Config.hpp:33-75 ----
    g["configs"] = []
    for kw in ({"level_count": 10, "top_level_count": 9}, {"level_count": 15, "top_level_count": 9},
               {"level_count": 17, "top_level_count": 9},
               {"level_count": 17, "top_level_count": 9, "word_bits_per_page": 14, "page_bits_per_bucket": 2,
                "bucket_bits_per_top_level": 7, "bucket_bits_per_bottom_level": 11},
               {"level_count": 4, "top_level_count": 2, "word_bits_per_page": 4, "page_bits_per_bucket": 0,
                "bucket_bits_per_top_level": 1, "bucket_bits_per_bottom_level": 3}):
        full = dict(level_count=17, top_level_count=9, word_bits_per_page=9, page_bits_per_bucket=2,
                    bucket_bits_per_top_level=10, bucket_bits_per_bottom_level=16)
        full.update(kw)
        cfg, ok = R.config_from_default(**full)
        g["configs"].append({"default": full, "valid": ok, "config": cfg_dict(cfg), "geometry": R.geometry(cfg)})
    # one config that fails Validate (> 2^32-2 words)
    cfg, ok = R.config_from_default(level_count=17, top_level_count=9, word_bits_per_page=14, page_bits_per_bucket=2,
                                    bucket_bits_per_top_level=10, bucket_bits_per_bottom_level=16)
    g["configs"].append({"default": None, "valid": ok, "config": cfg_dict(cfg), "geometry": None})

    # ---- NodePool.hpp: filled pointers, upsert, Edit (cfg1 = BASELINE config 1) ----
    cfg1 = B.default_config(level_count=10, top_level_count=9)
    rp = R.pool(cfg1)
    g["cfg1"] = {"config": cfg_dict(cfg1), "filled": rp.filled_nodes()}
    s1 = B.sphere((512, 512, 512), 341 ** 2)
    r1 = rp.edit(B.NULL, s1)
    root_words = rp.words_np(r1, 9).tolist()
    g["cfg1"]["sphere"] = {"edit": desc_dict(s1), "root": r1, "sum_bucket_words": int(rp.bucket_words_np().sum()),
                           "root_words": root_words, "same_edit_again_root": rp.edit(r1, s1),
                           "bucket_words_sha256": sha(rp.bucket_words_np()),
                           "canonical": O.canonical(rp.words_ptr, cfg1.node_levels, r1)}
    rays = [((0.5, 0.5, 1.5), (0, 0, -1)), ((0.5, 0.5, 1.5), (0.1, -0.2, -1)), ((-0.2, 0.3, 0.4), (1, 0.3, 0.2)),
            ((0.5, 0.5, 1.5), (0.6, 0, -1))]
    for _ in range(200):
        rays.append((tuple(float(v) for v in rng.uniform(-0.5, 1.5, 3)), tuple(float(v) for v in rng.normal(size=3))))
    trav = []
    for o, d in rays:
        dn = np.array(d, np.float32)
        dn = dn / np.float32(np.sqrt(np.float32(np.dot(dn, dn))))
        o32 = np.array(o, np.float32)
        h = rp.traverse(r1, o32, dn)
        trav.append({"o_bits": o32.view(np.uint32).tolist(), "d_bits": dn.view(np.uint32).tolist(),
                     "hit": h is not None, "pos_bits": None if h is None else h.view(np.uint32).tolist()})
    g["cfg1"]["traversal"] = trav
    # host-tracer frame (Traversal<float> per pixel, rays as trace.frag generates them)
    P = B.camera_params(cfg1, r1, (0.5, 0.5, 1.5), np.pi, 0.0, 160, 90)
    fr = rp.trace_frame_host(P, threads=4)
    g["cfg1"]["host_frame"] = {"width": 160, "height": 90, "pos": [0.5, 0.5, 1.5], "yaw": float(np.pi), "pitch": 0.0,
                               "n_hits": int(fr["n_hits"]), "hit_sha256": sha(fr["hit"]),
                               "pos_sha256": sha(fr["pos"].view(np.uint32) * fr["hit"][..., None])}

    # edit sequence: serial vs threaded must agree canonically
    seq = [B.sphere((512, 512, 300), 150 ** 2, dig=True), B.sphere((700, 600, 512), 200 ** 2),
           B.aabb((100, 50, 100), (400, 90, 900)), B.sphere((250, 70, 500), 60 ** 2, dig=True)]
    r2 = rp.edit_batch(r1, seq)
    can_serial = O.canonical(rp.words_ptr, cfg1.node_levels, r2)
    rp_t = R.pool(cfg1)
    rt = rp_t.edit_batch(B.NULL, [s1] + seq, threads=8, max_task_level=6)
    can_thr = O.canonical(rp_t.words_ptr, cfg1.node_levels, rt)
    assert can_serial == can_thr, "reference serial vs threaded canonical mismatch"
    g["cfg1"]["sequence"] = {"edits": [desc_dict(s1)] + [desc_dict(e) for e in seq], "serial_root": r2,
                             "canonical": can_serial}

    # ---- terrain + random spheres at 2^9 (small cfg3 analogue) ----
    cfg9 = B.default_config(level_count=9, top_level_count=9)
    rp9 = R.pool(cfg9)
    t9 = B.terrain(cfg9.voxel_level)
    rt9 = rp9.edit(B.NULL, t9)
    g["terrain9"] = {"config": cfg_dict(cfg9), "edit": desc_dict(t9), "root": rt9,
                     "canonical": O.canonical(rp9.words_ptr, cfg9.node_levels, rt9),
                     "heights": [[x, z, O.terrain_height(t9, x, z)] for x, z in rng.integers(0, 512, (24, 2)).tolist()]}
    sp = B.random_spheres(64, cfg9.voxel_level, seed=1234, rmin=4, rmax=32)
    rs9 = rp9.edit_batch(rt9, sp)
    g["terrain9"]["spheres"] = {"edits": [desc_dict(e) for e in sp], "root": rs9,
                                "canonical": O.canonical(rp9.words_ptr, cfg9.node_levels, rs9)}

    # ---- stale-test intents resurrected (test/test.cpp:118-199) on the current headers ----
    cfg5 = B.default_config(level_count=5, top_level_count=9)
    p5 = R.pool(cfg5)
    shift = cfg5.word_bits_per_page + cfg5.page_bits_per_bucket
    n0, n1 = [0b11, 0x2300, 0x4500], [0b110, 0x2300, 0x4400]
    a = p5.upsert(0, n0)
    b = p5.upsert(0, n0)
    c = p5.upsert(0, n1)
    d = p5.upsert(1, n1)
    n2 = [0xFF, 0x2300, 0x4400, 0x5500, 0x6600, 0x7700, 0x8800, 0x9900, 0x100]
    e = p5.upsert(2, n2)
    f = p5.upsert(3, [0x23, 0x55])
    bw = p5.bucket_words_np()
    g["upsert"] = {"config": cfg_dict(cfg5), "ptrs": [a, b, c, d, e, f],
                   "bucket_words": [int(bw[a >> shift]), int(bw[e >> shift]), int(bw[f >> shift])]}
    p5 = R.pool(cfg5)
    ra = p5.edit(B.NULL, B.aabb((0, 0, 0), (4, 4, 4)))
    rb = p5.edit(ra, B.aabb((0, 0, 0), (4, 4, 4)))
    rc = p5.edit(rb, B.aabb((1, 1, 1), (5, 5, 5)))
    rd = p5.edit(rc, B.aabb((1, 2, 3), (3, 5, 5)))
    g["aabb_edits"] = {"roots": [ra, rb, rc, rd],
                       "canonical_first": O.canonical(p5.words_ptr, cfg5.node_levels, ra),
                       "canonical_last": O.canonical(p5.words_ptr, cfg5.node_levels, rd)}

    with open(os.path.join(OUT, "reference_vectors.json"), "w") as fjson:
        json.dump(g, fjson, indent=1)

    # ---- colour: reference-built colour pool (VBREditorWrapper + VBRChunkWriter) on a 2^8 scene ----
    cfg8 = B.default_config(level_count=8, top_level_count=9)
    vl = cfg8.voxel_level
    pool = R.pool(cfg8)
    cp = R.color_pool(leaf_level=4)
    root = B.NULL
    root = pool.edit_color(cp, root, B.aabb((20, 10, 20), (200, 60, 220)), 0xFFFFFF)
    root = pool.edit_color(cp, root, B.aabb((0, 0, 0), (90, 90, 90)), 0x00FFFF)
    root = pool.edit_color(cp, root, B.sphere((128, 100, 128), 50 ** 2), 0x3060C0)
    root = pool.edit_color(cp, root, B.sphere((100, 80, 100), 40 ** 2), 0x007FFF, paint=True)
    root = pool.edit(root, B.sphere((150, 100, 150), 35 ** 2, dig=True))
    root = pool.edit_color(cp, root, B.sphere((60, 60, 160), 30 ** 2), 0x20C040)
    cn, cl = cp.arrays()
    samples = []
    pts = rng.integers(0, 1 << vl, (6000, 3)).tolist()
    for x, y, z in pts:
        if not O.voxel_get(pool.words_ptr, cfg8.node_levels, root, x, y, z):
            continue
        col = cp.color_at(vl, x, y, z)
        if col is not None:
            samples.append([x, y, z] + col.view(np.uint32).tolist())
    used = {}
    for off, cnt in pool.used_ranges():
        used[off] = pool.words_np(off, cnt).copy()
    offs = np.array(sorted(used), np.uint32)
    lens = np.array([len(used[o]) for o in offs], np.uint32)
    np.savez_compressed(os.path.join(OUT, "color_scene.npz"), color_nodes=cn, color_leaves=cl,
                        color_root=np.uint32(cp.root), leaf_level=np.uint32(4), node_root=np.uint32(root),
                        range_offsets=offs, range_lengths=lens, words=np.concatenate([used[o] for o in offs]),
                        bucket_words=pool.bucket_words_np().copy(), samples=np.array(samples, np.uint32),
                        bucket_bits=np.array(cfg8.bucket_bits(), np.uint32))
    print("colour scene: nodes", cn.size, "leaf words", cl.size, "samples", len(samples),
          "npz bytes", os.path.getsize(os.path.join(OUT, "color_scene.npz")))
    print("wrote", OUT)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
