"""oracle/bindings.py — TEST INFRASTRUCTURE: ctypes access to the CPU oracle and the compiled reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
The product (vkhashdag_b200) never does.

  Oracle  — oracle/liboracle.so, the CPU restatement (hashdag_oracle.cpp).
  Ref     — oracle/_ref/libhashdag_ref.so, the reference's own headers compiled in place (ref_harness.cpp);
            present only where it was built (this container) or shipped prebuilt (the GPU box).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_LEVELS = 22
NULL = 0xFFFFFFFF
COLOR_NULL = 0xC0000000

EDIT_AABB_FILL, EDIT_SPHERE_FILL, EDIT_SPHERE_DIG, EDIT_TERRAIN_FILL = 0, 1, 2, 3


class HdConfig(C.Structure):
    _fields_ = [("word_bits_per_page", C.c_uint32), ("page_bits_per_bucket", C.c_uint32),
                ("node_levels", C.c_uint32), ("bucket_bits_each_level", C.c_uint32 * MAX_LEVELS)]

    def bucket_bits(self):
        return [self.bucket_bits_each_level[i] for i in range(self.node_levels)]

    @property
    def voxel_level(self):
        return self.node_levels + 1

    def total_buckets(self):
        return sum(1 << b for b in self.bucket_bits())

    def total_words(self):
        return self.total_buckets() << (self.word_bits_per_page + self.page_bits_per_bucket)

    def level_bases(self):
        out, acc = [], 0
        for b in self.bucket_bits():
            out.append(acc)
            acc += 1 << b
        return out


class HdDefaultConfig(C.Structure):
    _fields_ = [("level_count", C.c_uint32), ("top_level_count", C.c_uint32), ("word_bits_per_page", C.c_uint32),
                ("page_bits_per_bucket", C.c_uint32), ("bucket_bits_per_top_level", C.c_uint32),
                ("bucket_bits_per_bottom_level", C.c_uint32)]


class HdEditDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("p0", C.c_uint32 * 3), ("p1", C.c_uint32 * 3), ("aux", C.c_uint32),
                ("r2", C.c_uint64)]


class HdTraceParams(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("look", C.c_float * 3), ("side", C.c_float * 3), ("up", C.c_float * 3),
                ("width", C.c_uint32), ("height", C.c_uint32), ("voxel_level", C.c_uint32), ("dag_root", C.c_uint32),
                ("dag_leaf_level", C.c_uint32), ("color_root", C.c_uint32), ("color_leaf_level", C.c_uint32),
                ("proj_factor", C.c_float), ("type", C.c_uint32)]


assert C.sizeof(HdTraceParams) == 84 and C.sizeof(HdEditDesc) == 40

HIT_DTYPE = np.dtype([("vox", np.uint32, 3), ("packed", np.uint32)])


def default_config(level_count=17, top_level_count=9, word_bits_per_page=9, page_bits_per_bucket=2,
                   bucket_bits_per_top_level=10, bucket_bits_per_bottom_level=16):
    """include/hashdag/Config.hpp:59-75 DefaultConfig{}() in pure Python."""
    cfg = HdConfig()
    cfg.word_bits_per_page, cfg.page_bits_per_bucket = word_bits_per_page, page_bits_per_bucket
    cfg.node_levels = level_count - 1
    for l in range(level_count - 1):
        cfg.bucket_bits_each_level[l] = bucket_bits_per_top_level if l < top_level_count else bucket_bits_per_bottom_level
    return cfg


def aabb(lo, hi):
    d = HdEditDesc()
    d.kind = EDIT_AABB_FILL
    d.p0[:] = lo
    d.p1[:] = hi
    return d


def sphere(center, r2, dig=False):
    d = HdEditDesc()
    d.kind = EDIT_SPHERE_DIG if dig else EDIT_SPHERE_FILL
    d.p0[:] = center
    d.r2 = int(r2)
    return d


def terrain(voxel_level, seed=0x5EED, octaves=4):
    """The synthetic noise terrain of SURVEY §8d cfg2 scaled to the resolution (oracle/terrain.h)."""
    res = 1 << voxel_level
    d = HdEditDesc()
    d.kind = EDIT_TERRAIN_FILL
    d.aux = seed
    d.p0[:] = (res // 4, voxel_level - 2, octaves)
    d.p1[:] = (res // 4, 0, 0)
    return d


def edit_array(edits):
    arr = (HdEditDesc * len(edits))()
    for i, e in enumerate(edits):
        arr[i] = e
    return arr


def random_spheres(n, voxel_level, seed=1234, rmin=16, rmax=256, y_lo=None, y_hi=None):
    """xorshift32 centres in the terrain band, radius uniform, alternating fill/dig."""
    res = 1 << voxel_level
    y_lo = res // 4 if y_lo is None else y_lo
    y_hi = res // 4 + res // 3 if y_hi is None else y_hi
    s = seed & 0xFFFFFFFF

    def nxt():
        nonlocal s
        s ^= (s << 13) & 0xFFFFFFFF
        s ^= s >> 17
        s ^= (s << 5) & 0xFFFFFFFF
        return s

    out = []
    for i in range(n):
        x, z = nxt() % res, nxt() % res
        y = y_lo + nxt() % max(1, y_hi - y_lo)
        r = rmin + nxt() % (rmax - rmin + 1)
        out.append(sphere((x, y, z), r * r, dig=bool(i & 1)))
    return out


def camera_params(cfg, root, pos, yaw, pitch, width, height, fov=np.pi / 3, color_root=COLOR_NULL,
                  color_leaf_level=10, type_=0, lod=True):
    """Push-constant block as src/rg/TracePass.cpp:106-134 + src/Camera.hpp:36-52 build it (float32 host maths)."""
    f = np.float32
    cy, sy, cp, sp = f(np.cos(f(yaw))), f(np.sin(f(yaw))), f(np.cos(f(pitch))), f(np.sin(f(pitch)))
    # trans = rotate(yaw, +Y) * rotate(pitch, -X); look = trans*(0,0,1); side = trans*(1,0,0)
    look = np.array([sy * cp, sp, cy * cp], dtype=f)
    side = np.array([cy, 0, -sy], dtype=f)
    look = look / f(np.sqrt(f(np.dot(look, look))))
    tg = f(np.tan(f(fov) * f(0.5)))
    aspect = f(width) / f(height)
    side = side / f(np.sqrt(f(np.dot(side, side)))) * tg * aspect
    up = np.cross(look, side).astype(f)
    up = up / f(np.sqrt(f(np.dot(up, up)))) * tg
    P = HdTraceParams()
    P.pos[:] = [float(v) for v in pos]
    P.look[:] = [float(v) for v in look]
    P.side[:] = [float(v) for v in side]
    P.up[:] = [float(v) for v in up]
    P.width, P.height = width, height
    P.voxel_level = cfg.node_levels + 1
    P.dag_root = root
    P.dag_leaf_level = cfg.node_levels
    P.color_root, P.color_leaf_level = color_root, color_leaf_level
    inv_2tan = f(1.0) / (f(2.0) * f(np.tan(f(0.5) * f(fov))))
    P.proj_factor = float(inv_2tan / (f(1.0) / f(height))) if lod else float("inf")
    P.type = type_
    return P


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32)) if a is not None else None


def build(ref=True):
    """(Re)build the oracle libraries in place.  Building the checker is not using it."""
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class _Lib:
    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)


class Oracle(_Lib):
    """The CPU restatement (liboracle.so)."""

    def __init__(self):
        super().__init__(os.path.join(HERE, "liboracle.so"))
        L = self.lib
        u32, u64, vp = C.c_uint32, C.c_uint64, C.c_void_p
        pu32 = C.POINTER(C.c_uint32)
        L.orc_hash_inner.restype = u32
        L.orc_hash_inner.argtypes = [pu32, u32]
        L.orc_hash_leaf.restype = u32
        L.orc_hash_leaf.argtypes = [pu32]
        L.orc_config_from_default.argtypes = [C.POINTER(HdDefaultConfig), C.POINTER(HdConfig)]
        L.orc_pool_create.restype = vp
        L.orc_pool_create.argtypes = [C.POINTER(HdConfig)]
        L.orc_pool_destroy.argtypes = [vp]
        L.orc_pool_words.restype = pu32
        L.orc_pool_words.argtypes = [vp]
        L.orc_pool_bucket_words.restype = pu32
        L.orc_pool_bucket_words.argtypes = [vp]
        L.orc_pool_total_words.restype = u64
        L.orc_pool_total_words.argtypes = [vp]
        L.orc_pool_total_buckets.restype = u32
        L.orc_pool_total_buckets.argtypes = [vp]
        L.orc_upsert.restype = u32
        L.orc_upsert.argtypes = [vp, u32, pu32, u32, u32]
        L.orc_filled_nodes.argtypes = [vp, pu32]
        L.orc_edit.restype = u32
        L.orc_edit.argtypes = [vp, u32, C.POINTER(HdEditDesc)]
        L.orc_edit_batch.restype = u32
        L.orc_edit_batch.argtypes = [vp, u32, C.POINTER(HdEditDesc), u32]
        L.orc_pool_stats.argtypes = [vp, C.POINTER(u64)]
        L.orc_pool_reset_stats.argtypes = [vp]
        L.orc_in_range_voxels.restype = u64
        L.orc_in_range_voxels.argtypes = [C.POINTER(HdEditDesc), u32]
        L.orc_terrain_height.restype = u32
        L.orc_terrain_height.argtypes = [C.POINTER(HdEditDesc), u32, u32]
        L.orc_canonical.argtypes = [pu32, u32, u32, C.POINTER(u64)]
        L.orc_voxel_get.restype = C.c_int
        L.orc_voxel_get.argtypes = [pu32, u32, u32, u32, u32, u32]
        L.orc_traverse.restype = C.c_int
        L.orc_traverse.argtypes = [pu32, u32, u32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_trace_frame.restype = u64
        L.orc_trace_frame.argtypes = [pu32, pu32, pu32, C.POINTER(HdTraceParams), u32, u32, u32, u32, pu32, vp, pu32]
        L.orc_color_fetch.argtypes = [pu32, pu32, u32, u32, u32, u32, u32, u32, C.POINTER(C.c_float)]

    def hash_inner(self, words):
        a = np.ascontiguousarray(words, dtype=np.uint32)
        return self.lib.orc_hash_inner(_u32p(a), len(a))

    def hash_leaf(self, words):
        a = np.ascontiguousarray(words, dtype=np.uint32)
        return self.lib.orc_hash_leaf(_u32p(a))

    def pool(self, cfg):
        return OraclePool(self, cfg)

    def canonical(self, words_ptr, node_levels, root):
        """dict(hash, by_ptr, by_content, voxels, per_level) of the DAG under root in a flat word array."""
        out = (C.c_uint64 * (4 + node_levels))()
        self.lib.orc_canonical(words_ptr, node_levels, root, out)
        return {"hash": out[0], "by_ptr": out[1], "by_content": out[2], "voxels": out[3],
                "per_level": [out[4 + i] for i in range(node_levels)]}

    def voxel_get(self, words_ptr, node_levels, root, x, y, z):
        return bool(self.lib.orc_voxel_get(words_ptr, node_levels, root, x, y, z))

    def traverse(self, words_ptr, node_levels, root, o, d):
        o3, d3, out = (C.c_float * 3)(*o), (C.c_float * 3)(*d), (C.c_float * 3)()
        hit = self.lib.orc_traverse(words_ptr, node_levels, root, o3, d3, out)
        return (np.array(out[:], dtype=np.float32) if hit else None)

    def trace_frame(self, words_ptr, params, color_nodes=None, color_leaves=None, rows=None, row_step=1,
                    threads=None, want=("rgba8", "hits", "iters")):
        W, H = params.width, params.height
        r0, r1 = rows if rows else (0, H)
        threads = threads or os.cpu_count()
        rgba = np.zeros((H, W), np.uint32) if "rgba8" in want else None
        hits = np.zeros((H, W), HIT_DTYPE) if "hits" in want else None
        iters = np.zeros((H, W), np.uint32) if "iters" in want else None
        cn = color_nodes if color_nodes is not None else np.zeros(8, np.uint32)
        cl = color_leaves if color_leaves is not None else np.zeros(8, np.uint32)
        cnp = cn if isinstance(cn, C.POINTER(C.c_uint32)) else _u32p(cn)
        clp = cl if isinstance(cl, C.POINTER(C.c_uint32)) else _u32p(cl)
        fetches = self.lib.orc_trace_frame(words_ptr, cnp, clp, C.byref(params), r0, r1, row_step, threads,
                                           _u32p(rgba), hits.ctypes.data if hits is not None else None, _u32p(iters))
        return {"rgba8": rgba, "hits": hits, "iters": iters, "fetches": fetches}

    def color_fetch(self, color_nodes, color_leaves, root, voxel_level, leaf_level, x, y, z):
        out = (C.c_float * 3)()
        cnp = color_nodes if isinstance(color_nodes, C.POINTER(C.c_uint32)) else _u32p(color_nodes)
        clp = color_leaves if isinstance(color_leaves, C.POINTER(C.c_uint32)) else _u32p(color_leaves)
        self.lib.orc_color_fetch(cnp, clp, root, voxel_level, leaf_level, x, y, z, out)
        return np.array(out[:], dtype=np.float32)

    def in_range_voxels(self, desc, voxel_level):
        return self.lib.orc_in_range_voxels(C.byref(desc), voxel_level)

    def terrain_height(self, desc, x, z):
        return self.lib.orc_terrain_height(C.byref(desc), x, z)


class _PoolBase:
    def words_np(self, offset=0, count=None):
        """numpy view (no copy) of the flat word space."""
        total = self.total_words
        count = total - offset if count is None else count
        buf = (C.c_uint32 * count).from_address(C.addressof(self.words_ptr.contents) + 4 * offset)
        return np.frombuffer(buf, dtype=np.uint32)

    def bucket_words_np(self):
        buf = (C.c_uint32 * self.total_buckets).from_address(C.addressof(self.bucket_words_ptr.contents))
        return np.frombuffer(buf, dtype=np.uint32)

    def used_ranges(self):
        """[(word_offset, count)] of the used prefix of every non-empty bucket."""
        bw = self.bucket_words_np()
        shift = self.cfg.word_bits_per_page + self.cfg.page_bits_per_bucket
        nz = np.nonzero(bw)[0]
        return [(int(b) << shift, int(bw[b])) for b in nz]


class OraclePool(_PoolBase):
    def __init__(self, orc, cfg):
        self.orc, self.cfg, self.L = orc, cfg, orc.lib
        self.h = self.L.orc_pool_create(C.byref(cfg))
        if not self.h:
            raise ValueError("invalid config")
        self.words_ptr = self.L.orc_pool_words(self.h)
        self.bucket_words_ptr = self.L.orc_pool_bucket_words(self.h)
        self.total_words = self.L.orc_pool_total_words(self.h)
        self.total_buckets = self.L.orc_pool_total_buckets(self.h)

    def close(self):
        if self.h:
            self.L.orc_pool_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def upsert(self, level, words, fallback=NULL):
        a = np.ascontiguousarray(words, dtype=np.uint32)
        return self.L.orc_upsert(self.h, level, _u32p(a), len(a), fallback)

    def filled_nodes(self):
        out = (C.c_uint32 * self.cfg.node_levels)()
        self.L.orc_filled_nodes(self.h, out)
        return list(out)

    def edit(self, root, desc):
        return self.L.orc_edit(self.h, root, C.byref(desc))

    def edit_batch(self, root, edits):
        arr = edit_array(edits)
        return self.L.orc_edit_batch(self.h, root, arr, len(edits))

    def stats(self):
        out = (C.c_uint64 * 8)()
        self.L.orc_pool_stats(self.h, out)
        keys = ["edit_nodes", "edit_leaves", "upserts", "appended_nodes", "appended_words", "overflow", "scan_words",
                "read_words"]
        return dict(zip(keys, out))

    def reset_stats(self):
        self.L.orc_pool_reset_stats(self.h)

    def canonical(self, root):
        return self.orc.canonical(self.words_ptr, self.cfg.node_levels, root)


class Ref(_Lib):
    """The reference's own headers, compiled (oracle/_ref/libhashdag_ref.so)."""

    PATH = os.path.join(HERE, "_ref", "libhashdag_ref.so")

    @staticmethod
    def available():
        return os.path.exists(Ref.PATH)

    def __init__(self):
        super().__init__(Ref.PATH)
        L = self.lib
        u32, u64, vp = C.c_uint32, C.c_uint64, C.c_void_p
        pu32 = C.POINTER(C.c_uint32)
        pf = C.POINTER(C.c_float)
        L.ref_hash_inner.restype = u32
        L.ref_hash_inner.argtypes = [pu32, u32]
        L.ref_hash_leaf.restype = u32
        L.ref_hash_leaf.argtypes = [pu32]
        L.ref_config_from_default.argtypes = [C.POINTER(HdDefaultConfig), C.POINTER(HdConfig)]
        L.ref_config_geometry.argtypes = [C.POINTER(HdConfig), C.POINTER(u64)]
        L.ref_pool_create.restype = vp
        L.ref_pool_create.argtypes = [C.POINTER(HdConfig)]
        L.ref_pool_destroy.argtypes = [vp]
        L.ref_pool_words.restype = pu32
        L.ref_pool_words.argtypes = [vp]
        L.ref_pool_bucket_words.restype = pu32
        L.ref_pool_bucket_words.argtypes = [vp]
        L.ref_pool_total_words.restype = u64
        L.ref_pool_total_words.argtypes = [vp]
        L.ref_pool_total_buckets.restype = u32
        L.ref_pool_total_buckets.argtypes = [vp]
        L.ref_upsert.restype = u32
        L.ref_upsert.argtypes = [vp, u32, pu32, u32, u32]
        L.ref_filled_nodes.argtypes = [vp, pu32]
        L.ref_edit.restype = u32
        L.ref_edit.argtypes = [vp, u32, C.POINTER(HdEditDesc), u32, u32]
        L.ref_edit_batch.restype = u32
        L.ref_edit_batch.argtypes = [vp, u32, C.POINTER(HdEditDesc), u32, u32, u32]
        L.ref_traverse.restype = C.c_int
        L.ref_traverse.argtypes = [vp, u32, pf, pf, pf]
        L.ref_trace_frame_host.restype = u64
        L.ref_trace_frame_host.argtypes = [vp, C.POINTER(HdTraceParams), u32, u32, u32, u32, C.POINTER(C.c_uint8), pf]
        L.ref_color_pool_create.restype = vp
        L.ref_color_pool_create.argtypes = [u32, u64, u64]
        L.ref_color_pool_destroy.argtypes = [vp]
        L.ref_color_root.restype = u32
        L.ref_color_root.argtypes = [vp]
        L.ref_color_nodes.restype = pu32
        L.ref_color_nodes.argtypes = [vp]
        L.ref_color_leaves.restype = pu32
        L.ref_color_leaves.argtypes = [vp]
        L.ref_color_node_words.restype = u64
        L.ref_color_node_words.argtypes = [vp]
        L.ref_color_leaf_words.restype = u64
        L.ref_color_leaf_words.argtypes = [vp]
        L.ref_edit_color.restype = u32
        L.ref_edit_color.argtypes = [vp, vp, u32, C.POINTER(HdEditDesc), u32, C.c_int]
        L.ref_color_at.restype = C.c_int
        L.ref_color_at.argtypes = [vp, u32, u32, u32, u32, pf]

    def hash_inner(self, words):
        a = np.ascontiguousarray(words, dtype=np.uint32)
        return self.lib.ref_hash_inner(_u32p(a), len(a))

    def hash_leaf(self, words):
        a = np.ascontiguousarray(words, dtype=np.uint32)
        return self.lib.ref_hash_leaf(_u32p(a))

    def config_from_default(self, **kw):
        dc = HdDefaultConfig(**kw)
        cfg = HdConfig()
        ok = self.lib.ref_config_from_default(C.byref(dc), C.byref(cfg)) == 0
        return cfg, ok

    def geometry(self, cfg):
        out = (C.c_uint64 * (6 + MAX_LEVELS))()
        self.lib.ref_config_geometry(C.byref(cfg), out)
        return {"node_levels": out[0], "words_per_page": out[1], "words_per_bucket": out[2], "total_buckets": out[3],
                "total_pages": out[4], "total_words": out[5], "level_bases": [out[6 + i] for i in range(cfg.node_levels)]}

    def pool(self, cfg):
        return RefPool(self, cfg)

    def color_pool(self, leaf_level, node_capacity=1 << 20, leaf_word_capacity=1 << 26):
        return RefColorPool(self, leaf_level, node_capacity, leaf_word_capacity)


class RefPool(_PoolBase):
    def __init__(self, ref, cfg):
        self.ref, self.cfg, self.L = ref, cfg, ref.lib
        self.h = self.L.ref_pool_create(C.byref(cfg))
        if not self.h:
            raise ValueError("invalid config")
        self.words_ptr = self.L.ref_pool_words(self.h)
        self.bucket_words_ptr = self.L.ref_pool_bucket_words(self.h)
        self.total_words = self.L.ref_pool_total_words(self.h)
        self.total_buckets = self.L.ref_pool_total_buckets(self.h)

    def close(self):
        if self.h:
            self.L.ref_pool_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def upsert(self, level, words, fallback=NULL):
        a = np.ascontiguousarray(words, dtype=np.uint32)
        return self.L.ref_upsert(self.h, level, _u32p(a), len(a), fallback)

    def filled_nodes(self):
        out = (C.c_uint32 * self.cfg.node_levels)()
        self.L.ref_filled_nodes(self.h, out)
        return list(out)

    def edit(self, root, desc, threads=0, max_task_level=10):
        return self.L.ref_edit(self.h, root, C.byref(desc), threads, max_task_level)

    def edit_batch(self, root, edits, threads=0, max_task_level=10):
        arr = edit_array(edits)
        return self.L.ref_edit_batch(self.h, root, arr, len(edits), threads, max_task_level)

    def traverse(self, root, o, d):
        o3, d3, out = (C.c_float * 3)(*o), (C.c_float * 3)(*d), (C.c_float * 3)()
        hit = self.L.ref_traverse(self.h, root, o3, d3, out)
        return np.array(out[:], dtype=np.float32) if hit else None

    def trace_frame_host(self, params, rows=None, row_step=1, threads=None, want_pos=True):
        W, H = params.width, params.height
        r0, r1 = rows if rows else (0, H)
        threads = threads or os.cpu_count()
        hit = np.zeros((H, W), np.uint8)
        pos = np.zeros((H, W, 3), np.float32) if want_pos else None
        n = self.L.ref_trace_frame_host(self.h, C.byref(params), r0, r1, row_step, threads,
                                        hit.ctypes.data_as(C.POINTER(C.c_uint8)),
                                        pos.ctypes.data_as(C.POINTER(C.c_float)) if want_pos else None)
        return {"hit": hit, "pos": pos, "n_hits": n}

    def edit_color(self, cpool, root, desc, rgb8, paint=False):
        return self.L.ref_edit_color(self.h, cpool.h, root, C.byref(desc), rgb8, int(paint))


class RefColorPool:
    def __init__(self, ref, leaf_level, node_capacity, leaf_word_capacity):
        self.L = ref.lib
        self.leaf_level = leaf_level
        self.h = self.L.ref_color_pool_create(leaf_level, node_capacity, leaf_word_capacity)

    def close(self):
        if self.h:
            self.L.ref_color_pool_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def root(self):
        return self.L.ref_color_root(self.h)

    def arrays(self):
        """(color_nodes, color_leaves) as numpy copies sized to what is used (min 8 words)."""
        nw, lw = self.L.ref_color_node_words(self.h), self.L.ref_color_leaf_words(self.h)
        n = np.ctypeslib.as_array(self.L.ref_color_nodes(self.h), shape=(max(nw, 8),)).copy()
        l = np.ctypeslib.as_array(self.L.ref_color_leaves(self.h), shape=(max(lw, 8),)).copy()
        return n, l

    def color_at(self, voxel_level, x, y, z):
        out = (C.c_float * 3)()
        ok = self.L.ref_color_at(self.h, voxel_level, x, y, z, out)
        return np.array(out[:], dtype=np.float32) if ok else None
