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
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

HERE = os.path.dirname(os.path.abspath(__file__))
from vkhashdag_b200.abi import *  # noqa: F401,F403 (ABI structs + descriptor helpers live with the product)
from vkhashdag_b200.abi import HdConfig, HdDefaultConfig, HdEditDesc, HdTraceParams, HIT_DTYPE, MAX_LEVELS, NULL, edit_array


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
        L.orc_canonical_fast.argtypes = [pu32, C.POINTER(HdConfig), u32, u32, C.POINTER(u64)]
        L.orc_count_stored_nodes.argtypes = [pu32, pu32, C.POINTER(HdConfig), C.POINTER(u64)]
        L.orc_voxel_get.restype = C.c_int
        L.orc_voxel_get.argtypes = [pu32, u32, u32, u32, u32, u32]
        L.orc_traverse.restype = C.c_int
        L.orc_traverse.argtypes = [pu32, u32, u32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_trace_frame.restype = u64
        L.orc_trace_frame.argtypes = [pu32, pu32, pu32, C.POINTER(HdTraceParams), u32, u32, u32, u32, pu32, vp, pu32]
        L.orc_trace_frame_beam.restype = u64
        L.orc_trace_frame_beam.argtypes = [pu32, pu32, pu32, C.POINTER(HdTraceParams), u32, u32, u32, u32, pu32, vp, pu32,
                                           C.POINTER(C.c_float), u32, u32]
        L.orc_beam_frame.argtypes = [pu32, C.POINTER(HdTraceParams), C.POINTER(C.c_float)]
        L.orc_trace_frame_host.restype = u64
        L.orc_trace_frame_host.argtypes = [pu32, u32, C.POINTER(HdTraceParams), u32, u32, u32, u32, C.POINTER(C.c_uint8),
                                           C.POINTER(C.c_float)]
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

    def canonical_fast(self, words_ptr, cfg, root, threads=0):
        """Same dict as canonical(), computed level by level on `threads` host threads (bench-scale pools)."""
        out = (C.c_uint64 * (4 + cfg.node_levels))()
        self.lib.orc_canonical_fast(words_ptr, C.byref(cfg), root, threads, out)
        return {"hash": out[0], "by_ptr": out[1], "by_content": out[2], "voxels": out[3],
                "per_level": [out[4 + i] for i in range(cfg.node_levels)]}

    def count_stored_nodes(self, pool):
        """[stored node count per level] by walking every bucket of a host pool/mirror."""
        out = (C.c_uint64 * pool.cfg.node_levels)()
        self.lib.orc_count_stored_nodes(pool.words_ptr, pool.bucket_words_ptr, C.byref(pool.cfg), out)
        return list(out)

    def voxel_get(self, words_ptr, node_levels, root, x, y, z):
        return bool(self.lib.orc_voxel_get(words_ptr, node_levels, root, x, y, z))

    def traverse(self, words_ptr, node_levels, root, o, d):
        o3, d3, out = (C.c_float * 3)(*o), (C.c_float * 3)(*d), (C.c_float * 3)()
        hit = self.lib.orc_traverse(words_ptr, node_levels, root, o3, d3, out)
        return (np.array(out[:], dtype=np.float32) if hit else None)

    def beam_frame(self, words_ptr, beam_params):
        """beam.frag over the beam image described by beam_params (abi.beam_params); float32 [bh, bw]."""
        out = np.zeros((beam_params.height, beam_params.width), np.float32)
        self.lib.orc_beam_frame(words_ptr, C.byref(beam_params), out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def trace_frame(self, words_ptr, params, color_nodes=None, color_leaves=None, rows=None, row_step=1,
                    threads=None, want=("rgba8", "hits", "iters"), beam=None):
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
        bp, bw, bh = None, 0, 0
        if beam is not None:
            beam = np.ascontiguousarray(beam, dtype=np.float32)
            bp, bw, bh = beam.ctypes.data_as(C.POINTER(C.c_float)), beam.shape[1], beam.shape[0]
        fetches = self.lib.orc_trace_frame_beam(words_ptr, cnp, clp, C.byref(params), r0, r1, row_step, threads,
                                                _u32p(rgba), hits.ctypes.data if hits is not None else None, _u32p(iters),
                                                bp, bw, bh)
        return {"rgba8": rgba, "hits": hits, "iters": iters, "fetches": fetches}

    def trace_frame_host(self, words_ptr, node_levels, params, rows=None, row_step=1, threads=None, want_pos=True):
        W, H = params.width, params.height
        r0, r1 = rows if rows else (0, H)
        threads = threads or os.cpu_count()
        hit = np.zeros((H, W), np.uint8)
        pos = np.zeros((H, W, 3), np.float32) if want_pos else None
        n = self.lib.orc_trace_frame_host(words_ptr, node_levels, C.byref(params), r0, r1, row_step, threads,
                                          hit.ctypes.data_as(C.POINTER(C.c_uint8)),
                                          pos.ctypes.data_as(C.POINTER(C.c_float)) if want_pos else None)
        return {"hit": hit, "pos": pos, "n_hits": n}

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
        L.ref_gc.restype = u32
        L.ref_gc.argtypes = [vp, u32, u32]
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
        L.ref_edit_color_mt.restype = u32
        L.ref_edit_color_mt.argtypes = [vp, vp, u32, C.POINTER(HdEditDesc), u32, C.c_int, u32]
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

    def gc(self, root, threads=4):
        """NodePoolThreadedGC::ThreadedGC: compacts the pool, returns the relocated root."""
        return self.L.ref_gc(self.h, root, threads)

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

    def edit_color(self, cpool, root, desc, rgb8, paint=False, threads=0):
        """vbr_edit: serial Edit (threads=0) or ThreadedEdit(busy_pool(threads), max_task_level = colour leaf level)."""
        return self.L.ref_edit_color_mt(self.h, cpool.h, root, C.byref(desc), rgb8, int(paint), threads)


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
