"""vkhashdag_b200/api.py — host-side mirror of the reference's pool interface over the C ABI (ctypes).

Names follow the reference so parity tests read like its own tests:
  DAGNodePool.Create / Edit / ThreadedEdit / Traversal / GetConfig / SetRoot / GetRoot / Flush
      (src/DAGNodePool.hpp:84-99, include/hashdag/NodePool.hpp:404-417, NodePoolThreadedEdit.hpp:104-126,
       NodePoolTraversal.hpp:93-256)
  AABBEditor / SphereEditor / TerrainEditor      editor structs (src/main.cpp:32-150) as POD descriptors

There is NO CPU fallback: importing works without a GPU (so the library can be inspected), every compute call
raises HashDagError(HD_ERR_NO_DEVICE) when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .abi import (COLOR_NULL, HIT_DTYPE, NULL, HdConfig, HdDefaultConfig, HdEditDesc, HdTraceParams, edit_array)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libhashdag_b200.so")

HD_OK, HD_ERR_INVALID, HD_ERR_CUDA, HD_ERR_OOM, HD_ERR_OVERFLOW, HD_ERR_NO_DEVICE = range(6)


class HdEditStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("visited_nodes", "visited_leaves", "upserts", "appended_nodes",
                                          "appended_words", "overflow_count", "in_range_voxels", "scan_words")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class HdTraceOutputs(C.Structure):
    _fields_ = [("rgba8", C.c_void_p), ("hits", C.c_void_p), ("iters", C.c_void_p), ("fetches", C.c_void_p)]


class HdTileShard(C.Structure):
    _fields_ = [("tile_w", C.c_uint32), ("tile_h", C.c_uint32), ("rank", C.c_uint32), ("world", C.c_uint32)]


class HdNodeRecord(C.Structure):
    _fields_ = [("ptr", C.c_uint32), ("level", C.c_uint32), ("n_words", C.c_uint32), ("words", C.c_uint32 * 9)]


class HdDirtyRange(C.Structure):
    _fields_ = [("word_offset", C.c_uint32), ("word_count", C.c_uint32)]


class HashDagError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"hd_status {status}: {msg}")
        self.status = status


_lib = None

# every symbol include/hashdag_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "hd_version", "hd_last_error", "hd_device_count", "hd_config_from_default", "hd_config_validate",
    "hd_config_total_buckets", "hd_config_total_words", "hd_config_level_base_bucket", "hd_pool_create",
    "hd_pool_destroy", "hd_pool_get_config", "hd_pool_clear", "hd_pool_set_root", "hd_pool_get_root",
    "hd_pool_words_dev", "hd_pool_bucket_words_dev", "hd_pool_stream", "hd_pool_upload_words", "hd_pool_read_words",
    "hd_pool_upload_bucket_words", "hd_pool_read_bucket_words", "hd_pool_filled_nodes", "hd_edit_batch",
    "hd_upsert_nodes", "hd_color_upload", "hd_trace", "hd_trace_dev", "hd_trace_tiles", "hd_trace_tiles_dev",
    "hd_tile_shard_pixels", "hd_traverse_ray", "hd_dirty_count", "hd_dirty_ranges", "hd_dirty_pack_dev",
    "hd_dirty_apply_dev", "hd_dirty_reset", "hd_pool_used_words", "hd_sync", "hd_kernel_launches", "hd_pool_save",
    "hd_pool_load", "hd_gc", "hd_trace_submit", "hd_trace_collect", "hd_beam_dev",
    "hd_trace_with_beam_dev", "hd_trace_with_beam", "hd_color_config", "hd_color_root", "hd_color_leaf_level",
    "hd_color_sizes", "hd_color_read", "hd_edit_color", "hd_edit_last_path",
    "hd_tile_shard_locate", "hd_pool_read_subtree", "hd_host_alloc", "hd_host_free", "hd_selftest_exact_arith", "hd_trace_table_info", "hd_selftest_edit_node8",
]


def lib():
    """Load libhashdag_b200.so (built in-tree by vkhashdag_b200/build.py).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -m vkhashdag_b200.build` (there is no fallback)")
    L = C.CDLL(LIB_PATH)
    u32, u64, vp, ci = C.c_uint32, C.c_uint64, C.c_void_p, C.c_int
    pu32 = C.POINTER(C.c_uint32)
    L.hd_version.restype = C.c_char_p
    L.hd_last_error.restype = C.c_char_p
    L.hd_device_count.restype = ci
    L.hd_config_from_default.argtypes = [C.POINTER(HdDefaultConfig), C.POINTER(HdConfig)]
    L.hd_config_validate.argtypes = [C.POINTER(HdConfig)]
    L.hd_config_total_buckets.restype = u32
    L.hd_config_total_buckets.argtypes = [C.POINTER(HdConfig)]
    L.hd_config_total_words.restype = u64
    L.hd_config_total_words.argtypes = [C.POINTER(HdConfig)]
    L.hd_config_level_base_bucket.restype = u32
    L.hd_config_level_base_bucket.argtypes = [C.POINTER(HdConfig), u32]
    L.hd_pool_create.argtypes = [C.POINTER(HdConfig), ci, C.POINTER(vp)]
    L.hd_pool_destroy.argtypes = [vp]
    L.hd_pool_destroy.restype = None
    L.hd_pool_get_config.argtypes = [vp, C.POINTER(HdConfig)]
    L.hd_pool_clear.argtypes = [vp]
    L.hd_pool_set_root.argtypes = [vp, u32]
    L.hd_pool_get_root.restype = u32
    L.hd_pool_get_root.argtypes = [vp]
    for f in ("hd_pool_words_dev", "hd_pool_bucket_words_dev", "hd_pool_stream"):
        getattr(L, f).restype = vp
        getattr(L, f).argtypes = [vp]
    L.hd_pool_upload_words.argtypes = [vp, u32, vp, u32]
    L.hd_pool_read_words.argtypes = [vp, u32, vp, u32]
    L.hd_pool_read_subtree.argtypes = [vp, u32, u32, u32, C.POINTER(HdNodeRecord), u32, pu32]
    L.hd_pool_upload_bucket_words.argtypes = [vp, u32, vp, u32]
    L.hd_pool_read_bucket_words.argtypes = [vp, u32, vp, u32]
    L.hd_pool_filled_nodes.argtypes = [vp, pu32]
    L.hd_edit_batch.argtypes = [vp, u32, C.POINTER(HdEditDesc), u32, pu32, C.POINTER(HdEditStats)]
    L.hd_upsert_nodes.argtypes = [vp, u32, vp, u32, u32, vp]
    L.hd_color_upload.argtypes = [vp, vp, u64, vp, u64]
    L.hd_trace.argtypes = [vp, C.POINTER(HdTraceParams), C.POINTER(HdTraceOutputs)]
    L.hd_trace_dev.argtypes = [vp, C.POINTER(HdTraceParams), C.POINTER(HdTraceOutputs)]
    L.hd_trace_tiles.argtypes = [vp, C.POINTER(HdTraceParams), C.POINTER(HdTileShard), C.POINTER(HdTraceOutputs)]
    L.hd_trace_tiles_dev.argtypes = [vp, C.POINTER(HdTraceParams), C.POINTER(HdTileShard), C.POINTER(HdTraceOutputs)]
    L.hd_color_config.argtypes = [vp, u32, u32]
    L.hd_color_root.restype = u32
    L.hd_color_root.argtypes = [vp]
    L.hd_color_leaf_level.restype = u32
    L.hd_color_leaf_level.argtypes = [vp]
    L.hd_color_sizes.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.hd_color_read.argtypes = [vp, vp, u64, vp, u64]
    L.hd_edit_color.argtypes = [vp, u32, C.POINTER(HdEditDesc), u32, u32, pu32, pu32, C.POINTER(HdEditStats)]
    L.hd_beam_dev.argtypes = [vp, C.POINTER(HdTraceParams), vp]
    L.hd_trace_with_beam_dev.argtypes = [vp, C.POINTER(HdTraceParams), vp, u32, u32, C.POINTER(HdTraceOutputs)]
    L.hd_trace_with_beam.argtypes = [vp, C.POINTER(HdTraceParams), C.POINTER(HdTraceParams), C.POINTER(HdTraceOutputs), vp]
    L.hd_trace_submit.argtypes = [vp, C.POINTER(HdTraceParams), C.POINTER(HdTileShard), vp, u32]
    L.hd_trace_collect.argtypes = [vp, u32]
    L.hd_tile_shard_pixels.restype = u64
    L.hd_tile_shard_pixels.argtypes = [C.POINTER(HdTraceParams), C.POINTER(HdTileShard)]
    L.hd_tile_shard_locate.argtypes = [C.POINTER(HdTraceParams), C.POINTER(HdTileShard), u32, pu32, pu32]
    L.hd_traverse_ray.argtypes = [vp, u32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(ci),
                                  C.POINTER(C.c_float)]
    L.hd_dirty_count.argtypes = [vp, pu32, C.POINTER(u64)]
    L.hd_edit_last_path.argtypes = [vp]
    L.hd_edit_last_path.restype = u32
    L.hd_dirty_ranges.argtypes = [vp, C.POINTER(HdDirtyRange), u32, pu32]
    L.hd_dirty_pack_dev.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.hd_dirty_apply_dev.argtypes = [vp, vp, u64]
    L.hd_dirty_reset.argtypes = [vp]
    L.hd_gc.argtypes = [vp, pu32, u32, pu32, C.POINTER(u64)]
    L.hd_pool_save.argtypes = [vp, C.c_char_p]
    L.hd_pool_load.argtypes = [C.c_char_p, ci, C.POINTER(vp)]
    L.hd_pool_used_words.argtypes = [vp, C.POINTER(u64)]
    L.hd_sync.argtypes = [vp]
    L.hd_host_alloc.argtypes = [u64, ci, C.POINTER(vp)]
    L.hd_host_free.argtypes = [vp]
    L.hd_kernel_launches.restype = u64
    L.hd_selftest_exact_arith.argtypes = [ci, C.POINTER(u64)]
    L.hd_selftest_edit_node8.argtypes = [ci, u32, C.POINTER(u64)]
    L.hd_trace_table_info.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    _lib = L
    return L


def _check(status):
    if status != HD_OK:
        raise HashDagError(status, lib().hd_last_error().decode(errors="replace"))


class HostBuffer:
    """Page-locked host memory from hd_host_alloc as a numpy uint32 array (`.array`); write_combined=True for frame
    read-back targets the CPU will not read much."""

    def __init__(self, n_words, write_combined=False):
        self._p = C.c_void_p()
        _check(lib().hd_host_alloc(int(n_words) * 4, int(write_combined), C.byref(self._p)))
        self.array = np.ctypeslib.as_array((C.c_uint32 * int(n_words)).from_address(self._p.value))

    def free(self):
        if self._p:
            self.array = None
            lib().hd_host_free(self._p)
            self._p = C.c_void_p()


def selftest_exact_arith(device=0):
    """Mismatches between the trace kernel's exact-arithmetic shortcuts and the IEEE operations they replace (0 = exact)."""
    n = C.c_uint64(0)
    _check(lib().hd_selftest_exact_arith(int(device), C.byref(n)))
    return int(n.value)


def selftest_edit_node8(n_cases=1 << 24, device=0):
    """Nodes (of n_cases pseudo-random editor / node pairs) whose eight child classifications differ between edit_node8 and
    eight edit_node calls (0 = identical)."""
    n = C.c_uint64(0)
    _check(lib().hd_selftest_edit_node8(int(device), int(n_cases), C.byref(n)))
    return int(n.value)


def kernel_launches():
    return int(lib().hd_kernel_launches())


# ---- editors (src/main.cpp:32-150) -----------------------------------------------------------------
class AABBEditor:
    def __init__(self, aabb_min, aabb_max):
        self.desc = abi.aabb(aabb_min, aabb_max)


class SphereEditor:
    """mode: 'fill' (EditMode::kFill) or 'dig' (EditMode::kDig)."""

    def __init__(self, center, r2, mode="fill"):
        if mode not in ("fill", "dig"):
            raise ValueError("GPU geometry editors: fill | dig (paint only changes colour: SURVEY §8f N2)")
        self.desc = abi.sphere(center, r2, dig=(mode == "dig"))


class TerrainEditor:
    def __init__(self, voxel_level, seed=0x5EED, octaves=4, amp_div=8, extent_bits=0):
        self.desc = abi.terrain(voxel_level, seed, octaves, amp_div, extent_bits)


def _desc(e):
    return e if isinstance(e, HdEditDesc) else e.desc


class DAGNodePool:
    """Device-resident hashed node pool (replaces src/DAGNodePool.{hpp,cpp} + the hashdag mix-ins it derives from)."""

    def __init__(self, config, device=0):
        self._h = C.c_void_p()
        self._L = lib()
        self.config = config
        _check(self._L.hd_pool_create(C.byref(config), device, C.byref(self._h)))
        self.device = device
        self.last_stats = None

    @classmethod
    def Create(cls, config, device=0):
        return cls(config, device)

    @classmethod
    def Load(cls, path, device=0):
        """hd_pool_load: a pool file written by Save (config, nodes, root, colour buffers)."""
        self = cls.__new__(cls)
        self._h, self._L, self.device, self.last_stats = C.c_void_p(), lib(), device, None
        _check(self._L.hd_pool_load(os.fsencode(path), device, C.byref(self._h)))
        self.config = HdConfig()
        _check(self._L.hd_pool_get_config(self._h, C.byref(self.config)))
        return self

    def ThreadedGC(self, roots):
        """NodePoolThreadedGC::ThreadedGC (NodePoolThreadedGC.hpp:394-403): accepts one root or a list; returns the
        remapped root(s).  self.last_gc_nodes = reachable unique nodes kept (filled nodes included)."""
        single = isinstance(roots, int)
        arr = np.ascontiguousarray([roots] if single else roots, dtype=np.uint32)
        out = np.empty_like(arr)
        kept = C.c_uint64()
        _check(self._L.hd_gc(self._h, arr.ctypes.data_as(C.POINTER(C.c_uint32)), arr.size,
                             out.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(kept)))
        self.last_gc_nodes = kept.value
        return int(out[0]) if single else [int(v) for v in out]

    def Save(self, path):
        _check(self._L.hd_pool_save(self._h, os.fsencode(path)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.hd_pool_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- NodePoolBase surface --
    def GetConfig(self):
        return self.config

    def SetRoot(self, root):
        _check(self._L.hd_pool_set_root(self._h, root))

    def GetRoot(self):
        return self._L.hd_pool_get_root(self._h)

    def Edit(self, root, editor):
        """NodePoolBase::Edit (NodePool.hpp:405-417): one edit, returns the new root pointer."""
        return self.EditBatch(root, [editor])

    def ThreadedEdit(self, root, editor, max_task_level=None):
        """NodePoolThreadedEdit::ThreadedEdit (NodePoolThreadedEdit.hpp:104-126).  The GPU path has no task level."""
        return self.EditBatch(root, [editor])

    def EditBatch(self, root, editors):
        """Apply editors in order in one GPU pass (same canonical DAG as sequential reference Edit calls)."""
        # a prebuilt (HdEditDesc * n) array is passed through untouched (marshalling 10^4 descriptors in Python costs
        # more than the GPU pass)
        arr = editors if isinstance(editors, C.Array) else edit_array([_desc(e) for e in editors])
        out = C.c_uint32(root)
        st = HdEditStats()
        _check(self._L.hd_edit_batch(self._h, root, arr, len(arr), C.byref(out), C.byref(st)))
        self.last_stats = st.as_dict()
        self.last_stats["path"] = ("general", "fused", "graph")[self._L.hd_edit_last_path(self._h)]
        return out.value

    def Upsert(self, level, nodes, words_each):
        """upsert_inner_node / upsert_leaf (NodePool.hpp:228-238) for n explicit packed nodes."""
        a = np.ascontiguousarray(nodes, dtype=np.uint32).reshape(-1, words_each)
        out = np.empty(a.shape[0], np.uint32)
        _check(self._L.hd_upsert_nodes(self._h, level, a.ctypes.data, words_each, a.shape[0], out.ctypes.data))
        return out

    def FilledNodes(self):
        out = (C.c_uint32 * self.config.node_levels)()
        _check(self._L.hd_pool_filled_nodes(self._h, out))
        return list(out)

    def Traversal(self, root, o, d):
        """NodePoolTraversal::Traversal<float> (NodePoolTraversal.hpp:93-256): float32 hit position or None."""
        o3, d3, out, hit = (C.c_float * 3)(*o), (C.c_float * 3)(*d), (C.c_float * 3)(), C.c_int(0)
        _check(self._L.hd_traverse_ray(self._h, root, o3, d3, C.byref(hit), out))
        return np.array(out[:], dtype=np.float32) if hit.value else None

    def Flush(self):
        """DAGNodePool::Flush (src/DAGNodePool.cpp:56-85): device memory is the pool, so this only synchronises."""
        _check(self._L.hd_sync(self._h))

    def Clear(self):
        _check(self._L.hd_pool_clear(self._h))

    # -- host mirror interop (ReadPage / WritePage) --
    def UploadWords(self, word_offset, words):
        a = np.ascontiguousarray(words, dtype=np.uint32)
        _check(self._L.hd_pool_upload_words(self._h, word_offset, a.ctypes.data, a.size))

    def ReadWords(self, word_offset, count):
        out = np.empty(count, np.uint32)
        _check(self._L.hd_pool_read_words(self._h, word_offset, out.ctypes.data, count))
        return out

    def ReadSubtree(self, root, level=0, depth=64, capacity=1 << 16):
        """hd_pool_read_subtree: [(ptr, level, [words])] of the subtree under `root`, breadth first, one device pass."""
        arr = (HdNodeRecord * capacity)()
        n = C.c_uint32()
        st = self._L.hd_pool_read_subtree(self._h, root, level, depth, arr, capacity, C.byref(n))
        if st not in (HD_OK, HD_ERR_OVERFLOW):
            _check(st)
        return [(r.ptr, r.level, list(r.words[:r.n_words])) for r in arr[:n.value]], st == HD_ERR_OVERFLOW

    def UploadBucketWords(self, first_bucket, values):
        a = np.ascontiguousarray(values, dtype=np.uint32)
        _check(self._L.hd_pool_upload_bucket_words(self._h, first_bucket, a.ctypes.data, a.size))

    def ReadBucketWords(self):
        out = np.empty(self.config.total_buckets(), np.uint32)
        _check(self._L.hd_pool_read_bucket_words(self._h, 0, out.ctypes.data, out.size))
        return out

    def UploadFrom(self, host_pool):
        """Mirror a host pool (anything with used_ranges()/words_np()/bucket_words_np()) onto the device."""
        for off, cnt in host_pool.used_ranges():
            self.UploadWords(off, host_pool.words_np(off, cnt))
        self.UploadBucketWords(0, host_pool.bucket_words_np())

    def Download(self):
        """(words_by_range dict, bucket_words): used prefix of every non-empty bucket, read back to the host."""
        bw = self.ReadBucketWords()
        shift = self.config.word_bits_per_page + self.config.page_bits_per_bucket
        return {int(b) << shift: self.ReadWords(int(b) << shift, int(bw[b])) for b in np.nonzero(bw)[0]}, bw

    def DownloadInto(self, host_pool, chunk_words=1 << 26):
        """Mirror the device pool into a host pool object (words_np(offset, count) / bucket_words_np() views, e.g. a
        ReadPage/WritePage-style mirror): runs of consecutive non-empty buckets are read as whole buckets straight into
        the mirror's memory (a bucket's unused tail is zero on the device).  Returns the words read."""
        bw = self.ReadBucketWords()
        host_pool.bucket_words_np()[:] = bw
        shift = self.config.word_bits_per_page + self.config.page_bits_per_bucket
        nz = np.flatnonzero(bw)
        if nz.size == 0:
            return 0
        cuts = np.flatnonzero(np.diff(nz) > 1) + 1
        total = 0
        for run in np.split(nz, cuts):
            off, cnt = int(run[0]) << shift, (int(run[-1]) - int(run[0]) + 1) << shift
            for o in range(off, off + cnt, chunk_words):
                c = min(chunk_words, off + cnt - o)
                dst = host_pool.words_np(o, c)
                _check(self._L.hd_pool_read_words(self._h, o, dst.ctypes.data, c))
                total += c
        return total

    def UsedWords(self):
        out = C.c_uint64()
        _check(self._L.hd_pool_used_words(self._h, C.byref(out)))
        return out.value

    def UploadColor(self, color_nodes, color_leaves):
        n = np.ascontiguousarray(color_nodes, dtype=np.uint32)
        l = np.ascontiguousarray(color_leaves, dtype=np.uint32)
        _check(self._L.hd_color_upload(self._h, n.ctypes.data, n.size, l.ctypes.data, l.size))

    # -- colour pool (DAGColorPool) --
    def ColorConfig(self, leaf_level, color_root=COLOR_NULL):
        """DAGColorPool::Config::leaf_level + SetRoot (src/main.cpp:204-211, DAGColorPool.hpp:209)."""
        _check(self._L.hd_color_config(self._h, leaf_level, color_root))

    def ColorRoot(self):
        return self._L.hd_color_root(self._h)

    def ReadColor(self):
        """(color_nodes, color_leaves) numpy copies of the used parts of the two colour buffers."""
        n, l = C.c_uint64(), C.c_uint64()
        _check(self._L.hd_color_sizes(self._h, C.byref(n), C.byref(l)))
        nodes, leaves = np.zeros(max(n.value, 8), np.uint32), np.zeros(max(l.value, 8), np.uint32)
        _check(self._L.hd_color_read(self._h, nodes.ctypes.data, n.value, leaves.ctypes.data, l.value))
        return nodes, leaves

    def EditColor(self, root, editor, rgb8, paint=False):
        """vbr_edit(editor) of src/main.cpp:224-230: geometry edit + colour update on the GPU.  Returns
        (new node root, new colour root)."""
        d = _desc(editor)
        out, cout, st = C.c_uint32(root), C.c_uint32(), HdEditStats()
        _check(self._L.hd_edit_color(self._h, root, C.byref(d), rgb8, int(paint), C.byref(out), C.byref(cout), C.byref(st)))
        self.last_stats = st.as_dict()
        return out.value, cout.value

    # -- trace --
    def Trace(self, params, want=("rgba8", "hits", "iters"), shard=None, out=None):
        """One frame through hd_trace / hd_trace_tiles with HOST outputs.  Returns dict of numpy arrays."""
        if shard is None:
            n = params.width * params.height
            shape = (params.height, params.width)
        else:
            shard = HdTileShard(*shard) if not isinstance(shard, HdTileShard) else shard
            n = int(self._L.hd_tile_shard_pixels(C.byref(params), C.byref(shard)))
            shape = (n,)
        res = out or {}
        if "rgba8" in want and "rgba8" not in res:
            res["rgba8"] = np.zeros(shape, np.uint32)
        if "hits" in want and "hits" not in res:
            res["hits"] = np.zeros(shape, HIT_DTYPE)
        if "iters" in want and "iters" not in res:
            res["iters"] = np.zeros(shape, np.uint32)
        if "fetches" in want and "fetches" not in res:
            res["fetches"] = np.zeros(shape, np.uint32)
        o = HdTraceOutputs(res["rgba8"].ctypes.data if "rgba8" in want else None,
                           res["hits"].ctypes.data if "hits" in want else None,
                           res["iters"].ctypes.data if "iters" in want else None,
                           res["fetches"].ctypes.data if "fetches" in want else None)
        if shard is None:
            _check(self._L.hd_trace(self._h, C.byref(params), C.byref(o)))
        else:
            _check(self._L.hd_trace_tiles(self._h, C.byref(params), C.byref(shard), C.byref(o)))
        return res

    def TraceTableInfo(self):
        """(root, node levels, nodes) of the staged top levels the trace kernel currently holds; root == NULL: none."""
        r, l, n = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(self._L.hd_trace_table_info(self._h, C.byref(r), C.byref(l), C.byref(n)))
        return int(r.value), int(l.value), int(n.value)

    def TraceDev(self, params, rgba8=0, hits=0, iters=0, fetches=0, shard=None):
        """Enqueue one frame with DEVICE output pointers (ints, e.g. torch.Tensor.data_ptr()); no synchronisation."""
        o = HdTraceOutputs(rgba8 or None, hits or None, iters or None, fetches or None)
        if shard is None:
            _check(self._L.hd_trace_dev(self._h, C.byref(params), C.byref(o)))
        else:
            shard = HdTileShard(*shard) if not isinstance(shard, HdTileShard) else shard
            _check(self._L.hd_trace_tiles_dev(self._h, C.byref(params), C.byref(shard), C.byref(o)))

    def TraceBeam(self, params, beam_params, want=("rgba8", "hits", "iters")):
        """Beam pre-pass + beam-optimised trace (BeamPass + trace.frag with BEAM_OPTIMIZATION), host outputs.
        Returns the planes plus 'beam' (float32 [bh, bw])."""
        shape = (params.height, params.width)
        res = {"beam": np.zeros((beam_params.height, beam_params.width), np.float32)}
        for name, dt in (("rgba8", np.uint32), ("hits", HIT_DTYPE), ("iters", np.uint32), ("fetches", np.uint32)):
            if name in want:
                res[name] = np.zeros(shape, dt)
        o = HdTraceOutputs(*[res[n].ctypes.data if n in res else None for n in ("rgba8", "hits", "iters", "fetches")])
        _check(self._L.hd_trace_with_beam(self._h, C.byref(params), C.byref(beam_params), C.byref(o), res["beam"].ctypes.data))
        return res

    def BeamDev(self, beam_params, beam_dev_ptr):
        _check(self._L.hd_beam_dev(self._h, C.byref(beam_params), beam_dev_ptr))

    def TraceBeamDev(self, params, beam_dev_ptr, bw, bh, rgba8=0, hits=0, iters=0):
        o = HdTraceOutputs(rgba8 or None, hits or None, iters or None, None)
        _check(self._L.hd_trace_with_beam_dev(self._h, C.byref(params), beam_dev_ptr, bw, bh, C.byref(o)))

    def TraceSubmit(self, params, host_rgba8, slot, shard=None):
        """hd_trace_submit: enqueue one frame + async read-back of its rgba8 plane into a host array (numpy, ideally
        backed by pinned memory).  Pair with TraceCollect(slot); two slots may be in flight."""
        sh = None
        if shard is not None:
            sh = HdTileShard(*shard) if not isinstance(shard, HdTileShard) else shard
        _check(self._L.hd_trace_submit(self._h, C.byref(params), C.byref(sh) if sh is not None else None,
                                       host_rgba8.ctypes.data, slot))

    def TraceCollect(self, slot):
        _check(self._L.hd_trace_collect(self._h, slot))

    def ShardPixels(self, params, shard):
        shard = HdTileShard(*shard) if not isinstance(shard, HdTileShard) else shard
        return int(self._L.hd_tile_shard_pixels(C.byref(params), C.byref(shard)))

    def Sync(self):
        _check(self._L.hd_sync(self._h))

    @property
    def stream(self):
        return self._L.hd_pool_stream(self._h)

    @property
    def words_dev(self):
        return self._L.hd_pool_words_dev(self._h)

    # -- replica sync --
    def DirtyCount(self):
        n, b = C.c_uint32(), C.c_uint64()
        _check(self._L.hd_dirty_count(self._h, C.byref(n), C.byref(b)))
        return n.value, b.value

    def DirtyRanges(self):
        n, _ = self.DirtyCount()
        arr = (HdDirtyRange * max(n, 1))()
        got = C.c_uint32()
        _check(self._L.hd_dirty_ranges(self._h, arr, n, C.byref(got)))
        return [(arr[i].word_offset, arr[i].word_count) for i in range(min(n, got.value))]

    def DirtyPack(self, staging_dev_ptr, capacity_bytes):
        """Pack the dirty ranges into a device staging buffer; raises replica.StagingTooSmall(need) if it cannot hold them."""
        b = C.c_uint64()
        st = self._L.hd_dirty_pack_dev(self._h, staging_dev_ptr, capacity_bytes, C.byref(b))
        if st == HD_ERR_OVERFLOW and b.value > capacity_bytes:
            from .replica import StagingTooSmall
            raise StagingTooSmall(b.value)
        _check(st)
        return b.value

    def DirtyApply(self, staging_dev_ptr, packed_bytes):
        _check(self._L.hd_dirty_apply_dev(self._h, staging_dev_ptr, packed_bytes))

    def DirtyReset(self):
        _check(self._L.hd_dirty_reset(self._h))
