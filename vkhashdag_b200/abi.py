"""vkhashdag_b200/abi.py — ctypes mirror of include/hashdag_b200.h plus the host-side helpers that build its PODs.

Mirrors, on the Python side, the reference's host types for this path:
  HdConfig / default_config   include/hashdag/Config.hpp:15-75
  aabb / sphere / terrain     editor structs of src/main.cpp:32-150 as hd_edit_desc
  camera_params               src/Camera.hpp:36-52 + src/rg/TracePass.cpp:106-134 (push-constant block)
"""
import ctypes as C

import numpy as np

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


def custom_config(bucket_bits_each_level, word_bits_per_page=9, page_bits_per_bucket=2):
    """Config with an explicit per-level bucket count (Config::bucket_bits_each_level, Config.hpp:21)."""
    cfg = HdConfig()
    cfg.word_bits_per_page, cfg.page_bits_per_bucket = word_bits_per_page, page_bits_per_bucket
    cfg.node_levels = len(bucket_bits_each_level)
    for l, b in enumerate(bucket_bits_each_level):
        cfg.bucket_bits_each_level[l] = b
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


def terrain(voxel_level, seed=0x5EED, octaves=4, amp_div=8, extent_bits=0):
    """The synthetic noise terrain of SURVEY §8d cfg2 (oracle/terrain.h).  With E = extent (the whole
    world, or a 2^extent_bits patch of it): base height E/4, octave o has lattice cell E/4^(o+1) and amplitude
    E/(amp_div*4^o)."""
    bits = extent_bits if extent_bits else voxel_level
    ext = 1 << bits
    d = HdEditDesc()
    d.kind = EDIT_TERRAIN_FILL
    d.aux = seed
    d.p0[:] = (ext // 4, bits - 2, octaves)
    d.p1[:] = (ext // amp_div, extent_bits, 0)
    return d


def edit_array(edits):
    arr = (HdEditDesc * len(edits))()
    for i, e in enumerate(edits):
        arr[i] = e
    return arr


def random_spheres(n, voxel_level, seed=1234, rmin=16, rmax=256, y_lo=None, y_hi=None, extent_bits=0):
    """xorshift32 centres in the terrain band (of the patch), radius uniform, alternating fill/dig."""
    res = 1 << (extent_bits if extent_bits else voxel_level)
    y_lo = res // 4 if y_lo is None else y_lo
    y_hi = res // 4 + res // 6 if y_hi is None else y_hi
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



def spheres_in_range_voxels(spheres, voxel_level):
    """Sum over sphere edits of |{v in world : VoxelInRange(v)}| (main.cpp:127-132) — the implementation-independent
    "edited voxels" unit of SURVEY §8d — computed on the host from the definition: per distinct r2 a table of y-column
    lengths over (dx, dz) with 2-D prefix sums, clipped per sphere to the world in x and z (y is clipped per column)."""
    res = 1 << voxel_level
    tables = {}
    total = 0
    for e in spheres:
        r2 = int(e.r2)
        cx, cy, cz = int(e.p0[0]), int(e.p0[1]), int(e.p0[2])
        r = int(np.sqrt(float(r2)))
        while (r + 1) * (r + 1) <= r2:
            r += 1
        while r * r > r2:
            r -= 1
        unclipped_y = cy - r >= 0 and cy + r < res
        key = (r2, unclipped_y and 0 or cy)  # y clipping depends on cy only when the ball pokes out of the world in y
        if key not in tables:
            d = np.arange(-r, r + 1, dtype=np.int64)
            rem = r2 - d[:, None] ** 2 - d[None, :] ** 2
            half = np.floor(np.sqrt(np.maximum(rem, 0).astype(np.float64))).astype(np.int64)
            half = np.where((half + 1) ** 2 <= rem, half + 1, half)
            half = np.where(half ** 2 > rem, half - 1, half)
            lo = np.maximum(cy - half, 0) if not unclipped_y else cy - half
            hi = np.minimum(cy + half, res - 1) if not unclipped_y else cy + half
            col = np.where(rem >= 0, np.maximum(hi - lo + 1, 0), 0)
            ps = np.zeros((2 * r + 2, 2 * r + 2), dtype=np.int64)
            ps[1:, 1:] = col.cumsum(0).cumsum(1)
            tables[key] = ps
        ps = tables[key]
        x0, x1 = max(-r, -cx), min(r, res - 1 - cx)
        z0, z1 = max(-r, -cz), min(r, res - 1 - cz)
        if x0 > x1 or z0 > z1:
            continue
        a0, a1, b0, b1 = x0 + r, x1 + r + 1, z0 + r, z1 + r + 1
        total += int(ps[a1, b1] - ps[a0, b1] - ps[a1, b0] + ps[a0, b0])
    return total

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


def beam_params(P, fov=np.pi / 3):
    """Parameter block of the beam pre-pass for a frame P (src/rg/BeamPass.cpp:11-17,81-111): same camera, image of
    ceil(W/8) x ceil(H/8) texels, projection factor with the 8-pixel screen tolerance (float32 host maths)."""
    f = np.float32
    B = HdTraceParams.from_buffer_copy(P)
    B.width, B.height = (P.width + 7) // 8, (P.height + 7) // 8
    inv_2tan = f(1.0) / (f(2.0) * f(np.tan(f(0.5) * f(fov))))
    tol = f(1.0) / (f(P.height) / f(8.0))
    B.proj_factor = float(inv_2tan / tol)
    return B
