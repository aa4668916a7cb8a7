"""The C-ABI library without a GPU: it loads, exports every symbol include/hashdag_b200.h declares, its host-only
helpers agree with the reference geometry, and compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import json
import os
import re

import pytest

import vkhashdag_b200 as v
from vkhashdag_b200 import abi, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "hashdag_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = v.lib()
    declared = header_symbols()
    assert len(declared) >= 35
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(api.SYMBOLS) == declared, "api.SYMBOLS out of sync with the header"
    assert b"sm_100a" in L.hd_version()


def test_struct_layouts_match_header():
    assert C.sizeof(abi.HdTraceParams) == 84        # tracer_pass::PC_Data, TracePass.cpp:10-18
    assert C.sizeof(abi.HdEditDesc) == 40
    assert C.sizeof(abi.HdConfig) == 4 * (3 + 22)
    assert abi.HIT_DTYPE.itemsize == 16
    assert C.sizeof(api.HdEditStats) == 64 and C.sizeof(api.HdTraceOutputs) == 32


def test_config_helpers_match_reference_geometry():
    L = v.lib()
    G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
    for c in G["configs"]:
        if c["default"] is None:
            cfg = abi.HdConfig()
            cfg.word_bits_per_page, cfg.page_bits_per_bucket = c["config"]["word_bits_per_page"], c["config"]["page_bits_per_bucket"]
            cfg.node_levels = len(c["config"]["bucket_bits"])
            for i, b in enumerate(c["config"]["bucket_bits"]):
                cfg.bucket_bits_each_level[i] = b
            assert L.hd_config_validate(C.byref(cfg)) == 0      # > 2^32-2 words (Config.hpp:48-56)
            continue
        dc = abi.HdDefaultConfig(**c["default"])
        cfg = abi.HdConfig()
        assert L.hd_config_from_default(C.byref(dc), C.byref(cfg)) == 0
        assert cfg.bucket_bits() == c["config"]["bucket_bits"]
        assert L.hd_config_validate(C.byref(cfg)) == 1
        assert L.hd_config_total_buckets(C.byref(cfg)) == c["geometry"]["total_buckets"]
        assert L.hd_config_total_words(C.byref(cfg)) == c["geometry"]["total_words"]
        for l, base in enumerate(c["geometry"]["level_bases"]):
            assert L.hd_config_level_base_bucket(C.byref(cfg), l) == base
    tiny = abi.HdConfig()
    tiny.word_bits_per_page, tiny.node_levels = 3, 2        # < kMinWordBitsPerPage
    assert L.hd_config_validate(C.byref(tiny)) == 0


def test_tile_shard_pixel_counts():
    L = v.lib()
    P = abi.HdTraceParams()
    P.width, P.height = 3840, 2160
    for world in (1, 2, 4, 8):
        total = 0
        for rank in range(world):
            sh = api.HdTileShard(64, 64, rank, world)
            total += L.hd_tile_shard_pixels(C.byref(P), C.byref(sh))
        assert total == 60 * 34 * 64 * 64
    assert L.hd_tile_shard_pixels(C.byref(P), C.byref(api.HdTileShard(64, 64, 3, 2))) == 0


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run; with one this test is a no-op."""
    L = v.lib()
    if L.hd_device_count() > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(v.HashDagError) as e:
        v.DAGNodePool(abi.default_config(level_count=6))
    assert e.value.status == api.HD_ERR_NO_DEVICE
    assert L.hd_edit_batch(None, 0, None, 0, None, None) == api.HD_ERR_INVALID


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vkhashdag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in text and "import oracle" not in text and "liboracle" not in text, f


def test_host_in_range_voxel_count_matches_oracle(oracle):
    """abi.spheres_in_range_voxels (the unit bench.py reports edited voxels/s in) == the oracle's lattice count,
    including balls clipped by the world in x, y and z."""
    from vkhashdag_b200 import abi
    for vl, kw in ((12, {}), (10, dict(rmin=3, rmax=200)), (8, dict(rmin=60, rmax=200, y_lo=0, y_hi=256))):
        sp = abi.random_spheres(60, vl, seed=7, **kw)
        assert abi.spheres_in_range_voxels(sp, vl) == sum(oracle.in_range_voxels(s, vl) for s in sp)
