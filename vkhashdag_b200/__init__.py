"""vkhashdag_b200 — B200-native HashDAG traversal + edit engine (CUDA sm_100a behind a C ABI).

The product is libhashdag_b200.so (csrc/, include/hashdag_b200.h).  This package is the thin host-side mirror of
the reference's pool interface used by tests and benches; it holds no compute and has no CPU fallback.
"""
from .abi import (COLOR_NULL, NULL, HdConfig, HdDefaultConfig, HdEditDesc, HdTraceParams, HIT_DTYPE, aabb,  # noqa: F401
                  camera_params, default_config, edit_array, random_spheres, sphere, terrain)
from .api import (AABBEditor, DAGNodePool, HashDagError, HostBuffer, SphereEditor, TerrainEditor, kernel_launches, lib)  # noqa: F401
