"""softbodyunity_b200 -- B200-native soft-body substep solver behind a C ABI.

Package contents are only what the hot path needs (tier framing): `csrc/` (CUDA
kernels + host planner + C ABI), the ctypes binding, the component mirror and the
synthetic mesh generators used by tests and benches.
"""
from . import meshgen  # noqa: F401
from ._abi import (FLAG_DAG, FLAG_FAST_MATH, FLAG_NO_FUSE, FLAG_NO_GRAPH, FLAG_NO_GROUND, FLAG_NO_NORMALS,  # noqa: F401
                   FLAG_NO_PDL, FLAG_NO_SNAKE, SbError, SbInfo, SbMeshDesc, SbParams, lib_path, load)
from .solver import SoftBody, default_params, lumped_inv_mass  # noqa: F401

__all__ = ["SoftBody", "default_params", "meshgen", "load", "lib_path", "SbError", "SbParams",
           "SbMeshDesc", "SbInfo", "FLAG_FAST_MATH", "FLAG_NO_GRAPH", "FLAG_NO_GROUND", "FLAG_NO_NORMALS"]
