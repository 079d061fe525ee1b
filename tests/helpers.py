"""Shared helpers for the tests: run the oracle and the product on the same inputs."""
import math

import numpy as np

from oracle import xpbd_oracle as orc
from softbodyunity_b200 import SoftBody


def oracle_params(sb: SoftBody, dt=None):
    p = sb.params
    return orc.params(dt=p.dt if dt is None else dt, substeps=p.substeps, iterations=p.iterations,
                      stiffness_distance=p.stiffness_distance, stiffness_volume=p.stiffness_volume,
                      damping=p.damping, friction=p.friction, gravity=tuple(p.gravity), ground_y=p.ground_y,
                      flags=p.flags & 1)


def oracle_for(sb: SoftBody, pos, tets, inv_mass=None, density=1000.0, dtype=np.float32):
    """Oracle model driven by the product's exported Gauss-Seidel order."""
    m = orc.Model(pos, tets, inv_mass=inv_mass, density=density, dtype=dtype, roles=sb.tet_roles())
    order, off = sb.schedule()
    return m, order, off


def rel_err(a, b):
    """max |a-b| / max(1e-30, bounding-box diagonal of b): the '1e-4 relative' of BASELINE.json:5."""
    a = np.asarray(a, np.float64)[:, :3]
    b = np.asarray(b, np.float64)[:, :3]
    scale = np.linalg.norm(b.max(0) - b.min(0))
    return float(np.abs(a - b).max() / max(scale, 1e-30))


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and bool((a.view(np.uint32) == b.view(np.uint32)).all())


def ulp_diff_count(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    return int((a != b).sum())


INF = math.inf
