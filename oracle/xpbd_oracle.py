"""ctypes binding of oracle/xpbd_oracle.c -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference mount holds only /root/reference/README.md:1; the
oracle restates published XPBD (see xpbd_oracle.c).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libxpbd_oracle.so")
_SRC = [os.path.join(HERE, f) for f in ("xpbd_oracle.c", "xpbd_oracle_impl.h", "Makefile")]


class OrcParams(C.Structure):
    _fields_ = [
        ("dt", C.c_float), ("substeps", C.c_int32), ("iterations", C.c_int32),
        ("stiffness_distance", C.c_float), ("stiffness_volume", C.c_float),
        ("damping", C.c_float), ("friction", C.c_float), ("gravity", C.c_float * 3),
        ("ground_y", C.c_float), ("flags", C.c_int32),
    ]


def params(dt=1.0 / 60.0, substeps=10, iterations=10, stiffness_distance=math.inf, stiffness_volume=math.inf,
           damping=0.0, friction=0.0, gravity=(0.0, -9.81, 0.0), ground_y=0.0, flags=0) -> OrcParams:
    p = OrcParams(dt, substeps, iterations, stiffness_distance, stiffness_volume, damping, friction)
    p.gravity = (C.c_float * 3)(*gravity)
    p.ground_y, p.flags = ground_y, flags
    return p


def build(force=False):
    stale = force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in _SRC)
    if stale:
        r = subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), capture_output=True, text=True)
        if r.returncode != 0 and not os.path.exists(LIB):
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_build_edges.restype = C.c_int64
        assert _lib.orc_abi_sizeof_params() == C.sizeof(OrcParams)
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


# same layout as sb_collider (include/softbody_b200.h): kind 0 sphere / 1 capsule / 2 box
COLLIDER = np.dtype([("kind", np.int32), ("friction", np.float32), ("p", np.float32, 10)])


def colliders(items):
    """[(kind, friction, params...)] -> array of COLLIDER; kind as int or "sphere" / "capsule" / "box"."""
    names = {"sphere": 0, "capsule": 1, "box": 2}
    out = np.zeros(len(items), COLLIDER)
    for k, (kind, fr, *p) in enumerate(items):
        p = np.asarray(p, np.float32).ravel()
        out[k]["kind"] = names.get(kind, kind)
        out[k]["friction"] = fr
        out[k]["p"][:len(p)] = p
    return out


def build_edges(n_verts, tets):
    tets = np.ascontiguousarray(tets, np.int32)
    n = lib().orc_build_edges(C.c_int32(n_verts), C.c_int32(len(tets)), _p(tets), None)
    if n < 0:
        raise ValueError(f"orc_build_edges failed: {n}")
    edges = np.empty((n, 2), np.int32)
    lib().orc_build_edges(C.c_int32(n_verts), C.c_int32(len(tets)), _p(tets), _p(edges))
    return edges


def lumped_inv_mass(pos, tets, density=1000.0):
    pos = np.ascontiguousarray(pos, np.float32)
    tets = np.ascontiguousarray(tets, np.int32)
    w = np.empty(len(pos), np.float32)
    lib().orc_lumped_inv_mass(C.c_int32(len(pos)), _p(pos), C.c_int32(len(tets)), _p(tets), C.c_float(density), _p(w))
    return w


class Model:
    """Mesh + derived rest data in the oracle's own derivation (independent of the product)."""

    def __init__(self, pos, tets, inv_mass=None, density=1000.0, dtype=np.float32, roles=None):
        """roles: the tets again, each row a permutation of the same row of `tets` -- the vertex order the
        volume projection is evaluated in (the product picks it, sb_get_tet_roles).  Edges and lumped masses
        are derived from `tets`; rest volumes and the projection follow `roles`."""
        self.dtype = np.dtype(dtype)
        self.sfx = "_f32" if self.dtype == np.float32 else "_f64"
        self.pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        self.tets = np.ascontiguousarray(tets, np.int32).reshape(-1, 4)
        self.V, self.T = len(self.pos), len(self.tets)
        self.roles = self.tets if roles is None else np.ascontiguousarray(roles, np.int32).reshape(-1, 4)
        assert self.roles.shape == self.tets.shape and (np.sort(self.roles, 1) == np.sort(self.tets, 1)).all()
        self.edges = build_edges(self.V, self.tets)
        self.E = len(self.edges)
        self.inv_mass = (lumped_inv_mass(self.pos, self.tets, density) if inv_mass is None
                         else np.ascontiguousarray(inv_mass, np.float32))
        self.x4 = np.zeros((self.V, 4), self.dtype)
        self.x4[:, :3] = self.pos
        self.x4[:, 3] = self.inv_mass
        self.v4 = np.zeros((self.V, 4), self.dtype)
        self.rest_len = np.empty(self.E, self.dtype)
        self.rest_vol6 = np.empty(self.T, self.dtype)
        getattr(lib(), "orc_rest_values" + self.sfx)(_p(self.x4), C.c_int32(self.E), _p(self.edges), _p(self.rest_len),
                                                     C.c_int32(self.T), _p(self.roles), _p(self.rest_vol6))

    def natural_order(self):
        """All edges in canonical order, then all tets: plain sequential Gauss-Seidel."""
        return np.concatenate([np.arange(self.E, dtype=np.int32),
                               (np.arange(self.T, dtype=np.int64) | 0x80000000).astype(np.uint32).view(np.int32)])

    def simulate(self, prm: OrcParams, n_frames=1, order=None, batch_off=None, spheres=None, threads=1, colliders=None,
                 order_odd=None, batch_off_odd=None):
        """order / batch_off: the Gauss-Seidel order of the iterations 0, 2, 4 ... of every substep; order_odd /
        batch_off_odd: that of the iterations 1, 3, 5 ... (None: the same order again)."""
        order = self.natural_order() if order is None else np.ascontiguousarray(order, np.int32)
        nb = 0 if batch_off is None else len(batch_off) - 1
        boff = None if batch_off is None else np.ascontiguousarray(batch_off, np.int64)
        odd = None if order_odd is None else np.ascontiguousarray(order_odd, np.int32)
        nb_odd = 0 if batch_off_odd is None else len(batch_off_odd) - 1
        boff_odd = None if batch_off_odd is None else np.ascontiguousarray(batch_off_odd, np.int64)
        if colliders is not None:
            assert spheres is None
            sph = np.ascontiguousarray(colliders, COLLIDER)
        elif spheres is not None:
            s4 = np.ascontiguousarray(spheres, np.float32).reshape(-1, 4)
            sph = np.zeros(len(s4), COLLIDER)
            sph["p"][:, :4] = s4
        else:
            sph = None
        rc = getattr(lib(), "orc_simulate2" + self.sfx)(
            C.c_int32(self.V), _p(self.x4), _p(self.v4), C.c_int32(self.E), _p(self.edges), _p(self.rest_len),
            C.c_int32(self.T), _p(self.roles), _p(self.rest_vol6), C.byref(prm), C.c_int64(len(order)), _p(order),
            C.c_int32(nb), _p(boff), C.c_int64(0 if odd is None else len(odd)), _p(odd), C.c_int32(nb_odd), _p(boff_odd),
            C.c_int32(0 if sph is None else len(sph)), _p(sph), C.c_int32(n_frames), C.c_int32(threads))
        if rc != 0:
            raise ValueError(f"orc_simulate failed: {rc}")
        return self.x4, self.v4

    def normals(self, tris):
        tris = np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
        out = np.empty((self.V, 3), self.dtype)
        getattr(lib(), "orc_normals" + self.sfx)(C.c_int32(self.V), _p(self.x4), C.c_int32(len(tris)), _p(tris), _p(out))
        return out

    def skin(self, tet_of, bary4, tris=None):
        """Render vertices bound to the tets (caller's vertex order): positions (n,3) and, with tris, normals (n,3)."""
        tet_of = np.ascontiguousarray(tet_of, np.int32)
        bary4 = np.ascontiguousarray(bary4, np.float32).reshape(-1, 4)
        out4 = np.zeros((len(tet_of), 4), self.dtype)
        getattr(lib(), "orc_skin" + self.sfx)(_p(self.x4), _p(self.tets), C.c_int32(len(tet_of)), _p(tet_of), _p(bary4), _p(out4))
        if tris is None:
            return out4[:, :3].copy()
        tris = np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
        nrm = np.empty((len(tet_of), 3), self.dtype)
        getattr(lib(), "orc_normals" + self.sfx)(C.c_int32(len(tet_of)), _p(out4), C.c_int32(len(tris)), _p(tris), _p(nrm))
        return out4[:, :3].copy(), nrm

    def diagnostics(self, gravity=(0.0, -9.81, 0.0)):
        assert self.dtype == np.float32
        g = np.asarray(gravity, np.float32)
        out = np.zeros(16, np.float64)
        rl = np.ascontiguousarray(self.rest_len, np.float32)
        lib().orc_diagnostics(C.c_int32(self.V), _p(self.x4), _p(self.v4), C.c_int32(self.E), _p(self.edges), _p(rl),
                              C.c_int32(self.T), _p(self.tets), _p(g), _p(out))
        return out
