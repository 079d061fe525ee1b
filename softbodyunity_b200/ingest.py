"""Mesh ingest and on-disk formats: thin calls into the host-only part of the C ABI
(include/softbody_b200.h: sb_tetmesh_*, sb_skin_compute, sb_state_*), the steps before and
after the substep path.  Reference: NOT IN MOUNT
(/root/reference/README.md:1 is the whole reference); formats are the public TetGen / Gmsh
ones.  Nothing here needs a GPU.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import SbError, SbParams


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _ck(rc):
    if rc != 0:
        raise SbError(rc, _abi.load().sb_ingest_last_error().decode())


def _take(handle):
    """(pos (V,3) f32, tets (T,4) i32, tris (F,3) i32) out of an sb_tetmesh, which is freed."""
    lib = _abi.load()
    try:
        V, T, F = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _ck(lib.sb_tetmesh_sizes(handle, C.byref(V), C.byref(T), C.byref(F)))
        pos = np.empty((V.value, 3), np.float32)
        tets = np.empty((T.value, 4), np.int32)
        tris = np.empty((F.value, 3), np.int32)
        _ck(lib.sb_tetmesh_copy(handle, _ptr(pos), _ptr(tets), _ptr(tris) if F.value else None))
        return pos, tets, tris
    finally:
        lib.sb_tetmesh_free(handle)


def tetrahedralize_surface(surf_pos, surf_tris, spacing, snap=False):
    """Closed triangle surface -> lattice tet mesh of the enclosed volume (cells of size `spacing`).
    snap=True (or a distance) then pulls the staircase boundary onto the surface (default reach: one spacing)."""
    sp = np.ascontiguousarray(surf_pos, np.float32).reshape(-1, 3)
    st = np.ascontiguousarray(surf_tris, np.int32).reshape(-1, 3)
    lib = _abi.load()
    h = C.c_void_p()
    _ck(lib.sb_tetmesh_from_surface(_ptr(sp), len(sp), _ptr(st), len(st), float(spacing), C.byref(h)))
    if snap:
        reach = float(spacing) if snap is True else float(snap)
        rc = lib.sb_tetmesh_snap_to_surface(h, _ptr(sp), len(sp), _ptr(st), len(st), reach, None)
        if rc != 0:
            lib.sb_tetmesh_free(h)
            _ck(rc)
    return _take(h)


def from_arrays(pos, tets, tris=None):
    """Orientation fixed (all tets positive), boundary triangles extracted when none are given."""
    p = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(tets, np.int32).reshape(-1, 4)
    f = None if tris is None else np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
    h = C.c_void_p()
    _ck(_abi.load().sb_tetmesh_from_arrays(_ptr(p), len(p), _ptr(t), len(t), _ptr(f), 0 if f is None else len(f), C.byref(h)))
    return _take(h)


def load_mesh(path):
    """TetGen (<base>.node/.ele[/.face], or the bare base path) or Gmsh MSH ASCII (.msh; versions 2.x and 4.1)."""
    h = C.c_void_p()
    _ck(_abi.load().sb_tetmesh_load(str(path).encode(), C.byref(h)))
    return _take(h)


def save_mesh(path, pos, tets, tris=None):
    p = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(tets, np.int32).reshape(-1, 4)
    f = np.zeros((0, 3), np.int32) if tris is None else np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
    lib = _abi.load()
    h = C.c_void_p()
    # a placeholder triangle list would be replaced by the boundary; keep the caller's (possibly empty) one
    _ck(lib.sb_tetmesh_from_arrays(_ptr(p), len(p), _ptr(t), len(t), _ptr(f) if len(f) else None, len(f), C.byref(h)))
    try:
        _ck(lib.sb_tetmesh_save(h, str(path).encode()))
    finally:
        lib.sb_tetmesh_free(h)


def skin_binding(tet_pos, tets, points):
    """(tet_of (n,) i32, bary (n,4) f32): enclosing or nearest tet of every point and its barycentric weights."""
    p = np.ascontiguousarray(tet_pos, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(tets, np.int32).reshape(-1, 4)
    q = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    tet_of = np.empty(len(q), np.int32)
    bary = np.empty((len(q), 4), np.float32)
    _ck(_abi.load().sb_skin_compute(_ptr(p), len(p), _ptr(t), len(t), _ptr(q), len(q), _ptr(tet_of), _ptr(bary)))
    return tet_of, bary


def topology_hash(n_verts, tets):
    t = np.ascontiguousarray(tets, np.int32).reshape(-1, 4)
    return int(_abi.load().sb_topology_hash(int(n_verts), _ptr(t), len(t)))


def write_state(path, x4, v4, params: SbParams, frame=0, topo_hash=0):
    x = np.ascontiguousarray(x4, np.float32).reshape(-1, 4)
    v = np.ascontiguousarray(v4, np.float32).reshape(-1, 4)
    assert x.shape == v.shape
    _ck(_abi.load().sb_state_write(str(path).encode(), _ptr(x), _ptr(v), len(x), C.byref(params), int(frame), int(topo_hash)))


def read_state(path):
    """dict(x4, v4, params, frame, topo_hash) of a .sbs snapshot (checksum verified)."""
    lib = _abi.load()
    n, fr, th, p = C.c_uint32(), C.c_uint64(), C.c_uint64(), SbParams()
    _ck(lib.sb_state_read(str(path).encode(), None, None, 0, C.byref(n), C.byref(p), C.byref(fr), C.byref(th)))
    if os.path.getsize(path) < 88 + 32 * n.value:
        raise SbError(_abi.SB_E_ARG, "snapshot is truncated")
    x = np.empty((n.value, 4), np.float32)
    v = np.empty((n.value, 4), np.float32)
    _ck(lib.sb_state_read(str(path).encode(), _ptr(x), _ptr(v), n.value, None, None, None, None))
    return dict(x4=x, v4=v, params=p, frame=fr.value, topo_hash=th.value)
