"""Host-side mirror of the soft-body solver component.

Reference: the C# MonoBehaviour is NOT IN MOUNT (/root/reference/README.md:1 is the
whole reference).  BASELINE.json:5 names its surface: a Step call plus the parameters
stiffness, damping, substeps, iterations.  `SoftBody` keeps those names; every method
is a thin call into the C ABI (include/softbody_b200.h), exactly what a P/Invoke shim
does.  There is no CPU path here.
"""
import ctypes as C
import math

import numpy as np

from . import _abi
from ._abi import SbError, SbInfo, SbMeshDesc, SbParams

# sb_collider (include/softbody_b200.h): kind 0 sphere (centre, r) / 1 capsule (A, r, B) / 2 box (centre, half, quat)
COLLIDER = np.dtype([("kind", np.int32), ("friction", np.float32), ("p", np.float32, 10)])


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _buf_ptr(a):
    """Pointer of a numpy array or a (CPU) torch tensor, without copying."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())  # torch tensor


def default_params(**over) -> SbParams:
    p = SbParams()
    _abi.load().sb_default_params(C.byref(p))
    for k, v in over.items():
        if k == "gravity":
            p.gravity = (C.c_float * 3)(*v)
        else:
            setattr(p, k, v)
    return p


def lumped_inv_mass(pos, tets, density=1000.0):
    """Inverse lumped masses of a whole mesh, exactly as sb_create derives them (host only)."""
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    tets = np.ascontiguousarray(tets, np.int32).reshape(-1, 4)
    out = np.empty(len(pos), np.float32)
    rc = _abi.load().sb_lumped_inv_mass(_ptr(pos), len(pos), _ptr(tets), len(tets), density, _ptr(out))
    if rc != 0:
        raise SbError(rc, "sb_lumped_inv_mass: bad arguments")
    return out


def ipc_export(device_ptr: int) -> bytes:
    buf = (C.c_ubyte * 64)()
    rc = _abi.load().sb_ipc_export(C.c_void_p(device_ptr), buf)
    if rc != 0:
        raise SbError(rc, "cudaIpcGetMemHandle failed")
    return bytes(buf)


def ipc_open(device: int, handle: bytes) -> int:
    buf = (C.c_ubyte * 64).from_buffer_copy(handle)
    out = C.c_void_p()
    rc = _abi.load().sb_ipc_open(device, buf, C.byref(out))
    if rc != 0:
        raise SbError(rc, _abi.load().sb_last_error(None).decode())
    return out.value


class SoftBody:
    """One soft body (or a batch of independent bodies in one mesh) on one GPU.

    Parameters mirror the component's inspector fields: `stiffness` (edge springs, N/m;
    inf = rigid), `volume_stiffness`, `damping`, `substeps`, `iterations`.
    """

    def __init__(self, pos, tets, surf_tris=None, inv_mass=None, *, edges=None, density=1000.0, device=0,
                 stiffness=math.inf, volume_stiffness=math.inf, damping=0.0, friction=0.0,
                 substeps=10, iterations=10, dt=1.0 / 60.0, gravity=(0.0, -9.81, 0.0), ground_y=0.0,
                 flags=0, tile_cap=0, max_tile_passes=-1, block_threads=0, later_tile_cap=0,
                 host_threads=0, round_width=0, attach_edges=0, tilings=0, n_ghost_verts=0, dist_ranks=0, stream=None,
                 host_only=False):
        self._lib = _abi.load()
        self._h = C.c_void_p()
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        tets = np.ascontiguousarray(tets, dtype=np.int32).reshape(-1, 4)
        tris = None if surf_tris is None else np.ascontiguousarray(surf_tris, dtype=np.int32).reshape(-1, 3)
        w = None if inv_mass is None else np.ascontiguousarray(inv_mass, dtype=np.float32).reshape(-1)
        if w is not None and w.shape[0] != pos.shape[0]:
            raise ValueError("inv_mass must have one entry per vertex")
        self.n_verts, self.n_tets = pos.shape[0], tets.shape[0]
        self.n_tris = 0 if tris is None else tris.shape[0]
        d = SbMeshDesc()
        d.pos_xyz, d.tets, d.surf_tris, d.inv_mass = _ptr(pos), _ptr(tets), _ptr(tris), _ptr(w)
        d.stream = C.c_void_p(stream) if stream else None
        e = None if edges is None else np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 2)
        d.edges, d.n_edges = _ptr(e), 0 if e is None else len(e)
        d.n_verts, d.n_tets, d.n_tris = self.n_verts, self.n_tets, self.n_tris
        d.density, d.device = density, device
        d.tile_cap, d.max_tile_passes, d.block_threads = tile_cap, max_tile_passes, block_threads
        d.later_tile_cap, d.host_threads = later_tile_cap, host_threads
        d.round_width, d.attach_edges, d.tilings = round_width, attach_edges, tilings
        d.n_ghost_verts = n_ghost_verts
        d.dist_ranks = dist_ranks
        self._params = default_params(
            dt=dt, substeps=substeps, iterations=iterations, stiffness_distance=stiffness,
            stiffness_volume=volume_stiffness, damping=damping, friction=friction, gravity=gravity,
            ground_y=ground_y, flags=flags)
        fn = self._lib.sb_plan if host_only else self._lib.sb_create
        rc = fn(C.byref(d), C.byref(self._params), C.byref(self._h))
        if rc != 0:
            raise SbError(rc, self._lib.sb_last_error(None).decode())
        self.host_only = host_only
        i = self.info()
        self.n_edges, self.n_surface = i["n_edges"], i["n_surface_verts"]

    # -- from the formats either side of the path (ingest.py) ---------------------------------
    @classmethod
    def from_surface(cls, surf_pos, surf_tris, spacing, snap=True, bind=True, **kw):
        """Body from a closed surface mesh (a Unity `Mesh`): lattice tets of `spacing`, boundary snapped onto the
        surface, and -- with bind -- the surface itself bound as the render mesh (read_skinned)."""
        from . import ingest
        pos, tets, tris = ingest.tetrahedralize_surface(surf_pos, surf_tris, spacing, snap=snap)
        sb = cls(pos, tets, tris, **kw)
        if bind:
            sb.skin_bind(surf_pos, surf_tris)
        return sb

    @classmethod
    def from_file(cls, path, **kw):
        """Body from a TetGen (.node/.ele) or Gmsh (.msh) tet mesh."""
        from . import ingest
        pos, tets, tris = ingest.load_mesh(path)
        return cls(pos, tets, tris, **kw)

    # -- lifecycle ---------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.sb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise SbError(rc, self._lib.sb_last_error(self._h).decode())

    # -- parameters (the inspector fields) ----------------------------------------
    def _push(self):
        self._ck(self._lib.sb_set_params(self._h, C.byref(self._params)))

    @property
    def params(self) -> SbParams:
        return self._params

    def set_params(self, **kw):
        alias = {"stiffness": "stiffness_distance", "volume_stiffness": "stiffness_volume"}
        for k, v in kw.items():
            k = alias.get(k, k)
            if k == "gravity":
                self._params.gravity = (C.c_float * 3)(*v)
            else:
                if not hasattr(self._params, k):
                    raise AttributeError(k)
                setattr(self._params, k, v)
        self._push()

    stiffness = property(lambda s: s._params.stiffness_distance, lambda s, v: s.set_params(stiffness=v))
    volume_stiffness = property(lambda s: s._params.stiffness_volume, lambda s, v: s.set_params(volume_stiffness=v))
    damping = property(lambda s: s._params.damping, lambda s, v: s.set_params(damping=v))
    substeps = property(lambda s: s._params.substeps, lambda s, v: s.set_params(substeps=v))
    iterations = property(lambda s: s._params.iterations, lambda s, v: s.set_params(iterations=v))

    def set_colliders(self, spheres_xyzr):
        s = np.ascontiguousarray(spheres_xyzr, dtype=np.float32).reshape(-1, 4)
        self._ck(self._lib.sb_set_colliders(self._h, _ptr(s) if len(s) else None, len(s)))

    def set_colliders_ex(self, colliders):
        """Mixed sphere / capsule / box colliders: an array of _abi.COLLIDER records (sb_collider)."""
        c = np.ascontiguousarray(colliders, dtype=COLLIDER)
        self._ck(self._lib.sb_set_colliders_ex(self._h, _ptr(c) if len(c) else None, len(c)))

    # -- the hot path ---------------------------------------------------------------
    def step(self, dt: float = 0.0, frames: int = 1):
        """FixedUpdate: advance `frames` frames of dt seconds (asynchronous)."""
        for _ in range(frames):
            self._ck(self._lib.sb_step(self._h, dt))

    def synchronize(self):
        self._ck(self._lib.sb_synchronize(self._h))

    # -- phased stepping (one rank of a partitioned mesh; see partition.py) ---------------
    OP_PREDICT, OP_PROJECT, OP_FINISH, OP_NORMALS = 0, 1, 2, 3

    def set_stream(self, stream_ptr: int):
        self._ck(self._lib.sb_set_stream(self._h, C.c_void_p(stream_ptr)))

    def prepare(self, dt: float = 0.0):
        self._ck(self._lib.sb_prepare(self._h, dt))

    def enqueue(self, op: int, arg: int = -1):
        self._ck(self._lib.sb_enqueue(self._h, op, arg))

    def halo_set(self, list_id: int, vertex_ids):
        ids = np.ascontiguousarray(vertex_ids, np.int32)
        self._ck(self._lib.sb_halo_set(self._h, list_id, _ptr(ids) if len(ids) else None, len(ids)))

    def halo_pack(self, list_id: int, dst_device_ptr: int):
        self._ck(self._lib.sb_halo_pack(self._h, list_id, C.c_void_p(dst_device_ptr)))

    def halo_unpack(self, list_id: int, src_device_ptr: int):
        self._ck(self._lib.sb_halo_unpack(self._h, list_id, C.c_void_p(src_device_ptr)))

    # -- halo exchange over peer memory ------------------------------------------------
    OP_EXCHANGE, OP_HALO_SEND, OP_HALO_RECV, OP_PASS, OP_LAUNCH = 4, 5, 6, 7, 8

    def halo_alloc(self, list_id: int) -> int:
        """Allocates this rank's receive buffer for a registered list; returns its device address."""
        base, nbytes = C.c_void_p(), C.c_uint64()
        self._ck(self._lib.sb_halo_alloc(self._h, list_id, C.byref(base), C.byref(nbytes)))
        return base.value

    def halo_connect(self, list_id: int, peer_base: int):
        self._ck(self._lib.sb_halo_connect(self._h, list_id, C.c_void_p(peer_base)))

    # -- one mesh over several GPUs through peer memory (sb_dist_*) ----------------------------
    def dist_setup(self, rank: int, n_ranks: int):
        """-> (device address of this rank's position array, of its control block)."""
        xb, cb = C.c_void_p(), C.c_void_p()
        self._ck(self._lib.sb_dist_setup(self._h, rank, n_ranks, C.byref(xb), C.byref(cb)))
        return xb.value, cb.value

    def dist_connect(self, peer: int, peer_x: int, peer_ctl: int):
        self._ck(self._lib.sb_dist_connect(self._h, peer, C.c_void_p(peer_x), C.c_void_p(peer_ctl)))

    def dist_owned(self):
        """-> (bool mask over the caller's vertices this rank owns, tiles this rank runs per pass)."""
        m = np.zeros(self.n_verts, np.uint8)
        t = np.zeros(8, np.uint32)
        self._ck(self._lib.sb_dist_owned(self._h, _ptr(m), _ptr(t)))
        return m.astype(bool), t[:self.info()["n_tile_passes"]].tolist()

    def dist_layout(self, rank: int, n_ranks: int, pass_index: int = 0):
        """Host only: (bool mask of the vertices `rank` would own, bool mask of the tiles of `pass_index` it would run)."""
        own = np.zeros(self.n_verts, np.uint8)
        tiles = np.zeros(self.info()["tiles_in_pass"][pass_index], np.int32)
        self._ck(self._lib.sb_dist_layout(self._h, rank, n_ranks, _ptr(own), _ptr(tiles), pass_index))
        return own.astype(bool), tiles.astype(bool)

    def dist_launch_order(self, rank: int, n_ranks: int, pass_index: int = 0):
        """Host only: (tiles of `pass_index` in the order `rank` launches them, how many of them -- the first -- are zone
        tiles: they wait for the neighbours' epoch and count towards this rank's)."""
        tiles = np.zeros(self.info()["tiles_in_pass"][pass_index], np.int32)
        self._ck(self._lib.sb_dist_layout(self._h, rank, n_ranks, None, _ptr(tiles), pass_index))
        mine = np.nonzero(tiles)[0]
        order = mine[np.argsort(np.abs(tiles[mine]))]
        return order, int((tiles < 0).sum())

    def dist_verify(self, n_ranks: int):
        """Host only: symbolic replay of one frame's hand-overs over `n_ranks` ranks -> (vertices loaded from the wrong
        rank's array, vertices not home when a per-vertex kernel / the frame end needs them, hand-overs no epoch orders,
        vertex values that changed rank during the frame); (0, 0, 0, > 0) for a correct layout of a mesh that is really cut."""
        a, b, c, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self._lib.sb_dist_verify(self._h, n_ranks, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, d.value

    def dist_error(self) -> bool:
        out = C.c_int32()
        self._ck(self._lib.sb_dist_error(self._h, C.byref(out)))
        return bool(out.value)

    def halo_error(self) -> bool:
        out = C.c_int32()
        self._ck(self._lib.sb_halo_error(self._h, C.byref(out)))
        return bool(out.value)

    # -- write-back -------------------------------------------------------------------
    def positions(self, out=None):
        out = np.empty((self.n_verts, 3), np.float32) if out is None else out
        self._ck(self._lib.sb_read_positions(self._h, _buf_ptr(out), self.n_verts))
        return out

    def normals(self, out=None):
        out = np.empty((self.n_verts, 3), np.float32) if out is None else out
        self._ck(self._lib.sb_read_normals(self._h, _buf_ptr(out), self.n_verts))
        return out

    def surface_vertices(self):
        ids = np.empty(self.n_surface, np.int32)
        n = C.c_uint32()
        self._ck(self._lib.sb_surface_vertices(self._h, _ptr(ids), self.n_surface, C.byref(n)))
        return ids

    def read_surface(self, pos_out=None, nrm_out=None):
        pos_out = np.empty((self.n_surface, 3), np.float32) if pos_out is None else pos_out
        nrm_out = np.empty((self.n_surface, 3), np.float32) if nrm_out is None else nrm_out
        self._ck(self._lib.sb_read_surface(self._h, _buf_ptr(pos_out), _buf_ptr(nrm_out), self.n_surface))
        return pos_out, nrm_out

    def get_state(self, x4=None, v4=None):
        x4 = np.empty((self.n_verts, 4), np.float32) if x4 is None else x4
        v4 = np.empty((self.n_verts, 4), np.float32) if v4 is None else v4
        self._ck(self._lib.sb_get_state(self._h, _buf_ptr(x4), _buf_ptr(v4), self.n_verts))
        return x4, v4

    def set_state(self, x4=None, v4=None):
        if isinstance(x4, np.ndarray):
            x4 = np.ascontiguousarray(x4, np.float32)
        if isinstance(v4, np.ndarray):
            v4 = np.ascontiguousarray(v4, np.float32)
        self._ck(self._lib.sb_set_state(self._h, _buf_ptr(x4), _buf_ptr(v4), self.n_verts))

    # -- one frame in one buffer ------------------------------------------------------------
    def packed_sizes(self):
        """(n_verts, n_surface, bytes_in, bytes_out) of sb_write_packed / sb_read_packed (owned vertices on one rank of
        a distributed mesh)."""
        n, ns, bi, bo = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_uint64()
        self._ck(self._lib.sb_packed_sizes(self._h, C.byref(n), C.byref(ns), C.byref(bi), C.byref(bo)))
        return n.value, ns.value, bi.value, bo.value

    def read_packed(self, out=None):
        """LateUpdate in one copy: a uint8 buffer [x4 | v4 | surface xyz | surface normals]; see unpack_frame."""
        n, ns, bi, bo = self.packed_sizes()
        out = np.empty(bo, np.uint8) if out is None else out
        self._ck(self._lib.sb_read_packed(self._h, _buf_ptr(out), bo))
        return out

    def write_packed(self, buf):
        n, ns, bi, bo = self.packed_sizes()
        self._ck(self._lib.sb_write_packed(self._h, _buf_ptr(buf), bi))

    def unpack_frame(self, buf):
        """Views into a packed frame: x4 (n,4), v4 (n,4), surface positions (ns,3), surface normals (ns,3)."""
        n, ns, bi, bo = self.packed_sizes()
        f = np.asarray(buf).view(np.uint8).reshape(-1)[:bo].view(np.float32)
        return (f[:4 * n].reshape(n, 4), f[4 * n:8 * n].reshape(n, 4), f[8 * n:8 * n + 3 * ns].reshape(ns, 3),
                f[8 * n + 3 * ns:8 * n + 6 * ns].reshape(ns, 3))

    # -- render mesh bound to the tets -----------------------------------------------------
    def skin_bind(self, render_pos, render_tris=None):
        """Bind a render mesh to the tets at rest (host); works on a host_only handle too."""
        rp = np.ascontiguousarray(render_pos, np.float32).reshape(-1, 3)
        rt = None if render_tris is None else np.ascontiguousarray(render_tris, np.int32).reshape(-1, 3)
        self._ck(self._lib.sb_skin_bind(self._h, _ptr(rp), len(rp), _ptr(rt), 0 if rt is None else len(rt)))
        self.n_render = len(rp)

    def skin_binding(self):
        tet_of = np.empty(self.n_render, np.int32)
        bary = np.empty((self.n_render, 4), np.float32)
        self._ck(self._lib.sb_skin_get_binding(self._h, _ptr(tet_of), _ptr(bary), self.n_render))
        return tet_of, bary

    def read_skinned(self, normals=True):
        """LateUpdate for an embedded render mesh: its vertices (and normals) from the current tets."""
        pos = np.empty((self.n_render, 3), np.float32)
        nrm = np.empty((self.n_render, 3), np.float32) if normals else None
        self._ck(self._lib.sb_read_skinned(self._h, _ptr(pos), _ptr(nrm), self.n_render))
        return (pos, nrm) if normals else pos

    # -- snapshots ---------------------------------------------------------------------------
    def save_state(self, path):
        self._ck(self._lib.sb_save_state(self._h, str(path).encode()))

    def load_state(self, path, apply_params=False):
        self._ck(self._lib.sb_load_state(self._h, str(path).encode(), int(apply_params)))
        if apply_params:
            self._ck(self._lib.sb_get_params(self._h, C.byref(self._params)))

    @property
    def frames_done(self) -> int:
        out = C.c_uint64()
        self._ck(self._lib.sb_frames_done(self._h, C.byref(out)))
        return out.value

    def diagnostics(self):
        out = np.zeros(16, np.float64)
        rc = self._lib.sb_diagnostics(self._h, _ptr(out))
        if rc not in (0, _abi.SB_E_NAN):
            self._ck(rc)
        return dict(kinetic=out[0], potential=out[1], volume=out[2], centroid=out[3:6].copy(),
                    momentum=out[6:9].copy(), angular_momentum=out[9:12].copy(), max_strain=out[12],
                    rms_strain=out[13], nonfinite=int(out[14]), min_y=out[15], raw=out)

    # -- build products ----------------------------------------------------------------
    def info(self) -> dict:
        i = SbInfo()
        self._ck(self._lib.sb_get_info(self._h, C.byref(i)))
        return i.as_dict()

    def topology(self):
        i = self.info()
        edges = np.empty((i["n_edges"], 2), np.int32)
        rest_len = np.empty(i["n_edges"], np.float32)
        rest_vol6 = np.empty(i["n_tets"], np.float32)
        inv_mass = np.empty(i["n_verts"], np.float32)
        self._ck(self._lib.sb_get_topology(self._h, _ptr(edges), _ptr(rest_len), _ptr(rest_vol6), _ptr(inv_mass)))
        return edges, rest_len, rest_vol6, inv_mass

    def tet_roles(self):
        """Tets in the vertex-role order the volume projection uses (an even permutation of the caller's order);
        a CPU replay of the schedule must use these roles to be bit-identical (see sb_get_tet_roles)."""
        roles = np.empty((self.n_tets, 4), np.int32)
        self._ck(self._lib.sb_get_tet_roles(self._h, _ptr(roles), None, None))
        return roles

    def tet_mates(self):
        """(mate, lead): the tet each tet forms a bi-tet with (or -1), and 1 for a single tet or the first tet of a pair."""
        mate = np.empty(self.n_tets, np.int32)
        lead = np.empty(self.n_tets, np.int32)
        self._ck(self._lib.sb_get_tet_mates(self._h, _ptr(mate), _ptr(lead)))
        return mate, lead

    def attached_edges(self):
        e01 = np.empty(self.n_tets, np.int32)
        e23 = np.empty(self.n_tets, np.int32)
        self._ck(self._lib.sb_get_tet_roles(self._h, None, _ptr(e01), _ptr(e23)))
        return e01, e23

    def schedule(self, odd=False):
        """(order, batch_off): the Gauss-Seidel order of the iterations 0, 2, 4 ... of a substep (sb_get_schedule),
        or with odd=True of the iterations 1, 3, 5 ... (sb_get_schedule_odd: the tile passes backwards, unless the
        snake is off)."""
        fn = self._lib.sb_get_schedule_odd if odd else self._lib.sb_get_schedule
        n, nb = C.c_int64(), C.c_int32()
        self._ck(fn(self._h, C.byref(n), None, C.byref(nb), None))
        order = np.empty(n.value, np.int32)
        off = np.empty(nb.value + 1, np.int64)
        self._ck(fn(self._h, C.byref(n), _ptr(order), C.byref(nb), _ptr(off)))
        return order, off

    def schedule_kw(self):
        """The whole Gauss-Seidel order as the keyword arguments of the oracle's Model.simulate."""
        order, off = self.schedule()
        order_odd, off_odd = self.schedule(odd=True)
        return dict(order=order, batch_off=off, order_odd=order_odd, batch_off_odd=off_odd)

    LAUNCH_KINDS = ("predict", "finish", "pass", "global", "group", "exchange", "normals", "dag")

    def frame_program(self):
        """The launches of one frame at the current parameters: (n, 6) int32 rows of
        (kind, arg, segments, repetitions, predict-before, finish-after); see sb_frame_program."""
        n = C.c_int32()
        self._ck(self._lib.sb_frame_program(self._h, C.byref(n), None, 0))
        ops = np.zeros((n.value, 6), np.int32)
        self._ck(self._lib.sb_frame_program(self._h, C.byref(n), _ptr(ops), n.value))
        return ops

    def tiles(self, p: int):
        n = C.c_uint32()
        t = np.empty(self.n_verts, np.int32)
        self._ck(self._lib.sb_get_tiles(self._h, p, _ptr(t), C.byref(n)))
        return t, n.value

    # -- device timing -------------------------------------------------------------------
    def time_frames(self, n_frames: int, dt: float = 0.0) -> float:
        ms = C.c_float()
        self._ck(self._lib.sb_time_frames(self._h, n_frames, dt, C.byref(ms)))
        return ms.value

    def verify_streams(self) -> int:
        """Debug (host only): records of the device constraint streams that disagree with the exported schedule."""
        n, wf, wfi = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self._lib.sb_debug_verify_streams(self._h, C.byref(n), C.byref(wf), C.byref(wfi)))
        self.smem_model = (wf.value, wfi.value)  # load wavefronts per sweep: with bank conflicts, conflict-free
        return n.value

    def trace_pass(self, p: int):
        """Debug: (64, 80) uint64 globaltimer stamps of one run of tile pass p."""
        out = np.zeros(64 * 80 + 256 + 3 * 4096, np.uint64)
        self._ck(self._lib.sb_debug_trace_pass(self._h, p, _ptr(out), out.size))
        self.last_cta_trace = out[64 * 80 + 256:].reshape(4096, 3)  # start ns, end ns, SM id per CTA
        return out[:64 * 80].reshape(64, 80)

    def time_kernel(self, which: int, reps: int = 20) -> float:
        ms = C.c_float()
        self._ck(self._lib.sb_time_kernel(self._h, which, reps, C.byref(ms)))
        return ms.value
