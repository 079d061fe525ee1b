"""ctypes binding of include/softbody_b200.h (the same calls a C# P/Invoke shim makes).

The product path has no CPU fallback: if libsoftbody_b200.so is missing or a device
call fails, this module raises.
"""
import ctypes as C
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))

SB_OK, SB_E_ARG, SB_E_CUDA, SB_E_NCCL, SB_E_NAN, SB_E_STATE, SB_E_NOMEM = 0, -1, -2, -3, -4, -5, -6
FLAG_NO_GROUND, FLAG_FAST_MATH, FLAG_NO_GRAPH, FLAG_NO_NORMALS, FLAG_NO_PDL, FLAG_DAG = 1, 2, 4, 8, 16, 32
FLAG_NO_SNAKE, FLAG_NO_FUSE = 64, 128
ABI_VERSION = 3
COLLIDER_SPHERE, COLLIDER_CAPSULE, COLLIDER_BOX, MAX_COLLIDERS = 0, 1, 2, 16

EXPORTS = [
    "sb_abi_check", "sb_default_params", "sb_create", "sb_plan", "sb_destroy", "sb_set_params",
    "sb_get_params", "sb_set_colliders", "sb_set_colliders_ex", "sb_step", "sb_synchronize", "sb_read_positions",
    "sb_read_normals", "sb_surface_vertices", "sb_read_surface", "sb_get_state", "sb_set_state",
    "sb_packed_sizes", "sb_read_packed", "sb_write_packed",
    "sb_diagnostics", "sb_get_info", "sb_get_topology", "sb_get_tet_roles", "sb_get_tet_mates", "sb_get_schedule", "sb_get_schedule_odd", "sb_frame_program", "sb_get_tiles",
    "sb_time_frames", "sb_time_kernel", "sb_debug_trace_pass", "sb_debug_verify_streams", "sb_last_error",
    "sb_set_stream", "sb_prepare", "sb_enqueue", "sb_halo_set", "sb_halo_pack", "sb_halo_unpack", "sb_lumped_inv_mass",
    "sb_halo_alloc", "sb_halo_connect", "sb_halo_error", "sb_ipc_export", "sb_ipc_open",
    "sb_dist_setup", "sb_dist_connect", "sb_dist_owned", "sb_dist_error", "sb_dist_layout", "sb_dist_verify",
    "sb_skin_bind", "sb_skin_get_binding", "sb_read_skinned", "sb_skin_compute",
    "sb_save_state", "sb_load_state", "sb_frames_done", "sb_state_write", "sb_state_read", "sb_topology_hash",
    "sb_tetmesh_from_surface", "sb_tetmesh_snap_to_surface", "sb_tetmesh_from_arrays", "sb_tetmesh_load", "sb_tetmesh_save", "sb_tetmesh_sizes",
    "sb_tetmesh_copy", "sb_tetmesh_desc", "sb_tetmesh_free", "sb_ingest_last_error",
]


class SbParams(C.Structure):
    _fields_ = [
        ("dt", C.c_float), ("substeps", C.c_int32), ("iterations", C.c_int32),
        ("stiffness_distance", C.c_float), ("stiffness_volume", C.c_float),
        ("damping", C.c_float), ("friction", C.c_float), ("gravity", C.c_float * 3),
        ("ground_y", C.c_float), ("flags", C.c_int32),
    ]


class SbMeshDesc(C.Structure):
    _fields_ = [
        ("pos_xyz", C.c_void_p), ("tets", C.c_void_p), ("surf_tris", C.c_void_p),
        ("inv_mass", C.c_void_p), ("stream", C.c_void_p), ("edges", C.c_void_p),
        ("n_verts", C.c_uint32), ("n_tets", C.c_uint32), ("n_tris", C.c_uint32),
        ("density", C.c_float), ("device", C.c_int32), ("tile_cap", C.c_int32),
        ("max_tile_passes", C.c_int32), ("block_threads", C.c_int32),
        ("later_tile_cap", C.c_int32), ("host_threads", C.c_int32), ("round_width", C.c_int32),
        ("attach_edges", C.c_int32), ("tilings", C.c_int32), ("n_ghost_verts", C.c_int32), ("n_edges", C.c_uint32), ("dist_ranks", C.c_int32),
    ]


class SbInfo(C.Structure):
    _fields_ = [
        ("n_verts", C.c_uint32), ("n_edges", C.c_uint32), ("n_tets", C.c_uint32), ("n_tris", C.c_uint32),
        ("n_surface_verts", C.c_uint32), ("n_tile_passes", C.c_uint32), ("n_tilings", C.c_uint32), ("n_ghost_verts", C.c_uint32), ("first_cut_pass", C.c_uint32), ("constraints_cut", C.c_uint64),
        ("n_global_batches", C.c_uint32), ("n_batches", C.c_uint32),
        ("tiles_in_pass", C.c_uint32 * 8), ("max_colours_in_pass", C.c_uint32 * 8),
        ("constraints_in_pass", C.c_uint64 * 8), ("edges_in_pass", C.c_uint64 * 8), ("runs_in_pass", C.c_uint64 * 8),
        ("constraints_global", C.c_uint64),
        ("tile_cap", C.c_uint32), ("block_threads", C.c_uint32), ("smem_bytes", C.c_uint32),
        ("round_width", C.c_uint32), ("reserved0", C.c_uint32), ("edges_attached", C.c_uint64), ("rounds_in_pass", C.c_uint64 * 8),
        ("launches_per_frame", C.c_uint32), ("device_bytes", C.c_uint64), ("build_seconds", C.c_double),
    ]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            val = getattr(self, name)
            d[name] = list(val) if hasattr(val, "__len__") else val
        return d


_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Loads (building first if the sources are newer) libsoftbody_b200.so."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if _build.stale():
        try:
            _build.build()
        except Exception as e:  # no nvcc on this box: a prebuilt library is still fine
            if not os.path.exists(path):
                raise RuntimeError(f"libsoftbody_b200.so is missing and could not be built: {e}") from e
    lib = C.CDLL(path)
    vp, i32, u32, f32 = C.c_void_p, C.c_int32, C.c_uint32, C.c_float
    P = C.POINTER
    sig = {
        "sb_abi_check": (C.c_int, [P(u32), P(u32), P(u32), P(u32)]),
        "sb_default_params": (None, [P(SbParams)]),
        "sb_create": (C.c_int, [P(SbMeshDesc), P(SbParams), P(vp)]),
        "sb_plan": (C.c_int, [P(SbMeshDesc), P(SbParams), P(vp)]),
        "sb_destroy": (C.c_int, [vp]),
        "sb_set_params": (C.c_int, [vp, P(SbParams)]),
        "sb_get_params": (C.c_int, [vp, P(SbParams)]),
        "sb_set_colliders": (C.c_int, [vp, vp, u32]),
        "sb_set_colliders_ex": (C.c_int, [vp, vp, u32]),
        "sb_step": (C.c_int, [vp, f32]),
        "sb_synchronize": (C.c_int, [vp]),
        "sb_read_positions": (C.c_int, [vp, vp, u32]),
        "sb_read_normals": (C.c_int, [vp, vp, u32]),
        "sb_surface_vertices": (C.c_int, [vp, vp, u32, P(u32)]),
        "sb_read_surface": (C.c_int, [vp, vp, vp, u32]),
        "sb_get_state": (C.c_int, [vp, vp, vp, u32]),
        "sb_set_state": (C.c_int, [vp, vp, vp, u32]),
        "sb_packed_sizes": (C.c_int, [vp, P(u32), P(u32), P(C.c_uint64), P(C.c_uint64)]),
        "sb_read_packed": (C.c_int, [vp, vp, C.c_uint64]),
        "sb_write_packed": (C.c_int, [vp, vp, C.c_uint64]),
        "sb_diagnostics": (C.c_int, [vp, vp]),
        "sb_get_info": (C.c_int, [vp, P(SbInfo)]),
        "sb_get_topology": (C.c_int, [vp, vp, vp, vp, vp]),
        "sb_get_schedule": (C.c_int, [vp, P(C.c_int64), vp, P(i32), vp]),
        "sb_get_schedule_odd": (C.c_int, [vp, P(C.c_int64), vp, P(i32), vp]),
        "sb_frame_program": (C.c_int, [vp, P(i32), vp, u32]),
        "sb_get_tiles": (C.c_int, [vp, u32, vp, P(u32)]),
        "sb_time_frames": (C.c_int, [vp, i32, f32, P(f32)]),
        "sb_time_kernel": (C.c_int, [vp, i32, i32, P(f32)]),
        "sb_debug_trace_pass": (C.c_int, [vp, u32, vp, u32]),
        "sb_get_tet_roles": (C.c_int, [vp, vp, vp, vp]),
        "sb_get_tet_mates": (C.c_int, [vp, vp, vp]),
        "sb_debug_verify_streams": (C.c_int, [vp, vp, vp, vp]),
        "sb_last_error": (C.c_char_p, [vp]),
        "sb_set_stream": (C.c_int, [vp, vp]),
        "sb_prepare": (C.c_int, [vp, f32]),
        "sb_enqueue": (C.c_int, [vp, i32, i32]),
        "sb_halo_set": (C.c_int, [vp, i32, vp, u32]),
        "sb_halo_pack": (C.c_int, [vp, i32, vp]),
        "sb_halo_unpack": (C.c_int, [vp, i32, vp]),
        "sb_lumped_inv_mass": (C.c_int, [vp, u32, vp, u32, f32, vp]),
        "sb_halo_alloc": (C.c_int, [vp, i32, P(vp), P(C.c_uint64)]),
        "sb_halo_connect": (C.c_int, [vp, i32, vp]),
        "sb_halo_error": (C.c_int, [vp, P(i32)]),
        "sb_ipc_export": (C.c_int, [vp, vp]),
        "sb_ipc_open": (C.c_int, [i32, vp, P(vp)]),
        "sb_dist_setup": (C.c_int, [vp, i32, i32, P(vp), P(vp)]),
        "sb_dist_connect": (C.c_int, [vp, i32, vp, vp]),
        "sb_dist_owned": (C.c_int, [vp, vp, vp]),
        "sb_dist_error": (C.c_int, [vp, P(i32)]),
        "sb_dist_layout": (C.c_int, [vp, i32, i32, vp, vp, u32]),
        "sb_dist_verify": (C.c_int, [vp, i32, vp, vp, vp, vp]),
        "sb_skin_bind": (C.c_int, [vp, vp, u32, vp, u32]),
        "sb_skin_get_binding": (C.c_int, [vp, vp, vp, u32]),
        "sb_read_skinned": (C.c_int, [vp, vp, vp, u32]),
        "sb_skin_compute": (C.c_int, [vp, u32, vp, u32, vp, u32, vp, vp]),
        "sb_save_state": (C.c_int, [vp, C.c_char_p]),
        "sb_load_state": (C.c_int, [vp, C.c_char_p, i32]),
        "sb_frames_done": (C.c_int, [vp, P(C.c_uint64)]),
        "sb_state_write": (C.c_int, [C.c_char_p, vp, vp, u32, P(SbParams), C.c_uint64, C.c_uint64]),
        "sb_state_read": (C.c_int, [C.c_char_p, vp, vp, u32, P(u32), P(SbParams), P(C.c_uint64), P(C.c_uint64)]),
        "sb_topology_hash": (C.c_uint64, [u32, vp, u32]),
        "sb_tetmesh_from_surface": (C.c_int, [vp, u32, vp, u32, f32, P(vp)]),
        "sb_tetmesh_from_arrays": (C.c_int, [vp, u32, vp, u32, vp, u32, P(vp)]),
        "sb_tetmesh_snap_to_surface": (C.c_int, [vp, vp, u32, vp, u32, f32, P(u32)]),
        "sb_tetmesh_load": (C.c_int, [C.c_char_p, P(vp)]),
        "sb_tetmesh_save": (C.c_int, [vp, C.c_char_p]),
        "sb_tetmesh_sizes": (C.c_int, [vp, P(u32), P(u32), P(u32)]),
        "sb_tetmesh_copy": (C.c_int, [vp, vp, vp, vp]),
        "sb_tetmesh_desc": (C.c_int, [vp, P(SbMeshDesc)]),
        "sb_tetmesh_free": (C.c_int, [vp]),
        "sb_ingest_last_error": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    ver, sp, sd, si = u32(), u32(), u32(), u32()
    lib.sb_abi_check(C.byref(ver), C.byref(sp), C.byref(sd), C.byref(si))
    if (ver.value, sp.value, sd.value, si.value) != (ABI_VERSION, C.sizeof(SbParams), C.sizeof(SbMeshDesc), C.sizeof(SbInfo)):
        raise RuntimeError(
            f"ABI mismatch: library v{ver.value} params={sp.value} desc={sd.value} info={si.value}; "
            f"binding v{ABI_VERSION} params={C.sizeof(SbParams)} desc={C.sizeof(SbMeshDesc)} info={C.sizeof(SbInfo)}")
    _lib = lib
    return lib


class SbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"softbody_b200 error {code}: {msg}")
        self.code = code
