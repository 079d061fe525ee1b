"""One large mesh over several GPUs: slab partition, ghost vertices, halo exchange.

BASELINE.json configs[4]: "single 8M-vertex synthetic tet mesh partitioned across 2/4/8 B200 with
NCCL-over-NVLink halo exchange per colour sweep".  Nothing of this exists upstream in the mount
(/root/reference/README.md:1 is the whole reference); the scheme:

* vertices are cut into `n_ranks` slabs of equal size along the longest axis;
* a constraint belongs to the LOWEST rank among its vertices; that rank keeps GHOST copies of the
  higher rank's vertices it touches (appended after its own vertices);
* per iteration: every rank sweeps its interior constraints (group 0) | exchange A: ghosts are
  refreshed from their owners | owners of cut constraints sweep them (group 1) | exchange B: the
  ghost values go back and overwrite the owners' copies.

The equivalent sequential Gauss-Seidel order is "all ranks' group-0 orders, then all ranks' group-1
orders" (`combined_order`), which the CPU oracle -- or a single GPU -- can replay.
"""
from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

from .solver import SoftBody, ipc_export, ipc_open, lumped_inv_mass

# halo list ids registered with sb_halo_set
H_GHOST, H_LOWER = 0, 1  # my ghost vertices (owned by rank+1) | my own vertices that are ghosts on rank-1


@dataclass
class RankMesh:
    rank: int
    n_ranks: int
    own: np.ndarray          # global ids of owned vertices (ascending)
    ghost: np.ndarray        # global ids of ghost vertices (ascending), all owned by rank + 1
    lower: np.ndarray        # global ids of OWN vertices that are ghosts on rank - 1 (ascending)
    pos: np.ndarray          # (n_own + n_ghost, 3) local rest pose, own first
    inv_mass: np.ndarray     # global lumped inverse masses restricted to the local vertices
    tets: np.ndarray         # (T_local, 4) local ids
    tet_gid: np.ndarray      # global tet id of each local tet
    tris: np.ndarray         # surface triangles whose three vertices are owned, local ids
    edges: np.ndarray        # (E_local, 2) local ids: the edges this rank sweeps (lowest endpoint rank == rank)
    local_of: Dict[int, int] = field(default_factory=dict, repr=False)

    @property
    def n_own(self):
        return len(self.own)

    @property
    def n_ghost(self):
        return len(self.ghost)

    @property
    def gids(self):
        return np.concatenate([self.own, self.ghost])


def slab_partition(pos, tets, tris, n_ranks, density=1000.0, inv_mass=None, only_rank=None) -> List[RankMesh]:
    """All ranks' meshes, or with only_rank=r a one-element list holding rank r's (what a process needs)."""
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    tets = np.ascontiguousarray(tets, np.int32).reshape(-1, 4)
    tris = np.zeros((0, 3), np.int32) if tris is None else np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
    V = len(pos)
    w_global = lumped_inv_mass(pos, tets, density) if inv_mass is None else np.ascontiguousarray(inv_mass, np.float32)
    axis = int(np.argmax(pos.max(0) - pos.min(0)))
    order = np.lexsort((np.arange(V), pos[:, axis]))
    rank_of = np.empty(V, np.int32)
    for r in range(n_ranks):
        lo, hi = (V * r) // n_ranks, (V * (r + 1)) // n_ranks
        rank_of[order[lo:hi]] = r
    tr = rank_of[tets]
    t_owner, t_top = tr.min(1), tr.max(1)
    if (t_top - t_owner).max(initial=0) > 1:
        raise ValueError("a tet spans more than two slabs: too many ranks for this mesh")
    out = []
    for r in (range(n_ranks) if only_rank is None else [only_rank]):
        own = np.flatnonzero(rank_of == r).astype(np.int64)
        sel = np.flatnonzero(t_owner == r)
        tv = tets[sel]
        ghost = np.unique(tv[rank_of[tv] > r]).astype(np.int64)
        below = tets[t_owner == r - 1] if r > 0 else np.zeros((0, 4), np.int32)
        lower = np.unique(below[rank_of[below] == r]).astype(np.int64)
        gids = np.concatenate([own, ghost])
        remap = -np.ones(V, np.int64)
        remap[gids] = np.arange(len(gids))
        ltets = remap[tv].astype(np.int32)
        tri_ok = (rank_of[tris] == r).all(1) if len(tris) else np.zeros(0, bool)
        # edges swept here: lowest endpoint rank == r.  They occur in my tets or in tets owned by rank r-1
        # (an edge between two of my vertices whose every tet also reaches down into slab r-1)
        cand = np.concatenate([tv, below]) if len(below) else tv
        pr = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]])
        e = np.sort(cand[:, pr].reshape(-1, 2).astype(np.int64), axis=1)
        e = e[rank_of[e].min(1) == r]
        e = np.unique(e[:, 0] * V + e[:, 1])
        ledges = np.stack([remap[e // V], remap[e % V]], 1).astype(np.int32)
        assert (ledges >= 0).all()
        out.append(RankMesh(r, n_ranks, own, ghost, lower, pos[gids], w_global[gids], ltets, sel.astype(np.int64),
                            remap[tris[tri_ok]].astype(np.int32), ledges))
    # exchange B writes a rank's lower-boundary vertices while it sweeps its own cut constraints on the upper side
    for m in out:
        upper = np.unique(m.tets[(m.tets >= m.n_own).any(1)])
        upper_own = m.own[upper[upper < m.n_own]]
        if len(np.intersect1d(upper_own, m.lower)):
            raise ValueError("slabs too thin: a vertex is on both interfaces of its slab")
    return out


class LocalComm:
    """All ranks in one process on one GPU ("virtual ranks"): the exchange is a device-to-device copy.
    Used by the tests; same orchestration code as the NCCL path."""

    def __init__(self, n_ranks):
        self.n, self.box = n_ranks, {}

    def exchange(self, phase, bodies):
        import torch
        assert all(b.stream is bodies[0].stream for b in bodies), "virtual ranks must share one stream"
        for b in bodies:  # everybody packs first; one shared stream, so stream order is the only sync needed
            b._pack(phase)
        with torch.cuda.stream(bodies[0].stream):
            for b in bodies:
                src = b.rank + 1 if phase == "A" else b.rank - 1
                if 0 <= src < self.n:
                    b._recv_buf(phase).copy_(bodies[src]._send_buf(phase))
        for b in bodies:
            b._unpack(phase)


class LocalPeerComm:
    """Virtual ranks in one process, exchanging through the SAME send/receive kernels as the
    multi-GPU path (direct stores into the neighbour's receive buffer + sequence flags).  All sends
    of an exchange are enqueued before any receive, so on the single shared stream no receive ever
    has to wait."""

    def __init__(self, bodies):
        self.n = len(bodies)
        for lo, hi in zip(bodies, bodies[1:]):
            base_ghost = lo.sb.halo_alloc(H_GHOST)   # lo receives its ghosts here (exchange A)
            base_lower = hi.sb.halo_alloc(H_LOWER)   # hi receives its lower-boundary vertices here (exchange B)
            hi.sb.halo_connect(H_LOWER, base_ghost)  # A: hi sends list LOWER into lo's GHOST buffer
            lo.sb.halo_connect(H_GHOST, base_lower)  # B: lo sends list GHOST into hi's LOWER buffer

    def exchange(self, phase, bodies):
        send, recv = (H_LOWER, H_GHOST) if phase == "A" else (H_GHOST, H_LOWER)
        for b in bodies:
            b.sb.enqueue(SoftBody.OP_HALO_SEND, send)
        for b in bodies:
            b.sb.enqueue(SoftBody.OP_HALO_RECV, recv)


def connect_peers(body, device):
    """Multi-process set-up of the peer-memory exchange: every rank allocates its receive buffers,
    the CUDA IPC handles travel through torch.distributed (plumbing only), and each rank opens its
    neighbours' buffers.  Afterwards body.sb.step() runs the exchanges inside the frame graph."""
    import torch.distributed as dist
    rank, n = body.rank, body.n_ranks
    mine = {}
    if body.mesh.n_ghost:
        mine["ghost"] = ipc_export(body.sb.halo_alloc(H_GHOST))
    if len(body.mesh.lower):
        mine["lower"] = ipc_export(body.sb.halo_alloc(H_LOWER))
    everyone = [None] * n
    dist.all_gather_object(everyone, mine)
    if rank > 0 and len(body.mesh.lower):       # exchange A: my LOWER list lands in (rank-1)'s GHOST buffer
        body.sb.halo_connect(H_LOWER, ipc_open(device, everyone[rank - 1]["ghost"]))
    if rank + 1 < n and body.mesh.n_ghost:      # exchange B: my GHOST list lands in (rank+1)'s LOWER buffer
        body.sb.halo_connect(H_GHOST, ipc_open(device, everyone[rank + 1]["lower"]))
    dist.barrier()


class TorchComm:
    """One rank per process: torch.distributed point-to-point (NCCL on GPUs)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.dist, self.torch = dist, torch
        self.n = dist.get_world_size()

    def exchange(self, phase, bodies):
        (b,) = bodies
        dist = self.dist
        b._pack(phase)
        dst = b.rank - 1 if phase == "A" else b.rank + 1
        src = b.rank + 1 if phase == "A" else b.rank - 1
        ops = []
        if 0 <= dst < self.n and b._send_buf(phase).numel():
            ops.append(dist.P2POp(dist.isend, b._send_buf(phase), dst))
        if 0 <= src < self.n and b._recv_buf(phase).numel():
            ops.append(dist.P2POp(dist.irecv, b._recv_buf(phase), src))
        if ops:
            with self.torch.cuda.stream(b.stream):  # NCCL orders itself after / before this stream's work
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
        b._unpack(phase)


class PartitionedBody:
    """One rank of a partitioned soft body: a SoftBody over own + ghost vertices plus its halo buffers."""

    def __init__(self, mesh: RankMesh, device=0, stream=None, **solver_kw):
        import torch
        self.mesh, self.rank, self.n_ranks = mesh, mesh.rank, mesh.n_ranks
        self.torch = torch
        # the solver kernels, the halo pack/unpack and the torch copies / NCCL calls must share ONE
        # explicit stream (the legacy default stream has handle 0, which the C ABI reads as "create one")
        self.stream = stream if stream is not None else torch.cuda.Stream(device)
        self.sb = SoftBody(mesh.pos, mesh.tets, mesh.tris if len(mesh.tris) else None, inv_mass=mesh.inv_mass,
                           edges=mesh.edges, n_ghost_verts=mesh.n_ghost, device=device, stream=self.stream.cuda_stream, **solver_kw)
        remap = {int(g): i for i, g in enumerate(mesh.gids)}
        self.sb.halo_set(H_GHOST, np.arange(mesh.n_own, mesh.n_own + mesh.n_ghost))
        self.sb.halo_set(H_LOWER, np.array([remap[int(g)] for g in mesh.lower], np.int32))
        dev = torch.device("cuda", device)
        self.buf_ghost = torch.zeros((mesh.n_ghost, 4), dtype=torch.float32, device=dev)
        self.buf_lower = torch.zeros((len(mesh.lower), 4), dtype=torch.float32, device=dev)

    # phase A: owners send their LOWER-boundary vertices down; the rank below receives its GHOSTS
    # phase B: ghost values go back up and overwrite the owner's LOWER-boundary vertices
    def _send_buf(self, phase):
        return self.buf_lower if phase == "A" else self.buf_ghost

    def _recv_buf(self, phase):
        return self.buf_ghost if phase == "A" else self.buf_lower

    def _pack(self, phase):
        lst, buf = (H_LOWER, self.buf_lower) if phase == "A" else (H_GHOST, self.buf_ghost)
        if buf.numel():
            self.sb.halo_pack(lst, buf.data_ptr())

    def _unpack(self, phase):
        lst, buf = (H_GHOST, self.buf_ghost) if phase == "A" else (H_LOWER, self.buf_lower)
        has_src = (self.rank + 1 < self.n_ranks) if phase == "A" else (self.rank > 0)
        if buf.numel() and has_src:
            self.sb.halo_unpack(lst, buf.data_ptr())


def step_partitioned(bodies: List[PartitionedBody], comm, dt: float = 0.0, frames: int = 1, prepared: bool = False):
    """Advance every given rank by `frames` frames (all ranks of a LocalComm, or the single local
    rank of a TorchComm)."""
    p = bodies[0].sb.params
    if not prepared:
        for b in bodies:
            b.sb.prepare(dt)
    S = SoftBody
    for _ in range(frames):
        for _ in range(p.substeps):
            for b in bodies:
                b.sb.enqueue(S.OP_PREDICT)
            for _ in range(p.iterations):
                for b in bodies:
                    b.sb.enqueue(S.OP_PROJECT, 0)
                comm.exchange("A", bodies)
                for b in bodies:
                    b.sb.enqueue(S.OP_PROJECT, 1)
                comm.exchange("B", bodies)
            for b in bodies:
                b.sb.enqueue(S.OP_FINISH)
        for b in bodies:
            b.sb.enqueue(S.OP_NORMALS)


class FrameRunner:
    """Steps the local rank(s) frame by frame; with use_graph the whole frame -- kernels, halo
    pack/unpack and the NCCL send/recv -- is captured once into a CUDA graph and replayed."""

    def __init__(self, bodies: List[PartitionedBody], comm, dt: float = 0.0, use_graph: bool = True):
        import torch
        self.bodies, self.comm, self.dt, self.torch = bodies, comm, dt, torch
        self.stream = bodies[0].stream
        self.graph = None
        self.graph_error = None
        for b in bodies:
            b.sb.prepare(dt)
        if use_graph:
            try:
                self._one_frame_eager()  # warm-up: NCCL connections, lazy allocations
                self.stream.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    self._one_frame_eager()
                self.graph = g
                self.warm_frames = 1
            except Exception as e:  # capture not supported for this transport: stay eager
                self.graph, self.graph_error = None, repr(e)[:200]
                self.warm_frames = 1

    def _one_frame_eager(self):
        step_partitioned(self.bodies, self.comm, self.dt, frames=1, prepared=True)

    def step(self, frames: int = 1):
        for _ in range(frames):
            if self.graph is not None:
                with self.torch.cuda.stream(self.stream):
                    self.graph.replay()
            else:
                self._one_frame_eager()


def combined_order(meshes: List[RankMesh], bodies_or_plans, global_edges):
    """The global Gauss-Seidel order equivalent to the partitioned run, in GLOBAL edge / tet ids:
    every rank's group-0 order, then every rank's group-1 order.  `bodies_or_plans[r]` is a SoftBody
    (device or host-only) planned on meshes[r]; `global_edges` is the canonical global edge list."""
    V = int(max(m.gids.max() for m in meshes)) + 1
    ge = np.asarray(global_edges, np.int64)
    ekey = ge[:, 0] * V + ge[:, 1]
    eorder = np.argsort(ekey)
    parts = {0: [], 1: []}
    batches = {0: [], 1: []}
    for m, sb in zip(meshes, bodies_or_plans):
        order, off = sb.schedule()
        ledges = sb.topology()[0]
        info = sb.info()
        n1 = info["constraints_cut"]
        n0 = len(order) - n1
        gids = m.gids
        is_tet = order < 0
        ids = order & 0x7fffffff
        out = np.empty(len(order), np.int64)
        out[is_tet] = m.tet_gid[ids[is_tet]] | 0x80000000
        le = np.sort(gids[ledges[ids[~is_tet]]], axis=1)
        k = le[:, 0] * V + le[:, 1]
        pos = np.searchsorted(ekey[eorder], k)
        assert (ekey[eorder][pos] == k).all(), "local edge not in the global edge list"
        out[~is_tet] = eorder[pos]
        out = out.astype(np.uint32).view(np.int32)
        cut_at = int(np.searchsorted(off, n0))
        assert off[cut_at] == n0, "group boundary must fall on a batch boundary"
        parts[0].append(out[:n0])
        parts[1].append(out[n0:])
        batches[0].append(np.diff(off[:cut_at + 1]))
        batches[1].append(np.diff(off[cut_at:]))
    order = np.concatenate(parts[0] + parts[1])
    sizes = np.concatenate(batches[0] + batches[1])
    return order, np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def combined_roles(meshes: List[RankMesh], bodies_or_plans, global_tets):
    """The vertex roles of every GLOBAL tet as the rank that owns it projects it (sb_get_tet_roles, mapped to
    global vertex ids): what a CPU replay of combined_order() needs to be bit-identical."""
    roles = np.ascontiguousarray(global_tets, np.int32).reshape(-1, 4).copy()
    for m, sb in zip(meshes, bodies_or_plans):
        roles[m.tet_gid] = m.gids[sb.tet_roles()]
    return roles


def gather_global(meshes: List[RankMesh], states):
    """Assemble a global (V, 4) array from per-rank (own + ghost) arrays: owners win."""
    V = int(max(m.gids.max() for m in meshes)) + 1
    out = np.zeros((V, 4), np.float32)
    for m, s in zip(meshes, states):
        out[m.own] = s[:m.n_own]
    return out
