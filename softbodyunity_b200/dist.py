"""ONE mesh over several GPUs through peer memory (sb_dist_*): no ghosts, no exchange
kernels.  Every rank plans the SAME whole mesh, owns a slab of the device numbering and runs its share of the
tiles of every pass; tiles read and write their vertex runs in the owner's memory over NVLink, and the kernels
of the ranks order themselves through epoch words.  torch.distributed carries only the CUDA-IPC handles.

  DistBody      one rank per process (torchrun)
  VirtualRanks  all ranks in one process on one GPU and one stream (tests): the same kernels and the same
                peer-pointer / epoch machinery, with the launches of the ranks interleaved kernel by kernel
"""
from typing import List

import numpy as np

from .solver import SoftBody, ipc_export, ipc_open

OP_LAUNCH = 8


class DistBody:
    def __init__(self, pos, tets, surf_tris=None, *, device=0, **solver_kw):
        import torch.distributed as dist
        self.rank, self.n_ranks = dist.get_rank(), dist.get_world_size()
        solver_kw.setdefault("dist_ranks", self.n_ranks)  # boxes numbered block by block: compact ranks
        self.sb = SoftBody(pos, tets, surf_tris, device=device, **solver_kw)
        xb, cb = self.sb.dist_setup(self.rank, self.n_ranks)
        everyone = [None] * self.n_ranks
        dist.all_gather_object(everyone, (ipc_export(xb), ipc_export(cb)))
        for p, (hx, hc) in enumerate(everyone):
            if p != self.rank:
                self.sb.dist_connect(p, ipc_open(device, hx), ipc_open(device, hc))
        self.owned, self.tiles = self.sb.dist_owned()
        dist.barrier()

    def step(self, dt: float = 0.0, frames: int = 1):
        self.sb.step(dt, frames=frames)

    def read_frame(self):
        """This rank's share of a frame in one copy (sb_read_packed): (ids, x4, v4, surface ids, positions, normals)
        of the vertices / surface vertices it owns, ascending vertex id."""
        x4, v4, sp, sn = self.sb.unpack_frame(self.sb.read_packed())
        surf = self.sb.surface_vertices()
        return np.nonzero(self.owned)[0], x4, v4, surf[self.owned[surf]], sp, sn

    def gather_state(self, with_surface=False):
        """(x4, v4) of the whole mesh on every rank: each rank contributes the vertices it owns (and, with_surface,
        the normals of the whole surface, (n_verts, 3), zero off the surface)."""
        import torch.distributed as dist
        parts = [None] * self.n_ranks
        dist.all_gather_object(parts, self.read_frame())
        n = len(self.owned)
        X, U, N = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 3), np.float32)
        for ids, xs, vs, sids, sp, sn in parts:
            X[ids], U[ids], N[sids] = xs, vs, sn
        if self.sb.dist_error():
            raise RuntimeError("a wait for a peer GPU timed out: the state is invalid")
        return (X, U, N) if with_surface else (X, U)


class VirtualRanks:
    def __init__(self, pos, tets, surf_tris, n_ranks, stream_ptr, **solver_kw):
        solver_kw.setdefault("dist_ranks", n_ranks)
        self.ranks: List[SoftBody] = [SoftBody(pos, tets, surf_tris, stream=stream_ptr, **solver_kw) for _ in range(n_ranks)]
        addr = [sb.dist_setup(r, n_ranks) for r, sb in enumerate(self.ranks)]
        for r, sb in enumerate(self.ranks):
            for p in range(n_ranks):
                if p != r:
                    sb.dist_connect(p, *addr[p])
        self.owned = [sb.dist_owned()[0] for sb in self.ranks]
        self.tiles = [sb.dist_owned()[1] for sb in self.ranks]

    def step(self, dt: float = 0.0, frames: int = 1):
        """One stream: the ranks' launches are interleaved kernel by kernel (a rank's kernel waits for its
        peers' previous kernel, which must therefore already be in the stream ahead of it)."""
        n_launches = len(self.ranks[0].frame_program())  # the same program on every rank
        for sb in self.ranks:
            sb.prepare(dt)
        for _ in range(frames):
            for i in range(n_launches):
                for sb in self.ranks:
                    sb.enqueue(OP_LAUNCH, i)

    def gather_state(self, with_surface=False):
        n = len(self.owned[0])
        X, U, N = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 3), np.float32)
        for sb, own in zip(self.ranks, self.owned):
            x4, v4, sp, sn = sb.unpack_frame(sb.read_packed())  # the vertices this rank owns, one copy
            surf = sb.surface_vertices()
            X[own], U[own], N[surf[own[surf]]] = x4, v4, sn
            if sb.dist_error():
                raise RuntimeError("a wait for a peer timed out: the state is invalid")
        return (X, U, N) if with_surface else (X, U)
