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

from ._abi import FLAG_NO_NORMALS
from .solver import SoftBody, ipc_export, ipc_open

OP_PREDICT, OP_FINISH, OP_PASS = 0, 2, 7


class DistBody:
    def __init__(self, pos, tets, surf_tris=None, *, device=0, **solver_kw):
        import torch.distributed as dist
        self.rank, self.n_ranks = dist.get_rank(), dist.get_world_size()
        solver_kw["flags"] = solver_kw.get("flags", 0) | FLAG_NO_NORMALS
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

    def gather_state(self):
        """(x4, v4) of the whole mesh on every rank: each rank contributes the vertices it owns."""
        import torch.distributed as dist
        x4, v4 = self.sb.get_state()
        parts = [None] * self.n_ranks
        dist.all_gather_object(parts, (np.nonzero(self.owned)[0], x4[self.owned], v4[self.owned]))
        X, U = np.zeros_like(x4), np.zeros_like(v4)
        for ids, xs, vs in parts:
            X[ids], U[ids] = xs, vs
        return X, U


class VirtualRanks:
    def __init__(self, pos, tets, surf_tris, n_ranks, stream_ptr, **solver_kw):
        solver_kw["flags"] = solver_kw.get("flags", 0) | FLAG_NO_NORMALS
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
        p = self.ranks[0].params
        n_pass = self.ranks[0].info()["n_tile_passes"]
        for sb in self.ranks:
            sb.prepare(dt)
        for _ in range(frames):
            for _ in range(p.substeps):
                for sb in self.ranks:
                    sb.enqueue(OP_PREDICT)
                for _ in range(p.iterations):
                    for k in range(n_pass):
                        for sb in self.ranks:
                            sb.enqueue(OP_PASS, k)
                for sb in self.ranks:
                    sb.enqueue(OP_FINISH)

    def gather_state(self):
        X = U = None
        for sb, own in zip(self.ranks, self.owned):
            x4, v4 = sb.get_state()
            if X is None:
                X, U = np.zeros_like(x4), np.zeros_like(v4)
            X[own], U[own] = x4[own], v4[own]
        return X, U
