"""Runs bench.py's NATIVE arm without a GPU, for the CPU test-suite: the solver is replaced by a stand-in that plans on the
host (a real sb_plan handle: info, frame program, layout) and returns made-up timings and state, CUDA calls are stubbed and
NCCL is swapped for gloo.  It checks the script's plumbing -- which keys the JSON line carries, that every rank count takes
the same path through the reductions -- and nothing else: the numbers it prints mean nothing.

    python tests/bench_mock.py [bench.py arguments]            (torchrun for N > 1)
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import softbodyunity_b200 as pkg  # noqa: E402
import softbodyunity_b200.dist as pdist  # noqa: E402

Real = pkg.SoftBody


class StandIn:
    def __init__(self, pos, tets, tris, device=0, **kw):
        self.h = Real(pos, tets, tris, host_only=True, **kw)
        self.n, self.own = len(pos), None
        self.n_surface = self.h.info()["n_surface_verts"]

    def info(self):
        return dict(self.h.info(), launches_per_frame=len(self.h.frame_program()))

    def step(self, dt=0.0, frames=1): pass
    def synchronize(self): pass
    def time_frames(self, n): return 6.0 * n
    def time_kernel(self, which, reps=20): return 0.015
    def get_state(self): return np.ones((self.n, 4), np.float32), np.zeros((self.n, 4), np.float32)
    def read_packed(self, out=None): return out
    def write_packed(self, buf): pass
    def read_surface(self, a=None, b=None): return a, b
    def frame_program(self): return self.h.frame_program()
    def schedule_kw(self): return self.h.schedule_kw()
    def tet_roles(self): return self.h.tet_roles()
    def halo_error(self): return False
    def dist_error(self): return False

    def packed_sizes(self):
        n = self.n if self.own is None else int(self.own.sum())
        return n, self.n_surface, 32 * n, 32 * n + 24 * self.n_surface

    def unpack_frame(self, buf):
        return np.ones((self.packed_sizes()[0], 4), np.float32), None, None, None


class StandInDist:
    def __init__(self, pos, tets, tris, device=0, **kw):
        self.sb = StandIn(pos, tets, tris, **kw)
        r, w = dist.get_rank(), dist.get_world_size()
        self.owned = self.sb.h.dist_layout(r, w)[0]
        self.sb.own = self.owned
        self.tiles = [int(self.sb.h.dist_layout(r, w, k)[1].sum()) for k in range(self.sb.info()["n_tile_passes"])]


def main(argv):
    pkg.SoftBody = StandIn
    pdist.DistBody = StandInDist
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda: None
    torch.Tensor.pin_memory = lambda self: self
    init = dist.init_process_group
    dist.init_process_group = lambda backend, device_id=None: init("gloo")
    tensor = torch.tensor
    torch.tensor = lambda *a, device=None, **k: tensor(*a, **k)
    import bench

    class NoSampler(bench.ClockSampler):
        def start(self):
            self.proc = None

    bench.ClockSampler = NoSampler
    bench.main(argv)


if __name__ == "__main__":
    main(sys.argv[1:])
