"""torchrun entry: one rank per GPU steps its slab of one block; rank 0 checks the assembled state
against the CPU oracle replaying the combined order (bitwise) and prints timings.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_run_partitioned.py --dims 40 80 40 --frames 6
"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, nargs=3, default=[24, 48, 24])
ap.add_argument("--frames", type=int, default=6)
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--transport", default="peer", choices=["peer", "nccl"])
ap.add_argument("--no-check", action="store_true")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from softbodyunity_b200 import SoftBody, meshgen
from softbodyunity_b200.partition import FrameRunner, PartitionedBody, TorchComm, combined_order, combined_roles, connect_peers, gather_global, slab_partition

pos, tets, tris = meshgen.block(*a.dims, spacing=0.02, origin=(0.0, 0.004, 0.0), seed=5)
meshes = slab_partition(pos, tets, tris, world)
mine = meshes[rank]
body = PartitionedBody(mine, device=local)
print(f'[rank {rank}] body ready: own {mine.n_own} ghost {mine.n_ghost} lower {len(mine.lower)}', flush=True)
class _Direct:  # peer transport: sb_step runs kernels AND exchanges from its own CUDA graph
    graph, graph_error = True, None
    def step(self, frames):
        body.sb.step(frames=frames)
if a.transport == "peer":
    connect_peers(body, local)
    runner, pre = _Direct(), 0
else:
    runner = FrameRunner([body], TorchComm(), use_graph=not a.no_graph)
    pre = getattr(runner, "warm_frames", 0)
print(f'[rank {rank}] runner ready graph={runner.graph is not None} err={runner.graph_error}', flush=True)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
runner.step(a.frames)
body.stream.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
print(f'[rank {rank}] stepped', flush=True)
x4, v4 = body.sb.get_state()
if body.sb.halo_error():
    print(f'[rank {rank}] HALO TIMEOUT', flush=True)
xs = [None] * world if rank == 0 else None
dist.gather_object((x4[:mine.n_own].copy(), v4[:mine.n_own].copy()), xs, dst=0)
if rank == 0:
    print(f"world {world}: V={len(pos)} frames={a.frames} graph={'yes' if runner.graph is not None else 'no (' + str(runner.graph_error) + ')'} "
          f"{1e3 * dt / a.frames:.2f} ms/frame")
    if not a.no_check:
        from oracle import xpbd_oracle as orc
        V = len(pos)
        X = np.zeros((V, 4), np.float32); U = np.zeros((V, 4), np.float32)
        for m, (xo, vo) in zip(meshes, xs):
            X[m.own] = xo; U[m.own] = vo
        plans = [SoftBody(m.pos, m.tets, m.tris if len(m.tris) else None, inv_mass=m.inv_mass, edges=m.edges, n_ghost_verts=m.n_ghost, host_only=True) for m in meshes]
        ref = orc.Model(pos, tets, roles=combined_roles(meshes, plans, tets))
        order, off = combined_order(meshes, plans, ref.edges)
        p = body.sb.params
        ref.simulate(orc.params(dt=p.dt, substeps=p.substeps, iterations=p.iterations), n_frames=a.frames + pre, order=order, batch_off=off, threads=os.cpu_count())
        same = np.array_equal(X.view(np.uint32), ref.x4.view(np.uint32)) and np.array_equal(U[:, :3].view(np.uint32), ref.v4[:, :3].view(np.uint32))
        print("bit-identical to the CPU oracle (combined order):", same, " min y", float(X[:, 1].min()))
        if not same:
            print("max |dx|", float(np.abs(X[:, :3] - ref.x4[:, :3]).max()))
            dist.destroy_process_group(); sys.exit(1)
dist.destroy_process_group()
