"""torchrun entry: ONE block over all ranks through peer memory (sb_dist_*); rank 0 checks the gathered state and the
surface normals against the CPU oracle replaying the plan's order (bitwise) and prints the time per frame.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/mp_run_dist.py --dims 40 40 80 --frames 6
"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, nargs=3, default=[40, 40, 80])
ap.add_argument("--frames", type=int, default=6)
ap.add_argument("--substeps", type=int, default=10)
ap.add_argument("--iterations", type=int, default=10)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--tile-cap", type=int, default=0)
ap.add_argument("--slabs", action="store_true", help="ranks as slabs of the default box order instead of compact blocks")
ap.add_argument("--no-check", action="store_true")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from softbodyunity_b200 import meshgen
from softbodyunity_b200.dist import DistBody

pos, tets, tris = meshgen.block(*a.dims, spacing=0.02, origin=(0.0, 0.004, 0.0), seed=5)
kw = dict(substeps=a.substeps, iterations=a.iterations, flags=a.flags, tile_cap=a.tile_cap)
if a.slabs:
    kw["dist_ranks"] = 0
body = DistBody(pos, tets, tris, device=local, **kw)
print(f"[rank {rank}] owns {int(body.owned.sum())} of {len(pos)} vertices, tiles per pass {body.tiles}", flush=True)
torch.cuda.synchronize(); dist.barrier()
body.step(frames=a.frames)          # first call: includes the graph capture
body.sb.synchronize(); dist.barrier()
ms = body.sb.time_frames(a.frames)  # CUDA events on the solver's stream
t = torch.tensor([ms], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
X, U, N = body.gather_state(with_surface=True)
err = body.sb.dist_error()
if rank == 0:
    print(f"world {world}: V={len(pos)} {a.substeps}x{a.iterations} frames={a.frames} {float(t.item()) / a.frames:.3f} ms/frame (device-timed, max over ranks)  "
          f"launches/frame {body.sb.info()['launches_per_frame']}  peer wait timed out: {err}")
    if not a.no_check:
        from oracle import xpbd_oracle as orc
        p = body.sb.params
        ref = orc.Model(pos, tets, roles=body.sb.tet_roles())
        ref.simulate(orc.params(dt=p.dt, substeps=p.substeps, iterations=p.iterations), n_frames=2 * a.frames, threads=os.cpu_count(), **body.sb.schedule_kw())
        ids = body.sb.surface_vertices()
        # (flags & 8, SB_FLAG_NO_NORMALS: no normals to compare; the frame then ends with the closing handshake, k_dist_sync)
        same = (np.array_equal(X.view(np.uint32), ref.x4.view(np.uint32)) and np.array_equal(U[:, :3].view(np.uint32), ref.v4[:, :3].view(np.uint32))
                and ((a.flags & 8) != 0 or np.array_equal(N[ids].view(np.uint32), ref.normals(tris)[ids].view(np.uint32))))
        print("bit-identical to the CPU oracle (state and surface normals):", same, " min y", float(X[:, 1].min()))
        if not same:
            print("max |dx|", float(np.abs(X[:, :3] - ref.x4[:, :3]).max()))
dist.destroy_process_group()
