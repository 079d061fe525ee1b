"""Replays a run between two state snapshots (.sbs) on the GPU and checks the result bit for bit.

  python tools/replay.py --mesh body.msh --from frame100.sbs --to frame160.sbs [--device 0]

The mesh file (TetGen / Gmsh) must be the one the snapshots were taken from (vertex count and topology hash are
checked by sb_load_state); the parameters stored in the first snapshot are applied.  A substep solver's state is
exactly (x, v), so the replay of an unmodified library is bit-identical; a difference means the kernels, the
planner's schedule or the parameters changed -- snapshots are the regression vectors the reference does not ship.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from softbodyunity_b200 import SoftBody, ingest  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", required=True)
    ap.add_argument("--from", dest="src", required=True)
    ap.add_argument("--to", dest="dst", required=True)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    pos, tets, tris = ingest.load_mesh(a.mesh)
    first, last = ingest.read_state(a.src), ingest.read_state(a.dst)
    frames = last["frame"] - first["frame"]
    if frames < 0:
        raise SystemExit("--to is older than --from")
    sb = SoftBody(pos, tets, tris, device=a.device)
    sb.load_state(a.src, apply_params=True)
    sb.step(frames=frames)
    x4, v4 = sb.get_state()
    dx = int((x4.view(np.uint32) != last["x4"].view(np.uint32)).sum())
    dv = int((v4.view(np.uint32) != last["v4"].view(np.uint32)).sum())
    err = float(np.abs(x4[:, :3].astype(np.float64) - last["x4"][:, :3]).max())
    print(f"replayed {frames} frames from frame {first['frame']}: {dx} position words and {dv} velocity words differ "
          f"(max |dx| = {err:.3e})")
    raise SystemExit(0 if dx == 0 and dv == 0 else 1)


if __name__ == "__main__":
    main()
