"""Short run for ncu: builds the 1 M block and steps a few projection sweeps (no graph)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from softbodyunity_b200 import FLAG_FAST_MATH, FLAG_NO_GRAPH, SoftBody, meshgen

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100)
ap.add_argument("--substeps", type=int, default=1)
ap.add_argument("--iterations", type=int, default=3)
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--block-threads", type=int, default=0)
ap.add_argument("--fast-math", action="store_true")
ap.add_argument("--tile-cap", type=int, default=0)
ap.add_argument("--round-width", type=int, default=0)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--info-out", default="")
a = ap.parse_args()
pos, tets, tris = meshgen.block(a.n, spacing=0.01, origin=(0.0, 0.002, 0.0))
sb = SoftBody(pos, tets, tris, substeps=a.substeps, iterations=a.iterations, block_threads=a.block_threads, tile_cap=a.tile_cap, round_width=a.round_width,
              flags=a.flags | FLAG_NO_GRAPH | (FLAG_FAST_MATH if a.fast_math else 0))
sb.step(frames=a.frames)
sb.synchronize()
d = sb.diagnostics()
i = sb.info()
print("ok", i["launches_per_frame"], d["min_y"], d["nonfinite"])
if a.info_out:  # what plan the capture is of: bench.py refuses to quote the traffic of another plan
    import json
    n = i["n_tile_passes"]
    json.dump({"rounds_per_sweep": int(sum(i["rounds_in_pass"][:n])), "tiles_in_pass": [int(t) for t in i["tiles_in_pass"][:n]],
               "substeps": a.substeps, "iterations": a.iterations, "program": sb.frame_program().tolist()}, open(a.info_out, "w"))
