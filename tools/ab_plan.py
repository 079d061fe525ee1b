"""A/B of planner switches on the GPU: the same mesh planned with different SB_* environment switches (read once
per process, hence one subprocess per variant), stepped in ground contact, device-timed (sb_time_frames).
usage: python tools/ab_plan.py [--n 100] [--frames 20] [--variants "SB_RECOLOUR=0 SB_ATTACH_AUGMENT=0;;SB_RECOLOUR=1"]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, json, time
sys.path.insert(0, %r)
from softbodyunity_b200 import SoftBody, meshgen
n, frames, bodies = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
if bodies:
    pos, tets, tris = meshgen.bodies(bodies, base_height=0.002)  # BASELINE.json configs[3]: 2028-vertex bodies, in ground contact
else:
    pos, tets, tris = meshgen.block(n, spacing=0.01, origin=(0.0, 0.002, 0.0))
t = time.time()
sb = SoftBody(pos, tets, tris)
build = time.time() - t
sb.step(frames=4)
sb.synchronize()
ms = min(sb.time_frames(frames) / frames for _ in range(3))
i = sb.info()
print(json.dumps(dict(n_verts=len(pos), ms_per_frame=ms, gvs=len(pos) * sb.params.substeps / ms / 1e6, rounds=sum(i["rounds_in_pass"]),
                      attached=i["edges_attached"], build_s=round(build, 2), nonfinite=sb.diagnostics()["nonfinite"],
                      passes=[round(sb.time_kernel(16 + p, 50) * 1e3, 2) for p in range(i["n_tile_passes"])])))
""" % ROOT

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100)
ap.add_argument("--frames", type=int, default=20)
ap.add_argument("--bodies", type=int, default=0, help="time a batch of this many 2028-vertex bodies instead of the n^3 block")
ap.add_argument("--variants", default="SB_RECOLOUR=0 SB_ATTACH_AUGMENT=0;SB_RECOLOUR=1 SB_ATTACH_AUGMENT=0;SB_RECOLOUR=0 SB_ATTACH_AUGMENT=1;")
a = ap.parse_args()
for v in a.variants.split(";"):
    env = dict(os.environ)
    env.update(dict(kv.split("=") for kv in v.split()))
    r = subprocess.run([sys.executable, "-c", CHILD, str(a.n), str(a.frames), str(a.bodies)], env=env, capture_output=True, text=True)
    print((v or "(defaults)") + ":", r.stdout.strip() or r.stderr[-400:], flush=True)
