"""CPU oracle run of a bench.py workload (default `--workload dist`: ONE n^3 block, the plan of an 8-way split; `block`:
the single-GPU headline mesh) -> the state checksum bench.py prints (`state_checksum.x4_words_hi_lo`) after given numbers
of frames, written to tests/golden/dist_checksum.json.  bench.py compares what the GPUs computed with this file (it reads the JSON, never the
oracle).  TEST INFRASTRUCTURE: this is the only place outside tests/ proper that runs the oracle at this size.

    python tests/golden/make_dist_checksum.py [--n 200] [--frames 13 23 25]                 # 8 M vertices: 1.5 hours on 8 cores
    python tests/golden/make_dist_checksum.py --workload block --n 100 --frames 23 25       # the 1 M headline mesh: 10 minutes
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import bench  # noqa: E402
from helpers import oracle_params  # noqa: E402
from oracle import xpbd_oracle as orc  # noqa: E402
from softbodyunity_b200 import SoftBody  # noqa: E402


def checksum(x4):
    w = np.ascontiguousarray(x4, np.float32).view(np.uint32).astype(np.int64)
    return "%x-%x" % (int((w >> 16).sum()), int((w & 0xffff).sum()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="dist", choices=["dist", "block", "sphere"])
    ap.add_argument("--n", type=int, default=200)
    ap.add_argument("--frames", type=int, nargs="+", default=[13, 25])
    ap.add_argument("--out", default=os.path.join(HERE, "dist_checksum.json"))
    a = ap.parse_args()
    args = bench.parse_args(["--workload", a.workload, "--n", str(a.n)])
    pos, tets, tris, name = bench.workload(args)
    plan = SoftBody(pos, tets, tris, host_only=True, substeps=args.substeps, iterations=args.iterations, flags=bench.solver_flags(args),
                    **bench.plan_options(args))
    info = plan.info()
    cfg = bench.describe_config(args, info, name, len(pos), 1)
    m = orc.Model(pos, tets, roles=plan.tet_roles())
    prm, sched = oracle_params(plan), plan.schedule_kw()
    key = bench.checksum_key(cfg, args)
    doc = json.load(open(a.out)) if os.path.exists(a.out) else {}
    entry = doc.setdefault(key, {"workload": cfg["workload"], "after_frames": {}})
    done, t0 = 0, time.time()
    for f in sorted(a.frames):
        m.simulate(prm, n_frames=f - done, threads=os.cpu_count() or 1, **sched)
        done = f
        entry["after_frames"][str(f)] = checksum(m.x4)
        entry["oracle_seconds"] = round(time.time() - t0, 1)
        json.dump(doc, open(a.out, "w"), indent=1, sort_keys=True)
        print(f, entry["after_frames"][str(f)], "%.0f s" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
