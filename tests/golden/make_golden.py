"""Generates tests/golden/*.npz: small pinned cases of the substep path.

There are no golden vectors in the reference (its mount is /root/reference/README.md:1 only), so these
are REGRESSION PINS produced by this repo's own CPU oracle (oracle/xpbd_oracle.c) under the Gauss-Seidel
order the planner exports; they pin the oracle, the planner's schedule and -- on a GPU -- the kernels
against silent change.  Run from the repo root:  python tests/golden/make_golden.py [case ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import xpbd_oracle as orc  # noqa: E402
from softbodyunity_b200 import SoftBody, meshgen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # config 1 substitute: soft cube dropped on the ground plane, 60 Hz, 10 x 10
    "sample_cube_6": dict(gen=lambda: meshgen.sample_cube(6, centre_height=0.55, jitter=0.05, seed=1234),
                          plan=dict(tile_cap=100), prm=dict(), frames=30),
    # soft body with compliance, damping, friction and a tilted gravity vector
    "soft_block": dict(gen=lambda: meshgen.block(7, 5, 6, spacing=0.05, origin=(0.0, 0.02, 0.0), seed=7),
                       plan=dict(tile_cap=64, later_tile_cap=48),
                       prm=dict(stiffness_distance=3.0e4, stiffness_volume=1.0e9, damping=0.5, friction=0.3,
                                gravity=(0.4, -9.81, -0.2), substeps=6, iterations=5), frames=20),
}


def _uv_sphere(r, n_lat, n_lon, centre):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_ingest import uv_sphere
    return uv_sphere(r, n_lat, n_lon, centre)


def _ingested_sphere():
    """Surface mesh -> tets through the ingest path; the surface itself is the render mesh."""
    from softbodyunity_b200 import ingest
    sp, st = _uv_sphere(0.25, 10, 20, (0.0, 0.5, 0.0))
    pos, tets, tris = ingest.tetrahedralize_surface(sp, st, 0.07)
    return pos, tets, tris, sp, st


# sphere / capsule / box colliders with friction, a body made by the ingest path, an embedded render mesh
CASES["ingested_sphere_on_colliders"] = dict(
    gen=_ingested_sphere, plan=dict(tile_cap=128),
    prm=dict(stiffness_distance=4.0e4, friction=0.25, substeps=5, iterations=6), frames=30,
    colliders=[("box", 0.3, -0.15, 0.12, 0.0, 0.2, 0.1, 0.3, 0.0, 0.0, 0.2588190451, 0.9659258263),
               ("capsule", 0.5, 0.05, 0.1, -0.3, 0.07, 0.3, 0.2, 0.3),
               ("sphere", 0.1, 0.0, 0.5, 0.05, 0.06)])


def main():
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        pos, tets, tris, *render = c["gen"]()
        kw = dict(c["prm"])
        sb = SoftBody(pos, tets, tris, host_only=True, **c["plan"],
                      **{{"stiffness_distance": "stiffness", "stiffness_volume": "volume_stiffness"}.get(k, k): v for k, v in kw.items()})
        order, off = sb.schedule()
        order_odd, off_odd = sb.schedule(odd=True)  # the tile passes backwards: iterations 1, 3, 5 ... of a substep
        m = orc.Model(pos, tets, roles=sb.tet_roles())
        extra = {}
        cols = None
        if "colliders" in c:
            cols = orc.colliders(c["colliders"])
            extra["colliders"] = cols
        m.simulate(orc.params(**kw), n_frames=c["frames"], order=order, batch_off=off, order_odd=order_odd, batch_off_odd=off_odd, colliders=cols)
        normals = m.normals(tris)
        if render:
            sb.skin_bind(*render)
            tet_of, bary = sb.skin_binding()
            skin_pos, skin_nrm = m.skin(tet_of, bary, render[1])
            extra.update(render_pos=render[0], render_tris=render[1], skin_tet=tet_of, skin_bary=bary, skin_pos=skin_pos, skin_nrm=skin_nrm)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), pos=pos, tets=tets, tris=tris, order=order, batch_off=off, order_odd=order_odd, batch_off_odd=off_odd, roles=sb.tet_roles(),
                            x4=m.x4, v4=m.v4, normals=normals, frames=c["frames"],
                            plan=np.array(sorted(c["plan"].items()), dtype=object), prm=np.array(sorted(kw.items()), dtype=object), **extra)
        print(name, pos.shape, tets.shape, "batches", len(off) - 1, "min y", float(m.x4[:, 1].min()))


if __name__ == "__main__":
    main()
