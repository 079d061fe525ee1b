"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from this repo's oracle;
the reference has none): the oracle, the planner's schedule and -- on a GPU -- the kernels must
reproduce them bit for bit."""
import glob
import os

import numpy as np
import pytest

from helpers import bits_equal
from oracle import xpbd_oracle as orc
from softbodyunity_b200 import SoftBody, ingest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
ALIAS = {"stiffness_distance": "stiffness", "stiffness_volume": "volume_stiffness"}


def load(path):
    z = np.load(path, allow_pickle=True)
    plan = {k: v for k, v in z["plan"]}
    prm = {k: v for k, v in z["prm"]}
    return z, plan, prm


def test_fixtures_exist():
    assert len(FIXTURES) >= 3


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_and_planner_reproduce_the_fixture(path):
    z, plan, prm = load(path)
    sb = SoftBody(z["pos"], z["tets"], z["tris"], host_only=True, **plan, **{ALIAS.get(k, k): v for k, v in prm.items()})
    order, off = sb.schedule()
    assert np.array_equal(order, z["order"]) and np.array_equal(off, z["batch_off"]), "the planner's schedule changed"
    order_odd, off_odd = sb.schedule(odd=True)
    assert np.array_equal(order_odd, z["order_odd"]) and np.array_equal(off_odd, z["batch_off_odd"]), "the planner's schedule changed"
    m = orc.Model(z["pos"], z["tets"], roles=z["roles"])
    assert np.array_equal(sb.tet_roles(), z["roles"]), "the planner's tet roles changed"
    cols = z["colliders"] if "colliders" in z else None
    m.simulate(orc.params(**prm), n_frames=int(z["frames"]), order=z["order"], batch_off=z["batch_off"], order_odd=z["order_odd"],
               batch_off_odd=z["batch_off_odd"], colliders=cols)
    assert bits_equal(m.x4, z["x4"]) and bits_equal(m.v4, z["v4"])
    assert bits_equal(m.normals(z["tris"]), z["normals"])
    if "render_pos" in z:  # the body came out of the ingest path and carries an embedded render mesh
        p, t, f = ingest.tetrahedralize_surface(z["render_pos"], z["render_tris"], 0.07)
        assert bits_equal(p, z["pos"]) and np.array_equal(t, z["tets"]) and np.array_equal(f, z["tris"]), "ingest changed"
        sb.skin_bind(z["render_pos"], z["render_tris"])
        tet_of, bary = sb.skin_binding()
        assert np.array_equal(tet_of, z["skin_tet"]) and bits_equal(bary, z["skin_bary"]), "the skin binding changed"
        sp, sn = m.skin(tet_of, bary, z["render_tris"])
        assert bits_equal(sp, z["skin_pos"]) and bits_equal(sn, z["skin_nrm"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_gpu_reproduces_the_fixture(path):
    z, plan, prm = load(path)
    sb = SoftBody(z["pos"], z["tets"], z["tris"], **plan, **{ALIAS.get(k, k): v for k, v in prm.items()})
    if "colliders" in z:
        sb.set_colliders_ex(z["colliders"])
    sb.step(frames=int(z["frames"]))
    x4, v4 = sb.get_state()
    assert bits_equal(x4, z["x4"]) and bits_equal(v4, z["v4"])
    assert bits_equal(sb.normals(), z["normals"])
    if "render_pos" in z:
        sb.skin_bind(z["render_pos"], z["render_tris"])
        sp, sn = sb.read_skinned()
        assert bits_equal(sp, z["skin_pos"]) and bits_equal(sn, z["skin_nrm"])
