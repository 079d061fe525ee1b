"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from this repo's oracle;
the reference has none): the oracle, the planner's schedule and -- on a GPU -- the kernels must
reproduce them bit for bit."""
import glob
import os

import numpy as np
import pytest

from helpers import bits_equal
from oracle import xpbd_oracle as orc
from softbodyunity_b200 import SoftBody

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
ALIAS = {"stiffness_distance": "stiffness", "stiffness_volume": "volume_stiffness"}


def load(path):
    z = np.load(path, allow_pickle=True)
    plan = {k: v for k, v in z["plan"]}
    prm = {k: v for k, v in z["prm"]}
    return z, plan, prm


def test_fixtures_exist():
    assert len(FIXTURES) >= 2


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_and_planner_reproduce_the_fixture(path):
    z, plan, prm = load(path)
    sb = SoftBody(z["pos"], z["tets"], z["tris"], host_only=True, **plan, **{ALIAS.get(k, k): v for k, v in prm.items()})
    order, off = sb.schedule()
    assert np.array_equal(order, z["order"]) and np.array_equal(off, z["batch_off"]), "the planner's schedule changed"
    m = orc.Model(z["pos"], z["tets"], roles=z["roles"])
    assert np.array_equal(sb.tet_roles(), z["roles"]), "the planner's tet roles changed"
    m.simulate(orc.params(**prm), n_frames=int(z["frames"]), order=z["order"], batch_off=z["batch_off"])
    assert bits_equal(m.x4, z["x4"]) and bits_equal(m.v4, z["v4"])
    assert bits_equal(m.normals(z["tris"]), z["normals"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_gpu_reproduces_the_fixture(path):
    z, plan, prm = load(path)
    sb = SoftBody(z["pos"], z["tets"], z["tris"], **plan, **{ALIAS.get(k, k): v for k, v in prm.items()})
    sb.step(frames=int(z["frames"]))
    x4, v4 = sb.get_state()
    assert bits_equal(x4, z["x4"]) and bits_equal(v4, z["v4"])
    assert bits_equal(sb.normals(), z["normals"])
