"""GPU side of the ingest / write-back widening: a render mesh driven by the tets (k_skin +
normals) and state snapshots, through the C ABI, against the CPU oracle.  Exact mode: bit-identical."""
import numpy as np
import pytest

from helpers import bits_equal, oracle_params
from oracle import xpbd_oracle as orc
from softbodyunity_b200 import SbError, SoftBody, ingest, meshgen
from test_ingest import uv_sphere

pytestmark = pytest.mark.gpu


def test_surface_to_tets_to_steps_to_skinned_render_mesh():
    # the whole chain a Unity caller runs: surface mesh -> tets -> solver -> render vertices and normals per frame
    sp, st = uv_sphere(0.4, 20, 40, centre=(0.0, 0.45, 0.0))
    p, t, f = ingest.tetrahedralize_surface(sp, st, 0.06)
    sb = SoftBody(p, t, f, tile_cap=512, stiffness=5e4, friction=0.2)
    sb.skin_bind(sp, st)
    tet_of, b = sb.skin_binding()
    m = orc.Model(p, t, roles=sb.tet_roles())
    sched = sb.schedule_kw()
    pos0, nrm0 = sb.read_skinned()
    ref0, rn0 = m.skin(tet_of, b, st)
    assert bits_equal(pos0, ref0) and bits_equal(nrm0, rn0)
    assert np.abs(pos0 - sp).max() < 1e-6
    for frames in (5, 20):
        sb.step(frames=frames)
        m.simulate(oracle_params(sb), n_frames=frames, threads=8, **sched)
        pos, nrm = sb.read_skinned()
        ref, rn = m.skin(tet_of, b, st)
        assert bits_equal(sb.get_state()[0], m.x4)
        assert bits_equal(pos, ref) and bits_equal(nrm, rn)
        np.testing.assert_allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-5)
    assert pos[:, 1].min() < 0.03 and np.abs(pos - pos0).max() > 0.03  # it fell and squashed on the ground
    assert bits_equal(sb.read_skinned(normals=False), ref)
    # rebinding a different render mesh replaces the first
    sp2, st2 = uv_sphere(0.2, 8, 16, centre=(0.0, 0.45, 0.0))
    sb.skin_bind(sp2, st2)
    t2, b2 = sb.skin_binding()
    pos2, nrm2 = sb.read_skinned()
    ref2, rn2 = m.skin(t2, b2, st2)
    assert pos2.shape == (len(sp2), 3) and bits_equal(pos2, ref2) and bits_equal(nrm2, rn2)


def test_skin_errors():
    pos, tets, tris = meshgen.block(6, 6, 6, spacing=0.05)
    sb = SoftBody(pos, tets, tris)
    sb.n_render = 3
    with pytest.raises(SbError, match="sb_skin_bind"):
        sb.read_skinned()
    with pytest.raises(SbError, match="out of range"):
        sb.skin_bind(pos[:4], np.array([[0, 1, 7]], np.int32))
    sb.skin_bind(pos[:10])  # no triangles: positions only, normals come back zero
    p, n = sb.read_skinned()
    assert bits_equal(p, pos[:10]) or np.abs(p - pos[:10]).max() < 1e-6
    assert not n.any()


def test_snapshot_resume_is_bit_identical(tmp_path):
    pos, tets, tris = meshgen.block(12, 10, 9, spacing=0.05, origin=(0, 0.05, 0))
    a = SoftBody(pos, tets, tris, tile_cap=400, stiffness=4e4, damping=0.3)
    a.step(frames=7)
    snap = tmp_path / "frame7.sbs"
    a.save_state(snap)
    assert a.frames_done == 7
    a.step(frames=9)
    xa, va = a.get_state()
    # a fresh handle resumes from the file
    b = SoftBody(pos, tets, tris, tile_cap=400)  # different parameters until the snapshot's are applied
    b.load_state(snap, apply_params=True)
    assert b.frames_done == 7 and b.params.damping == np.float32(0.3) and b.params.stiffness_distance == 4e4
    b.step(frames=9)
    xb, vb = b.get_state()
    assert bits_equal(xa, xb) and bits_equal(va, vb) and b.frames_done == 16
    # the file is what ingest.read_state reads; it matches the oracle at frame 7
    s = ingest.read_state(snap)
    m = orc.Model(pos, tets, roles=a.tet_roles())
    m.simulate(oracle_params(a), n_frames=7, threads=8, **a.schedule_kw())
    assert s["frame"] == 7 and bits_equal(s["x4"], m.x4) and bits_equal(s["v4"], m.v4)
    assert s["topo_hash"] == ingest.topology_hash(len(pos), tets)
    # a snapshot of another mesh is refused
    pos2, tets2, tris2 = meshgen.block(12, 10, 8, spacing=0.05)
    c = SoftBody(pos2, tets2, tris2)
    with pytest.raises(SbError, match="different mesh"):
        c.load_state(snap)


def test_replay_tool_between_two_snapshots(tmp_path):
    import subprocess
    import sys
    import os
    pos, tets, tris = meshgen.sphere(10, spacing=0.05)
    ingest.save_mesh(tmp_path / "body.msh", pos, tets, tris)
    p, t, f = ingest.load_mesh(tmp_path / "body.msh")
    sb = SoftBody(p, t, f, stiffness=6e4, damping=0.2, substeps=6, iterations=4)
    sb.step(frames=5)
    sb.save_state(tmp_path / "a.sbs")
    sb.step(frames=8)
    sb.save_state(tmp_path / "b.sbs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "replay.py"), "--mesh", str(tmp_path / "body.msh"),
                        "--from", str(tmp_path / "a.sbs"), "--to", str(tmp_path / "b.sbs")], capture_output=True, text=True)
    assert r.returncode == 0 and "0 position words and 0 velocity words differ" in r.stdout, r.stdout + r.stderr
