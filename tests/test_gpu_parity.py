"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle replaying
the same Gauss-Seidel order on the same seeded inputs.

Tolerances.  Exact mode (default): BIT-IDENTICAL fp32 state.  Fast-math mode and the
full-size runs: max position error <= 1e-4 of the body's bounding-box diagonal after
100 substeps x iterations worth of stepping (BASELINE.json:5: "within 1e-4 relative
(fp32) after 100 steps").  PARITY UNPINNED with respect to the upstream C# solver (not
in the mount): the oracle is the in-repo XPBD restatement.
"""
import math
import os

import numpy as np
import pytest

from helpers import INF, bits_equal, oracle_params, rel_err, ulp_diff_count
from oracle import xpbd_oracle as orc
from softbodyunity_b200 import FLAG_FAST_MATH, FLAG_NO_GRAPH, FLAG_NO_GROUND, SoftBody, meshgen

pytestmark = pytest.mark.gpu


def run_pair(pos, tets, tris, n_frames, spheres=None, inv_mass=None, colliders=None, **kw):
    sb = SoftBody(pos, tets, tris, inv_mass=inv_mass, **kw)
    if spheres is not None:
        sb.set_colliders(spheres)
    if colliders is not None:
        sb.set_colliders_ex(colliders)
    sched = sb.schedule_kw()
    m = orc.Model(pos, tets, inv_mass=inv_mass, density=kw.get("density", 1000.0), roles=sb.tet_roles())
    sb.step(frames=n_frames)
    x4, v4 = sb.get_state()
    m.simulate(oracle_params(sb), n_frames=n_frames, spheres=spheres, threads=8, **sched,
               colliders=colliders)
    return sb, m, x4, v4


CASES = {
    "one_tile": (lambda: meshgen.sample_cube(8, centre_height=0.7, jitter=0.05), dict()),
    "multi_pass": (lambda: meshgen.block(14, 12, 11, spacing=0.05, origin=(0, 0.03, 0)), dict(tile_cap=256, later_tile_cap=128)),
    "global_only": (lambda: meshgen.block(9, 9, 9, spacing=0.05, origin=(0, 0.03, 0)), dict(max_tile_passes=0)),
    "one_pass_then_global": (lambda: meshgen.block(12, 10, 9, spacing=0.05, origin=(0, 0.02, 0)), dict(tile_cap=200, max_tile_passes=1)),
    "sphere": (lambda: meshgen.sphere(16, spacing=0.05), dict(tile_cap=512)),
    "soft": (lambda: meshgen.block(10, 8, 8, spacing=0.05, origin=(0, 0.02, 0)), dict(stiffness=2.0e4, volume_stiffness=1.0e9, damping=0.5, friction=0.4, tile_cap=300)),
    "bt128": (lambda: meshgen.block(12, 12, 10, spacing=0.05, origin=(0, 0.02, 0)), dict(tile_cap=700, block_threads=128)),
    "bt32": (lambda: meshgen.block(12, 12, 10, spacing=0.05, origin=(0, 0.02, 0)), dict(tile_cap=300, block_threads=32)),
    "bt256": (lambda: meshgen.block(12, 12, 10, spacing=0.05, origin=(0, 0.02, 0)), dict(tile_cap=700, block_threads=256)),
    "no_pdl": (lambda: meshgen.block(12, 12, 10, spacing=0.05, origin=(0, 0.02, 0)), dict(tile_cap=300, flags=16)),
    "dag": (lambda: meshgen.block(14, 12, 11, spacing=0.05, origin=(0, 0.02, 0)), dict(tile_cap=256, flags=32)),
    "dag_bt32": (lambda: meshgen.block(14, 12, 11, spacing=0.05, origin=(0, 0.02, 0)), dict(tile_cap=128, flags=32, block_threads=32)),
    "dag_wide": (lambda: meshgen.block(16, 16, 16, spacing=0.05, origin=(0, 0.02, 0)), dict(tile_cap=512, flags=32, round_width=2, block_threads=128)),
    "wide_rounds": (lambda: meshgen.block(12, 12, 10, spacing=0.05, origin=(0, 0.02, 0)), dict(tile_cap=1500, round_width=2)),
    "wide_bt32": (lambda: meshgen.block(12, 12, 10, spacing=0.05, origin=(0, 0.02, 0)), dict(tile_cap=300, round_width=2, block_threads=32)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_exact_mode_is_bit_identical_to_the_oracle(name):
    gen, kw = CASES[name]
    pos, tets, tris = gen()
    sb, m, x4, v4 = run_pair(pos, tets, tris, n_frames=12, **kw)
    assert m.x4[:, 1].min() < 0.01 or name in ("one_tile", "sphere"), "case should be in ground contact"
    assert ulp_diff_count(x4, m.x4) == 0, f"{ulp_diff_count(x4, m.x4)} of {x4.size} position words differ"
    assert bits_equal(v4, m.v4)
    # positions() is the same data, unpermuted, xyz only
    assert bits_equal(sb.positions(), m.x4[:, :3])


def test_hundred_steps_within_1e4_both_modes():
    # "100 steps": 10 frames x 10 substeps, 10 iterations each, ground contact
    pos, tets, tris = meshgen.block(20, 16, 16, spacing=0.02, origin=(0, 0.01, 0))
    for flags, exact in ((0, True), (FLAG_FAST_MATH, False)):
        sb, m, x4, v4 = run_pair(pos, tets, tris, n_frames=10, tile_cap=1024, flags=flags)
        err = rel_err(x4, m.x4)
        assert err <= 1e-4, err
        assert m.x4[:, 1].min() == 0.0
        if exact:
            assert err == 0.0


def test_colliders_pins_and_gravity_vector():
    pos, tets, tris = meshgen.block(10, 10, 10, spacing=0.05, origin=(-0.225, 0.6, -0.225))
    w = orc.lumped_inv_mass(pos, tets)
    w[pos[:, 1] > pos[:, 1].max() - 0.01] = 0.0  # pin the top layer
    sph = np.array([[0.0, 0.45, 0.0, 0.2], [0.3, 0.3, 0.1, 0.15]], np.float32)
    sb, m, x4, v4 = run_pair(pos, tets, tris, n_frames=20, spheres=sph, inv_mass=w, gravity=(0.5, -9.81, -0.25),
                             tile_cap=400, stiffness=5e4)
    assert bits_equal(x4, m.x4) and bits_equal(v4, m.v4)
    pinned = w == 0
    assert np.array_equal(x4[pinned, :3], pos[pinned])
    d = np.linalg.norm(x4[:, :3] - sph[0, :3], axis=1)
    assert d.min() >= 0.2 - 1e-5


def test_capsule_box_sphere_colliders_with_friction_are_bit_identical():
    # a soft block impaled on a sphere, a degenerate capsule and an axis-aligned box falls onto a tilted box and a
    # capsule; every collider moves vertices (checked with the oracle alone when the scene was set up)
    pos, tets, tris = meshgen.block(12, 8, 12, spacing=0.05, origin=(-0.275, 0.45, -0.275))
    q = np.array([0.0, 0.0, math.sin(0.15), math.cos(0.15)]) * 1.7  # about z, deliberately not normalised
    cols = orc.colliders([
        ("box", 0.3, -0.2, 0.3, 0.0, 0.15, 0.1, 0.3, *q),
        ("capsule", 0.5, 0.1, 0.3, -0.3, 0.08, 0.35, 0.4, 0.3),
        ("sphere", 0.2, 0.1, 0.6, 0.1, 0.08),
        ("capsule", 0.0, -0.1, 0.65, -0.1, 0.06, -0.1, 0.65, -0.1),  # A == B: a sphere
        ("box", 0.0, 0.0, 0.6, -0.15, 0.06, 0.05, 0.07, 0, 0, 0, 0),  # zero quaternion: axis aligned
    ])
    sb, m, x4, v4 = run_pair(pos, tets, tris, n_frames=40, colliders=cols, tile_cap=400, stiffness=3e4, friction=0.3)
    assert bits_equal(x4, m.x4) and bits_equal(v4, m.v4)
    d = np.linalg.norm(x4[:, :3] - cols[2]["p"][:3], axis=1)
    assert d.min() >= 0.08 - 1e-5
    fresh = orc.Model(pos, tets, roles=sb.tet_roles())
    sched = sb.schedule_kw()
    fresh.simulate(oracle_params(sb), n_frames=40, threads=8, **sched)
    assert np.abs(fresh.x4 - m.x4).max() > 1e-2, "colliders should change the outcome"
    # replacing the list: sb_set_colliders (spheres, no friction) after sb_set_colliders_ex
    sb.set_colliders(np.array([[0.0, 0.2, 0.0, 0.15]], np.float32))
    sb.step(frames=3)
    m.simulate(oracle_params(sb), n_frames=3, spheres=np.array([[0.0, 0.2, 0.0, 0.15]], np.float32), threads=8, **sched)
    x4, v4 = sb.get_state()
    assert bits_equal(x4, m.x4) and bits_equal(v4, m.v4)


def test_normals_match_and_surface_readback():
    pos, tets, tris = meshgen.sphere(14, spacing=0.05)
    sb, m, x4, v4 = run_pair(pos, tets, tris, n_frames=6, tile_cap=600)
    n_ref = m.normals(tris)
    n_gpu = sb.normals()
    assert bits_equal(n_gpu, n_ref)
    ids = sb.surface_vertices()
    sp, sn = sb.read_surface()
    assert bits_equal(sp, m.x4[ids, :3]) and bits_equal(sn, n_ref[ids])
    np.testing.assert_allclose(np.linalg.norm(sn, axis=1), 1.0, atol=1e-6)


def test_graph_and_direct_launch_agree_and_params_update():
    pos, tets, tris = meshgen.block(9, 9, 8, spacing=0.05, origin=(0, 0.2, 0))
    a = SoftBody(pos, tets, tris, tile_cap=300)
    b = SoftBody(pos, tets, tris, tile_cap=300, flags=FLAG_NO_GRAPH)
    for sb in (a, b):
        sb.step(frames=3)
        sb.set_params(substeps=4, iterations=7, damping=1.0)   # new graph topology
        sb.step(frames=3)
        sb.step(dt=0.005, frames=2)                            # dt change: constants only
    xa, va = a.get_state()
    xb, vb = b.get_state()
    assert bits_equal(xa, xb) and bits_equal(va, vb)
    # oracle through the same parameter changes
    sched = a.schedule_kw()
    m = orc.Model(pos, tets, roles=a.tet_roles())
    m.simulate(orc.params(substeps=10, iterations=10), n_frames=3, **sched)
    m.simulate(orc.params(substeps=4, iterations=7, damping=1.0), n_frames=3, **sched)
    m.simulate(orc.params(dt=0.005, substeps=4, iterations=7, damping=1.0), n_frames=2, **sched)
    assert bits_equal(xa, m.x4) and bits_equal(va, m.v4)


def test_set_state_get_state_round_trip_and_restart():
    pos, tets, tris = meshgen.block(8, 8, 8, spacing=0.05, origin=(0, 0.3, 0))
    a = SoftBody(pos, tets, tris, tile_cap=200)
    a.step(frames=5)
    x4, v4 = a.get_state()
    a.step(frames=5)
    xa, va = a.get_state()
    b = SoftBody(pos, tets, tris, tile_cap=200)
    b.set_state(x4, v4)
    x4b, v4b = b.get_state()
    assert bits_equal(x4, x4b) and bits_equal(v4, v4b)
    b.step(frames=5)
    xb, vb = b.get_state()
    assert bits_equal(xa, xb) and bits_equal(va, vb)  # checkpoint/resume is exact


def test_diagnostics_match_oracle():
    pos, tets, tris = meshgen.block(10, 9, 8, spacing=0.05, origin=(0, 0.05, 0))
    sb, m, x4, v4 = run_pair(pos, tets, tris, n_frames=8, tile_cap=300, stiffness=1e5)
    d = sb.diagnostics()["raw"]
    ref = m.diagnostics()
    np.testing.assert_allclose(d[:12], ref[:12], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(d[12:14], ref[12:14], rtol=1e-9)
    assert d[14] == 0 and d[15] == ref[15]


def test_energy_and_volume_drift_track_the_oracle():
    # BASELINE.json:5: "energy and volume drift must match"; sampled every 10 frames over 100 frames
    pos, tets, tris = meshgen.sample_cube(7, centre_height=0.8, jitter=0.05)
    sb = SoftBody(pos, tets, tris, stiffness=1e5, substeps=5, iterations=4)
    sched = sb.schedule_kw()
    m = orc.Model(pos, tets, roles=sb.tet_roles())
    p = oracle_params(sb)
    for _ in range(10):
        sb.step(frames=10)
        m.simulate(p, n_frames=10, **sched)
        d, r = sb.diagnostics()["raw"], m.diagnostics()
        e_gpu, e_ref = d[0] + d[1], r[0] + r[1]
        assert abs(e_gpu - e_ref) <= 1e-9 * max(1.0, abs(e_ref))
        assert abs(d[2] - r[2]) <= 1e-12 + 1e-9 * abs(r[2])
    assert abs(r[2] - 1.0) < 0.05


def test_sphere_100k_hundred_substeps_both_modes():
    # BASELINE.json configs[1]: the 100 k-vertex tet-mesh sphere, 10 substeps x 10 iterations; 100 substeps (10 frames):
    # exact mode bit-identical to the oracle, fast math within 1e-4 of the bounding-box diagonal
    pos, tets, tris = meshgen.sphere(58, spacing=0.01)
    assert 95_000 < len(pos) < 110_000
    for flags, exact in ((0, True), (FLAG_FAST_MATH, False)):
        sb, m, x4, v4 = run_pair(pos, tets, tris, n_frames=10, flags=flags)
        assert m.x4[:, 1].min() == 0.0  # in ground contact
        if exact:
            assert ulp_diff_count(x4, m.x4) == 0 and bits_equal(v4, m.v4)
        else:
            assert rel_err(x4, m.x4) <= 1e-4, rel_err(x4, m.x4)


def test_thousand_substeps_energy_and_volume_drift_track_the_oracle():
    # BASELINE.json:5 "energy and volume drift must match over 1000 steps": the config-1 substitute (soft cube dropped on the
    # ground plane, 60 Hz, 10 x 10), 100 frames = 1000 substeps, diagnostics sampled every 10 frames; the state itself stays
    # bit-identical, so the curves agree to the rounding of the fp64 reductions
    pos, tets, tris = meshgen.sample_cube(6, centre_height=0.55, jitter=0.05, seed=1234)
    sb = SoftBody(pos, tets, tris, stiffness=2e5)
    sched = sb.schedule_kw()
    m = orc.Model(pos, tets, roles=sb.tet_roles())
    p = oracle_params(sb)
    e0 = None
    for _ in range(10):
        sb.step(frames=10)
        m.simulate(p, n_frames=10, **sched)
        d, r = sb.diagnostics()["raw"], m.diagnostics()
        e_gpu, e_ref = d[0] + d[1], r[0] + r[1]
        e0 = e_ref if e0 is None else e0
        assert abs(e_gpu - e_ref) <= 1e-9 * max(1.0, abs(e_ref))
        assert abs(d[2] - r[2]) <= 1e-12 + 1e-9 * abs(r[2])
    x4, v4 = sb.get_state()
    assert ulp_diff_count(x4, m.x4) == 0
    assert sb.frames_done == 100


def test_many_bodies_batch():
    pos, tets, tris = meshgen.bodies(40, dims=(6, 5, 5), spacing=0.04, base_height=0.03)
    sb, m, x4, v4 = run_pair(pos, tets, tris, n_frames=10, tile_cap=512)
    assert sb.info()["n_tile_passes"] == 1
    assert bits_equal(x4, m.x4)


def test_large_mesh_properties():
    # full-size path (1 M vertices, the headline config) through size-independent properties:
    # finite state, volume preserved, nothing under the ground, centre of mass falls as predicted
    # before contact, and the exact-mode checksum matches the oracle on a 1-frame replay.
    pos, tets, tris = meshgen.block(100)
    sb = SoftBody(pos, tets, tris)
    i = sb.info()
    assert i["n_verts"] == 1_000_000 and i["constraints_global"] == 0
    d0 = sb.diagnostics()
    sb.step(frames=2)
    x4, v4 = sb.get_state()
    sched = sb.schedule_kw()
    m = orc.Model(pos, tets, roles=sb.tet_roles())
    threads = os.cpu_count() or 16
    m.simulate(oracle_params(sb), n_frames=2, threads=threads, **sched)
    assert ulp_diff_count(x4, m.x4) == 0
    # BASELINE.json:5 "after 100 steps": 10 frames x 10 substeps x 10 iterations of the headline mesh, bit for bit
    sb.step(frames=8)
    x4, v4 = sb.get_state()
    m.simulate(oracle_params(sb), n_frames=8, threads=threads, **sched)
    assert ulp_diff_count(x4, m.x4) == 0 and bits_equal(v4, m.v4)
    d1 = sb.diagnostics()
    assert d1["nonfinite"] == 0 and d1["min_y"] >= 0.0
    # 10 sweeps per substep do not converge a 100-layer stack: it compresses a few percent on impact
    assert abs(d1["volume"] - d0["volume"]) / d0["volume"] < 0.05
    assert d1["rms_strain"] < 0.1  # (the bottom cells of the under-converged stack do collapse on impact)


@pytest.mark.parametrize("S,I", [(3, 4), (2, 5), (4, 1), (1, 6), (3, 2)])
@pytest.mark.parametrize("flags", [0, 128, 64, 64 | 128, 16, 4])
def test_snake_order_and_fused_launches_match_the_oracle(S, I, flags):
    # default: odd iterations run the passes backwards and consecutive occurrences of a pass share a launch that
    # also carries the substep boundary (finish + predict on the tile); 128 = no fusion, 64 = no snake,
    # 16 = no PDL, 4 = no graph.  Every variant is bit-identical to the oracle replaying ITS order, and fusion
    # never changes a bit.
    pos, tets, tris = meshgen.block(13, 11, 10, spacing=0.05, origin=(0, 0.001, 0))
    kw = dict(tile_cap=256, substeps=S, iterations=I, friction=0.3, damping=0.2, stiffness=5e4)
    sb, m, x4, v4 = run_pair(pos, tets, tris, n_frames=7, spheres=np.array([[0.3, 0.1, 0.25, 0.2]], np.float32), flags=flags, **kw)
    assert bits_equal(x4, m.x4) and bits_equal(v4, m.v4)  # (ground, a sphere collider inside the block, friction, damping)
    if flags == 0:
        other = SoftBody(pos, tets, tris, flags=128, **kw)
        other.set_colliders(np.array([[0.3, 0.1, 0.25, 0.2]], np.float32))
        other.step(frames=7)
        xo, vo = other.get_state()
        assert bits_equal(x4, xo) and bits_equal(v4, vo)
        assert len(sb.frame_program()) < len(other.frame_program())


def test_a_batch_of_bodies_runs_a_frame_in_one_launch():
    pos, tets, tris = meshgen.bodies(60, dims=(7, 6, 5), spacing=0.04, base_height=0.01)
    sb, m, x4, v4 = run_pair(pos, tets, tris, n_frames=8, tile_cap=512, substeps=4, iterations=5, friction=0.2)
    prog = sb.frame_program()
    assert len(prog) == 2 and tuple(prog[0]) == (2, 0, 4, 5, 1, 1) and prog[1][0] == 6
    assert m.x4[:, 1].min() == 0.0 and bits_equal(x4, m.x4) and bits_equal(v4, m.v4)
    # pinned vertices and a tile without constraints ride through the vertex stages too
    w = np.asarray(sb.topology()[3]).copy()
    w[::17] = 0.0
    sb2, m2, x2, v2 = run_pair(pos, tets, tris, n_frames=5, inv_mass=w, tile_cap=512, substeps=3, iterations=2)
    assert bits_equal(x2, m2.x4) and bits_equal(v2, m2.v4)
