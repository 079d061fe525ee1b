"""Mesh ingest and on-disk formats, host only: surface -> tets, orientation and
boundary extraction, TetGen / Gmsh round trips, render-mesh binding, state snapshots.  The reference has none of
this in the mount (/root/reference/README.md:1 is all there is), so the checks are geometric known answers and
independent numpy restatements.
"""
import os

import numpy as np
import pytest

from oracle import xpbd_oracle as orc
from softbodyunity_b200 import SbError, SoftBody, default_params, ingest, meshgen


def tet_volumes(pos, tets):
    a = pos[tets].astype(np.float64)
    return np.einsum("ij,ij->i", a[:, 1] - a[:, 0], np.cross(a[:, 2] - a[:, 0], a[:, 3] - a[:, 0])) / 6


def enclosed_volume(pos, tris):
    a = pos[tris].astype(np.float64)
    return np.einsum("ij,ij->i", a[:, 0], np.cross(a[:, 1], a[:, 2])).sum() / 6


def is_watertight(tris):
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])
    fwd = {(int(a), int(b)) for a, b in e}
    return len(fwd) == len(e) and all((b, a) in fwd for a, b in fwd)  # every directed edge once, its reverse once


def surface_of(pos, tris):
    used = np.unique(tris)
    remap = -np.ones(len(pos), np.int64)
    remap[used] = np.arange(len(used))
    return pos[used], remap[tris].astype(np.int32)


def uv_sphere(r=1.0, n_lat=24, n_lon=48, centre=(0, 0, 0)):
    v = [(0, 0, r)]
    for i in range(1, n_lat):
        th = np.pi * i / n_lat
        for j in range(n_lon):
            ph = 2 * np.pi * j / n_lon
            v.append((r * np.sin(th) * np.cos(ph), r * np.sin(th) * np.sin(ph), r * np.cos(th)))
    v.append((0, 0, -r))
    f = []
    ring = lambda i, j: 1 + (i - 1) * n_lon + j % n_lon
    for j in range(n_lon):
        f.append((0, ring(1, j), ring(1, j + 1)))
        f.append((len(v) - 1, ring(n_lat - 1, j + 1), ring(n_lat - 1, j)))
    for i in range(1, n_lat - 1):
        for j in range(n_lon):
            f.append((ring(i, j), ring(i + 1, j), ring(i + 1, j + 1)))
            f.append((ring(i, j), ring(i + 1, j + 1), ring(i, j + 1)))
    return np.asarray(v, np.float32) + np.float32(centre), np.asarray(f, np.int32)


def torus(R=1.0, r=0.35, nu=64, nv=32):
    u, w = np.meshgrid(np.arange(nu) * 2 * np.pi / nu, np.arange(nv) * 2 * np.pi / nv, indexing="ij")
    p = np.stack([(R + r * np.cos(w)) * np.cos(u), r * np.sin(w), (R + r * np.cos(w)) * np.sin(u)], -1).reshape(-1, 3)
    idx = lambda i, j: (i % nu) * nv + j % nv
    f = []
    for i in range(nu):
        for j in range(nv):
            f.append((idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)))
            f.append((idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)))
    return p.astype(np.float32), np.asarray(f, np.int32)


# ---- surface -> tets -------------------------------------------------------------------------------------

def test_cube_surface_gives_the_exact_lattice():
    # the block's own surface triangulation has diagonals that pass exactly through the ray positions
    pos, tets, tris = meshgen.block(6, 6, 6, spacing=0.2, origin=(0, 0, 0), jitter=0.0)
    sp, st = surface_of(pos, tris)
    p, t, f = ingest.tetrahedralize_surface(sp, st, 0.1)
    assert p.shape == (11 ** 3, 3) and t.shape == (5 * 10 ** 3, 4) and f.shape == (12 * 100, 3)
    v = tet_volumes(p, t)
    assert v.min() > 0 and abs(v.sum() - 1.0) < 1e-6
    assert is_watertight(f) and abs(enclosed_volume(p, f) - 1.0) < 1e-6  # outward winding: positive volume
    assert np.abs(p.min(0)).max() < 1e-6 and np.abs(p.max(0) - 1).max() < 1e-6
    # the planner takes it as it is
    sb = SoftBody(p, t, f, host_only=True, tile_cap=512)
    assert sb.info()["n_tets"] == len(t) and sb.verify_streams() == 0


@pytest.mark.parametrize("flip", [False, True])
def test_sphere_and_torus_volumes(flip):
    sp, st = uv_sphere(1.0)
    if flip:
        st = st[:, ::-1].copy()  # inward-wound surface: the winding number is -1 inside, still non-zero
    p, t, f = ingest.tetrahedralize_surface(sp, st, 0.08)
    v = tet_volumes(p, t)
    assert v.min() > 0
    assert abs(v.sum() / (4 / 3 * np.pi) - 1) < 0.03
    assert is_watertight(f) and abs(enclosed_volume(p, f) - v.sum()) < 1e-6 * v.sum() + 1e-9
    # every kept cell centre is inside the sphere, every dropped one outside (up to the faceting of the surface)
    c = p[t].astype(np.float64).reshape(-1, 5, 4, 3).mean((1, 2))[:, :]
    assert (np.linalg.norm(c, axis=1) < 1.0 + 1e-6).all()
    tp, tt = torus()
    p, t, f = ingest.tetrahedralize_surface(tp, tt, 0.06)
    v = tet_volumes(p, t)
    assert abs(v.sum() / (2 * np.pi ** 2 * 1.0 * 0.35 ** 2) - 1) < 0.04
    assert np.linalg.norm(p[:, [0, 2]], axis=1).min() > 1.0 - 0.35 - 0.06 * 1.5  # the hole stays open
    assert is_watertight(f)


def test_snapping_pulls_the_staircase_boundary_onto_the_surface():
    sp, st = uv_sphere(1.0, 32, 64)
    p0, t0, f0 = ingest.tetrahedralize_surface(sp, st, 0.1)
    p1, t1, f1 = ingest.tetrahedralize_surface(sp, st, 0.1, snap=True)
    assert np.array_equal(t0, t1) and np.array_equal(f0, f1)  # topology untouched
    b = np.unique(f0)
    inner = np.setdiff1d(np.arange(len(p0)), b)
    assert np.array_equal(p0[inner], p1[inner])  # only boundary vertices move
    dev0 = np.sqrt(((np.linalg.norm(p0[b], axis=1) - 1) ** 2).mean())
    dev1 = np.sqrt(((np.linalg.norm(p1[b], axis=1) - 1) ** 2).mean())
    assert dev1 < 0.5 * dev0, (dev0, dev1)
    v0, v1 = tet_volumes(p0, t0), tet_volumes(p1, t1)
    assert (v1 > 0.29 * v0).all()  # no tet collapses
    assert abs(v1.sum() / (4 / 3 * np.pi) - 1) < abs(v0.sum() / (4 / 3 * np.pi) - 1)  # and the volume is closer
    # a lattice that already lies on the surface stays put
    pos, tets, tris = meshgen.block(6, 6, 6, spacing=0.2, origin=(0, 0, 0), jitter=0.0)
    cs, ct = surface_of(pos, tris)
    a = ingest.tetrahedralize_surface(cs, ct, 0.1)
    b2 = ingest.tetrahedralize_surface(cs, ct, 0.1, snap=True)
    assert np.abs(a[0] - b2[0]).max() < 1e-6
    # the snapped body plans and steps (oracle, a few frames resting on the ground)
    q = p1.copy()
    q[:, 1] -= q[:, 1].min() - 0.01  # lowest vertex 1 cm above the ground plane
    sb = SoftBody(q, t1, f1, host_only=True)
    order, off = sb.schedule()
    m = orc.Model(q, t1, roles=sb.tet_roles())
    m.simulate(orc.params(stiffness_distance=1e5), n_frames=5, order=order, batch_off=off, threads=4)
    assert np.isfinite(m.x4).all() and m.x4[:, 1].min() >= 0.0


def test_triangle_soup_with_split_vertices_gives_the_same_tets():
    # Unity meshes split vertices at UV seams and hard edges: the surface is closed geometrically, not by index
    sp, st = uv_sphere(0.6, 14, 28, centre=(0.1, 0.2, 0.3))
    ref = ingest.tetrahedralize_surface(sp, st, 0.09)
    soup_pos = sp[st].reshape(-1, 3)
    soup_tri = np.arange(len(soup_pos), dtype=np.int32).reshape(-1, 3)
    got = ingest.tetrahedralize_surface(soup_pos, soup_tri, 0.09)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    # and every split copy binds like the vertex it copies
    t_ref, b_ref = ingest.skin_binding(ref[0], ref[1], sp)
    t_soup, b_soup = ingest.skin_binding(ref[0], ref[1], soup_pos)
    assert np.array_equal(t_soup.reshape(-1, 3), t_ref[st]) and np.array_equal(b_soup.reshape(-1, 3, 4), b_ref[st])


def test_two_components_and_bad_input():
    a, fa = uv_sphere(0.5, 12, 24, centre=(0, 0, 0))
    b, fb = uv_sphere(0.3, 12, 24, centre=(2, 0.1, 0))
    p, t, f = ingest.tetrahedralize_surface(np.concatenate([a, b]), np.concatenate([fa, fb + len(a)]), 0.07)
    c = p[t].astype(np.float64).mean(1)
    near_a, near_b = np.linalg.norm(c, axis=1) < 0.6, np.linalg.norm(c - [2, 0.1, 0], axis=1) < 0.4
    assert near_a.any() and near_b.any() and (near_a | near_b).all()
    with pytest.raises(SbError, match="spacing"):
        ingest.tetrahedralize_surface(a, fa, 0.0)
    with pytest.raises(SbError, match="1e9 lattice cells"):
        ingest.tetrahedralize_surface(a, fa, 1e-5)
    with pytest.raises(SbError, match="not closed"):
        ingest.tetrahedralize_surface(a, fa[a[fa].mean(1)[:, 0] < 0.3], 0.07)  # a hole on the +x side of the sphere
    with pytest.raises(SbError, match="out of range"):
        ingest.tetrahedralize_surface(a, fa + 1000, 0.1)
    with pytest.raises(SbError, match="inside"):
        ingest.tetrahedralize_surface(a[:4] * 0, np.array([[0, 1, 2]] * 4, np.int32), 0.1)  # degenerate: nothing enclosed


def test_from_arrays_fixes_orientation_and_extracts_the_boundary():
    pos, tets, tris = meshgen.sphere(10, spacing=0.1, jitter=0.05)
    bad = tets.copy()
    bad[::3, [0, 1]] = bad[::3, [1, 0]]  # every third tet inside out
    assert (tet_volumes(pos, bad) < 0).any()
    p, t, f = ingest.from_arrays(pos, bad)
    assert np.array_equal(p, pos) and tet_volumes(p, t).min() > 0
    assert np.array_equal(np.sort(t, 1), np.sort(tets, 1))
    key = lambda a: {tuple(sorted(map(int, r))) for r in a}
    assert key(f) == key(tris) and len(f) == len(tris) and is_watertight(f)
    assert abs(enclosed_volume(p, f) - tet_volumes(p, t).sum()) < 1e-9
    # a given triangle list is kept
    p2, t2, f2 = ingest.from_arrays(pos, tets, tris[:7])
    assert np.array_equal(f2, tris[:7])
    with pytest.raises(SbError, match="out of range"):
        ingest.from_arrays(pos, tets + len(pos))


# ---- files ---------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["m.msh", "m.node", "m"])
def test_mesh_files_round_trip_bit_exactly(tmp_path, name):
    pos, tets, tris = meshgen.sphere(9, spacing=0.0371, jitter=0.1)
    path = tmp_path / name
    ingest.save_mesh(path, pos, tets, tris)
    p, t, f = ingest.load_mesh(path)
    assert np.array_equal(p.view(np.uint32), pos.view(np.uint32)) and np.array_equal(t, tets) and np.array_equal(f, tris)
    if name == "m.node":  # any of the three TetGen files names the set
        p2, t2, f2 = ingest.load_mesh(tmp_path / "m.ele")
        assert np.array_equal(t2, tets) and np.array_equal(f2, tris)
        os.remove(tmp_path / "m.face")  # without a .face file the boundary is extracted
        p3, t3, f3 = ingest.load_mesh(tmp_path / "m.node")
        assert {tuple(sorted(map(int, r))) for r in f3} == {tuple(sorted(map(int, r))) for r in tris}


def test_hand_written_tetgen_and_gmsh_files(tmp_path):
    # 1-based ids, comments, attributes and boundary markers, an inverted tet
    (tmp_path / "a.node").write_text("# two tets\n5 3 1 1\n1 0 0 0 7.5 1\n2 1 0 0 7.5 1\n3 0 1 0 7.5 1\n4 0 0 1 7.5 0\n5 1 1 1 7.5 1 # apex\n")
    (tmp_path / "a.ele").write_text("2 4 1\n1  1 2 3 4  9\n2  5 2 3 4  9\n")
    p, t, f = ingest.load_mesh(tmp_path / "a.node")
    assert p.shape == (5, 3) and np.array_equal(np.sort(t, 1), [[0, 1, 2, 3], [1, 2, 3, 4]])
    assert tet_volumes(p, t).min() > 0 and len(f) == 6 and is_watertight(f)
    (tmp_path / "b.msh").write_text(
        "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n4\n10 0 0 0\n20 1 0 0\n30 0 1 0\n40 0 0 1\n$EndNodes\n"
        "$Elements\n3\n1 15 2 0 1 10\n2 2 2 0 1 10 30 20\n3 4 2 0 1 10 20 30 40\n$EndElements\n")
    p, t, f = ingest.load_mesh(tmp_path / "b.msh")
    assert np.array_equal(t, [[0, 1, 2, 3]]) and np.array_equal(f, [[0, 2, 1]])  # point element skipped, ids remapped
    for text, msg in (("$MeshFormat\n4.0 0 8\n$EndMeshFormat\n", "2.x and 4.1"), ("$MeshFormat\n4.1 1 8\n$EndMeshFormat\n", "binary"),
                      ("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n", "no nodes"), ("$Nodes\n1\n1 0 0 0\n$EndNodes\n", "MeshFormat")):
        (tmp_path / "c.msh").write_text(text)
        with pytest.raises(SbError, match=msg):
            ingest.load_mesh(tmp_path / "c.msh")
    # MSH 4.1 (what current Gmsh writes): entity blocks -- a point entity, a parametric curve block (one extra coordinate
    # per node), a volume block; node tags first, coordinates after; point and line elements skipped; the same two tets as a.node
    (tmp_path / "e.msh").write_text(
        "$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$Entities\n0 0 0 0\n$EndEntities\n"
        "$Nodes\n3 5 10 50\n0 1 0 1\n10\n0 0 0\n1 7 1 2\n20\n30\n1 0 0 0.25\n0 1 0 0.75\n3 1 0 2\n40\n50\n0 0 1\n1 1 1\n$EndNodes\n"
        "$Elements\n4 6 1 6\n0 1 15 1\n1 10\n1 7 1 1\n2 20 30\n2 3 2 2\n3 10 30 20\n4 20 30 40\n3 1 4 2\n5 10 20 30 40\n6 50 20 30 40\n$EndElements\n")
    p4, t4, f4 = ingest.load_mesh(tmp_path / "e.msh")
    assert p4.shape == (5, 3)
    pa, ta, fa = ingest.load_mesh(tmp_path / "a.node")
    assert np.array_equal(p4, pa) and np.array_equal(np.sort(t4, 1), np.sort(ta, 1)) and tet_volumes(p4, t4).min() > 0
    assert np.array_equal(f4, [[0, 2, 1], [1, 2, 3]])           # the file's own triangles, ids remapped
    # what the library writes (2.2) and this 4.1 file describe the same body
    ingest.save_mesh(tmp_path / "e22.msh", p4, t4, f4)
    p22, t22, f22 = ingest.load_mesh(tmp_path / "e22.msh")
    assert np.array_equal(p22, p4) and np.array_equal(t22, t4) and np.array_equal(f22, f4)
    for text, msg in (("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$Nodes\n1 2 1 2\n3 1 0 2\n1\n2\n0 0 0\n$EndNodes\n", "malformed node"),
                      ("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$Nodes\n1 3 1 2\n3 1 0 2\n1\n2\n0 0 0\n1 0 0\n$EndNodes\n", "fewer nodes"),
                      ("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$Nodes\n1 2 1 2\n3 1 0 2\n1\n1\n0 0 0\n1 0 0\n$EndNodes\n", "duplicate"),
                      ("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$Nodes\n1 4 1 4\n3 1 0 4\n1\n2\n3\n4\n0 0 0\n1 0 0\n0 1 0\n0 0 1\n$EndNodes\n"
                       "$Elements\n1 1 1 1\n3 1 4 1\n1 1 2 3 9\n$EndElements\n", "unknown node"),
                      ("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$Nodes\n1 4 1 4\n3 1 0 4\n1\n2\n3\n4\n0 0 0\n1 0 0\n0 1 0\n0 0 1\n$EndNodes\n"
                       "$Elements\n1 2 1 2\n3 1 4 2\n1 1 2 3 4\n", "fewer elements")):
        (tmp_path / "g.msh").write_text(text)
        with pytest.raises(SbError, match=msg):
            ingest.load_mesh(tmp_path / "g.msh")
    (tmp_path / "d.node").write_text("2 3 0 0\n0 0 0 0\n1 1 0 0\n")
    (tmp_path / "d.ele").write_text("1 4 0\n0 0 1 2 3\n")
    with pytest.raises(SbError, match="unknown point"):
        ingest.load_mesh(tmp_path / "d.node")
    with pytest.raises(SbError, match="cannot open"):
        ingest.load_mesh(tmp_path / "nope.node")
    with pytest.raises(SbError, match="extension"):
        ingest.load_mesh(tmp_path / "x.obj")


def test_state_snapshot_round_trip_and_corruption(tmp_path):
    rng = np.random.default_rng(3)
    x4 = rng.normal(size=(1000, 4)).astype(np.float32)
    v4 = rng.normal(size=(1000, 4)).astype(np.float32)
    prm = default_params(substeps=7, iterations=3, damping=0.25, gravity=(0.1, -3.0, 0.2))
    path = tmp_path / "s.sbs"
    ingest.write_state(path, x4, v4, prm, frame=1234567890123, topo_hash=0xDEADBEEFCAFE)
    s = ingest.read_state(path)
    assert np.array_equal(s["x4"].view(np.uint32), x4.view(np.uint32)) and np.array_equal(s["v4"].view(np.uint32), v4.view(np.uint32))
    assert s["frame"] == 1234567890123 and s["topo_hash"] == 0xDEADBEEFCAFE
    assert (s["params"].substeps, s["params"].iterations, s["params"].damping) == (7, 3, 0.25)
    assert os.path.getsize(path) == 80 + 2 * 16 * 1000 + 8
    raw = bytearray(path.read_bytes())
    raw[80 + 4321] ^= 0x10  # one flipped bit in the payload
    (tmp_path / "bad.sbs").write_bytes(raw)
    with pytest.raises(SbError, match="checksum"):
        ingest.read_state(tmp_path / "bad.sbs")
    (tmp_path / "short.sbs").write_bytes(bytes(raw[:5000]))
    with pytest.raises(SbError, match="truncated"):
        ingest.read_state(tmp_path / "short.sbs")
    (tmp_path / "junk.sbs").write_bytes(b"not a snapshot at all" * 10)
    with pytest.raises(SbError, match="magic"):
        ingest.read_state(tmp_path / "junk.sbs")
    # the topology hash tells meshes apart
    pos, tets, _ = meshgen.block(4, 4, 4)
    h = ingest.topology_hash(len(pos), tets)
    other = tets.copy()
    other[5, [1, 2]] = other[5, [2, 1]]
    assert h == ingest.topology_hash(len(pos), tets.copy()) and h != ingest.topology_hash(len(pos), other)
    assert h != ingest.topology_hash(len(pos) + 1, tets)


# ---- render mesh -> tets ----------------------------------------------------------------------------------

def _bary_numpy(pos, tet, p):
    a = pos[tet].astype(np.float64)
    m = np.stack([a[1] - a[0], a[2] - a[0], a[3] - a[0]], 1)
    b = np.linalg.solve(m, p.astype(np.float64) - a[0])
    return np.array([1 - b.sum(), *b])


def test_skin_binding_encloses_and_reconstructs():
    pos, tets, tris = meshgen.sphere(12, spacing=0.1, jitter=0.1)
    rng = np.random.default_rng(5)
    c = pos.mean(0)
    inside = (c + rng.uniform(-0.25, 0.25, size=(600, 3))).astype(np.float32)  # the sphere's radius is 0.55
    tet_of, b = ingest.skin_binding(pos, tets, inside)
    assert b.min() >= -1e-6 and np.abs(b.sum(1) - 1).max() < 1e-6
    rec = np.einsum("nk,nkj->nj", b.astype(np.float64), pos[tets[tet_of]].astype(np.float64))
    assert np.abs(rec - inside).max() < 1e-6
    for i in range(0, 600, 37):  # independent restatement
        np.testing.assert_allclose(b[i], _bary_numpy(pos, tets[tet_of[i]], inside[i]), atol=2e-6)
    # points outside bind to a nearby tet with extrapolating weights; the affine reconstruction still holds
    r = np.linalg.norm(pos - c, axis=1).max()
    d = rng.normal(size=(200, 3))
    outside = (c + d / np.linalg.norm(d, axis=1)[:, None] * r * 1.15).astype(np.float32)
    tet_o, bo = ingest.skin_binding(pos, tets, outside)
    assert (bo.min(1) < 0).all() and np.abs(bo.sum(1) - 1).max() < 1e-5
    rec = np.einsum("nk,nkj->nj", bo.astype(np.float64), pos[tets[tet_o]].astype(np.float64))
    assert np.abs(rec - outside).max() < 1e-5
    cent = pos[tets[tet_o]].mean(1)
    assert np.linalg.norm(cent - outside, axis=1).max() < 0.45  # a tet close by, not an arbitrary one
    # far away: still an answer
    far_t, far_b = ingest.skin_binding(pos, tets, np.float32([[50, 50, 50]]))
    assert abs(far_b.sum() - 1) < 1e-3
    # surface vertices of the mesh itself: weight one on themselves (to rounding)
    sv = np.unique(tris)[:300]
    ts, bs = ingest.skin_binding(pos, tets, pos[sv])
    hit = (tets[ts] == sv[:, None])
    assert hit.any(1).all() and np.abs(bs[hit] - 1).max() < 1e-5


def test_skin_follows_affine_motion_in_the_oracle_and_binds_through_a_plan_handle():
    sp, st = uv_sphere(0.5, 16, 32, centre=(0, 1, 0))
    p, t, f = ingest.tetrahedralize_surface(sp, st, 0.09)
    sb = SoftBody(p, t, f, host_only=True, tile_cap=512)
    sb.skin_bind(sp, st)
    tet_of, b = sb.skin_binding()
    t2, b2 = ingest.skin_binding(p, t, sp)
    assert np.array_equal(tet_of, t2) and np.array_equal(b, b2)
    m = orc.Model(p, t)
    rest, n0 = m.skin(tet_of, b, st)
    assert np.abs(rest - sp).max() < 1e-6
    rad = sp - np.float32([0, 1, 0])
    assert ((n0 * rad).sum(1) / np.linalg.norm(rad, axis=1) > 0.95).all()  # outward normals of a sphere
    # an affine map of the tets maps the render mesh the same way
    A = np.array([[0.9, 0.2, 0.0], [-0.1, 1.1, 0.3], [0.0, -0.2, 0.8]])
    m.x4[:, :3] = (p.astype(np.float64) @ A.T + [0.3, -0.2, 0.1]).astype(np.float32)
    moved = m.skin(tet_of, b)
    assert np.abs(moved - (sp.astype(np.float64) @ A.T + [0.3, -0.2, 0.1])).max() < 2e-6
    with pytest.raises(SbError):
        sb.read_skinned()  # no device behind an sb_plan handle


def test_parsers_survive_garbage(tmp_path):
    """Truncated, mutated and random files give an error code, never a crash or an exception across the ABI."""
    pos, tets, tris = meshgen.block(4, 4, 4)
    good = {}
    for name in ("g.msh", "g.node"):
        ingest.save_mesh(tmp_path / name, pos, tets, tris)
    for f in ("g.msh", "g.node", "g.ele", "g.face"):
        good[f] = (tmp_path / f).read_bytes()
    ingest.write_state(tmp_path / "g.sbs", np.zeros((64, 4), np.float32), np.zeros((64, 4), np.float32), default_params())
    good["g.sbs"] = (tmp_path / "g.sbs").read_bytes()
    # the same mesh as Gmsh 4.1 writes it: two node blocks (the second parametric), a triangle block and a tet block
    half = len(pos) // 2
    lines = ["$MeshFormat", "4.1 0 8", "$EndMeshFormat", "$Nodes", f"2 {len(pos)} 1 {len(pos)}", f"3 1 0 {half}"]
    lines += [str(i + 1) for i in range(half)] + ["%.9g %.9g %.9g" % tuple(q) for q in pos[:half]]
    lines += [f"2 5 1 {len(pos) - half}"] + [str(i + 1) for i in range(half, len(pos))] + ["%.9g %.9g %.9g 0.5 0.25" % tuple(q) for q in pos[half:]]
    lines += ["$EndNodes", "$Elements", f"2 {len(tris) + len(tets)} 1 {len(tris) + len(tets)}", f"2 5 2 {len(tris)}"]
    lines += ["%d %d %d %d" % (i + 1, *(t + 1)) for i, t in enumerate(tris)] + [f"3 1 4 {len(tets)}"]
    lines += ["%d %d %d %d %d" % (len(tris) + i + 1, *(t + 1)) for i, t in enumerate(tets)] + ["$EndElements", ""]
    good["g41.msh"] = "\n".join(lines).encode()
    (tmp_path / "g41.msh").write_bytes(good["g41.msh"])
    p41, t41, f41 = ingest.load_mesh(tmp_path / "g41.msh")
    p22, t22, f22 = ingest.load_mesh(tmp_path / "g.msh")
    assert np.array_equal(p41, p22) and np.array_equal(t41, t22) and np.array_equal(f41, f22)
    rng = np.random.default_rng(11)
    outcomes = {"ok": 0, "error": 0}
    for trial in range(150):
        f = ("g.msh", "g.node", "g.ele", "g.sbs", "g41.msh")[trial % 5]
        raw = bytearray(good[f])
        kind = trial % 3
        if kind == 0:
            raw = raw[:int(rng.integers(0, len(raw)))]
        elif kind == 1:
            for _ in range(8):
                raw[int(rng.integers(0, len(raw)))] = int(rng.integers(32, 127))
        else:
            k = int(rng.integers(0, max(1, len(raw) - 40)))
            raw[k:k + 20] = b"99999999999999999999"  # absurd counts / ids
        # siblings of a TetGen set stay intact
        for g in ("g.node", "g.ele", "g.face"):
            (tmp_path / g.replace("g.", "t.")).write_bytes(good[g])
        target = tmp_path / f.replace("g.", "t.")
        target.write_bytes(bytes(raw))
        try:
            if f.endswith(".sbs"):
                ingest.read_state(target)
            else:
                ingest.load_mesh(target)
            outcomes["ok"] += 1
        except SbError:
            outcomes["error"] += 1
    assert outcomes["error"] > 50 and outcomes["ok"] + outcomes["error"] == 150


def test_softbody_from_surface_and_from_file(tmp_path):
    sp, st = uv_sphere(0.3, 12, 24, centre=(0, 0.5, 0))
    sb = SoftBody.from_surface(sp, st, 0.06, host_only=True, tile_cap=256)
    assert sb.n_render == len(sp) and sb.verify_streams() == 0
    tet_of, b = sb.skin_binding()
    assert np.abs(b.sum(1) - 1).max() < 1e-5 and b.min() > -0.6  # snapped boundary: the render vertices sit on or near it
    pos, tets, tris = ingest.tetrahedralize_surface(sp, st, 0.06, snap=True)
    ingest.save_mesh(tmp_path / "s.msh", pos, tets, tris)
    sb2 = SoftBody.from_file(tmp_path / "s.msh", host_only=True, tile_cap=256)
    assert (sb2.n_verts, sb2.n_tets) == (sb.n_verts, sb.n_tets)
    o1, f1 = sb.schedule()
    o2, f2 = sb2.schedule()
    assert np.array_equal(o1, o2) and np.array_equal(f1, f2)
