"""One mesh over several ranks through peer memory (sb_dist_*): virtual ranks on ONE GPU run the same kernels, peer
pointers and epoch words as the multi-GPU path, and must reproduce the single-handle run -- and therefore the oracle
-- bit for bit, because the execution order IS the single-GPU order.  (Real GPUs: tests/test_multigpu.py.)"""
import numpy as np
import pytest

from helpers import bits_equal, oracle_params
from oracle import xpbd_oracle as orc
from softbodyunity_b200 import SoftBody, meshgen
from softbodyunity_b200.dist import VirtualRanks


@pytest.mark.gpu
@pytest.mark.parametrize("n_ranks,kw", [(2, dict(tile_cap=256)), (3, dict(tile_cap=200, block_threads=32)),
                                         (4, dict(tile_cap=300, flags=16)), (8, dict(tile_cap=300, iterations=5)),
                                         (2, dict(tile_cap=256, flags=128)), (4, dict(tile_cap=256, dist_ranks=0)),
                                         (2, dict(tile_cap=300, dist_ranks=8)), (4, dict(tile_cap=300, dist_ranks=8))])  # bench.py: the 8-way plan on 2 / 4 GPUs
def test_virtual_ranks_over_peer_memory_match_one_gpu_and_the_oracle(n_ranks, kw):
    import torch
    # (eight ranks: a mesh with room for eight blocks of boxes -- four tilings and one leftover pass)
    pos, tets, tris = meshgen.block(*((18, 18, 36) if 8 in (n_ranks, kw.get("dist_ranks")) else (14, 12, 26)), spacing=0.05, origin=(0, 0.02, 0))
    kw = dict(dict(substeps=5, iterations=6, dist_ranks=n_ranks), **kw)
    # one GPU, the same plan (dist_ranks is a planner hint: box grid rounded to the cuts, boxes numbered block by block)
    one = SoftBody(pos, tets, tris, **kw)
    one.step(frames=6)
    x1, v1 = one.get_state()
    stream = torch.cuda.Stream()
    vr = VirtualRanks(pos, tets, tris, n_ranks, stream.cuda_stream, **kw)
    owned = np.stack(vr.owned)
    assert (owned.sum(0) == 1).all(), "every vertex has exactly one owner"
    info = vr.ranks[0].info()
    for k in range(info["n_tile_passes"]):  # a tile of a pass runs on exactly one rank (pass 0: every tile runs)
        assert sum(t[k] for t in vr.tiles) <= info["tiles_in_pass"][k]
    assert sum(t[0] for t in vr.tiles) == info["tiles_in_pass"][0]
    vr.step(frames=6)
    stream.synchronize()
    X, U, N = vr.gather_state(with_surface=True)
    assert not any(sb.dist_error() for sb in vr.ranks)
    assert bits_equal(X, x1) and bits_equal(U[:, :3], v1[:, :3])
    m = orc.Model(pos, tets, roles=vr.ranks[0].tet_roles())
    m.simulate(oracle_params(one), n_frames=6, threads=8, **vr.ranks[0].schedule_kw())
    assert m.x4[:, 1].min() == 0.0
    assert bits_equal(X, m.x4) and bits_equal(U[:, :3], m.v4[:, :3])
    ids = one.surface_vertices()
    assert bits_equal(N[ids], m.normals(tris)[ids])


@pytest.mark.gpu
@pytest.mark.parametrize("n_ranks", [2, 3, 4])
def test_an_ingested_body_is_cut_by_recursive_bisection_and_matches_the_oracle(n_ranks):
    # not a lattice block: a torus surface -> tets (boundary snapped onto the surface), cut into n_ranks compact parts by
    # recursive coordinate bisection of the boxes' centroids weighted by their vertex counts (the partitioner of this path)
    import torch
    from softbodyunity_b200 import ingest
    from test_ingest import torus
    sp, st = torus(0.5, 0.2, 48, 24)
    sp = sp + np.float32([0.0, 0.25, 0.0])  # lying flat, 5 cm above the ground
    pos, tets, tris = ingest.tetrahedralize_surface(sp, st, 0.035, snap=True)
    assert len(pos) > 8000
    kw = dict(substeps=4, iterations=5, tile_cap=400, dist_ranks=n_ranks, stiffness=3e5)  # (four tilings and one leftover pass)
    stream = torch.cuda.Stream()
    vr = VirtualRanks(pos, tets, tris, n_ranks, stream.cuda_stream, **kw)
    owned = np.stack(vr.owned)
    assert (owned.sum(0) == 1).all()
    share = owned.sum(1) / len(pos)
    assert share.min() > 0.6 / n_ranks and share.max() < 1.5 / n_ranks, share  # balanced by vertex count
    vr.step(frames=5)
    stream.synchronize()
    X, U, N = vr.gather_state(with_surface=True)
    assert not any(sb.dist_error() for sb in vr.ranks)
    m = orc.Model(pos, tets, roles=vr.ranks[0].tet_roles())
    m.simulate(oracle_params(vr.ranks[0]), n_frames=5, threads=8, **vr.ranks[0].schedule_kw())
    assert bits_equal(X, m.x4) and bits_equal(U[:, :3], m.v4[:, :3])
    ids = vr.ranks[0].surface_vertices()
    assert bits_equal(N[ids], m.normals(tris)[ids])


@pytest.mark.gpu
def test_dist_setup_rejects_what_it_cannot_split():
    pos, tets, tris = meshgen.sample_cube(6)
    sb = SoftBody(pos, tets, tris)  # one tile: not a tiled mesh
    with pytest.raises(Exception):
        sb.dist_setup(0, 2)
    pos, tets, tris = meshgen.block(14, 12, 26, spacing=0.05)
    sb = SoftBody(pos, tets, tris, tile_cap=256)
    sb.dist_setup(0, 2)
    with pytest.raises(Exception):
        sb.time_kernel(16, 2)  # a launch the peers do not run would part the epochs


@pytest.mark.gpu
def test_packed_frame_is_the_separate_read_backs_in_one_copy():
    pos, tets, tris = meshgen.sphere(12, spacing=0.05)
    sb = SoftBody(pos, tets, tris, tile_cap=400)
    sb.step(frames=3)
    n, ns, b_in, b_out = sb.packed_sizes()
    assert (n, ns, b_in, b_out) == (len(pos), sb.n_surface, 32 * len(pos), 32 * len(pos) + 24 * sb.n_surface)
    buf = sb.read_packed()
    x4, v4, sp, sn = sb.unpack_frame(buf)
    X, U = sb.get_state()
    SP, SN = sb.read_surface()
    assert bits_equal(x4, X) and bits_equal(v4, U) and bits_equal(sp, SP) and bits_equal(sn, SN)
    # state in through the same door: a second body continues from the packed state bit for bit
    other = SoftBody(pos, tets, tris, tile_cap=400)
    other.write_packed(buf)
    sb.step(frames=2)
    other.step(frames=2)
    assert bits_equal(sb.get_state()[0], other.get_state()[0]) and bits_equal(sb.get_state()[1], other.get_state()[1])
