"""One mesh over several ranks through peer memory (sb_dist_*): virtual ranks on ONE GPU run the same kernels, peer
pointers and epoch words as the multi-GPU path, and must reproduce the single-handle run -- and therefore the oracle
-- bit for bit, because the execution order IS the single-GPU order."""
import numpy as np
import pytest

from helpers import bits_equal, oracle_params
from oracle import xpbd_oracle as orc
from softbodyunity_b200 import SoftBody, meshgen
from softbodyunity_b200.dist import VirtualRanks


@pytest.mark.gpu
@pytest.mark.parametrize("n_ranks,kw", [(2, dict(tile_cap=256)), (3, dict(tile_cap=200, block_threads=32)),
                                         (4, dict(tile_cap=300, flags=16))])
def test_virtual_ranks_over_peer_memory_match_one_gpu_and_the_oracle(n_ranks, kw):
    import torch
    pos, tets, tris = meshgen.block(14, 12, 26, spacing=0.05, origin=(0, 0.02, 0))
    one = SoftBody(pos, tets, tris, substeps=5, iterations=6, **kw)
    one.step(frames=6)
    x1, v1 = one.get_state()
    stream = torch.cuda.Stream()
    vr = VirtualRanks(pos, tets, tris, n_ranks, stream.cuda_stream, substeps=5, iterations=6, **kw)
    owned = np.stack(vr.owned)
    assert (owned.sum(0) == 1).all(), "every vertex has exactly one owner"
    info = one.info()
    for k in range(info["n_tile_passes"]):  # every non-empty tile of every pass runs on exactly one rank
        assert sum(t[k] for t in vr.tiles) <= info["tiles_in_pass"][k]
        assert all(t[k] > 0 for t in vr.tiles)
    vr.step(frames=6)
    stream.synchronize()
    X, U = vr.gather_state()
    assert not any(sb.dist_error() for sb in vr.ranks)
    assert bits_equal(X, x1) and bits_equal(U[:, :3], v1[:, :3])
    order, off = one.schedule()
    m = orc.Model(pos, tets, roles=one.tet_roles())
    m.simulate(oracle_params(one), n_frames=6, order=order, batch_off=off, threads=8)
    assert m.x4[:, 1].min() == 0.0
    assert bits_equal(X, m.x4)


@pytest.mark.gpu
def test_dist_setup_rejects_what_it_cannot_split():
    pos, tets, tris = meshgen.sample_cube(6)
    sb = SoftBody(pos, tets, tris)  # one tile: not a tiled mesh
    with pytest.raises(Exception):
        sb.dist_setup(0, 2)
