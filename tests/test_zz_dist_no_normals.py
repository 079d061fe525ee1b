"""One mesh over several ranks (sb_dist_*) WITHOUT a normals launch: the frame then ends with the closing handshake
(k_dist_sync) instead -- the program is one launch longer than the single-GPU one, the ranks do not deadlock or time
out, and state read back right after the frame is the oracle's.  (Virtual ranks on one stream cannot show the cross-GPU
race the handshake closes -- a read-back before a neighbour's last stores have landed -- only that the sequence is sound.)"""
import numpy as np
import pytest

from helpers import bits_equal, oracle_params
from oracle import xpbd_oracle as orc
from softbodyunity_b200 import FLAG_NO_NORMALS, SoftBody, meshgen
from softbodyunity_b200.dist import VirtualRanks


@pytest.mark.gpu
@pytest.mark.parametrize("n_ranks", [2, 4])
def test_a_distributed_frame_without_normals_ends_with_the_closing_handshake(n_ranks):
    import torch
    pos, tets, tris = meshgen.block(14, 12, 26, spacing=0.05, origin=(0, 0.02, 0))
    kw = dict(substeps=4, iterations=5, tile_cap=256, dist_ranks=n_ranks, flags=FLAG_NO_NORMALS)
    one = SoftBody(pos, tets, tris, **kw)
    stream = torch.cuda.Stream()
    vr = VirtualRanks(pos, tets, tris, n_ranks, stream.cuda_stream, **kw)
    assert len(vr.ranks[0].frame_program()) == len(one.frame_program()) + 1
    assert int(vr.ranks[0].frame_program()[-1][0]) == 6   # the normals slot of the program carries the handshake
    vr.step(frames=5)
    stream.synchronize()
    X, U = vr.gather_state()
    assert not any(sb.dist_error() for sb in vr.ranks)
    m = orc.Model(pos, tets, roles=vr.ranks[0].tet_roles())
    m.simulate(oracle_params(vr.ranks[0]), n_frames=5, threads=8, **vr.ranks[0].schedule_kw())
    assert bits_equal(X, m.x4) and bits_equal(U[:, :3], m.v4[:, :3])


@pytest.mark.gpu
def test_whole_mesh_reads_and_stray_launches_are_refused_on_one_rank_of_a_distributed_mesh():
    # a rank holds current values for the vertices it owns only, and a launch its peers do not run would part the epochs:
    # SB_E_STATE with a message, before any device work (sb_read_packed / sb_write_packed are the per-rank doors)
    pos, tets, tris = meshgen.block(14, 12, 26, spacing=0.05)
    sb = SoftBody(pos, tets, tris, tile_cap=256)
    sb.dist_setup(0, 2)
    for call in (sb.diagnostics, sb.read_surface, sb.normals, lambda: sb.save_state("/tmp/never_written.sbs"), lambda: sb.trace_pass(0),
                 lambda: sb.time_kernel(16, 2)):
        with pytest.raises(Exception, match="distributed"):
            call()
