"""One mesh over several ranks (BASELINE.json configs[4]): slab partition, ghost vertices, the two
constraint groups and the halo exchange.  CPU tests check the host logic; the GPU tests run all
ranks as "virtual ranks" on one device (same orchestration code, device-local copies standing in
for NCCL) and demand the SAME BITS as the CPU oracle replaying the combined order."""
import numpy as np
import pytest

from helpers import bits_equal, oracle_params
from oracle import xpbd_oracle as orc
from softbodyunity_b200 import SoftBody, lumped_inv_mass, meshgen
from softbodyunity_b200.partition import (LocalComm, LocalPeerComm, PartitionedBody, combined_order, combined_roles, gather_global, slab_partition,
                                          step_partitioned)


def make(n_ranks, dims=(10, 33, 9)):
    pos, tets, tris = meshgen.block(*dims, spacing=0.05, origin=(0.0, 0.02, 0.0), seed=11)
    return pos, tets, tris, slab_partition(pos, tets, tris, n_ranks)


def test_lumped_mass_helper_matches_the_oracle():
    pos, tets, _ = meshgen.block(6, 7, 5, spacing=0.1)
    assert np.array_equal(lumped_inv_mass(pos, tets, 900.0).view(np.uint32), orc.lumped_inv_mass(pos, tets, 900.0).view(np.uint32))


@pytest.mark.parametrize("n_ranks", [2, 3, 4])
def test_partition_and_combined_order_are_valid(n_ranks):
    pos, tets, tris, meshes = make(n_ranks)
    V = len(pos)
    owned = np.concatenate([m.own for m in meshes])
    assert np.array_equal(np.sort(owned), np.arange(V)), "every vertex has exactly one owner"
    sizes = [m.n_own for m in meshes]
    assert max(sizes) - min(sizes) <= 1
    for m in meshes:
        assert (m.n_ghost > 0) == (m.rank < n_ranks - 1) and (len(m.lower) > 0) == (m.rank > 0)
        if m.rank + 1 < n_ranks:  # my ghosts are exactly the next rank's lower-boundary vertices
            assert np.array_equal(m.ghost, meshes[m.rank + 1].lower)
    plans = [SoftBody(m.pos, m.tets, m.tris, inv_mass=m.inv_mass, edges=m.edges, n_ghost_verts=m.n_ghost, host_only=True, tile_cap=256) for m in meshes]
    ref = orc.Model(pos, tets, roles=combined_roles(meshes, plans, tets))
    order, off = combined_order(meshes, plans, ref.edges)
    ids, kind = order & 0x7fffffff, order < 0
    assert np.array_equal(np.sort(ids[~kind]), np.arange(ref.E)) and np.array_equal(np.sort(ids[kind]), np.arange(ref.T))
    for b in range(len(off) - 1):  # every batch is an independent set in GLOBAL numbering
        i, k = ids[off[b]:off[b + 1]], kind[off[b]:off[b + 1]]
        verts = (ref.tets[i] if k[0] else ref.edges[i]).reshape(-1)
        assert len(np.unique(verts)) == len(verts)
    # group 0 never touches a ghost; group 1 always does
    for m, p in zip(meshes, plans):
        o, _ = p.schedule()
        n1 = p.info()["constraints_cut"]
        le = p.topology()[0]

        def touches_ghost(ent):
            vs = m.tets[ent & 0x7fffffff] if ent < 0 else le[ent]
            return (vs >= m.n_own).any()
        assert not any(touches_ghost(e) for e in o[:len(o) - n1])
        assert all(touches_ghost(e) for e in o[len(o) - n1:])
    ref.simulate(orc.params(substeps=4, iterations=4), n_frames=5, order=order, batch_off=off)
    assert np.isfinite(ref.x4).all() and ref.x4[:, 1].min() >= 0.0


def test_too_many_ranks_is_rejected():
    pos, tets, tris = meshgen.block(4, 6, 4, spacing=0.1)
    with pytest.raises(ValueError):
        slab_partition(pos, tets, tris, 5)


@pytest.mark.gpu
@pytest.mark.parametrize("n_ranks,peer", [(2, False), (3, False), (2, True), (4, True)])
def test_virtual_ranks_match_the_oracle_bitwise(n_ranks, peer):
    # peer=False: pack / device copy / unpack (the NCCL transport's shape); peer=True: the peer-memory
    # send / receive kernels of the multi-GPU path (stores into the neighbour's buffer + sequence flags)
    pos, tets, tris, meshes = make(n_ranks)
    import torch
    stream = torch.cuda.Stream()
    bodies = [PartitionedBody(m, stream=stream, tile_cap=256, substeps=5, iterations=6) for m in meshes]
    comm = LocalPeerComm(bodies) if peer else LocalComm(n_ranks)
    step_partitioned(bodies, comm, frames=8)
    stream.synchronize()
    states = [b.sb.get_state() for b in bodies]
    x4 = gather_global(meshes, [s[0] for s in states])
    v4 = gather_global(meshes, [s[1] for s in states])
    ref = orc.Model(pos, tets, roles=combined_roles(meshes, [b.sb for b in bodies], tets))
    order, off = combined_order(meshes, [b.sb for b in bodies], ref.edges)
    ref.simulate(oracle_params(bodies[0].sb), n_frames=8, order=order, batch_off=off, threads=8)
    assert ref.x4[:, 1].min() == 0.0
    assert bits_equal(x4, ref.x4)
    assert bits_equal(v4[:, :3], ref.v4[:, :3])
    # ghosts hold their owner's final value only after exchange A of the next sweep; owners are authoritative
    for m, s in zip(meshes, states):
        assert s[1][m.n_own:, 3].min(initial=1.0) == 1.0  # ghost flag kept in v.w
    assert not any(b.sb.halo_error() for b in bodies) if peer else True


@pytest.mark.gpu
def test_partitioned_equals_single_handle_phased_run():
    # one rank, no ghosts: the phased API is the same sequence as sb_step
    pos, tets, tris = meshgen.block(9, 9, 8, spacing=0.05, origin=(0, 0.02, 0))
    a = SoftBody(pos, tets, tris, tile_cap=300, flags=64)  # (the phased API sweeps the passes forwards in every iteration: no snake)
    a.step(frames=4)
    (m,) = slab_partition(pos, tets, tris, 1)
    b = PartitionedBody(m, tile_cap=300)
    step_partitioned([b], LocalComm(1), frames=4)
    xa, va = a.get_state()
    xb, vb = b.sb.get_state()
    assert bits_equal(xa, xb) and bits_equal(va, vb)
