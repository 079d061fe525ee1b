"""The N > 1 path on CPU: two gloo ranks shard a batch of independent bodies, plan their shards
(host-only handles) and reduce counters/timings the way bench.py does under torchrun."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from softbodyunity_b200 import SoftBody, meshgen
from softbodyunity_b200.shard import reduce_max, reduce_sum, shard_range


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_bodies, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_bodies, rank, world)
        pos, tets, tris = meshgen.bodies(hi - lo, dims=(5, 4, 4), seed=1234 + rank)
        sb = SoftBody(pos, tets, tris, host_only=True, tile_cap=256)
        info = sb.info()
        order, off = sb.schedule()
        ok = info["n_tile_passes"] == 1 and info["constraints_global"] == 0 and len(order) == info["n_edges"] + info["n_tets"]
        total_verts = reduce_sum(info["n_verts"])
        total_cons = reduce_sum(len(order))
        slowest = reduce_max(1.0 + rank)          # stands in for the per-rank elapsed time
        all_ok = reduce_sum(1.0 if ok else 0.0)
        out[rank] = (lo, hi, total_verts, total_cons, slowest, all_ok)
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_shard_plan_and_reduce():
    world, n_bodies = 2, 11
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_bodies, out), nprocs=world, join=True)
        res = dict(out)
    assert sorted(res) == [0, 1]
    (lo0, hi0, tv0, tc0, slow0, ok0), (lo1, hi1, tv1, tc1, slow1, ok1) = res[0], res[1]
    assert (lo0, hi0, lo1, hi1) == (0, 6, 6, 11)
    assert tv0 == tv1 == n_bodies * 80          # every body planned exactly once across the ranks
    one = SoftBody(*meshgen.bodies(1, dims=(5, 4, 4)), host_only=True)
    per_body = one.info()["n_edges"] + one.info()["n_tets"]
    assert tc0 == tc1 == n_bodies * per_body
    assert slow0 == slow1 == 2.0 and ok0 == ok1 == 2.0


def _dist_worker(rank, world, port, out):
    """One mesh over `world` ranks through peer memory (sb_dist_*): the host-side layout every rank derives from
    the SAME plan -- which vertices it owns, which tiles of every pass it runs -- reduced the way the ranks would."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pos, tets, tris = meshgen.block(14, 12, 26, spacing=0.05, origin=(0, 0.02, 0))
        sb = SoftBody(pos, tets, tris, host_only=True, tile_cap=256)
        info = sb.info()
        own = None
        covered = []
        for p in range(info["n_tile_passes"]):
            own, tiles = sb.dist_layout(rank, world, p)
            t = torch.from_numpy(tiles.astype(np.int32))
            dist.all_reduce(t)            # how many ranks run each tile of the pass
            tile_of, n_tiles = sb.tiles(p)
            busy = np.zeros(n_tiles, bool)
            busy[np.unique(tile_of[tile_of >= 0])] = True
            covered.append(bool(((t.numpy() == 1) | ~busy).all() and (t.numpy() <= 1).all() and tiles.any()))
        o = torch.from_numpy(own.astype(np.int32))
        dist.all_reduce(o)                # how many ranks own each vertex
        out[rank] = (int(own.sum()), bool((o.numpy() == 1).all()), all(covered), reduce_sum(float(own.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ranks_agree_on_the_peer_memory_layout(world):
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_dist_worker, args=(world, port, out), nprocs=world, join=True)
        res = [out[r] for r in range(world)]
    n_verts = 14 * 12 * 26
    sizes = [r[0] for r in res]
    assert sum(sizes) == n_verts and all(r[3] == n_verts for r in res)
    assert max(sizes) - min(sizes) <= 2 * 256          # slabs are whole tiles of the unshifted tiling
    assert all(r[1] for r in res), "every vertex has exactly one owner"
    assert all(r[2] for r in res), "every tile with constraints runs on exactly one rank, and every rank has work in every pass"
