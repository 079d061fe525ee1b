"""debug: which of (single GPU with the dist plan, virtual ranks) departs from the oracle"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import bits_equal, oracle_params
from oracle import xpbd_oracle as orc
from softbodyunity_b200 import SoftBody, meshgen
from softbodyunity_b200.dist import VirtualRanks
pos, tets, tris = meshgen.block(14, 12, 26, spacing=0.05, origin=(0, 0.02, 0))
for n_ranks, extra in ((2, dict(tile_cap=256)), (2, dict(tile_cap=256, flags=128)), (2, dict(tile_cap=256, flags=16)), (4, dict(tile_cap=300, flags=16))):
    kw = dict(dict(substeps=5, iterations=6, dist_ranks=n_ranks), **extra)
    for frames in (1, 6):
        one = SoftBody(pos, tets, tris, **kw)
        one.step(frames=frames)
        x1, v1 = one.get_state()
        m = orc.Model(pos, tets, roles=one.tet_roles())
        m.simulate(oracle_params(one), n_frames=frames, threads=8, **one.schedule_kw())
        stream = torch.cuda.Stream()
        vr = VirtualRanks(pos, tets, tris, n_ranks, stream.cuda_stream, **kw)
        vr.step(frames=frames)
        stream.synchronize()
        X, U = vr.gather_state()
        bad = np.nonzero((X.view(np.uint32) != m.x4.view(np.uint32)).any(1))[0]
        print(n_ranks, extra, "frames", frames, "one==oracle", bits_equal(x1, m.x4), "vr==oracle", bits_equal(X, m.x4), "vr==one", bits_equal(X, x1),
              "n_bad", len(bad), "tiles", vr.tiles, "info", one.info()["tiles_in_pass"][:5], flush=True)
        if len(bad):
            own = np.stack(vr.owned)
            print("   bad verts owned by ranks:", [int(own[r][bad].sum()) for r in range(n_ranks)], "first bad", bad[:8], "z of bad", np.round(pos[bad[:8], 2], 2))
