"""debug: which of (single GPU with the dist plan, virtual ranks) departs from the oracle"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from helpers import bits_equal, oracle_params
from oracle import xpbd_oracle as orc
from softbodyunity_b200 import SoftBody, meshgen, ingest
from softbodyunity_b200.dist import VirtualRanks
from test_ingest import torus
sp, st = torus(0.5, 0.2, 48, 24); sp = sp + np.float32([0.0, 0.25, 0.0])
tor = ingest.tetrahedralize_surface(sp, st, 0.035, snap=True)
tor_nosnap = ingest.tetrahedralize_surface(sp, st, 0.035, snap=False)
blk = meshgen.block(14, 12, 26, spacing=0.05, origin=(0, 0.02, 0))
cases = [("torus", tor, 4, dict(substeps=4, iterations=5, tile_cap=400, stiffness=3e5)),
         ("torus I=4", tor, 4, dict(substeps=4, iterations=4, tile_cap=400, stiffness=3e5)),
         ("torus nofuse", tor, 4, dict(substeps=4, iterations=5, tile_cap=400, stiffness=3e5, flags=128)),
         ("torus nosnake", tor, 4, dict(substeps=4, iterations=5, tile_cap=400, stiffness=3e5, flags=64)),
         ("torus nosnap", tor_nosnap, 4, dict(substeps=4, iterations=5, tile_cap=400, stiffness=3e5)),
         ("torus rigid", tor, 4, dict(substeps=4, iterations=5, tile_cap=400)),
         ("block I=5", blk, 4, dict(substeps=4, iterations=5, tile_cap=300)),
         ("block I=5 r2", blk, 2, dict(substeps=4, iterations=5, tile_cap=256))]
for name, (pos, tets, tris), n_ranks, kw in cases:
    kw = dict(kw, dist_ranks=n_ranks)
    for frames in (1, 5):
        stream = torch.cuda.Stream()
        vr = VirtualRanks(pos, tets, tris, n_ranks, stream.cuda_stream, **kw)
        vr.step(frames=frames)
        stream.synchronize()
        X, U = vr.gather_state()
        one = vr.ranks[0]
        m = orc.Model(pos, tets, roles=one.tet_roles())
        m.simulate(oracle_params(one), n_frames=frames, threads=8, **one.schedule_kw())
        bad = np.nonzero((X.view(np.uint32) != m.x4.view(np.uint32)).any(1))[0]
        i = one.info()
        print(name, n_ranks, "frames", frames, "vr==oracle", len(bad) == 0, "n_bad", len(bad), "of", len(pos), "passes", i["n_tile_passes"], i["tiles_in_pass"][:5], "tiles", vr.tiles,
              "err", [sb.dist_error() for sb in vr.ranks], flush=True)
        if len(bad):
            own = np.stack(vr.owned)
            print("   bad verts owned by ranks:", [int(own[r][bad].sum()) for r in range(n_ranks)], "max |dx|", float(np.abs(X[bad, :3] - m.x4[bad, :3]).max()))
