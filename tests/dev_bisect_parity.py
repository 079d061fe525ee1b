"""Debug: where does the GPU state first leave the oracle's?  (1 substep x 1 iteration, by constraint kind)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import oracle_params
from oracle import xpbd_oracle as orc
from softbodyunity_b200 import SoftBody, meshgen
cases = {
 "bt256": (lambda: meshgen.block(12, 12, 10, spacing=0.05, origin=(0, 0.02, 0)), dict(tile_cap=700, block_threads=256)),
 "bt128": (lambda: meshgen.block(12, 12, 10, spacing=0.05, origin=(0, 0.02, 0)), dict(tile_cap=700, block_threads=128)),
 "bodies": (lambda: meshgen.bodies(40, dims=(6, 5, 5), spacing=0.04, base_height=0.03), dict(tile_cap=512)),
 "normals": (lambda: meshgen.sphere(14, spacing=0.05), dict(tile_cap=600)),
 "wide": (lambda: meshgen.block(12, 12, 10, spacing=0.05, origin=(0, 0.02, 0)), dict(tile_cap=1500, round_width=2)),
}
for name, (gen, kw) in cases.items():
    pos, tets, tris = gen()
    for label, extra in (("both", {}), ("edges only", dict(volume_stiffness=0.0)), ("tets only", dict(stiffness=0.0)),
                         ("no graph", dict(flags=4)), ("2 iters", dict(iterations=2))):
        k = dict(substeps=1, iterations=1)
        k.update(kw); k.update(extra)
        sb = SoftBody(pos, tets, tris, **k)
        sched = sb.schedule_kw()
        m = orc.Model(pos, tets, roles=sb.tet_roles())
        sb.step(frames=1)
        x4, v4 = sb.get_state()
        m.simulate(oracle_params(sb), n_frames=1, threads=1, **sched)
        badv = np.nonzero((x4.view(np.uint32) != m.x4.view(np.uint32)).any(1))[0]
        msg = f"{name:8s} {label:11s}: {len(badv)} of {len(pos)} vertices differ, max |dx| {np.abs(x4 - m.x4).max():.3e}"
        if len(badv):
            info = sb.info()
            for p in range(info["n_tile_passes"]):
                tile_of, nt = sb.tiles(p)
                tb = np.unique(tile_of[badv])
                sizes = np.bincount(tile_of[tile_of >= 0], minlength=nt)
                msg += f" | pass {p}: bad tiles {tb[:8].tolist()} of {nt} sizes {sizes[tb[tb >= 0]][:8].tolist()}"
            msg += f" | first bad verts {badv[:6].tolist()}"
        print(msg, flush=True)
        sb.close()
