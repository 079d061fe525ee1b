"""Per-tile timeline of one tile pass (debug stamps): where a CTA's time goes."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from softbodyunity_b200 import SoftBody, meshgen, FLAG_FAST_MATH
ap = argparse.ArgumentParser()
for k in ("tile-cap", "block-threads", "round-width"):
    ap.add_argument("--" + k, type=int, default=0)
ap.add_argument("--fast-math", action="store_true")
ap.add_argument("--pass-index", type=int, default=1)
a = ap.parse_args()
pos, tets, tris = meshgen.block(100, spacing=0.01, origin=(0.0, 0.002, 0.0))
sb = SoftBody(pos, tets, tris, tile_cap=a.tile_cap, block_threads=a.block_threads, round_width=a.round_width,
              flags=FLAG_FAST_MATH if a.fast_math else 0)
sb.step(frames=3); sb.synchronize()
tr = sb.trace_pass(a.pass_index).astype(np.int64)
t0 = tr[:, 0].min()
for cta in (0, 1, 7, 31, 63):
    r = tr[cta]
    rounds = r[4:][r[4:] > 0]
    d = np.diff(np.concatenate([rounds, [r[2]]]))
    print(f"cta {cta}: start +{r[0]-t0} ns, prologue {r[1]-r[0]} ns, loop {r[2]-r[1]} ns ({len(rounds)} rounds, per round min/med/max {d.min()}/{int(np.median(d))}/{d.max()} ns), epilogue {r[3]-r[2]} ns, total {r[3]-r[0]} ns")
print("round durations cta 7:", np.diff(np.concatenate([tr[7][4:][tr[7][4:] > 0], [tr[7][2]]])).tolist())
print("pass time (kernel alone, ms):", sb.time_kernel(16 + a.pass_index, 20))

ct = sb.last_cta_trace.astype(np.int64)
ct = ct[ct[:, 0] > 0]
t0 = ct[:, 0].min()
st, en, sm = ct[:, 0] - t0, ct[:, 1] - t0, ct[:, 2]
print(f"all {len(ct)} CTAs: start min/median/p90/max {st.min()}/{int(np.median(st))}/{int(np.percentile(st,90))}/{st.max()} ns; end min/median/p90/max {en.min()}/{int(np.median(en))}/{int(np.percentile(en,90))}/{en.max()} ns; lifetime median/max {int(np.median(en-st))}/{(en-st).max()} ns")
per_sm = np.bincount(sm, minlength=148)
print("CTAs per SM min/max:", per_sm[per_sm > 0].min(), per_sm.max(), " SMs used:", (per_sm > 0).sum())
late = np.argsort(-en)[:8]
print("latest-ending CTAs (id, sm, start, end, ctas on that sm):", [(int(i), int(sm[i]), int(st[i]), int(en[i]), int(per_sm[sm[i]])) for i in late])
