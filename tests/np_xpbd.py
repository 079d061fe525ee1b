"""Independent numpy restatement of the XPBD substep (second opinion on the C oracle).

Written from the published formulas (Macklin et al. 2016; Mueller's soft-body demo),
vectorised per independent batch.  It does NOT follow the oracle's operation order or
FMA placement, so agreement is to rounding (not bitwise); that is the point: two
independently written implementations of the same equations.  PARITY UNPINNED with
respect to the upstream C# solver, which is not in the mount.
"""
import numpy as np


def compliance(k):
    if np.isinf(k) and k > 0:
        return 0.0
    return 1.0 / k if k > 0 else -1.0


def simulate(x, v, w, edges, rest_len, tets, rest_vol6, *, dt, substeps, iterations, k_d, k_v, damping, friction,
             gravity, ground_y, use_ground, batches, n_frames, spheres=(), dtype=np.float64):
    """batches: list of (kind, ids) with kind 'e' or 't'; ids vertex-disjoint within a batch."""
    x = x.astype(dtype).copy()
    v = v.astype(dtype).copy()
    w = w.astype(dtype)
    g = np.asarray(gravity, dtype)
    h = dtype(dt) / dtype(substeps)
    cd, cv = compliance(k_d), compliance(k_v)
    a_d = dtype(cd) / (h * h)
    a_v = dtype(cv) / (h * h)
    dyn = w > 0
    for _ in range(n_frames):
        for _ in range(substeps):
            v[dyn] += h * g
            xp = x.copy()
            x[dyn] += h * v[dyn]
            for _ in range(iterations):
                for kind, ids in batches:
                    if kind == "e" and cd >= 0:
                        a, b = edges[ids, 0], edges[ids, 1]
                        d = x[a] - x[b]
                        ln = np.sqrt((d * d).sum(1))
                        ws = w[a] + w[b]
                        ok = (ws > 0) & (ln > 0)
                        lam = np.zeros_like(ln)
                        lam[ok] = -(ln[ok] - rest_len[ids][ok]) / (ws[ok] + a_d)
                        n = np.zeros_like(d)
                        n[ok] = d[ok] / ln[ok, None]
                        x[a] += (w[a] * lam)[:, None] * n
                        x[b] -= (w[b] * lam)[:, None] * n
                    elif kind == "t" and cv >= 0:
                        q = tets[ids]
                        p0, p1, p2, p3 = x[q[:, 0]], x[q[:, 1]], x[q[:, 2]], x[q[:, 3]]
                        # gradients of V = det/6 with respect to each vertex
                        g1 = np.cross(p2 - p0, p3 - p0) / 6
                        g2 = np.cross(p3 - p0, p1 - p0) / 6
                        g3 = np.cross(p1 - p0, p2 - p0) / 6
                        g0 = -(g1 + g2 + g3)
                        vol = np.einsum("ij,ij->i", p1 - p0, np.cross(p2 - p0, p3 - p0)) / 6
                        den = sum(w[q[:, k]] * (gk * gk).sum(1) for k, gk in enumerate((g0, g1, g2, g3))) + a_v
                        ok = den > 0
                        lam = np.zeros_like(den)
                        lam[ok] = -(vol[ok] - rest_vol6[ids][ok] / 6) / den[ok]
                        for k, gk in enumerate((g0, g1, g2, g3)):
                            x[q[:, k]] += (w[q[:, k]] * lam)[:, None] * gk
            if use_ground:
                below = dyn & (x[:, 1] < ground_y)
                x[below, 1] = ground_y
                x[below, 0] = xp[below, 0] + (1 - friction) * (x[below, 0] - xp[below, 0])
                x[below, 2] = xp[below, 2] + (1 - friction) * (x[below, 2] - xp[below, 2])
            for s in spheres:
                c, r = np.asarray(s[:3], dtype), dtype(s[3])
                d = x - c
                l = np.sqrt((d * d).sum(1))
                inside = dyn & (l > 0) & (l < r)
                x[inside] = c + d[inside] * (r / l[inside])[:, None]
            v[dyn] = (x[dyn] - xp[dyn]) / h * max(0.0, 1 - h * damping)
    return x, v
