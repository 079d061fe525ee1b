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


def quat_matrix(q):
    """Rotation matrix (box -> world) of the quaternion (x, y, z, w); zero quaternion = identity."""
    x, y, z, w = (float(c) for c in q)
    n = np.sqrt(x * x + y * y + z * z + w * w)
    if n == 0:
        return np.eye(3)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def collide(x, xp, kind, fr, p):
    """One collider on all points (sphere 0 / capsule 1 / box 2): textbook closest-point push-out, then the share
    `fr` of the tangential motion since xp is removed.  Returns the new positions."""
    x = x.copy()
    if kind in (0, "sphere", 1, "capsule"):
        r = p[3]
        if kind in (0, "sphere"):
            c = np.broadcast_to(p[:3], x.shape)
        else:
            a, b = p[:3], p[4:7]
            ab = b - a
            l2 = ab @ ab
            t = np.clip(((x - a) @ ab) / l2, 0, 1) if l2 > 0 else np.zeros(len(x), x.dtype)
            c = a + t[:, None] * ab
        d = x - c
        l = np.sqrt((d * d).sum(1))
        hit = (l > 0) & (l < r)
        n = np.zeros_like(x)
        n[hit] = d[hit] / l[hit][:, None]
        x[hit] = c[hit] + n[hit] * r
    else:
        c, half, R = p[:3], p[3:6], quat_matrix(p[6:10]).astype(x.dtype)
        loc = (x - c) @ R  # components along the box axes (columns of R)
        pen = half - np.abs(loc)
        hit = (pen > 0).all(1)
        k = np.argmin(pen, axis=1)
        n = R.T[k]
        sgn = np.where(loc[np.arange(len(x)), k] >= 0, 1.0, -1.0)
        x[hit] += (sgn * pen[np.arange(len(x)), k])[hit][:, None] * n[hit]
    if fr > 0:
        m = x - xp
        mt = m - (m * n).sum(1)[:, None] * n
        x[hit] -= fr * mt[hit]
    return x


def simulate(x, v, w, edges, rest_len, tets, rest_vol6, *, dt, substeps, iterations, k_d, k_v, damping, friction,
             gravity, ground_y, use_ground, batches, n_frames, spheres=(), colliders=(), dtype=np.float64):
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
            for kind, fr, *p in colliders:
                x[dyn] = collide(x[dyn], xp[dyn], kind, fr, np.asarray(p, dtype).ravel())
            v[dyn] = (x[dyn] - xp[dyn]) / h * max(0.0, 1 - h * damping)
    return x, v
