"""Synthetic tet meshes for the configurations BASELINE.json names.

The reference ships no sample assets in the mount (/root/reference/README.md:1 is the
whole reference), so the sample cube, the sphere, the 1 M block, the 4096-body batch
and the 8 M block are all generated here, deterministically from (shape, n, seed).
Lattice cells are split into 5 tets with alternating parity so faces match.
"""
import numpy as np

# local corner (a,b,c) -> bit index a + 2b + 4c
_A = np.array([[0, 3, 5, 6], [1, 0, 3, 5], [2, 0, 6, 3], [4, 0, 5, 6], [7, 3, 6, 5]])  # even cells
_B = np.array([[1, 2, 4, 7], [0, 1, 2, 4], [3, 1, 7, 2], [5, 1, 4, 7], [6, 2, 7, 4]])  # odd cells


def _orient(pos, tets):
    p = pos[tets]
    e1, e2, e3 = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0], p[:, 3] - p[:, 0]
    det = np.einsum("ij,ij->i", e1, np.cross(e2, e3))
    flip = det < 0
    tets[flip, 2], tets[flip, 3] = tets[flip, 3].copy(), tets[flip, 2].copy()
    return tets


def lattice_tets(nx, ny, nz, cell_mask=None):
    """5-tet split of the (nx-1)(ny-1)(nz-1) cells of an nx*ny*nz vertex lattice
    (vertex id = i + nx*(j + ny*k)).  Returns int32 (T,4), cell-major."""
    ci, cj, ck = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), np.arange(nz - 1), indexing="ij")
    ci, cj, ck = (a.transpose(2, 1, 0).reshape(-1) for a in (ci, cj, ck))  # x fastest
    if cell_mask is not None:
        keep = cell_mask(ci, cj, ck)
        ci, cj, ck = ci[keep], cj[keep], ck[keep]
    corner = np.empty((ci.size, 8), np.int64)
    for b in range(8):
        a, bb, c = b & 1, (b >> 1) & 1, (b >> 2) & 1
        corner[:, b] = (ci + a) + nx * ((cj + bb) + ny * (ck + c))
    odd = ((ci + cj + ck) & 1).astype(bool)
    tets = np.empty((ci.size, 5, 4), np.int64)
    tets[~odd] = corner[~odd][:, _A]
    tets[odd] = corner[odd][:, _B]
    return tets.reshape(-1, 4).astype(np.int32)


def lattice_positions(nx, ny, nz, spacing, origin=(0.0, 0.0, 0.0), jitter=0.1, seed=1234):
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    pos = np.stack([i, j, k], -1).reshape(-1, 3).astype(np.float64) * spacing
    if jitter > 0:
        rng = np.random.default_rng(seed)
        pos += rng.uniform(-jitter, jitter, pos.shape) * spacing
    pos += np.asarray(origin, np.float64)
    return pos.astype(np.float32)


def surface_triangles(tets, n_verts):
    """Faces that belong to exactly one tet, wound outward for positively oriented tets."""
    t = np.asarray(tets, np.int64)
    faces = np.concatenate([t[:, [0, 2, 1]], t[:, [0, 1, 3]], t[:, [0, 3, 2]], t[:, [1, 2, 3]]])
    s = np.sort(faces, axis=1)
    if n_verts < (1 << 21):
        key = (s[:, 0] << 42) | (s[:, 1] << 21) | s[:, 2]
        _, idx, cnt = np.unique(key, return_index=True, return_counts=True)
    else:
        _, idx, cnt = np.unique(s, axis=0, return_index=True, return_counts=True)
    keep = np.sort(idx[cnt == 1])
    return faces[keep].astype(np.int32)


def _lattice_surface(tets, nx, ny, nz):
    """Surface of a full lattice block: faces whose three vertices share a boundary plane."""
    t = np.asarray(tets, np.int64)
    vid = np.arange(nx * ny * nz)
    i, j, k = vid % nx, (vid // nx) % ny, vid // (nx * ny)
    m = ((i == 0) * 1 | (i == nx - 1) * 2 | (j == 0) * 4 | (j == ny - 1) * 8 | (k == 0) * 16 | (k == nz - 1) * 32).astype(np.uint8)
    out = []
    for f in ([0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]):
        fv = t[:, f]
        on = (m[fv[:, 0]] & m[fv[:, 1]] & m[fv[:, 2]]) != 0
        out.append(fv[on])
    return np.concatenate(out).astype(np.int32)


def block(nx, ny=None, nz=None, spacing=0.01, origin=(0.0, 0.05, 0.0), jitter=0.1, seed=1234):
    """Rectangular block of nx*ny*nz vertices.  Returns (pos f32 (V,3), tets i32 (T,4), tris i32 (F,3))."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    pos = lattice_positions(nx, ny, nz, spacing, origin, jitter, seed)
    tets = _orient(lattice_positions(nx, ny, nz, 1.0, jitter=0.0).astype(np.float64), lattice_tets(nx, ny, nz))
    tris = _lattice_surface(tets, nx, ny, nz)
    return pos, tets, tris


def sphere(n=58, spacing=0.01, centre_height=None, jitter=0.1, seed=1234):
    """Block of n^3 vertices clipped to the inscribed sphere by cell centre; unused vertices dropped."""
    r = (n - 1) / 2.0

    def mask(ci, cj, ck):
        return (ci + 0.5 - r) ** 2 + (cj + 0.5 - r) ** 2 + (ck + 0.5 - r) ** 2 <= r * r

    tets = lattice_tets(n, n, n, mask)
    used = np.unique(tets)
    remap = -np.ones(n ** 3, np.int64)
    remap[used] = np.arange(used.size)
    h = r * spacing + 0.05 if centre_height is None else centre_height
    pos = lattice_positions(n, n, n, spacing, (-r * spacing, h - r * spacing, -r * spacing), jitter, seed)[used]
    flat = lattice_positions(n, n, n, 1.0, jitter=0.0)[used].astype(np.float64)
    tets = _orient(flat, remap[tets].astype(np.int32))
    return pos, tets, surface_triangles(tets, used.size)


def bodies(count, dims=(13, 13, 12), spacing=0.02, gap=0.1, jitter=0.1, seed=1234, base_height=0.05):
    """`count` independent blocks laid out on a square grid in x/z (one disconnected mesh)."""
    nx, ny, nz = dims
    p0, t0, s0 = block(nx, ny, nz, spacing, (0.0, 0.0, 0.0), 0.0, seed)
    nv = p0.shape[0]
    side = int(np.ceil(np.sqrt(count)))
    b = np.arange(count)
    off = np.stack([(b % side) * (nx * spacing + gap), np.full(count, base_height), (b // side) * (nz * spacing + gap)], -1)
    pos = (p0[None].astype(np.float64) + off[:, None, :])
    if jitter > 0:
        rng = np.random.default_rng(seed)
        pos += rng.uniform(-jitter, jitter, pos.shape) * spacing
    pos = pos.reshape(-1, 3).astype(np.float32)
    shift = (b * nv)[:, None, None]
    tets = (t0[None].astype(np.int64) + shift).reshape(-1, 4).astype(np.int32)
    tris = (s0[None].astype(np.int64) + shift).reshape(-1, 3).astype(np.int32)
    return pos, tets, tris


def sample_cube(n=10, side=1.0, centre_height=2.0, jitter=0.0, seed=1234):
    """Config 1 substitute: soft cube of n^3 vertices, centre 2 m above y = 0."""
    sp = side / (n - 1)
    return block(n, n, n, sp, (-side / 2, centre_height - side / 2, -side / 2), jitter, seed)
