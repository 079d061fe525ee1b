// ingest.cpp -- the data formats either side of the substep path, host only:
//
//   * a tet mesh from a closed surface mesh (Unity hands the solver a surface `Mesh`): lattice cells whose centre
//     is inside the surface, five tets per cell with alternating parity, boundary faces wound outward;
//   * the binding of a render mesh to the tets (enclosing or nearest tet + barycentric weights), which
//     sb_skin_bind uploads and k_skin evaluates every frame;
//   * tet-mesh files: TetGen .node/.ele(/.face) and Gmsh MSH 2.2 ASCII, read and written;
//   * state snapshots (.sbs): positions + velocities + parameters + frame number, checksummed, for replay.
//
// Reference: NOT IN MOUNT (/root/reference/README.md:1 is the whole reference); formats are the public ones,
// everything else is [SPEC].  No CUDA in this file; every entry point works without a device.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/softbody_b200.h"

struct sb_tetmesh {
  std::vector<float> pos;    // 3V
  std::vector<int32_t> tets; // 4T, positively oriented
  std::vector<int32_t> tris; // 3F, outward
};

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

inline double det6(const double *a, const double *b, const double *c, const double *d) {
  const double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]},
               e3[3] = {d[0] - a[0], d[1] - a[1], d[2] - a[2]};
  return e1[0] * (e2[1] * e3[2] - e2[2] * e3[1]) + e1[1] * (e2[2] * e3[0] - e2[0] * e3[2]) +
         e1[2] * (e2[0] * e3[1] - e2[1] * e3[0]);
}

// every tet positively oriented (e1 . (e2 x e3) > 0); returns the number of degenerate ones
size_t orient_tets(const std::vector<float> &pos, std::vector<int32_t> &tets) {
  size_t flat = 0;
  for (size_t t = 0; t + 3 < tets.size(); t += 4) {
    double p[4][3];
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 3; k++) p[j][k] = pos[3 * (size_t)tets[t + j] + k];
    const double d = det6(p[0], p[1], p[2], p[3]);
    if (d < 0) std::swap(tets[t + 2], tets[t + 3]);
    if (d == 0) flat++;
  }
  return flat;
}

// faces that belong to exactly one tet, wound outward, in (tet, face) order
void boundary_faces(const std::vector<int32_t> &tets, std::vector<int32_t> &tris) {
  static const int F[4][3] = {{0, 2, 1}, {0, 1, 3}, {0, 3, 2}, {1, 2, 3}};
  struct Key {
    int32_t a, b, c;
    bool operator==(const Key &o) const { return a == o.a && b == o.b && c == o.c; }
  };
  struct Hash {
    size_t operator()(const Key &k) const {
      uint64_t h = (uint64_t)(uint32_t)k.a * 0x9e3779b97f4a7c15ull;
      h ^= ((uint64_t)(uint32_t)k.b + 0x7f4a7c15ull) * 0xbf58476d1ce4e5b9ull;
      h ^= ((uint64_t)(uint32_t)k.c + 0x1ce4e5b9ull) * 0x94d049bb133111ebull;
      return (size_t)(h ^ (h >> 31));
    }
  };
  std::unordered_map<Key, uint32_t, Hash> count;
  count.reserve(tets.size());
  auto key_of = [&](size_t t, int f) {
    int32_t v[3] = {tets[t + F[f][0]], tets[t + F[f][1]], tets[t + F[f][2]]};
    std::sort(v, v + 3);
    return Key{v[0], v[1], v[2]};
  };
  for (size_t t = 0; t + 3 < tets.size(); t += 4)
    for (int f = 0; f < 4; f++) count[key_of(t, f)]++;
  tris.clear();
  for (size_t t = 0; t + 3 < tets.size(); t += 4)
    for (int f = 0; f < 4; f++)
      if (count[key_of(t, f)] == 1)
        for (int j = 0; j < 3; j++) tris.push_back(tets[t + F[f][j]]);
}

std::string check_mesh(const sb_tetmesh &m) {
  const size_t V = m.pos.size() / 3;
  for (float p : m.pos)
    if (!std::isfinite(p)) return "vertex position is not finite";
  for (int32_t i : m.tets)
    if (i < 0 || (size_t)i >= V) return "tet vertex index out of range";
  for (int32_t i : m.tris)
    if (i < 0 || (size_t)i >= V) return "triangle vertex index out of range";
  return "";
}

// ---- surface -> tets -------------------------------------------------------------------------------

// Winding number of the closed surface around the centres of one lattice column (fixed y, z; x varies), by
// signed crossings of the +x ray.  Point-in-projected-triangle uses exact-sign edge functions with a top-left
// tie rule, so a ray through an edge or vertex shared by several triangles is counted once.
struct Crossing {
  double x;
  int sign;
};

inline double edge_fn(double ay, double az, double by, double bz, double py, double pz) {
  return (by - ay) * (pz - az) - (bz - az) * (py - ay);
}
inline bool top_left(double ay, double az, double by, double bz) {
  const double dy = by - ay, dz = bz - az;
  return (dz == 0 && dy < 0) || dz > 0; // for counter-clockwise triangles in the (y, z) plane
}

} // namespace

extern "C" {

const char *sb_ingest_last_error(void) { return g_err.c_str(); }

int sb_tetmesh_free(sb_tetmesh_handle m) {
  delete m;
  return SB_OK;
}

int sb_tetmesh_sizes(sb_tetmesh_handle m, uint32_t *V, uint32_t *T, uint32_t *F) {
  if (!m) return fail(SB_E_ARG, "null mesh");
  if (V) *V = (uint32_t)(m->pos.size() / 3);
  if (T) *T = (uint32_t)(m->tets.size() / 4);
  if (F) *F = (uint32_t)(m->tris.size() / 3);
  return SB_OK;
}

int sb_tetmesh_copy(sb_tetmesh_handle m, float *pos_xyz, int32_t *tets, int32_t *tris) {
  if (!m) return fail(SB_E_ARG, "null mesh");
  if (pos_xyz && !m->pos.empty()) std::memcpy(pos_xyz, m->pos.data(), m->pos.size() * sizeof(float));
  if (tets && !m->tets.empty()) std::memcpy(tets, m->tets.data(), m->tets.size() * sizeof(int32_t));
  if (tris && !m->tris.empty()) std::memcpy(tris, m->tris.data(), m->tris.size() * sizeof(int32_t));
  return SB_OK;
}

int sb_tetmesh_desc(sb_tetmesh_handle m, sb_mesh_desc *d) {
  if (!m || !d) return fail(SB_E_ARG, "null argument");
  std::memset(d, 0, sizeof *d);
  d->pos_xyz = m->pos.data();
  d->tets = m->tets.data();
  d->surf_tris = m->tris.empty() ? nullptr : m->tris.data();
  d->n_verts = (uint32_t)(m->pos.size() / 3);
  d->n_tets = (uint32_t)(m->tets.size() / 4);
  d->n_tris = (uint32_t)(m->tris.size() / 3);
  d->density = 1000.0f;
  d->max_tile_passes = -1;
  return SB_OK;
}

int sb_tetmesh_from_arrays(const float *pos_xyz, uint32_t V, const int32_t *tets, uint32_t T, const int32_t *tris,
                           uint32_t F, sb_tetmesh_handle *out) {
  if (!out || !pos_xyz || !tets || !V || !T || (F && !tris)) return fail(SB_E_ARG, "null or empty mesh arrays");
  try {
    sb_tetmesh *m = new sb_tetmesh;
    m->pos.assign(pos_xyz, pos_xyz + 3 * (size_t)V);
    m->tets.assign(tets, tets + 4 * (size_t)T);
    if (F) m->tris.assign(tris, tris + 3 * (size_t)F);
    const std::string e = check_mesh(*m);
    if (!e.empty()) { delete m; return fail(SB_E_ARG, e); }
    orient_tets(m->pos, m->tets);
    if (!F) boundary_faces(m->tets, m->tris);
    *out = m;
    return SB_OK;
  } catch (const std::bad_alloc &) { return fail(SB_E_NOMEM, "out of memory");
  } catch (const std::exception &e) { return fail(SB_E_ARG, e.what()); }
}

int sb_tetmesh_from_surface(const float *sp, uint32_t nv, const int32_t *st, uint32_t nt, float spacing,
                            sb_tetmesh_handle *out) {
  if (!out || !sp || !st || nv < 4 || nt < 4) return fail(SB_E_ARG, "a closed surface needs at least 4 vertices and 4 triangles");
  if (!(spacing > 0) || !std::isfinite(spacing)) return fail(SB_E_ARG, "spacing must be positive");
  try {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (uint32_t i = 0; i < nv; i++)
      for (int k = 0; k < 3; k++) {
        const double p = sp[3 * (size_t)i + k];
        if (!std::isfinite(p)) return fail(SB_E_ARG, "surface vertex is not finite");
        lo[k] = std::min(lo[k], p);
        hi[k] = std::max(hi[k], p);
      }
    for (uint32_t i = 0; i < 3 * (size_t)nt; i++)
      if (st[i] < 0 || (uint32_t)st[i] >= nv) return fail(SB_E_ARG, "surface triangle index out of range");
    const double h = spacing;
    int64_t n[3];
    for (int k = 0; k < 3; k++) {
      n[k] = (int64_t)std::ceil((hi[k] - lo[k]) / h - 1e-9);
      if (n[k] < 1) n[k] = 1;
      // centre the lattice on the bounding box
      lo[k] = 0.5 * (lo[k] + hi[k]) - 0.5 * (double)n[k] * h;
    }
    if ((double)n[0] * (double)n[1] * (double)n[2] > 1.0e9) return fail(SB_E_ARG, "spacing too small: more than 1e9 lattice cells");
    const int64_t nx = n[0], ny = n[1], nz = n[2];
    // triangles bucketed by the (y, z) columns their projection overlaps
    std::vector<std::vector<uint32_t>> bucket((size_t)(ny * nz));
    for (uint32_t t = 0; t < nt; t++) {
      double ylo = 1e300, yhi = -1e300, zlo = 1e300, zhi = -1e300;
      for (int j = 0; j < 3; j++) {
        const float *p = sp + 3 * (size_t)st[3 * (size_t)t + j];
        ylo = std::min(ylo, (double)p[1]); yhi = std::max(yhi, (double)p[1]);
        zlo = std::min(zlo, (double)p[2]); zhi = std::max(zhi, (double)p[2]);
      }
      const int64_t j0 = std::max<int64_t>(0, (int64_t)std::floor((ylo - lo[1]) / h - 0.5)),
                    j1 = std::min<int64_t>(ny - 1, (int64_t)std::ceil((yhi - lo[1]) / h - 0.5)),
                    k0 = std::max<int64_t>(0, (int64_t)std::floor((zlo - lo[2]) / h - 0.5)),
                    k1 = std::min<int64_t>(nz - 1, (int64_t)std::ceil((zhi - lo[2]) / h - 0.5));
      for (int64_t k = k0; k <= k1; k++)
        for (int64_t j = j0; j <= j1; j++) bucket[(size_t)(k * ny + j)].push_back(t);
    }
    std::vector<uint8_t> inside((size_t)(nx * ny * nz), 0);
    std::vector<Crossing> cr;
    for (int64_t k = 0; k < nz; k++)
      for (int64_t j = 0; j < ny; j++) {
        const double py = lo[1] + ((double)j + 0.5) * h, pz = lo[2] + ((double)k + 0.5) * h;
        cr.clear();
        for (uint32_t t : bucket[(size_t)(k * ny + j)]) {
          const float *A = sp + 3 * (size_t)st[3 * (size_t)t], *B = sp + 3 * (size_t)st[3 * (size_t)t + 1],
                      *C = sp + 3 * (size_t)st[3 * (size_t)t + 2];
          double a[3] = {A[0], A[1], A[2]}, b[3] = {B[0], B[1], B[2]}, c[3] = {C[0], C[1], C[2]};
          double area = edge_fn(a[1], a[2], b[1], b[2], c[1], c[2]);
          if (area == 0) continue; // edge-on to the ray
          const int sign = area > 0 ? 1 : -1;
          if (area < 0) { std::swap(b[0], c[0]); std::swap(b[1], c[1]); std::swap(b[2], c[2]); area = -area; }
          const double w0 = edge_fn(b[1], b[2], c[1], c[2], py, pz), w1 = edge_fn(c[1], c[2], a[1], a[2], py, pz),
                       w2 = edge_fn(a[1], a[2], b[1], b[2], py, pz);
          if (w0 < 0 || w1 < 0 || w2 < 0) continue;
          if ((w0 == 0 && !top_left(b[1], b[2], c[1], c[2])) || (w1 == 0 && !top_left(c[1], c[2], a[1], a[2])) ||
              (w2 == 0 && !top_left(a[1], a[2], b[1], b[2])))
            continue;
          cr.push_back({(w0 * a[0] + w1 * b[0] + w2 * c[0]) / area, sign});
        }
        if (cr.empty()) continue;
        int total = 0;
        for (const Crossing &q : cr) total += q.sign;
        if (total != 0) { // a ray that enters a closed surface leaves it again
          char msg[160];
          std::snprintf(msg, sizeof msg, "the surface is not closed (a ray at y = %.6g, z = %.6g crosses it %+d times net)", py, pz, total);
          return fail(SB_E_ARG, msg);
        }
        std::sort(cr.begin(), cr.end(), [](const Crossing &p, const Crossing &q) { return p.x < q.x; });
        // winding number at a point = signed crossings of the ray beyond it
        int wind = 0;
        size_t q = cr.size();
        for (int64_t i = nx - 1; i >= 0; i--) {
          const double px = lo[0] + ((double)i + 0.5) * h;
          while (q > 0 && cr[q - 1].x > px) wind += cr[--q].sign;
          if (wind != 0) inside[(size_t)((k * ny + j) * nx + i)] = 1;
        }
      }
    // lattice vertices used by the kept cells, numbered in order of first use (cells x-fastest); five tets per cell
    static const int EVEN[5][4] = {{0, 3, 5, 6}, {1, 0, 3, 5}, {2, 0, 6, 3}, {4, 0, 5, 6}, {7, 3, 6, 5}};
    const int64_t vx = nx + 1, vy = ny + 1;
    std::unordered_map<int64_t, int32_t> vid;
    sb_tetmesh *m = new sb_tetmesh;
    auto vertex = [&](int64_t i, int64_t j, int64_t k) -> int32_t {
      const int64_t key = (k * vy + j) * vx + i;
      auto it = vid.find(key);
      if (it != vid.end()) return it->second;
      const int32_t id = (int32_t)(m->pos.size() / 3);
      m->pos.push_back((float)(lo[0] + (double)i * h));
      m->pos.push_back((float)(lo[1] + (double)j * h));
      m->pos.push_back((float)(lo[2] + (double)k * h));
      vid.emplace(key, id);
      return id;
    };
    for (int64_t k = 0; k < nz; k++)
      for (int64_t j = 0; j < ny; j++)
        for (int64_t i = 0; i < nx; i++) {
          if (!inside[(size_t)((k * ny + j) * nx + i)]) continue;
          int32_t c[8];
          for (int b = 0; b < 8; b++) c[b] = vertex(i + (b & 1), j + ((b >> 1) & 1), k + ((b >> 2) & 1));
          const int flip = (int)((i + j + k) & 1); // odd cells are the even split mirrored in x
          for (int t = 0; t < 5; t++)
            for (int q = 0; q < 4; q++) m->tets.push_back(c[EVEN[t][q] ^ flip]);
        }
    if (m->tets.empty()) { delete m; return fail(SB_E_ARG, "no lattice cell centre is inside the surface (open surface or spacing too large)"); }
    orient_tets(m->pos, m->tets);
    boundary_faces(m->tets, m->tris);
    *out = m;
    return SB_OK;
  } catch (const std::bad_alloc &) { return fail(SB_E_NOMEM, "out of memory");
  } catch (const std::exception &e) { return fail(SB_E_ARG, e.what()); }
}

// ---- boundary vertices onto the surface -----------------------------------------------------------------------

namespace {

// closest point of triangle abc to p (Ericson, Real-Time Collision Detection 5.1.5), in double
void closest_on_triangle(const double *p, const double *a, const double *b, const double *c, double *out) {
  double ab[3], ac[3], ap[3];
  for (int k = 0; k < 3; k++) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
  auto dot = [](const double *u, const double *v) { return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]; };
  const double d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= 0 && d2 <= 0) { for (int k = 0; k < 3; k++) out[k] = a[k]; return; }
  double bp[3];
  for (int k = 0; k < 3; k++) bp[k] = p[k] - b[k];
  const double d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= 0 && d4 <= d3) { for (int k = 0; k < 3; k++) out[k] = b[k]; return; }
  const double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) {
    const double v = d1 / (d1 - d3);
    for (int k = 0; k < 3; k++) out[k] = a[k] + v * ab[k];
    return;
  }
  double cp[3];
  for (int k = 0; k < 3; k++) cp[k] = p[k] - c[k];
  const double d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= 0 && d5 <= d6) { for (int k = 0; k < 3; k++) out[k] = c[k]; return; }
  const double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) {
    const double w = d2 / (d2 - d6);
    for (int k = 0; k < 3; k++) out[k] = a[k] + w * ac[k];
    return;
  }
  const double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    for (int k = 0; k < 3; k++) out[k] = b[k] + w * (c[k] - b[k]);
    return;
  }
  const double den = 1.0 / (va + vb + vc), v = vb * den, w = vc * den;
  for (int k = 0; k < 3; k++) out[k] = a[k] + ab[k] * v + ac[k] * w;
}

} // namespace

/*
 * The lattice of sb_tetmesh_from_surface does not conform to the surface it fills: its boundary is a staircase up
 * to half a cell inside or outside.  This moves every boundary vertex of the tet mesh to the closest point of the
 * surface when that point is nearer than max_dist (half the lattice spacing is the natural choice), unless the move
 * would shrink a tet at the vertex below 30 % of the volume it had in the lattice (the vertex then stays).  Returns the number of
 * vertices moved through *n_moved (may be NULL).
 */
int sb_tetmesh_snap_to_surface(sb_tetmesh_handle m, const float *sp, uint32_t nv, const int32_t *st, uint32_t nt, float max_dist,
                               uint32_t *n_moved) {
  if (!m || !sp || !st || !nv || !nt) return fail(SB_E_ARG, "null or empty argument");
  if (!(max_dist > 0) || !std::isfinite(max_dist)) return fail(SB_E_ARG, "max_dist must be positive");
  try {
    for (size_t i = 0; i < 3 * (size_t)nt; i++)
      if (st[i] < 0 || (uint32_t)st[i] >= nv) return fail(SB_E_ARG, "surface triangle index out of range");
    const size_t V = m->pos.size() / 3, T = m->tets.size() / 4;
    // boundary vertices and the tets at each of them
    std::vector<uint8_t> on_boundary(V, 0);
    for (int32_t i : m->tris) on_boundary[i] = 1;
    std::vector<uint32_t> toff(V + 1, 0), tlist(4 * T);
    for (size_t i = 0; i < 4 * T; i++) toff[(size_t)m->tets[i] + 1]++;
    for (size_t v = 0; v < V; v++) toff[v + 1] += toff[v];
    {
      std::vector<uint32_t> cur(toff.begin(), toff.end() - 1);
      for (size_t t = 0; t < T; t++)
        for (int j = 0; j < 4; j++) tlist[cur[m->tets[4 * t + j]]++] = (uint32_t)t;
    }
    // uniform grid over the surface triangles, cells of max_dist: a query looks at the 27 cells around it
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (uint32_t i = 0; i < nv; i++)
      for (int k = 0; k < 3; k++) {
        const double p = sp[3 * (size_t)i + k];
        if (!std::isfinite(p)) return fail(SB_E_ARG, "surface vertex is not finite");
        lo[k] = std::min(lo[k], p); hi[k] = std::max(hi[k], p);
      }
    double cell = max_dist;
    int64_t g[3];
    for (;;) {
      double cells = 1;
      for (int k = 0; k < 3; k++) { g[k] = std::max<int64_t>(1, (int64_t)std::ceil((hi[k] - lo[k]) / cell) + 1); cells *= (double)g[k]; }
      if (cells <= 32.0e6) break;
      cell *= 1.5;
    }
    const int64_t reach = (int64_t)std::ceil(max_dist / cell);
    auto cell_of = [&](double p, int k) { return std::min<int64_t>(g[k] - 1, std::max<int64_t>(0, (int64_t)std::floor((p - lo[k]) / cell))); };
    std::vector<uint32_t> off((size_t)(g[0] * g[1] * g[2]) + 1, 0), items;
    for (int phase = 0; phase < 2; phase++) {
      for (uint32_t t = 0; t < nt; t++) {
        int64_t c0[3], c1[3];
        for (int k = 0; k < 3; k++) {
          double tl = 1e300, th = -1e300;
          for (int j = 0; j < 3; j++) { const double p = sp[3 * (size_t)st[3 * (size_t)t + j] + k]; tl = std::min(tl, p); th = std::max(th, p); }
          c0[k] = cell_of(tl, k); c1[k] = cell_of(th, k);
        }
        for (int64_t z = c0[2]; z <= c1[2]; z++)
          for (int64_t y = c0[1]; y <= c1[1]; y++)
            for (int64_t x = c0[0]; x <= c1[0]; x++) {
              const size_t c = (size_t)((z * g[1] + y) * g[0] + x);
              if (phase == 0) off[c + 1]++;
              else items[off[c]++] = t;
            }
      }
      if (phase == 0) {
        for (size_t c = 0; c + 1 < off.size(); c++) off[c + 1] += off[c];
        items.resize(off.back());
      } else {
        for (size_t c = off.size() - 1; c > 0; c--) off[c] = off[c - 1];
        off[0] = 0;
      }
    }
    auto vol6 = [&](size_t t, size_t moved_v, const double *np) {
      double q[4][3];
      for (int j = 0; j < 4; j++) {
        const size_t v = (size_t)m->tets[4 * t + j];
        for (int k = 0; k < 3; k++) q[j][k] = v == moved_v ? np[k] : (double)m->pos[3 * v + k];
      }
      return det6(q[0], q[1], q[2], q[3]);
    };
    std::vector<double> vol0(T); // volumes before any vertex moves: the bound is relative to these
    {
      const double none[3] = {0, 0, 0};
      for (size_t t = 0; t < T; t++) vol0[t] = vol6(t, (size_t)-1, none);
    }
    uint32_t moved = 0;
    for (size_t v = 0; v < V; v++) {
      if (!on_boundary[v]) continue;
      const double p[3] = {m->pos[3 * v], m->pos[3 * v + 1], m->pos[3 * v + 2]};
      double best[3] = {0, 0, 0}, best_d2 = (double)max_dist * max_dist;
      bool found = false;
      const int64_t c[3] = {cell_of(p[0], 0), cell_of(p[1], 1), cell_of(p[2], 2)};
      for (int64_t z = std::max<int64_t>(0, c[2] - reach); z <= std::min(g[2] - 1, c[2] + reach); z++)
        for (int64_t y = std::max<int64_t>(0, c[1] - reach); y <= std::min(g[1] - 1, c[1] + reach); y++)
          for (int64_t x = std::max<int64_t>(0, c[0] - reach); x <= std::min(g[0] - 1, c[0] + reach); x++) {
            const size_t cc = (size_t)((z * g[1] + y) * g[0] + x);
            for (uint32_t k = off[cc]; k < off[cc + 1]; k++) {
              const uint32_t t = items[k];
              double a[3], b[3], cq[3], q[3];
              for (int j = 0; j < 3; j++) {
                a[j] = sp[3 * (size_t)st[3 * (size_t)t] + j];
                b[j] = sp[3 * (size_t)st[3 * (size_t)t + 1] + j];
                cq[j] = sp[3 * (size_t)st[3 * (size_t)t + 2] + j];
              }
              closest_on_triangle(p, a, b, cq, q);
              const double d2 = (q[0] - p[0]) * (q[0] - p[0]) + (q[1] - p[1]) * (q[1] - p[1]) + (q[2] - p[2]) * (q[2] - p[2]);
              if (d2 < best_d2) { best_d2 = d2; best[0] = q[0]; best[1] = q[1]; best[2] = q[2]; found = true; }
            }
          }
      if (!found || best_d2 == 0) continue;
      // the float the vertex becomes must keep every tet at it well shaped: all the way, else part of the way
      for (int step = 4; step >= 1; step--) {
        const double f = 0.25 * step;
        const double np[3] = {(double)(float)(p[0] + f * (best[0] - p[0])), (double)(float)(p[1] + f * (best[1] - p[1])),
                              (double)(float)(p[2] + f * (best[2] - p[2]))};
        bool ok = true;
        for (uint32_t k = toff[v]; k < toff[v + 1] && ok; k++) ok = vol6(tlist[k], v, np) > 0.3 * vol0[tlist[k]];
        if (!ok) continue;
        for (int k = 0; k < 3; k++) m->pos[3 * v + k] = (float)np[k];
        moved++;
        break;
      }
    }
    if (n_moved) *n_moved = moved;
    return SB_OK;
  } catch (const std::bad_alloc &) { return fail(SB_E_NOMEM, "out of memory");
  } catch (const std::exception &e) { return fail(SB_E_ARG, e.what()); }
}

// ---- render mesh -> tets ------------------------------------------------------------------------------

int sb_skin_compute(const float *tet_pos_xyz, uint32_t V, const int32_t *tets, uint32_t T, const float *pts_xyz, uint32_t n,
                    int32_t *tet_of, float *bary4) {
  if (!tet_pos_xyz || !tets || !T || !V || (n && (!pts_xyz || !tet_of || !bary4))) return fail(SB_E_ARG, "null or empty argument");
  try {
    for (size_t i = 0; i < 4 * (size_t)T; i++)
      if (tets[i] < 0 || (uint32_t)tets[i] >= V) return fail(SB_E_ARG, "tet vertex index out of range");
    // uniform grid over the tets' bounding boxes, about one tet-size per cell
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, mean_ext = 0;
    for (uint32_t t = 0; t < T; t++) {
      double tl[3] = {1e300, 1e300, 1e300}, th[3] = {-1e300, -1e300, -1e300};
      for (int j = 0; j < 4; j++)
        for (int k = 0; k < 3; k++) {
          const double p = tet_pos_xyz[3 * (size_t)tets[4 * (size_t)t + j] + k];
          tl[k] = std::min(tl[k], p); th[k] = std::max(th[k], p);
        }
      for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], tl[k]); hi[k] = std::max(hi[k], th[k]); mean_ext += th[k] - tl[k]; }
    }
    mean_ext /= 3.0 * T;
    double cell = std::max(mean_ext, 1e-30);
    int64_t g[3];
    for (;;) {
      double cells = 1;
      for (int k = 0; k < 3; k++) { g[k] = std::max<int64_t>(1, (int64_t)std::ceil((hi[k] - lo[k]) / cell)); cells *= (double)g[k]; }
      if (cells <= 64.0e6) break;
      cell *= 1.5;
    }
    auto cell_of = [&](double p, int k) { return std::min<int64_t>(g[k] - 1, std::max<int64_t>(0, (int64_t)std::floor((p - lo[k]) / cell))); };
    std::vector<uint32_t> off((size_t)(g[0] * g[1] * g[2]) + 1, 0), items;
    for (int phase = 0; phase < 2; phase++) {
      for (uint32_t t = 0; t < T; t++) {
        int64_t c0[3], c1[3];
        for (int k = 0; k < 3; k++) {
          double tl = 1e300, th = -1e300;
          for (int j = 0; j < 4; j++) { const double p = tet_pos_xyz[3 * (size_t)tets[4 * (size_t)t + j] + k]; tl = std::min(tl, p); th = std::max(th, p); }
          c0[k] = cell_of(tl, k); c1[k] = cell_of(th, k);
        }
        for (int64_t z = c0[2]; z <= c1[2]; z++)
          for (int64_t y = c0[1]; y <= c1[1]; y++)
            for (int64_t x = c0[0]; x <= c1[0]; x++) {
              const size_t c = (size_t)((z * g[1] + y) * g[0] + x);
              if (phase == 0) off[c + 1]++;
              else items[off[c]++] = t;
            }
      }
      if (phase == 0) {
        for (size_t c = 0; c + 1 < off.size(); c++) off[c + 1] += off[c];
        items.resize(off.back());
      } else {
        for (size_t c = off.size() - 1; c > 0; c--) off[c] = off[c - 1];
        off[0] = 0;
      }
    }
    auto bary = [&](uint32_t t, const double *p, double *b) -> bool {
      double q[4][3];
      for (int j = 0; j < 4; j++)
        for (int k = 0; k < 3; k++) q[j][k] = tet_pos_xyz[3 * (size_t)tets[4 * (size_t)t + j] + k];
      const double d = det6(q[0], q[1], q[2], q[3]);
      if (d == 0) return false;
      b[0] = det6(p, q[1], q[2], q[3]) / d;
      b[1] = det6(q[0], p, q[2], q[3]) / d;
      b[2] = det6(q[0], q[1], p, q[3]) / d;
      b[3] = det6(q[0], q[1], q[2], p) / d;
      return true;
    };
    for (uint32_t i = 0; i < n; i++) {
      const double p[3] = {pts_xyz[3 * (size_t)i], pts_xyz[3 * (size_t)i + 1], pts_xyz[3 * (size_t)i + 2]};
      if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) return fail(SB_E_ARG, "render vertex is not finite");
      const int64_t c[3] = {cell_of(p[0], 0), cell_of(p[1], 1), cell_of(p[2], 2)};
      double best_min = -1e300, best_b[4] = {1, 0, 0, 0};
      int64_t best_t = -1;
      // the point's own cell first, then growing shells of cells until a shell beyond the best candidate adds nothing
      const int64_t max_ring = std::max(g[0], std::max(g[1], g[2]));
      int64_t found_ring = -1;
      for (int64_t ring = 0; ring <= max_ring; ring++) {
        for (int64_t z = std::max<int64_t>(0, c[2] - ring); z <= std::min(g[2] - 1, c[2] + ring); z++)
          for (int64_t y = std::max<int64_t>(0, c[1] - ring); y <= std::min(g[1] - 1, c[1] + ring); y++)
            for (int64_t x = std::max<int64_t>(0, c[0] - ring); x <= std::min(g[0] - 1, c[0] + ring); x++) {
              if (std::max(std::llabs(x - c[0]), std::max(std::llabs(y - c[1]), std::llabs(z - c[2]))) != ring) continue;
              const size_t cc = (size_t)((z * g[1] + y) * g[0] + x);
              for (uint32_t k = off[cc]; k < off[cc + 1]; k++) {
                double b[4];
                if (!bary(items[k], p, b)) continue;
                const double mn = std::min(std::min(b[0], b[1]), std::min(b[2], b[3]));
                if (mn > best_min || (mn == best_min && (int64_t)items[k] < best_t)) {
                  best_min = mn; best_t = items[k];
                  for (int j = 0; j < 4; j++) best_b[j] = b[j];
                }
              }
            }
        if (best_t >= 0 && found_ring < 0) found_ring = ring;
        // enclosed (to rounding): done.  Outside every tet: the best of the first shell that has any, and one more
        if (best_t >= 0 && (best_min >= -1e-7 || ring > found_ring)) break;
      }
      if (best_t < 0) return fail(SB_E_ARG, "no usable tet for a render vertex (all tets degenerate?)");
      tet_of[i] = (int32_t)best_t;
      // weights as floats that sum to one in the kernel's sense: b0 is what is left
      const float b1 = (float)best_b[1], b2 = (float)best_b[2], b3 = (float)best_b[3];
      bary4[4 * (size_t)i + 1] = b1; bary4[4 * (size_t)i + 2] = b2; bary4[4 * (size_t)i + 3] = b3;
      bary4[4 * (size_t)i] = (float)(1.0 - ((double)b1 + (double)b2 + (double)b3));
    }
    return SB_OK;
  } catch (const std::bad_alloc &) { return fail(SB_E_NOMEM, "out of memory");
  } catch (const std::exception &e) { return fail(SB_E_ARG, e.what()); }
}

// ---- files ---------------------------------------------------------------------------------------------------

namespace {

bool next_data_line(std::istream &in, std::string &line) {
  while (std::getline(in, line)) {
    const size_t h = line.find('#');
    if (h != std::string::npos) line.erase(h);
    if (line.find_first_not_of(" \t\r\n") != std::string::npos) return true;
  }
  return false;
}

std::string strip_ext(const std::string &p, std::string *ext) {
  const size_t dot = p.find_last_of('.'), slash = p.find_last_of('/');
  if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) { if (ext) ext->clear(); return p; }
  if (ext) *ext = p.substr(dot);
  return p.substr(0, dot);
}

// a count read from a header cannot exceed the bytes of its file (every entry takes several): guards the resize
long long file_bytes(const std::string &path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  return f ? (long long)f.tellg() : 0;
}

// TetGen: <base>.node + <base>.ele (+ optional <base>.face)
int load_tetgen(const std::string &base, sb_tetmesh &m) {
  std::ifstream fn(base + ".node"), fe(base + ".ele");
  if (!fn) return fail(SB_E_ARG, "cannot open " + base + ".node");
  if (!fe) return fail(SB_E_ARG, "cannot open " + base + ".ele");
  std::string line;
  long long np = 0, dim = 0, nattr = 0, nmark = 0;
  if (!next_data_line(fn, line)) return fail(SB_E_ARG, ".node: empty file");
  { std::istringstream s(line); s >> np >> dim >> nattr >> nmark; }
  if (np <= 0 || np > 0x7fffffffLL || np > file_bytes(base + ".node") || dim != 3) return fail(SB_E_ARG, ".node: expected '<points> 3 <attrs> <markers>'");
  std::map<long long, int32_t> id_of;
  m.pos.resize(3 * (size_t)np);
  for (long long i = 0; i < np; i++) {
    if (!next_data_line(fn, line)) return fail(SB_E_ARG, ".node: fewer points than the header says");
    std::istringstream s(line);
    long long id; double x, y, z;
    if (!(s >> id >> x >> y >> z)) return fail(SB_E_ARG, ".node: malformed point line");
    if (!id_of.emplace(id, (int32_t)i).second) return fail(SB_E_ARG, ".node: duplicate point id");
    m.pos[3 * (size_t)i] = (float)x; m.pos[3 * (size_t)i + 1] = (float)y; m.pos[3 * (size_t)i + 2] = (float)z;
  }
  long long ne = 0, npt = 0;
  if (!next_data_line(fe, line)) return fail(SB_E_ARG, ".ele: empty file");
  { std::istringstream s(line); s >> ne >> npt; }
  if (ne <= 0 || ne > 0x7fffffffLL || ne > file_bytes(base + ".ele") || (npt != 4 && npt != 10)) return fail(SB_E_ARG, ".ele: expected '<tets> 4|10 <attrs>'");
  m.tets.resize(4 * (size_t)ne);
  for (long long t = 0; t < ne; t++) {
    if (!next_data_line(fe, line)) return fail(SB_E_ARG, ".ele: fewer tets than the header says");
    std::istringstream s(line);
    long long id, v[4];
    if (!(s >> id >> v[0] >> v[1] >> v[2] >> v[3])) return fail(SB_E_ARG, ".ele: malformed tet line");
    for (int j = 0; j < 4; j++) {
      auto it = id_of.find(v[j]);
      if (it == id_of.end()) return fail(SB_E_ARG, ".ele: tet refers to an unknown point id");
      m.tets[4 * (size_t)t + j] = it->second;
    }
  }
  std::ifstream ff(base + ".face");
  if (ff && next_data_line(ff, line)) {
    long long nf = 0;
    { std::istringstream s(line); s >> nf; }
    if (nf < 0 || nf > 0x7fffffffLL) return fail(SB_E_ARG, ".face: bad face count");
    for (long long f = 0; f < nf; f++) {
      if (!next_data_line(ff, line)) return fail(SB_E_ARG, ".face: fewer faces than the header says");
      std::istringstream s(line);
      long long id, v[3];
      if (!(s >> id >> v[0] >> v[1] >> v[2])) return fail(SB_E_ARG, ".face: malformed face line");
      for (int j = 0; j < 3; j++) {
        auto it = id_of.find(v[j]);
        if (it == id_of.end()) return fail(SB_E_ARG, ".face: face refers to an unknown point id");
        m.tris.push_back(it->second);
      }
    }
  }
  return SB_OK;
}

// Gmsh MSH ASCII, versions 2.x and 4.1: element type 4 = 4-node tet, type 2 = 3-node triangle; other types are skipped.
// 2.x lists nodes as "id x y z" and elements as "id type ntags tags... nodes..."; 4.1 groups both into entity blocks
// ("dim tag parametric n" + n node tags + n coordinate lines; "dim tag type n" + n lines "id nodes...").
int load_msh(const std::string &path, sb_tetmesh &m) {
  std::ifstream in(path);
  if (!in) return fail(SB_E_ARG, "cannot open " + path);
  const long long bytes = file_bytes(path);
  std::string line;
  std::map<long long, int32_t> id_of;
  int major = 0;
  auto add_node = [&](long long id, size_t i, double x, double y, double z) {
    if (!id_of.emplace(id, (int32_t)i).second) return false;
    m.pos[3 * i] = (float)x; m.pos[3 * i + 1] = (float)y; m.pos[3 * i + 2] = (float)z;
    return true;
  };
  // the node ids at the end of an element line -> tets / tris (nn = 0: an element type this library has no use for)
  auto add_element = [&](std::istringstream &s, int nn) -> const char * {
    for (int j = 0; j < nn; j++) {
      long long v;
      if (!(s >> v)) return ".msh: element with too few nodes";
      auto it = id_of.find(v);
      if (it == id_of.end()) return ".msh: element refers to an unknown node id";
      (nn == 4 ? m.tets : m.tris).push_back(it->second);
    }
    return nullptr;
  };
  while (std::getline(in, line)) {
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
    if (line == "$MeshFormat") {
      double ver = 0; int type = -1;
      if (!std::getline(in, line)) break;
      std::istringstream s(line);
      s >> ver >> type;
      if (type != 0) return fail(SB_E_ARG, ".msh: binary files are not supported (save as ASCII)");
      if (ver >= 2.0 && ver < 3.0) major = 2;
      else if (ver > 4.05 && ver < 4.15) major = 4;
      else return fail(SB_E_ARG, ".msh: only MSH 2.x and 4.1 ASCII are supported");
    } else if (line == "$Nodes" && major == 2) {
      long long nn = 0;
      in >> nn;
      if (nn <= 0 || nn > 0x7fffffffLL || nn > bytes) return fail(SB_E_ARG, ".msh: no nodes");
      m.pos.resize(3 * (size_t)nn);
      for (long long i = 0; i < nn; i++) {
        long long id; double x, y, z;
        if (!(in >> id >> x >> y >> z)) return fail(SB_E_ARG, ".msh: malformed node");
        if (!add_node(id, (size_t)i, x, y, z)) return fail(SB_E_ARG, ".msh: duplicate node id");
      }
    } else if (line == "$Nodes" && major == 4) {
      long long nblocks = 0, nn = 0, tmin = 0, tmax = 0;
      in >> nblocks >> nn >> tmin >> tmax;
      if (!in || nn <= 0 || nn > 0x7fffffffLL || nn > bytes || nblocks < 0 || nblocks > bytes) return fail(SB_E_ARG, ".msh: no nodes");
      m.pos.resize(3 * (size_t)nn);
      long long done = 0;
      std::vector<long long> tags;
      for (long long b = 0; b < nblocks; b++) {
        long long dim = 0, tag = 0, parametric = 0, nb = 0;
        if (!(in >> dim >> tag >> parametric >> nb) || nb < 0 || done + nb > nn || dim < 0 || dim > 3)
          return fail(SB_E_ARG, ".msh: malformed node block");
        tags.resize((size_t)nb);
        for (long long i = 0; i < nb; i++)
          if (!(in >> tags[(size_t)i])) return fail(SB_E_ARG, ".msh: malformed node block (tags)");
        for (long long i = 0; i < nb; i++) {
          double x, y, z, skip;
          if (!(in >> x >> y >> z)) return fail(SB_E_ARG, ".msh: malformed node");
          for (long long k = 0; parametric && k < dim; k++)
            if (!(in >> skip)) return fail(SB_E_ARG, ".msh: malformed parametric node");
          if (!add_node(tags[(size_t)i], (size_t)(done + i), x, y, z)) return fail(SB_E_ARG, ".msh: duplicate node id");
        }
        done += nb;
      }
      if (done != nn) return fail(SB_E_ARG, ".msh: fewer nodes than the header says");
    } else if (line == "$Elements" && major == 2) {
      long long ne = 0;
      in >> ne;
      if (!in || ne < 0) return fail(SB_E_ARG, ".msh: bad element count");
      std::getline(in, line);
      for (long long e = 0; e < ne; e++) {
        if (!std::getline(in, line)) return fail(SB_E_ARG, ".msh: fewer elements than the header says");
        std::istringstream s(line);
        long long id, type, ntags, tag;
        if (!(s >> id >> type >> ntags) || ntags < 0 || ntags > 64) return fail(SB_E_ARG, ".msh: malformed element");
        for (long long k = 0; k < ntags; k++) s >> tag;
        if (const char *err = add_element(s, type == 4 ? 4 : type == 2 ? 3 : 0)) return fail(SB_E_ARG, err);
      }
    } else if (line == "$Elements" && major == 4) {
      long long nblocks = 0, ne = 0, tmin = 0, tmax = 0;
      in >> nblocks >> ne >> tmin >> tmax;
      if (!in || ne < 0 || nblocks < 0 || nblocks > bytes) return fail(SB_E_ARG, ".msh: bad element count");
      long long done = 0;
      for (long long b = 0; b < nblocks; b++) {
        long long dim = 0, tag = 0, type = 0, nb = 0;
        if (!(in >> dim >> tag >> type >> nb) || nb < 0 || done + nb > ne) return fail(SB_E_ARG, ".msh: malformed element block");
        std::getline(in, line);
        for (long long e = 0; e < nb; e++) {
          if (!std::getline(in, line)) return fail(SB_E_ARG, ".msh: fewer elements than the header says");
          std::istringstream s(line);
          long long id;
          if (!(s >> id)) return fail(SB_E_ARG, ".msh: malformed element");
          if (const char *err = add_element(s, type == 4 ? 4 : type == 2 ? 3 : 0)) return fail(SB_E_ARG, err);
        }
        done += nb;
      }
    }
  }
  if (!major) return fail(SB_E_ARG, ".msh: no $MeshFormat section");
  if (m.pos.empty() || m.tets.empty()) return fail(SB_E_ARG, ".msh: no nodes or no tetrahedra (element type 4)");
  return SB_OK;
}

uint64_t fnv1a(const void *data, size_t n, uint64_t h = 0xcbf29ce484222325ull) {
  const unsigned char *p = (const unsigned char *)data;
  for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
  return h;
}

struct StateHeader {
  char magic[8]; // "SBSTATE1"
  uint32_t version, n_verts;
  uint64_t frame, topo_hash;
  sb_params params;
};
static_assert(sizeof(StateHeader) == 8 + 8 + 16 + 48, "snapshot header layout");

} // namespace

int sb_tetmesh_load(const char *path, sb_tetmesh_handle *out) {
  if (!path || !out) return fail(SB_E_ARG, "null argument");
  try {
    std::string ext;
    const std::string base = strip_ext(path, &ext);
    sb_tetmesh *m = new sb_tetmesh;
    int rc;
    if (ext == ".msh") rc = load_msh(path, *m);
    else if (ext == ".node" || ext == ".ele" || ext == ".face") rc = load_tetgen(base, *m);
    else if (ext.empty()) rc = load_tetgen(path, *m);
    else rc = fail(SB_E_ARG, "unknown mesh file extension '" + ext + "' (.msh, .node/.ele)");
    if (rc == SB_OK) {
      const std::string e = check_mesh(*m);
      if (!e.empty()) rc = fail(SB_E_ARG, e);
    }
    if (rc != SB_OK) { delete m; return rc; }
    orient_tets(m->pos, m->tets);
    if (m->tris.empty()) boundary_faces(m->tets, m->tris);
    *out = m;
    return SB_OK;
  } catch (const std::bad_alloc &) { return fail(SB_E_NOMEM, "out of memory");
  } catch (const std::exception &e) { return fail(SB_E_ARG, e.what()); }
}

int sb_tetmesh_save(sb_tetmesh_handle m, const char *path) {
  if (!m || !path) return fail(SB_E_ARG, "null argument");
  std::string ext;
  const std::string base = strip_ext(path, &ext);
  const size_t V = m->pos.size() / 3, T = m->tets.size() / 4, F = m->tris.size() / 3;
  auto open = [&](const std::string &p) { return std::fopen(p.c_str(), "w"); };
  if (ext == ".msh") {
    FILE *f = open(path);
    if (!f) return fail(SB_E_ARG, std::string("cannot write ") + path);
    std::fprintf(f, "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%zu\n", V);
    for (size_t i = 0; i < V; i++) std::fprintf(f, "%zu %.9g %.9g %.9g\n", i + 1, m->pos[3 * i], m->pos[3 * i + 1], m->pos[3 * i + 2]);
    std::fprintf(f, "$EndNodes\n$Elements\n%zu\n", F + T);
    for (size_t i = 0; i < F; i++)
      std::fprintf(f, "%zu 2 2 0 1 %d %d %d\n", i + 1, m->tris[3 * i] + 1, m->tris[3 * i + 1] + 1, m->tris[3 * i + 2] + 1);
    for (size_t i = 0; i < T; i++)
      std::fprintf(f, "%zu 4 2 0 1 %d %d %d %d\n", F + i + 1, m->tets[4 * i] + 1, m->tets[4 * i + 1] + 1, m->tets[4 * i + 2] + 1, m->tets[4 * i + 3] + 1);
    std::fprintf(f, "$EndElements\n");
    return std::fclose(f) == 0 ? SB_OK : fail(SB_E_ARG, "write failed");
  }
  if (ext == ".node" || ext == ".ele" || ext.empty()) {
    const std::string b = ext.empty() ? std::string(path) : base;
    FILE *f = open(b + ".node");
    if (!f) return fail(SB_E_ARG, "cannot write " + b + ".node");
    std::fprintf(f, "# written by libsoftbody_b200\n%zu 3 0 0\n", V);
    for (size_t i = 0; i < V; i++) std::fprintf(f, "%zu %.9g %.9g %.9g\n", i, m->pos[3 * i], m->pos[3 * i + 1], m->pos[3 * i + 2]);
    if (std::fclose(f) != 0) return fail(SB_E_ARG, "write failed");
    f = open(b + ".ele");
    if (!f) return fail(SB_E_ARG, "cannot write " + b + ".ele");
    std::fprintf(f, "%zu 4 0\n", T);
    for (size_t i = 0; i < T; i++) std::fprintf(f, "%zu %d %d %d %d\n", i, m->tets[4 * i], m->tets[4 * i + 1], m->tets[4 * i + 2], m->tets[4 * i + 3]);
    if (std::fclose(f) != 0) return fail(SB_E_ARG, "write failed");
    f = open(b + ".face");
    if (!f) return fail(SB_E_ARG, "cannot write " + b + ".face");
    std::fprintf(f, "%zu 0\n", F);
    for (size_t i = 0; i < F; i++) std::fprintf(f, "%zu %d %d %d\n", i, m->tris[3 * i], m->tris[3 * i + 1], m->tris[3 * i + 2]);
    return std::fclose(f) == 0 ? SB_OK : fail(SB_E_ARG, "write failed");
  }
  return fail(SB_E_ARG, "unknown mesh file extension '" + ext + "' (.msh, .node/.ele)");
}

uint64_t sb_topology_hash(uint32_t n_verts, const int32_t *tets, uint32_t n_tets) {
  uint64_t h = fnv1a(&n_verts, sizeof n_verts);
  h = fnv1a(&n_tets, sizeof n_tets, h);
  if (tets && n_tets) h = fnv1a(tets, 4 * (size_t)n_tets * sizeof(int32_t), h);
  return h;
}

int sb_state_write(const char *path, const float *x4, const float *v4, uint32_t n_verts, const sb_params *params,
                   uint64_t frame, uint64_t topo_hash) {
  if (!path || !x4 || !v4 || !params || !n_verts) return fail(SB_E_ARG, "null or empty argument");
  FILE *f = std::fopen(path, "wb");
  if (!f) return fail(SB_E_ARG, std::string("cannot write ") + path);
  StateHeader h{};
  std::memcpy(h.magic, "SBSTATE1", 8);
  h.version = 1; h.n_verts = n_verts; h.frame = frame; h.topo_hash = topo_hash; h.params = *params;
  const size_t bytes = 16 * (size_t)n_verts;
  uint64_t sum = fnv1a(&h, sizeof h);
  sum = fnv1a(x4, bytes, sum);
  sum = fnv1a(v4, bytes, sum);
  bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 && std::fwrite(x4, 1, bytes, f) == bytes &&
            std::fwrite(v4, 1, bytes, f) == bytes && std::fwrite(&sum, sizeof sum, 1, f) == 1;
  ok = (std::fclose(f) == 0) && ok;
  return ok ? SB_OK : fail(SB_E_ARG, "write failed");
}

int sb_state_read(const char *path, float *x4, float *v4, uint32_t capacity, uint32_t *n_verts, sb_params *params,
                  uint64_t *frame, uint64_t *topo_hash) {
  if (!path) return fail(SB_E_ARG, "null path");
  FILE *f = std::fopen(path, "rb");
  if (!f) return fail(SB_E_ARG, std::string("cannot open ") + path);
  StateHeader h{};
  if (std::fread(&h, sizeof h, 1, f) != 1 || std::memcmp(h.magic, "SBSTATE1", 8) != 0 || h.version != 1) {
    std::fclose(f);
    return fail(SB_E_ARG, "not a state snapshot (bad magic or version)");
  }
  if (n_verts) *n_verts = h.n_verts;
  if (params) *params = h.params;
  if (frame) *frame = h.frame;
  if (topo_hash) *topo_hash = h.topo_hash;
  if (!x4 && !v4) { std::fclose(f); return SB_OK; } // header only
  if (!x4 || !v4) { std::fclose(f); return fail(SB_E_ARG, "x4 and v4 must both be given (or both NULL for the header only)"); }
  if (capacity < h.n_verts) { std::fclose(f); return fail(SB_E_ARG, "snapshot holds more vertices than the buffers"); }
  const size_t bytes = 16 * (size_t)h.n_verts;
  uint64_t sum = 0;
  const bool ok = std::fread(x4, 1, bytes, f) == bytes && std::fread(v4, 1, bytes, f) == bytes && std::fread(&sum, sizeof sum, 1, f) == 1;
  std::fclose(f);
  if (!ok) return fail(SB_E_ARG, "snapshot is truncated");
  uint64_t want = fnv1a(&h, sizeof h);
  want = fnv1a(x4, bytes, want);
  want = fnv1a(v4, bytes, want);
  if (want != sum) return fail(SB_E_ARG, "snapshot checksum mismatch");
  return SB_OK;
}

} // extern "C"
