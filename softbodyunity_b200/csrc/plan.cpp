// plan.cpp -- see plan.h.  Compiled with -ffp-contract=off: the rest values use the
// same operation order as the kernels (explicit fmaf only).
#include "plan.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <numeric>
#include <thread>

namespace sb {

namespace {

int resolve_threads(int t) {
  if (t > 0) return t;
  unsigned hc = std::thread::hardware_concurrency();
  return hc ? (int)std::min(hc, 32u) : 4;
}

// Dynamic-chunk parallel loop over [0, n).
template <class F>
void parallel_for(size_t n, int threads, size_t chunk, F fn) {
  if (threads <= 1 || n <= chunk) {
    for (size_t i = 0; i < n; i++) fn(i, 0);
    return;
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  int nt = (int)std::min<size_t>((size_t)threads, (n + chunk - 1) / chunk);
  for (int w = 0; w < nt; w++)
    pool.emplace_back([&, w]() {
      for (;;) {
        size_t lo = next.fetch_add(chunk);
        if (lo >= n) break;
        size_t hi = std::min(n, lo + chunk);
        for (size_t i = lo; i < hi; i++) fn(i, w);
      }
    });
  for (auto &t : pool) t.join();
}

inline uint32_t f2u(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}

// ---- canonical topology ------------------------------------------------------

std::string build_edges(Plan &P, int threads) {
  static const int pr[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
  const uint32_t V = P.V, T = P.T;
  std::vector<uint64_t> off((size_t)V + 1, 0);
  for (uint32_t t = 0; t < T; t++) {
    const int32_t *q = &P.tets[4 * (size_t)t];
    for (int j = 0; j < 4; j++)
      if (q[j] < 0 || (uint32_t)q[j] >= V) return "tet vertex index out of range";
    for (int k = 0; k < 6; k++) {
      int32_t a = q[pr[k][0]], b = q[pr[k][1]];
      if (a == b) return "tet with a repeated vertex";
      off[(size_t)std::min(a, b) + 1]++;
    }
  }
  for (uint32_t v = 0; v < V; v++) off[v + 1] += off[v];
  std::vector<int32_t> nb(off[V]);
  {
    std::vector<uint64_t> cur(off.begin(), off.end() - 1);
    for (uint32_t t = 0; t < T; t++) {
      const int32_t *q = &P.tets[4 * (size_t)t];
      for (int k = 0; k < 6; k++) {
        int32_t a = q[pr[k][0]], b = q[pr[k][1]];
        if (a > b) std::swap(a, b);
        nb[cur[a]++] = b;
      }
    }
  }
  std::vector<uint32_t> ucnt(V);
  parallel_for(V, threads, 16384, [&](size_t v, int) {
    auto lo = nb.begin() + off[v], hi = nb.begin() + off[v + 1];
    std::sort(lo, hi);
    ucnt[v] = (uint32_t)(std::unique(lo, hi) - lo);
  });
  std::vector<uint64_t> eoff((size_t)V + 1, 0);
  for (uint32_t v = 0; v < V; v++) eoff[v + 1] = eoff[v] + ucnt[v];
  if (eoff[V] >= 0x7fffffffull) return "too many edges";
  P.E = (uint32_t)eoff[V];
  P.edges.resize(2 * (size_t)P.E);
  parallel_for(V, threads, 16384, [&](size_t v, int) {
    for (uint32_t k = 0; k < ucnt[v]; k++) {
      P.edges[2 * (eoff[v] + k)] = (int32_t)v;
      P.edges[2 * (eoff[v] + k) + 1] = nb[off[v] + k];
    }
  });
  return "";
}

void lumped_inv_mass(Plan &P, float density) {
  // m_i = sum over tets (ascending) of (density * |det|/6) * 0.25, in double.
  std::vector<double> m(P.V, 0.0);
  for (uint32_t t = 0; t < P.T; t++) {
    const int32_t *q = &P.tets[4 * (size_t)t];
    double p[4][3];
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 3; k++) p[j][k] = (double)P.pos[3 * (size_t)q[j] + k];
    double e1[3], e2[3], e3[3];
    for (int k = 0; k < 3; k++) {
      e1[k] = p[1][k] - p[0][k];
      e2[k] = p[2][k] - p[0][k];
      e3[k] = p[3][k] - p[0][k];
    }
    double cx = e2[1] * e3[2] - e2[2] * e3[1];
    double cy = e2[2] * e3[0] - e2[0] * e3[2];
    double cz = e2[0] * e3[1] - e2[1] * e3[0];
    double det = e1[0] * cx + e1[1] * cy + e1[2] * cz;
    double share = ((double)density * (std::fabs(det) / 6.0)) * 0.25;
    for (int j = 0; j < 4; j++) m[q[j]] += share;
  }
  P.inv_mass.resize(P.V);
  for (uint32_t i = 0; i < P.V; i++) P.inv_mass[i] = m[i] > 0 ? (float)(1.0 / m[i]) : 0.0f;
}

inline void cross_c(float *o, const float *a, const float *b) {
  o[0] = std::fmaf(a[1], b[2], -(a[2] * b[1]));
  o[1] = std::fmaf(a[2], b[0], -(a[0] * b[2]));
  o[2] = std::fmaf(a[0], b[1], -(a[1] * b[0]));
}

void rest_values(Plan &P, int threads) {
  P.rest_len.resize(P.E);
  P.rest_vol6.resize(P.T);
  parallel_for(P.E, threads, 65536, [&](size_t e, int) {
    const float *a = &P.pos[3 * (size_t)P.edges[2 * e]], *b = &P.pos[3 * (size_t)P.edges[2 * e + 1]];
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    P.rest_len[e] = std::sqrt(std::fmaf(dz, dz, std::fmaf(dy, dy, dx * dx)));
  });
  parallel_for(P.T, threads, 65536, [&](size_t t, int) {
    const int32_t *q = &P.tet_roles[4 * t];
    const float *p0 = &P.pos[3 * (size_t)q[0]], *p1 = &P.pos[3 * (size_t)q[1]];
    const float *p2 = &P.pos[3 * (size_t)q[2]], *p3 = &P.pos[3 * (size_t)q[3]];
    float e1[3], e2[3], e3[3], G1[3];
    for (int k = 0; k < 3; k++) {
      e1[k] = p1[k] - p0[k];
      e2[k] = p2[k] - p0[k];
      e3[k] = p3[k] - p0[k];
    }
    cross_c(G1, e2, e3);
    P.rest_vol6[t] = std::fmaf(e1[2], G1[2], std::fmaf(e1[1], G1[1], e1[0] * G1[0]));
  });
}

// Edge ownership ("compounds").  Every edge of a tet joins two of its four vertices, so a thread that holds
// a tet's vertices in registers can project that edge as well, at no shared-memory cost.  Each edge is
// attached to at most one tet, each tet takes at most two edges and they must be opposite ones (vertex-
// disjoint): they become the edges between roles (0,1) and (2,3) after an even permutation of the tet's
// vertices (swapping the two ends of an edge is bit-neutral for the edge and flips the permutation's
// parity, so an even one always exists and the tet keeps its orientation).  Greedy, scarcest edges first.
void attach_edges(Plan &P, bool enable) {
  static const int pr[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
  const uint32_t T = P.T, E = P.E;
  P.tet_roles = P.tets;
  P.tet_e01.assign(T, -1);
  P.tet_e23.assign(T, -1);
  P.edge_owner.assign(E, -1);
  P.edges_attached = 0;
  P.tet_mate.assign(T, -1);
  P.tet_lead.assign(T, 1);
  if (!enable || !T || !E) return;
  // first edge of every vertex a in the sorted list (a < b)
  std::vector<uint32_t> first((size_t)P.V + 1, 0);
  for (uint32_t e = 0; e < E; e++) first[(size_t)P.edges[2 * (size_t)e] + 1]++;
  for (uint32_t v = 0; v < P.V; v++) first[v + 1] += first[v];
  auto edge_id = [&](int32_t a, int32_t b) -> int32_t {
    if (a > b) std::swap(a, b);
    uint32_t lo = first[a], hi = first[(size_t)a + 1];
    while (lo < hi) {
      const uint32_t mid = (lo + hi) / 2;
      if (P.edges[2 * (size_t)mid + 1] < b) lo = mid + 1;
      else hi = mid;
    }
    return lo < first[(size_t)a + 1] && P.edges[2 * (size_t)lo + 1] == b ? (int32_t)lo : -1;
  };
  // incidence: edge -> (tet, which of its six vertex pairs)
  std::vector<int32_t> te(6 * (size_t)T);
  std::vector<uint32_t> ioff((size_t)E + 1, 0);
  for (uint32_t t = 0; t < T; t++)
    for (int k = 0; k < 6; k++) {
      const int32_t e = edge_id(P.tets[4 * (size_t)t + pr[k][0]], P.tets[4 * (size_t)t + pr[k][1]]);
      te[6 * (size_t)t + k] = e;
      if (e >= 0) ioff[(size_t)e + 1]++;
    }
  uint32_t max_inc = 0;
  for (uint32_t e = 0; e < E; e++) {
    max_inc = std::max(max_inc, ioff[e + 1]);
    ioff[e + 1] += ioff[e];
  }
  std::vector<uint32_t> inc(ioff[E]); // tet * 8 + pair
  {
    std::vector<uint32_t> cur(ioff.begin(), ioff.end() - 1);
    for (uint32_t t = 0; t < T; t++)
      for (int k = 0; k < 6; k++)
        if (te[6 * (size_t)t + k] >= 0) inc[cur[te[6 * (size_t)t + k]]++] = t * 8u + (uint32_t)k;
  }
  // edges by ascending number of incident tets (counting sort, stable in edge id)
  std::vector<uint32_t> by_inc(E);
  {
    std::vector<uint32_t> cnt((size_t)max_inc + 2, 0);
    for (uint32_t e = 0; e < E; e++) cnt[(size_t)(ioff[e + 1] - ioff[e]) + 1]++;
    for (uint32_t k = 0; k <= max_inc; k++) cnt[k + 1] += cnt[k];
    for (uint32_t e = 0; e < E; e++) by_inc[cnt[ioff[e + 1] - ioff[e]]++] = e;
  }
  std::vector<int8_t> pair01(T, -1), pair23(T, -1);
  // partitioned meshes: an edge rides only with a tet of its own constraint group (both touch a ghost vertex
  // or neither does), and edges / tets among ghosts only belong to another rank
  const uint32_t first_ghost = P.V - P.n_ghost;
  auto n_ghosts = [&](const int32_t *v, int n) {
    int g = 0;
    for (int k = 0; k < n; k++) g += (uint32_t)v[k] >= first_ghost;
    return g;
  };
  std::vector<uint8_t> tet_cut(P.n_ghost ? T : 0);
  for (uint32_t t = 0; t < (uint32_t)tet_cut.size(); t++) {
    const int g = n_ghosts(&P.tets[4 * (size_t)t], 4);
    tet_cut[t] = g == 0 ? 0 : g == 4 ? 2 : 1;
  }
  for (uint32_t oi = 0; oi < E; oi++) {
    const uint32_t e = by_inc[oi];
    uint8_t e_cut = 0;
    if (P.n_ghost) {
      const int g = n_ghosts(&P.edges[2 * (size_t)e], 2);
      if (g == 2) continue;
      e_cut = (uint8_t)g;
    }
    auto same_group = [&](uint32_t t) { return !P.n_ghost || tet_cut[t] == e_cut; };
    int64_t pick = -1;
    for (uint32_t k = ioff[e]; k < ioff[e + 1] && pick < 0; k++)
      if (pair01[inc[k] >> 3] < 0 && same_group(inc[k] >> 3)) pick = inc[k];
    if (pick >= 0) {
      pair01[pick >> 3] = (int8_t)(pick & 7);
      P.tet_e01[pick >> 3] = (int32_t)e;
      P.edge_owner[e] = (int32_t)(pick >> 3);
      continue;
    }
    for (uint32_t k = ioff[e]; k < ioff[e + 1] && pick < 0; k++) {
      const uint32_t t = inc[k] >> 3, pk = inc[k] & 7;
      if (pair23[t] < 0 && pair01[t] == (int8_t)(5 - pk) && same_group(t)) pick = inc[k]; // pair 5 - k is the opposite edge
    }
    if (pick >= 0) {
      pair23[pick >> 3] = (int8_t)(pick & 7);
      P.tet_e23[pick >> 3] = (int32_t)e;
      P.edge_owner[e] = (int32_t)(pick >> 3);
    }
  }
  // Augmentation.  The greedy pass leaves ~8 % of the edges of a lattice mesh without a tet although slots remain:
  // an edge left over takes the slot of an edge that can move to another of its own tets (an augmenting path of
  // length two in the edge -> slot matching).  Fewer free edges = fewer edge rounds per tile.
  static const int augment = getenv("SB_ATTACH_AUGMENT") ? atoi(getenv("SB_ATTACH_AUGMENT")) : 1; // depth of the search (0: off, A/B switch for benches)
  if (augment) {
    auto group_ok = [&](uint32_t t, uint32_t e) {
      if (!P.n_ghost) return true;
      return tet_cut[t] == (uint8_t)n_ghosts(&P.edges[2 * (size_t)e], 2);
    };
    // Put edge e into a slot of one of its tets other than `not_t`: a free compatible slot if there is one, else
    // (depth > 0) the slot of an occupant that can itself be placed elsewhere -- an augmenting path, depth first.
    std::function<bool(uint32_t, uint32_t, int)> place = [&](uint32_t e, uint32_t not_t, int depth) -> bool {
      for (uint32_t k = ioff[e]; k < ioff[e + 1]; k++) {
        const uint32_t t = inc[k] >> 3, pk = inc[k] & 7;
        if (t == not_t || !group_ok(t, e)) continue;
        if (pair01[t] < 0) {
          pair01[t] = (int8_t)pk; P.tet_e01[t] = (int32_t)e; P.edge_owner[e] = (int32_t)t;
          return true;
        }
        if (pair23[t] < 0 && pair01[t] == (int8_t)(5 - pk)) {
          pair23[t] = (int8_t)pk; P.tet_e23[t] = (int32_t)e; P.edge_owner[e] = (int32_t)t;
          return true;
        }
      }
      if (depth <= 0) return false;
      for (uint32_t k = ioff[e]; k < ioff[e + 1]; k++) {
        const uint32_t t = inc[k] >> 3, pk = inc[k] & 7;
        if (t == not_t || !group_ok(t, e) || pair01[t] < 0) continue;
        if (pair23[t] < 0) {
          // t holds one edge e1 (not opposite to e, or e would have fitted): e1 moves on, e takes its slot
          const uint32_t e1 = (uint32_t)P.tet_e01[t];
          const int8_t p1 = pair01[t];
          pair01[t] = (int8_t)pk; P.tet_e01[t] = (int32_t)e; P.edge_owner[e] = (int32_t)t; // (so that e1 cannot come back here)
          P.edge_owner[e1] = -1;
          if (place(e1, t, depth - 1)) return true;
          pair01[t] = p1; P.tet_e01[t] = (int32_t)e1; P.edge_owner[e1] = (int32_t)t; P.edge_owner[e] = -1;
        } else if (pair23[t] == (int8_t)(5 - pk) || pair01[t] == (int8_t)(5 - pk)) {
          // t is full and one of its two edges is opposite to e: the OTHER one moves on and e takes its slot
          const bool keep01 = pair01[t] == (int8_t)(5 - pk);
          const uint32_t e_keep = (uint32_t)(keep01 ? P.tet_e01[t] : P.tet_e23[t]);
          const uint32_t e_go = (uint32_t)(keep01 ? P.tet_e23[t] : P.tet_e01[t]);
          const int8_t p_keep = keep01 ? pair01[t] : pair23[t], p_go = keep01 ? pair23[t] : pair01[t];
          pair01[t] = p_keep; P.tet_e01[t] = (int32_t)e_keep;
          pair23[t] = (int8_t)pk; P.tet_e23[t] = (int32_t)e; P.edge_owner[e] = (int32_t)t;
          P.edge_owner[e_go] = -1;
          if (place(e_go, t, depth - 1)) return true;
          pair23[t] = p_go; P.tet_e23[t] = (int32_t)e_go; P.edge_owner[e_go] = (int32_t)t; P.edge_owner[e] = -1;
        }
      }
      return false;
    };
    for (uint32_t oi = 0; oi < E; oi++) {
      const uint32_t e = by_inc[oi];
      if (P.edge_owner[e] >= 0) continue;
      if (P.n_ghost && n_ghosts(&P.edges[2 * (size_t)e], 2) == 2) continue;
      place(e, 0xffffffffu, augment);
    }
  }
  // roles: (a, b, c, d) with (a, b) the first attached edge and (c, d) the other two vertices in their
  // original relative order; if that permutation is odd, swap a and b
  for (uint32_t t = 0; t < T; t++) {
    if (pair01[t] < 0) continue;
    const int i0 = pr[pair01[t]][0], i1 = pr[pair01[t]][1];
    int perm[4] = {i0, i1, -1, -1};
    int at = 2;
    for (int j = 0; j < 4; j++)
      if (j != i0 && j != i1) perm[at++] = j;
    int inversions = 0;
    for (int x = 0; x < 4; x++)
      for (int y = x + 1; y < 4; y++) inversions += perm[x] > perm[y];
    if (inversions & 1) std::swap(perm[0], perm[1]);
    for (int j = 0; j < 4; j++) P.tet_roles[4 * (size_t)t + j] = P.tets[4 * (size_t)t + perm[j]];
    P.edges_attached += 1 + (pair23[t] >= 0);
  }
}

// ---- bi-tets ---------------------------------------------------------------------------------------------
//
// Two tets that share a face have five vertices between them.  A thread that projects both keeps the three shared
// vertices in registers: 5 loads + 5 stores of shared memory for two tets instead of 8 + 8, half the records, rounds
// and barriers.  The pair is one COMPOUND: tet A = (a, s0, s1, s2), tet B = (b, s1, s0, s2) -- the apex first, the
// shared face in the cyclic order that makes both an even permutation of the caller's tets (A and B lie on opposite
// sides of the face, hence the swap) -- and up to four edges ride along, between the roles (0,1) and (2,3) of each:
//   A: (a, s0) and (s1, s2)      B: (b, s1) and (s0, s2)
// Rotating the face (s0 -> s1 -> s2 -> s0) keeps the parities, so a pair offers three such slot sets and the
// attachment below picks the rotation that fits the edges it wants to place.  The kernel projects A, its edges, B,
// its edges, in this order; the exported schedule says the same.
//
// tet_mate[t] = partner or -1; tet_lead[t] = 1 for a single tet or the A of a pair, 0 for a B.
void pair_and_attach(Plan &P, int threads, bool attach) {
  static const int pr[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
  const uint32_t T = P.T, E = P.E, V = P.V;
  P.tet_roles = P.tets;
  P.tet_e01.assign(T, -1);
  P.tet_e23.assign(T, -1);
  P.edge_owner.assign(E, -1);
  P.edges_attached = 0;
  P.tet_mate.assign(T, -1);
  P.tet_lead.assign(T, 1);
  if (!T) return;
  // ---- face neighbours through the tets around each tet's vertices ----
  std::vector<uint32_t> voff((size_t)V + 1, 0), vt(4 * (size_t)T);
  for (size_t i = 0; i < 4 * (size_t)T; i++) voff[(size_t)P.tets[i] + 1]++;
  for (uint32_t v = 0; v < V; v++) voff[v + 1] += voff[v];
  {
    std::vector<uint32_t> cur(voff.begin(), voff.end() - 1);
    for (uint32_t t = 0; t < T; t++)
      for (int j = 0; j < 4; j++) vt[cur[P.tets[4 * (size_t)t + j]]++] = t; // ascending tet id per vertex
  }
  std::vector<int32_t> nb(4 * (size_t)T, -1); // nb[4 t + j] = the tet across the face opposite vertex j
  parallel_for(T, threads, 4096, [&](size_t t, int) {
    const int32_t *q = &P.tets[4 * t];
    for (int j = 0; j < 4; j++) {
      const int32_t x = q[(j + 1) & 3], y = q[(j + 2) & 3], z = q[(j + 3) & 3];
      for (uint32_t k = voff[x]; k < voff[(size_t)x + 1]; k++) {
        const uint32_t u = vt[k];
        if (u == t) continue;
        const int32_t *r = &P.tets[4 * (size_t)u];
        const bool hy = r[0] == y || r[1] == y || r[2] == y || r[3] == y;
        const bool hz = r[0] == z || r[1] == z || r[2] == z || r[3] == z;
        if (hy && hz) { nb[4 * t + j] = (int32_t)u; break; }
      }
    }
  });
  vt = std::vector<uint32_t>();
  // The pair pattern for (A = t, apex index ja; B = u): the shared face in the cyclic order that makes (a, s0, s1, s2)
  // an even permutation of A; usable when (b, s1, s0, s2) is then an even permutation of B (always, in a
  // consistently oriented mesh).
  auto pattern = [&](uint32_t t, int ja, uint32_t u, int32_t *out5) -> bool { // out5 = a, s0, s1, s2, b
    const int32_t *q = &P.tets[4 * (size_t)t], *r = &P.tets[4 * (size_t)u];
    int32_t s[3];
    int n = 0;
    for (int j = 0; j < 4; j++)
      if (j != ja) s[n++] = q[j];
    if (ja & 1) std::swap(s[1], s[2]); // moving the apex to the front is ja transpositions
    int jb = -1;
    for (int j = 0; j < 4; j++)
      if (r[j] != s[0] && r[j] != s[1] && r[j] != s[2]) jb = jb < 0 ? j : 4;
    if (jb < 0 || jb > 3) return false;
    const int32_t want[4] = {r[jb], s[1], s[0], s[2]};
    int idx[4];
    for (int k = 0; k < 4; k++) {
      idx[k] = -1;
      for (int j = 0; j < 4; j++)
        if (r[j] == want[k]) idx[k] = j;
      if (idx[k] < 0) return false;
    }
    int inv = 0;
    for (int x = 0; x < 4; x++)
      for (int y = x + 1; y < 4; y++) inv += idx[x] > idx[y];
    if (inv & 1) return false;
    out5[0] = q[ja]; out5[1] = s[0]; out5[2] = s[1]; out5[3] = s[2]; out5[4] = r[jb];
    return true;
  };
  // ---- greedy matching on the face-adjacency graph: a tet takes the free neighbour that has the fewest free
  // neighbours itself (ties: the lowest id); sequential, hence independent of the thread count ----
  static const int pairs_on = getenv("SB_BITETS") ? atoi(getenv("SB_BITETS")) : 1;
  std::vector<int8_t> apex_of(T, -1); // for a leader with a mate: index of its apex in the caller's tet
  if (pairs_on && !P.n_ghost) {
    auto free_deg = [&](uint32_t u) {
      int d = 0;
      for (int j = 0; j < 4; j++) d += nb[4 * (size_t)u + j] >= 0 && P.tet_mate[nb[4 * (size_t)u + j]] < 0;
      return d;
    };
    for (uint32_t t = 0; t < T; t++) {
      if (P.tet_mate[t] >= 0) continue;
      int best_j = -1, best_d = 99;
      for (int j = 0; j < 4; j++) {
        const int32_t u = nb[4 * (size_t)t + j];
        if (u < 0 || P.tet_mate[u] >= 0) continue;
        int32_t five[5];
        if (!pattern(t, j, (uint32_t)u, five)) continue;
        const int d = free_deg((uint32_t)u);
        if (d < best_d || (d == best_d && u < nb[4 * (size_t)t + best_j])) { best_d = d; best_j = j; }
      }
      if (best_j < 0) continue;
      const uint32_t u = (uint32_t)nb[4 * (size_t)t + best_j];
      P.tet_mate[t] = (int32_t)u;
      P.tet_mate[u] = (int32_t)t;
      P.tet_lead[u] = 0;
      apex_of[t] = (int8_t)best_j;
    }
  }
  nb = std::vector<int32_t>();
  for (uint32_t t = 0; t < T; t++) P.tets_paired += P.tet_mate[t] >= 0;
  // ---- edges -> compounds ----
  std::vector<uint32_t> first((size_t)V + 1, 0);
  for (uint32_t e = 0; e < E; e++) first[(size_t)P.edges[2 * (size_t)e] + 1]++;
  for (uint32_t v = 0; v < V; v++) first[v + 1] += first[v];
  auto edge_id = [&](int32_t a, int32_t b) -> int32_t {
    if (a > b) std::swap(a, b);
    uint32_t lo = first[a], hi = first[(size_t)a + 1];
    while (lo < hi) {
      const uint32_t mid = (lo + hi) / 2;
      if (P.edges[2 * (size_t)mid + 1] < b) lo = mid + 1;
      else hi = mid;
    }
    return lo < first[(size_t)a + 1] && P.edges[2 * (size_t)lo + 1] == b ? (int32_t)lo : -1;
  };
  // local edges of an owner (a leader): single tet: its six vertex pairs; pair: 0..2 (a, s_j), 3..5 (b, s_j),
  // 6..8 the face edge opposite s_j
  auto local_edges = [&](uint32_t t, int32_t (*ab)[2]) -> int {
    if (P.tet_mate[t] < 0) {
      for (int k = 0; k < 6; k++) { ab[k][0] = P.tets[4 * (size_t)t + pr[k][0]]; ab[k][1] = P.tets[4 * (size_t)t + pr[k][1]]; }
      return 6;
    }
    int32_t f[5];
    pattern(t, apex_of[t], (uint32_t)P.tet_mate[t], f);
    for (int j = 0; j < 3; j++) {
      ab[j][0] = f[0]; ab[j][1] = f[1 + j];
      ab[3 + j][0] = f[4]; ab[3 + j][1] = f[1 + j];
      ab[6 + j][0] = f[1 + (j + 1) % 3]; ab[6 + j][1] = f[1 + (j + 2) % 3];
    }
    return 9;
  };
  // slot sets: a compound may hold any subset of ONE of its options
  static const uint16_t opt_single[3] = {(1 << 0) | (1 << 5), (1 << 1) | (1 << 4), (1 << 2) | (1 << 3)};
  uint16_t opt_pair[3];
  for (int i = 0; i < 3; i++) opt_pair[i] = (uint16_t)((1 << i) | (1 << (3 + (i + 1) % 3)) | (1 << (6 + i)) | (1 << (6 + (i + 1) % 3)));
  auto feasible = [&](uint32_t t, uint16_t mask) {
    const uint16_t *o = P.tet_mate[t] < 0 ? opt_single : opt_pair;
    return (mask & ~o[0]) == 0 || (mask & ~o[1]) == 0 || (mask & ~o[2]) == 0;
  };
  // incidence: edge -> (owner, local edge)
  std::vector<uint32_t> ioff((size_t)E + 1, 0);
  for (uint32_t t = 0; t < T && attach; t++) {
    if (!P.tet_lead[t]) continue;
    int32_t ab[9][2];
    const int n = local_edges(t, ab);
    for (int k = 0; k < n; k++) {
      const int32_t e = edge_id(ab[k][0], ab[k][1]);
      if (e >= 0) ioff[(size_t)e + 1]++;
    }
  }
  uint32_t max_inc = 0;
  for (uint32_t e = 0; e < E; e++) {
    max_inc = std::max(max_inc, ioff[e + 1]);
    ioff[e + 1] += ioff[e];
  }
  std::vector<uint64_t> inc(ioff[E]); // owner * 16 + local edge
  {
    std::vector<uint32_t> cur(ioff.begin(), ioff.end() - 1);
    for (uint32_t t = 0; t < T && attach; t++) {
      if (!P.tet_lead[t]) continue;
      int32_t ab[9][2];
      const int n = local_edges(t, ab);
      for (int k = 0; k < n; k++) {
        const int32_t e = edge_id(ab[k][0], ab[k][1]);
        if (e >= 0) inc[cur[e]++] = (uint64_t)t * 16u + (uint32_t)k;
      }
    }
  }
  std::vector<uint32_t> by_inc(E);
  {
    std::vector<uint32_t> cnt((size_t)max_inc + 2, 0);
    for (uint32_t e = 0; e < E; e++) cnt[(size_t)(ioff[e + 1] - ioff[e]) + 1]++;
    for (uint32_t k = 0; k <= max_inc; k++) cnt[k + 1] += cnt[k];
    for (uint32_t e = 0; e < E; e++) by_inc[cnt[ioff[e + 1] - ioff[e]]++] = e;
  }
  std::vector<uint16_t> taken(T, 0);
  std::vector<int8_t> slot_of(E, -1); // local edge the edge occupies in its owner (edge_owner = the LEADER here)
  auto put = [&](uint32_t e, uint32_t t, int l) { taken[t] |= (uint16_t)(1u << l); P.edge_owner[e] = (int32_t)t; slot_of[e] = (int8_t)l; };
  auto drop = [&](uint32_t e) { taken[P.edge_owner[e]] &= (uint16_t)~(1u << slot_of[e]); P.edge_owner[e] = -1; slot_of[e] = -1; };
  static const int augment = getenv("SB_ATTACH_AUGMENT") ? atoi(getenv("SB_ATTACH_AUGMENT")) : 1;
  // Put edge e into a compound other than `not_t`: where it fits as things are, else (depth > 0) in the place of an
  // occupant that can itself move elsewhere (an augmenting path in the edge -> slot matching, depth first).
  std::function<bool(uint32_t, uint32_t, int)> place = [&](uint32_t e, uint32_t not_t, int depth) -> bool {
    for (uint32_t k = ioff[e]; k < ioff[e + 1]; k++) {
      const uint32_t t = (uint32_t)(inc[k] >> 4);
      const int l = (int)(inc[k] & 15);
      if (t == not_t) continue;
      if (feasible(t, (uint16_t)(taken[t] | (1u << l)))) { put(e, t, l); return true; }
    }
    if (depth <= 0) return false;
    for (uint32_t k = ioff[e]; k < ioff[e + 1]; k++) {
      const uint32_t t = (uint32_t)(inc[k] >> 4);
      const int l = (int)(inc[k] & 15);
      if (t == not_t) continue;
      int32_t ab[9][2];
      const int n = local_edges(t, ab);
      for (int l1 = 0; l1 < n; l1++) {
        if (!(taken[t] >> l1 & 1)) continue;
        if (!feasible(t, (uint16_t)((taken[t] & ~(1u << l1)) | (1u << l)))) continue;
        const int32_t e1 = edge_id(ab[l1][0], ab[l1][1]);
        drop((uint32_t)e1);
        put(e, t, l); // (so that e1 cannot come back here)
        if (place((uint32_t)e1, t, depth - 1)) return true;
        drop(e);
        put((uint32_t)e1, t, l1);
      }
    }
    return false;
  };
  for (uint32_t oi = 0; oi < E && attach; oi++) place(by_inc[oi], 0xffffffffu, 0);
  if (augment && attach)
    for (uint32_t oi = 0; oi < E; oi++)
      if (P.edge_owner[by_inc[oi]] < 0) place(by_inc[oi], 0xffffffffu, augment);
  // ---- roles and attached edges per tet ----
  for (uint32_t t = 0; t < T; t++) {
    if (!P.tet_lead[t]) continue;
    int32_t ab[9][2];
    const int n = local_edges(t, ab);
    auto held = [&](int l) -> int32_t { return (taken[t] >> l & 1) ? edge_id(ab[l][0], ab[l][1]) : -1; };
    if (P.tet_mate[t] < 0) {
      if (!taken[t]) continue;
      int k0 = -1;
      for (int k = 0; k < 6 && k0 < 0; k++)
        if (taken[t] >> k & 1) k0 = k;
      // roles: (a, b, c, d) with (a, b) the first attached edge and (c, d) the other two vertices in their
      // original relative order; if that permutation is odd, swap a and b
      const int i0 = pr[k0][0], i1 = pr[k0][1];
      int perm[4] = {i0, i1, -1, -1};
      int at = 2;
      for (int j = 0; j < 4; j++)
        if (j != i0 && j != i1) perm[at++] = j;
      int inversions = 0;
      for (int x = 0; x < 4; x++)
        for (int y = x + 1; y < 4; y++) inversions += perm[x] > perm[y];
      if (inversions & 1) std::swap(perm[0], perm[1]);
      for (int j = 0; j < 4; j++) P.tet_roles[4 * (size_t)t + j] = P.tets[4 * (size_t)t + perm[j]];
      P.tet_e01[t] = held(k0);
      P.tet_e23[t] = held(5 - k0);
      (void)n;
    } else {
      const uint32_t u = (uint32_t)P.tet_mate[t];
      int rot = 0;
      for (int i = 2; i >= 0; i--)
        if ((taken[t] & ~opt_pair[i]) == 0) rot = i;
      int32_t f[5];
      pattern(t, apex_of[t], u, f);
      const int32_t s0 = f[1 + rot], s1 = f[1 + (rot + 1) % 3], s2 = f[1 + (rot + 2) % 3];
      int32_t *ra = &P.tet_roles[4 * (size_t)t], *rb = &P.tet_roles[4 * (size_t)u];
      ra[0] = f[0]; ra[1] = s0; ra[2] = s1; ra[3] = s2;
      rb[0] = f[4]; rb[1] = s1; rb[2] = s0; rb[3] = s2;
      P.tet_e01[t] = held(rot);                    // (a, s0)
      P.tet_e23[t] = held(6 + rot);                // (s1, s2): the face edge opposite s0
      P.tet_e01[u] = held(3 + (rot + 1) % 3);      // (b, s1)
      P.tet_e23[u] = held(6 + (rot + 1) % 3);      // (s0, s2): the face edge opposite s1
    }
  }
  // edge_owner: the tet whose roles carry the edge
  for (uint32_t e = 0; e < E; e++) P.edge_owner[e] = -1;
  for (uint32_t t = 0; t < T; t++) {
    if (P.tet_e01[t] >= 0) { P.edge_owner[P.tet_e01[t]] = (int32_t)t; P.edges_attached++; }
    if (P.tet_e23[t] >= 0) { P.edge_owner[P.tet_e23[t]] = (int32_t)t; P.edges_attached++; }
  }
}

std::string build_surface(Plan &P) {
  const uint32_t V = P.V, F = P.F;
  std::vector<uint32_t> cnt((size_t)V + 1, 0);
  for (size_t i = 0; i < 3 * (size_t)F; i++) {
    if (P.tris[i] < 0 || (uint32_t)P.tris[i] >= V) return "surface triangle index out of range";
    cnt[P.tris[i]]++;
  }
  std::vector<int32_t> sidx(V, -1);
  P.surf_ids.clear();
  for (uint32_t v = 0; v < V; v++)
    if (cnt[v]) {
      sidx[v] = (int32_t)P.surf_ids.size();
      P.surf_ids.push_back((int32_t)v);
    }
  size_t ns = P.surf_ids.size();
  P.surf_tri_off.assign(ns + 1, 0);
  for (size_t s = 0; s < ns; s++) P.surf_tri_off[s + 1] = P.surf_tri_off[s] + cnt[P.surf_ids[s]];
  P.surf_tri_ids.resize(P.surf_tri_off[ns]);
  std::vector<uint32_t> cur(P.surf_tri_off.begin(), P.surf_tri_off.end() - 1);
  for (uint32_t f = 0; f < F; f++)
    for (int j = 0; j < 3; j++) {
      int32_t s = sidx[P.tris[3 * (size_t)f + j]];
      // a triangle listing the same vertex twice contributes twice, like the sequential sum does
      P.surf_tri_ids[cur[s]++] = f;
    }
  return "";
}

// ---- vertex tiling -----------------------------------------------------------

struct Uf {
  std::vector<uint32_t> p;
  explicit Uf(uint32_t n) : p(n) { std::iota(p.begin(), p.end(), 0u); }
  uint32_t find(uint32_t x) {
    while (p[x] != x) {
      p[x] = p[p[x]];
      x = p[x];
    }
    return x;
  }
  void unite(uint32_t a, uint32_t b) {
    a = find(a);
    b = find(b);
    if (a != b) p[std::max(a, b)] = std::min(a, b); // root = smallest id of the component
  }
};

// Recursive coordinate bisection of idx[lo,hi) into K parts of near-equal size
// along the longest axis; appends the part boundaries (as end offsets) to `ends`.
void rcb(std::vector<uint32_t> &idx, size_t lo, size_t hi, uint32_t K, const float *pos3,
         std::vector<size_t> &ends) {
  if (K <= 1 || hi - lo <= 1) {
    ends.push_back(hi);
    return;
  }
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (size_t i = lo; i < hi; i++)
    for (int k = 0; k < 3; k++) {
      float c = pos3[3 * (size_t)idx[i] + k];
      mn[k] = std::min(mn[k], c);
      mx[k] = std::max(mx[k], c);
    }
  int ax = 0;
  for (int k = 1; k < 3; k++)
    if (mx[k] - mn[k] > mx[ax] - mn[ax]) ax = k;
  uint32_t KL = K / 2;
  size_t nL = (size_t)(((unsigned __int128)(hi - lo) * KL + K / 2) / K);
  nL = std::min(std::max<size_t>(nL, 1), hi - lo - 1);
  auto cmp = [&](uint32_t a, uint32_t b) {
    float ca = pos3[3 * (size_t)a + ax], cb = pos3[3 * (size_t)b + ax];
    return ca < cb || (ca == cb && a < b);
  };
  std::nth_element(idx.begin() + lo, idx.begin() + lo + nL, idx.begin() + hi, cmp);
  rcb(idx, lo, lo + nL, KL, pos3, ends);
  rcb(idx, lo + nL, hi, K - KL, pos3, ends);
}

// First-level tiling in the caller's numbering.  Fills P.perm / P.inv and returns
// the tile boundaries in device numbering.
std::vector<uint32_t> tile_vertices(Plan &P, uint32_t cap, int n_sm) {
  const uint32_t V = P.V;
  Uf uf(V);
  for (uint32_t t = 0; t < P.T; t++) {
    const int32_t *q = &P.tets[4 * (size_t)t];
    uf.unite(q[0], q[1]);
    uf.unite(q[0], q[2]);
    uf.unite(q[0], q[3]);
  }
  // components in order of their smallest vertex id
  std::vector<uint32_t> root(V), csize(V, 0);
  for (uint32_t v = 0; v < V; v++) {
    root[v] = uf.find(v);
    csize[root[v]]++;
  }
  std::vector<uint64_t> coff((size_t)V + 1, 0);
  for (uint32_t v = 0; v < V; v++) coff[v + 1] = coff[v] + (root[v] == v ? csize[v] : 0);
  std::vector<uint32_t> members(V);
  {
    std::vector<uint64_t> cur(coff.begin(), coff.end() - 1);
    for (uint32_t v = 0; v < V; v++) members[cur[root[v]]++] = v; // ascending within a component
  }
  std::vector<uint32_t> tile_end; // ends in `members` order (which becomes the device order)
  size_t open_lo = 0, open_n = 0;
  auto flush = [&]() {
    if (open_n) tile_end.push_back((uint32_t)(open_lo + open_n));
    open_n = 0;
  };
  for (uint32_t r = 0; r < V; r++) {
    if (root[r] != r) continue;
    size_t lo = coff[r], n = csize[r];
    if (n <= cap) {
      if (open_n + n > cap) flush();
      if (!open_n) open_lo = lo;
      open_n += n;
    } else {
      flush();
      uint32_t K = (uint32_t)((n + cap - 1) / cap);
      if (K > (uint32_t)n_sm) K = ((K + n_sm - 1) / n_sm) * n_sm; // whole waves
      while ((n + K - 1) / K > cap) K++;
      std::vector<size_t> ends;
      rcb(members, lo, lo + n, K, P.pos.data(), ends);
      size_t prev = lo;
      for (size_t e : ends) {
        std::sort(members.begin() + prev, members.begin() + e);
        tile_end.push_back((uint32_t)e);
        prev = e;
      }
    }
  }
  flush();
  P.perm = members;
  P.inv.resize(V);
  for (uint32_t d = 0; d < V; d++) P.inv[P.perm[d]] = d;
  std::vector<uint32_t> off;
  off.push_back(0);
  for (uint32_t e : tile_end) off.push_back(e);
  return off;
}


// ---- balanced shifted grid tilings ----------------------------------------------
//
// N_T tilings of space by the same box grid, tiling s shifted by s/N_T of a box along
// every axis.  A constraint (an element spans far less than a quarter box) straddles
// the planes of at most one tiling per axis, so with 4 tilings it is interior to a
// tile of at least one; constraints are dealt to the tilings they are interior to so
// that every pass carries about 1/N_T of the work and 1/N_T of each vertex's colours.
struct Tiling {
  std::vector<int32_t> part; // caller numbering -> tile
  uint32_t n_tiles = 0;
};

// `atom_key` (per caller vertex) orders vertices inside a box of tiling 0 by the 1/N_T-box
// sub-box ("atom") they fall in, z-major, so that every box of every shifted tiling is a
// union of a few contiguous runs of the device numbering.
bool grid_tilings(const Plan &P, uint32_t cap, int n_tilings, std::vector<Tiling> &out, std::vector<uint32_t> &atom_key, int dist_ranks = 0) {
  const uint32_t V = P.V;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (uint32_t v = 0; v < V; v++)
    for (int k = 0; k < 3; k++) {
      lo[k] = std::min(lo[k], P.pos[3 * (size_t)v + k]);
      hi[k] = std::max(hi[k], P.pos[3 * (size_t)v + k]);
    }
  double ext[3], vol = 1.0;
  int flat = 0;
  for (int k = 0; k < 3; k++) {
    ext[k] = (double)hi[k] - lo[k];
    if (!(ext[k] > 0)) { ext[k] = 0; flat++; }
    else vol *= ext[k];
  }
  if (flat == 3) return false;
  // start from the box edge that would hold `cap` vertices at uniform density, then shrink
  // until no box of any tiling exceeds cap (bounded number of attempts)
  double w = std::pow((double)cap * vol / V, 1.0 / (3 - flat));
  std::vector<uint32_t> count;
  for (int attempt = 0; attempt < 24; attempt++, w *= 0.94) {
    int n[3];
    double wa[3];
    for (int k = 0; k < 3; k++) {
      n[k] = ext[k] > 0 ? std::max(1, (int)std::ceil(ext[k] / w - 1e-9)) : 1;
    }
    // One mesh over 2 / 4 / 8 GPUs: the boxes are later cut into compact blocks by recursive bisection along the
    // longest axis (make_plan).  An odd number of boxes along a cut axis would leave one side a whole layer of
    // boxes heavier (17 boxes: 8 | 9, i.e. 9^3 against an average of 8.5^3 boxes per rank at 8 ranks: +19 % on the
    // slowest rank), so the count along every axis is rounded up to a multiple of the cuts it will take.
    if (dist_ranks >= 2 && (dist_ranks & (dist_ranks - 1)) == 0) {
      double e[3] = {ext[0], ext[1], ext[2]};
      int mult[3] = {1, 1, 1};
      for (int r = dist_ranks; r > 1; r /= 2) {
        int ax = 2; // ties: z first, like the bisection
        for (int k = 1; k >= 0; k--)
          if (e[k] > e[ax]) ax = k;
        mult[ax] *= 2;
        e[ax] /= 2;
      }
      for (int k = 0; k < 3; k++)
        if (ext[k] > 0) n[k] = (n[k] + mult[k] - 1) / mult[k] * mult[k];
    }
    for (int k = 0; k < 3; k++) wa[k] = ext[k] > 0 ? ext[k] / n[k] : 1.0;
    if ((double)(n[0] + 1) * (n[1] + 1) * (n[2] + 1) > 64e6) return false;
    out.assign(n_tilings, Tiling());
    bool ok = true;
    // fine-grid ("atom") coordinates: boxes of every tiling are unions of atoms
    std::vector<int32_t> atom(3 * (size_t)V);
    atom_key.resize(V);
    // Order of the atoms inside a box of the unshifted tiling: z-major, but as a snake (every other row of x runs
    // backwards, every other slice of rows runs backwards in y).  The part of a shifted box that falls into one
    // unshifted box is a product of prefixes / suffixes of the atom ranges; in plain z-major order it is one run of
    // device ids per (z, y) row, in the snake two consecutive rows join where the snake turns: half the runs
    // (bulk copies) per tile.
    static const int snake = getenv("SB_ATOM_SNAKE") ? atoi(getenv("SB_ATOM_SNAKE")) : 1;
    for (uint32_t v = 0; v < V; v++) {
      uint32_t sub[3];
      for (int k = 2; k >= 0; k--) {
        const double f = ext[k] > 0 ? ((double)P.pos[3 * (size_t)v + k] - lo[k]) / wa[k] * n_tilings : 0.0;
        const int a = std::min(n[k] * n_tilings - 1, std::max(0, (int)std::floor(f)));
        atom[3 * (size_t)v + k] = a;
        sub[k] = (uint32_t)(a % n_tilings);
      }
      const uint32_t nt_ = (uint32_t)n_tilings;
      uint32_t zi = sub[2], yi = sub[1], xi = sub[0];
      if (snake) {
        if (zi & 1u) yi = nt_ - 1 - yi;
        if ((zi * nt_ + yi) & 1u) xi = nt_ - 1 - xi;
      }
      atom_key[v] = (zi * nt_ + yi) * nt_ + xi;
    }
    for (int s = 0; s < n_tilings && ok; s++) {
      const int m[3] = {n[0] + (s > 0), n[1] + (s > 0), n[2] + (s > 0)};
      count.assign((size_t)m[0] * m[1] * m[2], 0u);
      Tiling &T = out[s];
      T.part.resize(V);
      for (uint32_t v = 0; v < V; v++) {
        int c[3];
        for (int k = 0; k < 3; k++) c[k] = std::min(m[k] - 1, (atom[3 * (size_t)v + k] + s) / n_tilings);
        const uint32_t cell = (uint32_t)c[0] + (uint32_t)m[0] * ((uint32_t)c[1] + (uint32_t)m[1] * (uint32_t)c[2]);
        T.part[v] = (int32_t)cell;
        count[cell]++;
      }
      for (size_t c = 0; c < count.size(); c++)
        if (count[c] > cap) ok = false;
      // Rims.  A shifted tiling ends in slices of boxes (1/4, 1/2, 3/4 of a box thick, and their products at the
      // edges and corners): a quarter of its tiles hold a few hundred vertices, take a CTA slot each and run
      // valence-many nearly empty rounds.  Small boxes are merged with a face neighbour while the union stays
      // within the cap: a union of boxes of one tiling keeps every element that was interior to one of them (and
      // gains those that crossed the plane between them).
      static const int merge_rims = getenv("SB_MERGE_RIMS") ? atoi(getenv("SB_MERGE_RIMS")) : 1;
      static const int merge_pct = getenv("SB_MERGE_PCT") ? atoi(getenv("SB_MERGE_PCT")) : 130;
      uint32_t biggest = 0;
      for (uint32_t c : count) biggest = std::max(biggest, c);
      const uint32_t limit = (uint32_t)std::min<uint64_t>(cap, (uint64_t)std::min(cap, biggest) * (uint32_t)merge_pct / 100);
      std::vector<uint32_t> root(count.size());
      std::iota(root.begin(), root.end(), 0u);
      if (ok && s > 0 && merge_rims) {
        auto find = [&](uint32_t c) {
          while (root[c] != c) c = root[c] = root[root[c]];
          return c;
        };
        std::vector<uint32_t> size(count.begin(), count.end());
        std::vector<uint32_t> small;
        for (uint32_t c = 0; c < (uint32_t)count.size(); c++)
          if (count[c] && count[c] <= limit / 2) small.push_back(c);
        std::stable_sort(small.begin(), small.end(), [&](uint32_t a, uint32_t b) { return count[a] < count[b]; });
        for (int sweep = 0; sweep < 3; sweep++)
          for (uint32_t c : small) {
            const uint32_t r = find(c);
            if (size[r] > limit / 2) continue; // grown enough
            const int cx = (int)(c % (uint32_t)m[0]), cy = (int)(c / (uint32_t)m[0] % (uint32_t)m[1]), cz = (int)(c / ((uint32_t)m[0] * (uint32_t)m[1]));
            uint32_t best = 0xffffffffu, best_size = 0;
            static const int D6[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
            for (const auto &d : D6) {
              const int x = cx + d[0], y = cy + d[1], z = cz + d[2];
              if (x < 0 || y < 0 || z < 0 || x >= m[0] || y >= m[1] || z >= m[2]) continue;
              const uint32_t q0 = (uint32_t)x + (uint32_t)m[0] * ((uint32_t)y + (uint32_t)m[1] * (uint32_t)z);
              if (!count[q0]) continue;
              const uint32_t q = find(q0);
              if (q == r || size[q] + size[r] > limit) continue;
              if (best == 0xffffffffu || size[q] < best_size) { best = q; best_size = size[q]; }
            }
            if (best == 0xffffffffu) continue;
            const uint32_t keep = std::min(r, best), gone = std::max(r, best); // deterministic: the lower cell index names the tile
            root[gone] = keep;
            size[keep] = size[r] + size[best];
          }
        for (uint32_t c = 0; c < (uint32_t)count.size(); c++) root[c] = find(c);
      }
      // compact the non-empty tiles
      std::vector<int32_t> remap(count.size(), -1);
      uint32_t nt = 0;
      for (size_t c = 0; c < count.size(); c++)
        if (count[c] && root[c] == c) remap[c] = (int32_t)nt++;
      for (uint32_t v = 0; v < V; v++) T.part[v] = remap[root[T.part[v]]];
      T.n_tiles = nt;
    }
    if (ok) return true;
  }
  return false;
}

// ---- tile passes -------------------------------------------------------------

struct Mask128 {
  uint64_t lo = 0, hi = 0;
};

inline int first_free(const Mask128 &m) {
  if (~m.lo) return __builtin_ctzll(~m.lo);
  if (~m.hi) return 64 + __builtin_ctzll(~m.hi);
  return -1;
}
inline void set_bit(Mask128 &m, int c) {
  if (c < 64) m.lo |= 1ull << c;
  else m.hi |= 1ull << (c - 64);
}

struct TileOut {
  std::vector<uint32_t> verts; // device ids (non-contiguous passes)
  uint32_t n_ecol = 0, n_tcol = 0, n_verts = 0;
  std::vector<uint32_t> stream;  // words: edge rounds, then tet rounds
  std::vector<float> aux;        // per tet round
  std::vector<uint8_t> flags;        // per tet colour: bit 0 some (0,1) edge attached, 1 some (2,3) edge, 2 some mate, 3 / 4 some edge of a mate
  std::vector<int32_t> ents;     // processing order
  std::vector<uint32_t> col_cnt; // per colour, edge colours first
  uint64_t n_edges = 0, n_tets = 0;
  std::string err;
};

struct DevTopo {
  std::vector<int32_t> edges; // 2E device ids, roles (a,b) as in the canonical list
  std::vector<int32_t> tets;  // 4T device ids, roles as given
  std::vector<int32_t> apex;  // T: for the first tet of a bi-tet, the device id of its mate's apex (the fifth vertex), else -1
};

inline int ent_verts(const DevTopo &D, int32_t ent, int32_t *vs, bool with_mate = true) {
  if (ent >= 0) {
    vs[0] = D.edges[2 * (size_t)ent];
    vs[1] = D.edges[2 * (size_t)ent + 1];
    return 2;
  }
  const int32_t *q = &D.tets[4 * (size_t)(ent & 0x7fffffff)];
  vs[0] = q[0]; vs[1] = q[1]; vs[2] = q[2]; vs[3] = q[3];
  if (with_mate && !D.apex.empty() && D.apex[ent & 0x7fffffff] >= 0) { // a bi-tet: the mate rides with its leader
    vs[4] = D.apex[ent & 0x7fffffff];
    return 5;
  }
  return 4;
}

// Builds one pass.  part[v] = tile of device vertex v or -1.  Constraints of
// `in` whose vertices all lie in one tile are consumed; the rest go to `out`.
std::string build_pass(const Plan &P, const DevTopo &D, const std::vector<int32_t> &part, uint32_t n_tiles,
                       const std::vector<uint32_t> *contig_off, const std::vector<int32_t> &in,
                       std::vector<int32_t> &out, TilePass &TP, int threads, uint32_t bt_opt, uint32_t width, int n_sm,
                       const std::vector<int32_t> *box0 = nullptr, bool wide = false) {
  // 1. classify and bucket by tile (stable)
  std::vector<uint64_t> toff((size_t)n_tiles + 1, 0);
  std::vector<int32_t> owner(in.size());
  parallel_for(in.size(), threads, 1 << 16, [&](size_t i, int) {
    int32_t vs[5];
    int n = ent_verts(D, in[i], vs);
    int32_t p = part[vs[0]];
    for (int k = 1; k < n; k++)
      if (part[vs[k]] != p) p = -1;
    owner[i] = p;
  });
  out.clear();
  for (size_t i = 0; i < in.size(); i++) {
    if (owner[i] >= 0) toff[(size_t)owner[i] + 1]++;
    else out.push_back(in[i]);
  }
  for (uint32_t t = 0; t < n_tiles; t++) toff[t + 1] += toff[t];
  std::vector<int32_t> tent(toff[n_tiles]);
  {
    std::vector<uint64_t> cur(toff.begin(), toff.end() - 1);
    for (size_t i = 0; i < in.size(); i++)
      if (owner[i] >= 0) tent[cur[owner[i]]++] = in[i];
  }
  owner.clear();
  owner.shrink_to_fit();

  // CTA width of this pass: as given, else by how many constraints a tile holds (a round should be
  // full, and a tile should not need many more rounds than its valence asks for)
  uint32_t bt = bt_opt;
  if (!bt) { // by the fullest tile: partial tiles at the rim of a tiling must not talk the pass into narrow CTAs
    uint64_t per_tile = 0;
    for (uint32_t t = 0; t < n_tiles; t++) {
      uint64_t w = 0; // in single constraints: a bi-tet counts as two tets
      for (uint64_t i = toff[t]; i < toff[t + 1]; i++) w += (tent[i] < 0 && !D.apex.empty() && D.apex[tent[i] & 0x7fffffff] >= 0) ? 2 : 1;
      per_tile = std::max<uint64_t>(per_tile, w);
    }
    // `wide`: every pass fits the SMs in one wave even at six 160-thread CTAs each -- five warps per tile then beat
    // four (more warps in flight per SM, a round or two fewer per tile; measured 6.16 -> 5.87 ms per frame at 1 M
    // vertices); with several waves per pass four fuller warps are better (4 M vertices: 20.1 against 20.7 ms)
    bt = per_tile < 1600 ? 64u : per_tile < 6400 ? (wide ? 160u : 128u) : 256u;
  }
  // a round: 2 * width * bt free edges, or bt compounds (width 1: a tet with its attached edges in one 16-byte
  // word; width 2: a bi-tet -- or a single tet -- with its attached edges in two words)
  const uint32_t cap_e = 2 * width * bt, cap_t = bt;
  const bool bitet = width == 2;

  // A tile of a shifted tiling stages ALL the vertices of its boxes, not only those its constraints of this pass
  // touch: a vertex whose constraints all went to other tilings would otherwise split a run of device ids in two
  // (measured: 66 runs per tile instead of the ~30 the geometry gives).  No other tile of the pass touches a vertex
  // of this tile's boxes, so loading it and storing it back unchanged is safe.
  static const int whole_boxes = getenv("SB_WHOLE_BOXES") ? atoi(getenv("SB_WHOLE_BOXES")) : 1;
  std::vector<uint32_t> mem_off, mem_list;
  if (box0 && whole_boxes && !contig_off) {
    mem_off.assign((size_t)n_tiles + 1, 0);
    for (uint32_t d = 0; d < P.V; d++)
      if (part[d] >= 0) mem_off[(size_t)part[d] + 1]++;
    for (uint32_t t = 0; t < n_tiles; t++) mem_off[t + 1] += mem_off[t];
    mem_list.resize(mem_off[n_tiles]);
    std::vector<uint32_t> cur(mem_off.begin(), mem_off.end() - 1);
    for (uint32_t d = 0; d < P.V; d++)
      if (part[d] >= 0) mem_list[cur[part[d]]++] = d; // ascending device id inside a tile
  }

  // 2. per tile: local numbering, capacity-limited greedy colouring, rounds
  std::vector<TileOut> outs(n_tiles);
  int nt = std::max(1, threads);
  std::vector<std::vector<uint32_t>> loc_scratch(nt);
  parallel_for(n_tiles, threads, 1, [&](size_t t, int w) {
    TileOut &O = outs[t];
    const int32_t *ents = tent.data() + toff[t];
    size_t ne = toff[t + 1] - toff[t];
    uint32_t base = 0, nv = 0;
    std::vector<uint32_t> &loc = loc_scratch[w];
    if (contig_off) {
      base = (*contig_off)[t];
      nv = (*contig_off)[t + 1] - base;
    } else {
      if (loc.size() != P.V) loc.assign(P.V, 0xffffffffu);
      if (!mem_off.empty()) {
        if (ne) O.verts.assign(mem_list.begin() + mem_off[t], mem_list.begin() + mem_off[t + 1]);
      } else {
        for (size_t i = 0; i < ne; i++) {
          int32_t vs[5];
          int n = ent_verts(D, ents[i], vs);
          for (int k = 0; k < n; k++)
            if (loc[vs[k]] == 0xffffffffu) {
              loc[vs[k]] = 0;
              O.verts.push_back((uint32_t)vs[k]);
            }
        }
        std::sort(O.verts.begin(), O.verts.end());
      }
      nv = (uint32_t)O.verts.size();
      for (uint32_t k = 0; k < nv; k++) loc[O.verts[k]] = k;
    }
    O.n_verts = nv;
    if (nv > 65536) {
      O.err = "tile exceeds 65536 vertices";
      return;
    }
    auto local = [&](int32_t v) -> uint32_t { return contig_off ? (uint32_t)v - base : loc[v]; };
    // colour = round.  Per kind: a bit set per vertex of the colours it already carries, plus the set
    // of full colours; a constraint takes the first colour that is free at all its vertices and not full.
    size_t n_kind[2] = {0, 0};
    for (size_t i = 0; i < ne; i++) n_kind[ents[i] < 0]++;
    const size_t words[2] = {(n_kind[0] / cap_e + 256) / 64 + 1, (n_kind[1] / cap_t + 256) / 64 + 1};
    std::vector<uint64_t> mask[2], full[2], soft[2];
    for (int k = 0; k < 2; k++) {
      if (n_kind[k]) mask[k].assign((size_t)nv * words[k], 0);
      full[k].assign(words[k], 0);
      soft[k].assign(words[k], 0);
    }
    std::vector<uint32_t> col(ne);
    std::vector<uint32_t> ecount, tcount;
    // Constraints are coloured in a scattered order (a fixed hash of the constraint id): colouring them in
    // id order, which is spatial order, fills the early colours from one corner of the tile and leaves
    // the last corner a long tail of nearly empty colours.
    // First-fit is sensitive to that order: a tile whose colouring ends well above its lower bound
    // max(valence, ceil(n / capacity)) is coloured again in other scattered orders and the best is kept
    // (the slowest tile of a pass ends the pass, so the worst tiles matter most).
    std::vector<uint32_t> cord(ne);
    std::vector<uint32_t> try_col(ne), try_e, try_t;
    size_t lower = 0;
    {
      std::vector<uint32_t> val((size_t)nv * 2, 0);
      for (size_t i = 0; i < ne; i++) {
        int32_t vs[5];
        const int n = ent_verts(D, ents[i], vs);
        for (int k = 0; k < n; k++) val[(size_t)local(vs[k]) * 2 + (ents[i] < 0)]++;
      }
      uint32_t mv[2] = {0, 0};
      for (uint32_t v = 0; v < nv; v++) { mv[0] = std::max(mv[0], val[2 * v]); mv[1] = std::max(mv[1], val[2 * v + 1]); }
      lower = std::max<size_t>(mv[0], (n_kind[0] + cap_e - 1) / cap_e) + std::max<size_t>(mv[1], (n_kind[1] + cap_t - 1) / cap_t);
    }
    auto colour_once = [&](uint32_t seed, std::vector<uint32_t> &col_o, std::vector<uint32_t> &ecount_o, std::vector<uint32_t> &tcount_o) -> bool {
      ecount_o.clear();
      tcount_o.clear();
      for (int k = 0; k < 2; k++) {
        std::fill(mask[k].begin(), mask[k].end(), 0);
        std::fill(full[k].begin(), full[k].end(), 0);
        std::fill(soft[k].begin(), soft[k].end(), 0);
      }
      std::iota(cord.begin(), cord.end(), 0u);
      {
        auto hash = [](uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; };
        std::vector<uint64_t> key(ne);
        for (size_t i = 0; i < ne; i++) key[i] = ((uint64_t)hash((uint32_t)ents[i] ^ (seed * 0x9e3779b9u)) << 32) | (uint32_t)i;
        std::sort(key.begin(), key.end());
        for (size_t i = 0; i < ne; i++) cord[i] = (uint32_t)key[i];
      }
      for (size_t oi = 0; oi < ne; oi++) {
        const size_t i = cord[oi];
        int32_t vs[5];
        const int n = ent_verts(D, ents[i], vs);
        const int kind = ents[i] < 0;
        const size_t W = words[kind];
        uint64_t *M = mask[kind].data();
        // first colour free at every vertex that is below the soft capacity; failing that, one below the
        // hard capacity (the slack the soft limit left is what absorbs the stragglers); failing that, a new one
        static const uint32_t slack_div = getenv("SB_SLACK_DIV") ? (uint32_t)atoi(getenv("SB_SLACK_DIV")) : 16u;
        const uint32_t hard = kind ? cap_t : cap_e, slack = std::max(1u, hard / slack_div);
        std::vector<uint32_t> &cnt = kind ? tcount_o : ecount_o;
        auto first_free_colour = [&](const std::vector<uint64_t> &closed) {
          for (size_t wd = 0; wd < W; wd++) {
            uint64_t u = closed[wd];
            for (int k = 0; k < n; k++) u |= M[(size_t)local(vs[k]) * W + wd];
            if (~u) return (int)(wd * 64) + __builtin_ctzll(~u);
          }
          return -1;
        };
        const size_t k_target = (n_kind[kind] + (hard - slack) - 1) / (hard - slack);
        int c = first_free_colour(soft[kind]);
        if ((c < 0 || (size_t)c >= cnt.size()) && cnt.size() >= k_target) {
          const int c2 = first_free_colour(full[kind]);
          if (c2 >= 0 && (size_t)c2 < cnt.size()) c = c2;
        }
        if (c < 0) return false;
        for (int k = 0; k < n; k++) M[(size_t)local(vs[k]) * W + (size_t)c / 64] |= 1ull << (c & 63);
        col_o[i] = (uint32_t)c;
        if ((size_t)c >= cnt.size()) cnt.resize(c + 1, 0);
        ++cnt[c];
        if (cnt[c] >= hard - slack) soft[kind][(size_t)c / 64] |= 1ull << (c & 63);
        if (cnt[c] == hard) full[kind][(size_t)c / 64] |= 1ull << (c & 63);
      }
      return true;
    };
    {
      static const int n_try = getenv("SB_COLOUR_TRIES") ? atoi(getenv("SB_COLOUR_TRIES")) : 1;
      static const int try_margin = getenv("SB_COLOUR_MARGIN") ? atoi(getenv("SB_COLOUR_MARGIN")) : 2;
      bool have = false;
      for (int a = 0; a < std::max(1, n_try); a++) {
        if (!colour_once((uint32_t)a, try_col, try_e, try_t)) {
          if (a == 0) { O.err = "vertex valence needs more than 256 colours beyond the capacity bound"; return; }
          continue;
        }
        if (!have || try_e.size() + try_t.size() < ecount.size() + tcount.size()) {
          col.swap(try_col); ecount.swap(try_e); tcount.swap(try_t);
          try_col.resize(ne);
          have = true;
        }
        if (ecount.size() + tcount.size() <= lower + (size_t)try_margin) break; // close enough to the bound
      }
    }
    // Stragglers.  First-fit leaves a tail of nearly empty colours (e.g. ... 124 103 86 48 21 2): every one of them
    // costs the tile a full round trip (gather, project, scatter, barrier).  Empty the smallest colours into the room
    // the others have left: a member moves to a colour where it conflicts with nobody, or with exactly one member
    // that can itself move elsewhere (one Kempe step).  Colours that end up empty are dropped.
    static const int recolour = getenv("SB_RECOLOUR") ? atoi(getenv("SB_RECOLOUR")) : 1; // (0: A/B switch for benches)
    for (int kind = 0; kind < 2 && recolour; kind++) {
      std::vector<uint32_t> &cnt = kind ? tcount : ecount;
      const uint32_t K = (uint32_t)cnt.size(), hard = kind ? cap_t : cap_e;
      if (K < 2) continue;
      const uint32_t NONE = 0xffffffffu;
      std::vector<uint32_t> owner((size_t)nv * K, NONE); // who holds colour c at local vertex v
      std::vector<std::vector<uint32_t>> members(K);
      auto verts_of = [&](uint32_t i, uint32_t *lv) {
        int32_t vs[5];
        const int n = ent_verts(D, ents[i], vs);
        for (int k = 0; k < n; k++) lv[k] = local(vs[k]);
        return n;
      };
      for (size_t i = 0; i < ne; i++) {
        if ((ents[i] < 0) != (kind != 0)) continue;
        uint32_t lv[5];
        const int n = verts_of((uint32_t)i, lv);
        for (int k = 0; k < n; k++) owner[(size_t)lv[k] * K + col[i]] = (uint32_t)i;
        members[col[i]].push_back((uint32_t)i);
      }
      auto move = [&](uint32_t i, uint32_t to) {
        uint32_t lv[5];
        const int n = verts_of(i, lv);
        for (int k = 0; k < n; k++) { owner[(size_t)lv[k] * K + col[i]] = NONE; owner[(size_t)lv[k] * K + to] = i; }
        --cnt[col[i]];
        ++cnt[to];
        col[i] = to;
      };
      // conflicts of member i in colour c: distinct holders of c at i's vertices, ignoring `skip`; returns count (<= 4)
      auto conflicts = [&](uint32_t i, uint32_t c, uint32_t skip, uint32_t *out) {
        uint32_t lv[5];
        const int n = verts_of(i, lv);
        int m = 0;
        for (int k = 0; k < n; k++) {
          const uint32_t o = owner[(size_t)lv[k] * K + c];
          if (o == NONE || o == skip) continue;
          bool seen = false;
          for (int j = 0; j < m; j++) seen |= out[j] == o;
          if (!seen) out[m++] = o;
        }
        return m;
      };
      static const int rc_sweeps = getenv("SB_RC_SWEEPS") ? atoi(getenv("SB_RC_SWEEPS")) : 2;
      static const int rc_frac = getenv("SB_RC_FRAC") ? atoi(getenv("SB_RC_FRAC")) : 75; // % of the capacity
      std::vector<uint8_t> closed(K, 0); // colours being emptied (or already empty) take no new member
      for (int sweep = 0; sweep < rc_sweeps; sweep++) {
      std::vector<uint32_t> by_size(K);
      std::iota(by_size.begin(), by_size.end(), 0u);
      std::stable_sort(by_size.begin(), by_size.end(), [&](uint32_t a, uint32_t b) { return cnt[a] < cnt[b]; });
      for (uint32_t c = 0; c < K; c++) closed[c] = cnt[c] == 0;
      for (uint32_t src : by_size) {
        if (cnt[src] == 0 || cnt[src] > hard * (uint32_t)rc_frac / 100) continue;
        // is there room elsewhere at all?
        size_t room = 0;
        for (uint32_t c = 0; c < K; c++)
          if (c != src && !closed[c]) room += hard - cnt[c];
        if (room < cnt[src]) continue;
        closed[src] = 1;
        std::vector<uint32_t> todo;
        for (uint32_t i : members[src])
          if (col[i] == src) todo.push_back(i);
        for (uint32_t i : todo) {
          uint32_t cf[5];
          bool done = false;
          for (uint32_t c = 0; c < K && !done; c++) { // a free seat
            if (closed[c] || cnt[c] >= hard) continue;
            if (conflicts(i, c, NONE, cf) == 0) { move(i, c); members[c].push_back(i); done = true; }
          }
          for (uint32_t c = 0; c < K && !done; c++) { // one Kempe step: the only conflicting member moves on
            if (closed[c] || cnt[c] >= hard) continue;
            if (conflicts(i, c, NONE, cf) != 1) continue;
            const uint32_t x = cf[0];
            for (uint32_t c2 = 0; c2 < K && !done; c2++) {
              if (c2 == c || closed[c2] || cnt[c2] >= hard) continue;
              uint32_t cf2[5];
              if (conflicts(x, c2, NONE, cf2) != 0) continue;
              move(x, c2); members[c2].push_back(x);
              move(i, c); members[c].push_back(i);
              done = true;
            }
          }
          for (uint32_t c = 0; c < K && !done; c++) { // ... or the only two (the colour may be full: two leave, one joins)
            if (closed[c]) continue;
            if (conflicts(i, c, NONE, cf) != 2) continue;
            const uint32_t x = cf[0], y = cf[1];
            uint32_t cx = NONE;
            for (uint32_t c2 = 0; c2 < K && cx == NONE; c2++) {
              uint32_t cf2[5];
              if (c2 != c && !closed[c2] && cnt[c2] < hard && conflicts(x, c2, NONE, cf2) == 0) cx = c2;
            }
            if (cx == NONE) continue;
            const uint32_t from = col[x];
            move(x, cx); // y must not collide with x in its new colour: evaluated after x has moved
            uint32_t cy = NONE;
            for (uint32_t c2 = 0; c2 < K && cy == NONE; c2++) {
              uint32_t cf2[5];
              if (c2 != c && !closed[c2] && cnt[c2] < hard && conflicts(y, c2, NONE, cf2) == 0) cy = c2;
            }
            if (cy == NONE) { move(x, from); continue; }
            members[cx].push_back(x);
            move(y, cy); members[cy].push_back(y);
            move(i, c); members[c].push_back(i);
            done = true;
          }
        }
        if (cnt[src] != 0) closed[src] = 0; // not emptied: it stays a colour like any other
      }
      }
      // drop the empty colours (order of the others kept)
      std::vector<uint32_t> renum(K, NONE);
      uint32_t nk = 0;
      for (uint32_t c = 0; c < K; c++)
        if (cnt[c]) renum[c] = nk++;
      if (nk != K) {
        for (size_t i = 0; i < ne; i++)
          if ((ents[i] < 0) == (kind != 0)) col[i] = renum[col[i]];
        std::vector<uint32_t> nc(nk);
        for (uint32_t c = 0; c < K; c++)
          if (cnt[c]) nc[renum[c]] = cnt[c];
        cnt.swap(nc);
      }
    }
    if (getenv("SB_PLAN_STATS")) { // debug: rounds against the lower bounds (capacity, valence)
      std::vector<uint32_t> val((size_t)nv * 2, 0);
      for (size_t i = 0; i < ne; i++) {
        int32_t vs[5];
        const int n = ent_verts(D, ents[i], vs);
        for (int k = 0; k < n; k++) val[(size_t)local(vs[k]) * 2 + (ents[i] < 0)]++;
      }
      uint32_t mv[2] = {0, 0};
      for (uint32_t v = 0; v < nv; v++) { mv[0] = std::max(mv[0], val[2 * v]); mv[1] = std::max(mv[1], val[2 * v + 1]); }
      if (t % 97 == 0) {
        fprintf(stderr, "TCOUNT %zu:", t);
        for (uint32_t c : tcount) fprintf(stderr, " %u", c);
        fprintf(stderr, " | E:");
        for (uint32_t c : ecount) fprintf(stderr, " %u", c);
        fprintf(stderr, "\n");
      }
      fprintf(stderr, "TILE %zu nv %u edges %zu tets %zu ecol %zu (cap %zu val %u) tcol %zu (cap %zu val %u)\n", t, nv, n_kind[0],
              n_kind[1], ecount.size(), (n_kind[0] + cap_e - 1) / cap_e, mv[0], tcount.size(), (n_kind[1] + cap_t - 1) / cap_t, mv[1]);
    }
    mask[0] = std::vector<uint64_t>();
    mask[1] = std::vector<uint64_t>();
    O.n_ecol = (uint32_t)ecount.size();
    O.n_tcol = (uint32_t)tcount.size();
    // bucket by colour (stable): edge colours first, then tet colours
    const uint32_t ncol = O.n_ecol + O.n_tcol;
    std::vector<uint32_t> coff(ncol + 1, 0);
    for (uint32_t c = 0; c < O.n_ecol; c++) coff[c + 1] = coff[c] + ecount[c];
    for (uint32_t c = 0; c < O.n_tcol; c++) coff[O.n_ecol + c + 1] = coff[O.n_ecol + c] + tcount[c];
    O.col_cnt.resize(ncol);
    for (uint32_t c = 0; c < ncol; c++) O.col_cnt[c] = coff[c + 1] - coff[c];
    O.ents.resize(ne);
    {
      std::vector<uint32_t> cur(coff.begin(), coff.end() - 1);
      for (size_t i = 0; i < ne; i++) O.ents[cur[(ents[i] < 0 ? O.n_ecol : 0) + col[i]]++] = ents[i];
    }
    // Within a colour any order is equivalent; pick one where each run of 8 records (the
    // quarter-warp an LDS.128 serves per wavefront) touches 8 different 16-byte bank groups
    // (local id mod 8) in every vertex slot.  Compounds with a second attached edge come first,
    // so that the warps behind them skip that code.
    {
      std::vector<int32_t> tmp;
      std::vector<uint8_t> cls;
      std::vector<uint32_t> bucket[8];
      for (uint32_t c = 0; c < ncol; c++) {
        const uint32_t lo = coff[c], n = coff[c + 1] - coff[c];
        const bool tet = c >= O.n_ecol;
        const int nvs = tet ? (bitet ? 5 : 4) : 2;
        // first class: compounds that run the longer code (a mate; else a second attached edge), so that the warps
        // behind them skip it
        auto second_edge = [&](int32_t e) {
          return tet && (bitet ? P.tet_mate[e & 0x7fffffff] >= 0 : P.tet_e23[e & 0x7fffffff] >= 0);
        };
        if (tet) std::stable_partition(O.ents.begin() + lo, O.ents.begin() + lo + n, second_edge);
        if (n < 16) continue;
        std::vector<uint8_t> res((size_t)n * 5, 8); // bank group (local id mod 8) per vertex slot; 8 = no such vertex
        for (uint32_t k = 0; k < n; k++) {
          int32_t vs[5];
          const int nv_k = ent_verts(D, O.ents[lo + k], vs);
          for (int j = 0; j < nv_k; j++) res[(size_t)k * 5 + j] = (uint8_t)(local(vs[j]) & 7u);
        }
        uint32_t n_first = 0;
        while (n_first < n && second_edge(O.ents[lo + n_first])) n_first++;
        tmp.clear();
        tmp.reserve(n);
        cls.clear();
        uint8_t used[5] = {0, 0, 0, 0, 0};
        // greedy octets, class by class; the octet that straddles the class boundary keeps its residues
        for (int pass = 0; pass < 2; pass++) {
          const uint32_t k0 = pass ? n_first : 0, k1 = pass ? n : n_first;
          for (auto &bk : bucket) bk.clear();
          for (uint32_t k = k0; k < k1; k++) bucket[res[(size_t)k * 5]].push_back(k);
          for (auto &bk : bucket) std::reverse(bk.begin(), bk.end()); // pop_back takes ascending ids first
          uint32_t left = k1 - k0;
          while (left) {
            const uint32_t pos = (uint32_t)tmp.size() & 7u;
            if (pos == 0) used[0] = used[1] = used[2] = used[3] = used[4] = 0;
            // prefer the bucket whose slot-0 residue is still free in this octet
            int b = -1;
            for (int r = 0; r < 8; r++) {
              const int cand = (int)((pos + r) & 7);
              if (!bucket[cand].empty() && !(used[0] >> cand & 1)) { b = cand; break; }
            }
            if (b < 0) {
              size_t best = 0;
              for (int r = 0; r < 8; r++)
                if (bucket[r].size() > best) { best = bucket[r].size(); b = r; }
            }
            std::vector<uint32_t> &B = bucket[b];
            // among the next few candidates take the one with the fewest conflicts in the other slots
            size_t pick = B.size() - 1;
            int pick_conf = 99;
            for (size_t q = 0; q < 12 && q < B.size(); q++) {
              const uint32_t k = B[B.size() - 1 - q];
              int conf = 0;
              for (int j = 1; j < nvs; j++) conf += used[j] >> res[(size_t)k * 5 + j] & 1;
              if (conf < pick_conf) { pick_conf = conf; pick = B.size() - 1 - q; if (!conf) break; }
            }
            const uint32_t k = B[pick];
            B.erase(B.begin() + (ptrdiff_t)pick);
            for (int j = 0; j < nvs; j++) used[j] |= (uint8_t)(1u << res[(size_t)k * 5 + j]);
            tmp.push_back(O.ents[lo + k]);
            cls.push_back((uint8_t)pass);
            left--;
          }
        }
        // local search: swap records of the same class between octets while that lowers the wavefront
        // count (per octet and vertex slot: the largest number of records in one bank group)
        {
          const uint32_t n_oct = (n + 7) / 8;
          std::vector<uint8_t> rr((size_t)n * 5, 8);
          for (uint32_t k = 0; k < n; k++) {
            int32_t vs[5];
            const int nv_k = ent_verts(D, tmp[k], vs);
            for (int j = 0; j < nv_k; j++) rr[(size_t)k * 5 + j] = (uint8_t)(local(vs[j]) & 7u);
          }
          auto oct_cost = [&](uint32_t o, uint32_t swap_pos, uint32_t swap_with) {
            // cost of octet o with the record at swap_pos replaced by the one at swap_with (swap_pos == ~0u: as is)
            uint32_t cost = 0;
            const uint32_t k0 = o * 8, k1 = std::min(n, k0 + 8);
            static const int ls_pairs = getenv("SB_LS_PAIRS") ? atoi(getenv("SB_LS_PAIRS")) : 0;
            for (int j = 0; j < nvs; j++) {
              uint8_t cntb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
              uint8_t m = 0;
              uint32_t pairs = 0;
              for (uint32_t k = k0; k < k1; k++) {
                const uint32_t src = k == swap_pos ? swap_with : k;
                if (rr[(size_t)src * 5 + j] > 7) continue; // a single tet among bi-tets: no fifth vertex
                const uint8_t v = ++cntb[rr[(size_t)src * 5 + j]];
                pairs += v - 1u; // records already on this residue: the sum is the number of conflicting pairs
                m = std::max(m, v);
              }
              cost += ls_pairs ? 32u * m + pairs : m;
            }
            return cost;
          };
          static const int ls_pairs_on = getenv("SB_LS_PAIRS") ? atoi(getenv("SB_LS_PAIRS")) : 0;
          std::vector<uint32_t> oc(n_oct);
          for (uint32_t o = 0; o < n_oct; o++) oc[o] = oct_cost(o, ~0u, 0);
          uint32_t rng = 0x9e3779b9u ^ (uint32_t)(t * 2654435761u) ^ (c * 40503u);
          static const int ls_sweeps = getenv("SB_LS_SWEEPS") ? atoi(getenv("SB_LS_SWEEPS")) : 6;
          static const int ls_tries = getenv("SB_LS_TRIES") ? atoi(getenv("SB_LS_TRIES")) : 24;
          for (int sweep = 0; sweep < ls_sweeps; sweep++) {
            bool improved = false;
            for (uint32_t a = 0; a < n; a++) {
              const uint32_t oa = a / 8;
              if (oc[oa] <= (uint32_t)nvs * (ls_pairs_on ? 32u : 1u)) continue; // already conflict-free
              for (int tries = 0; tries < ls_tries; tries++) {
                rng = rng * 1664525u + 1013904223u;
                const uint32_t b = (uint32_t)(((uint64_t)(rng >> 8) * n) >> 24);
                const uint32_t ob = b / 8;
                if (ob == oa || b >= n || cls[a] != cls[b]) continue;
                const uint32_t na = oct_cost(oa, a, b), nb = oct_cost(ob, b, a);
                static const int ls_plateau = getenv("SB_LS_PLATEAU") ? atoi(getenv("SB_LS_PLATEAU")) : 0;
                if (na + nb < oc[oa] + oc[ob] || (ls_plateau && na + nb == oc[oa] + oc[ob] && na < oc[oa])) {
                  std::swap(tmp[a], tmp[b]);
                  for (int j = 0; j < 5; j++) std::swap(rr[(size_t)a * 5 + j], rr[(size_t)b * 5 + j]);
                  oc[oa] = na;
                  oc[ob] = nb;
                  improved = true;
                  break;
                }
              }
            }
            if (!improved) break;
          }
        }
        std::copy(tmp.begin(), tmp.end(), O.ents.begin() + lo);
      }
    }
    // Rounds.  Record k of a colour goes to the slot round_slot() names: packed towards the first warps,
    // the octets above on the eight lanes of one LDS wavefront.
    const uint32_t round_words = 4 * width * bt;
    O.stream.assign((size_t)ncol * round_words, 0u);
    O.aux.assign((size_t)O.n_tcol * width * bt, std::numeric_limits<float>::quiet_NaN());
    O.flags.assign(O.n_tcol, 0);
    for (uint32_t c = 0; c < ncol; c++) {
      const bool tet = c >= O.n_ecol;
      uint32_t *rw = O.stream.data() + (size_t)c * round_words;
      for (uint32_t k = 0; k < O.col_cnt[c]; k++) {
        int32_t vs[5];
        const int32_t e = O.ents[coff[c] + k];
        const int nv_k = ent_verts(D, e, vs);
        uint32_t thr, sub;
        round_slot(k, tet ? 1 : 2 * width, thr, sub);
        if (!tet) {
          uint32_t *r = rw + (size_t)thr * 4 * width + 2 * sub;
          r[0] = local(vs[0]) | (local(vs[1]) << 16);
          r[1] = f2u(P.rest_len[e]);
        } else {
          uint32_t *r = rw + (size_t)thr * 4 * width;
          const int32_t id = e & 0x7fffffff, e01 = P.tet_e01[id], e23 = P.tet_e23[id];
          const uint32_t qnan = 0x7fc00000u;
          r[0] = local(vs[0]) | (local(vs[1]) << 16);
          r[1] = local(vs[2]) | (local(vs[3]) << 16);
          // TIMING EXPERIMENT ONLY (results are garbage): every lane's role k reads slot thr + k * bt, so that the eight
          // lanes of a quarter-warp never share a 16-byte bank group -- what a conflict-free layout would cost
          static const int debug_noconflict = getenv("SB_DEBUG_NOCONFLICT") ? atoi(getenv("SB_DEBUG_NOCONFLICT")) : 0;
          if (debug_noconflict && O.n_verts >= 4 * bt) {
            r[0] = thr | ((thr + bt) << 16);
            r[1] = (thr + 2 * bt) | ((thr + 3 * bt) << 16);
          }
          r[2] = f2u(P.rest_vol6[id]);
          r[3] = e01 >= 0 ? f2u(P.rest_len[e01]) : qnan;
          O.aux[(size_t)(c - O.n_ecol) * width * bt + (size_t)thr * width] =
              e23 >= 0 ? P.rest_len[e23] : std::numeric_limits<float>::quiet_NaN();
          O.n_edges += (e01 >= 0) + (e23 >= 0);
          uint8_t &fl = O.flags[c - O.n_ecol];
          if (e01 >= 0) fl |= 1;
          if (e23 >= 0) fl |= 2;
          if (bitet) {
            // second word: the mate B = (b, s1, s0, s2) on the registers (4, 2, 1, 3): its apex, its rest volume, the
            // rest lengths of its attached edges (b, s1) and (s0, s2); a NaN volume = no mate
            const int32_t mate = nv_k == 5 ? P.tet_mate[id] : -1;
            r[4] = mate >= 0 ? local(vs[4]) : 0u;
            r[5] = mate >= 0 ? f2u(P.rest_vol6[mate]) : qnan;
            r[6] = mate >= 0 && P.tet_e01[mate] >= 0 ? f2u(P.rest_len[P.tet_e01[mate]]) : qnan;
            r[7] = mate >= 0 && P.tet_e23[mate] >= 0 ? f2u(P.rest_len[P.tet_e23[mate]]) : qnan;
            if (mate >= 0) {
              fl |= 4;
              if (P.tet_e01[mate] >= 0) fl |= 8;
              if (P.tet_e23[mate] >= 0) fl |= 16;
              O.n_edges += (P.tet_e01[mate] >= 0) + (P.tet_e23[mate] >= 0);
              O.n_tets++;
            }
          }
        }
      }
      (tet ? O.n_tets : O.n_edges) += O.col_cnt[c];
    }
    if (!contig_off)
      for (uint32_t v : O.verts) loc[v] = 0xffffffffu;
  });

  // 3. concatenate
  TP = TilePass();
  TP.contiguous = contig_off != nullptr;
  TP.vert_off.push_back(0);
  TP.bt = bt;
  TP.width = width;
  TP.run_off.push_back(0);
  TP.ent_off.push_back(0);
  TP.col_off.push_back(0);
  size_t nw = 0, nen = 0, nv = 0;
  for (auto &O : outs) {
    if (!O.err.empty()) return O.err;
    nw += O.stream.size();
    nen += O.ents.size();
    nv += O.verts.size();
  }
  if (nw / 4 >= 0xffffffffull) return "constraint stream of one pass exceeds 64 GiB";
  TP.stream.reserve(nw);
  TP.ents.reserve(nen);
  TP.tile_verts.reserve(nv);
  for (uint32_t t = 0; t < n_tiles; t++) {
    TileOut &O = outs[t];
    TP.rounds.push_back({(uint32_t)(TP.stream.size() / 4), O.n_ecol, O.n_tcol, (uint32_t)TP.aux.size()});
    TP.aux.insert(TP.aux.end(), O.aux.begin(), O.aux.end());
    if (TP.col_flags.size() < O.n_tcol) TP.col_flags.resize(O.n_tcol, 0);
    for (uint32_t c = 0; c < O.n_tcol; c++) TP.col_flags[c] |= O.flags[c];
    TP.stream.insert(TP.stream.end(), O.stream.begin(), O.stream.end());
    TP.ents.insert(TP.ents.end(), O.ents.begin(), O.ents.end());
    TP.ent_off.push_back(TP.ents.size());
    TP.col_cnt.insert(TP.col_cnt.end(), O.col_cnt.begin(), O.col_cnt.end());
    TP.col_off.push_back((uint32_t)TP.col_cnt.size());
    TP.n_ecol.push_back(O.n_ecol);
    if (TP.contiguous) {
      TP.vert_off.push_back((*contig_off)[t + 1]);
    } else {
      TP.tile_verts.insert(TP.tile_verts.end(), O.verts.begin(), O.verts.end());
      TP.vert_off.push_back((uint32_t)TP.tile_verts.size());
      // maximal runs of consecutive device ids: {first device id, first local id}, closed by {0, n}.  A run never
      // crosses a box of the unshifted tiling (box0): those boxes are the units a mesh spread over several GPUs is
      // cut along, and a run is read and written in the memory of the rank that owns its first vertex.
      for (uint32_t k = 0; k < O.n_verts; k++)
        if (k == 0 || O.verts[k] != O.verts[k - 1] + 1 || (box0 && (*box0)[O.verts[k]] != (*box0)[O.verts[k - 1]]))
          TP.runs.push_back({O.verts[k], k});
      TP.runs.push_back({0u, O.n_verts});
      TP.run_off.push_back((uint32_t)TP.runs.size());
      if (getenv("SB_PLAN_STATS") && t == n_tiles / 2) {
        fprintf(stderr, "RUNS tile %u nv %u:", t, O.n_verts);
        const size_t r0 = TP.run_off[TP.run_off.size() - 2], r1 = TP.runs.size() - 1;
        for (size_t r = r0; r < r1; r++) fprintf(stderr, " %u+%u", TP.runs[r].x, TP.runs[r + 1].y - TP.runs[r].y);
        fprintf(stderr, "\n");
      }
    }
    TP.max_ecol = std::max(TP.max_ecol, O.n_ecol);
    TP.max_tcol = std::max(TP.max_tcol, O.n_tcol);
    TP.max_tile_verts = std::max(TP.max_tile_verts, O.n_verts);
    TP.n_edges += O.n_edges;
    TP.n_tets += O.n_tets;
    TP.rounds_total += O.n_ecol + O.n_tcol;
    O = TileOut();
  }
  if (TP.contiguous) TP.vert_off[0] = (*contig_off)[0];
  // Launch order.  All CTAs of a pass are resident at once and the hardware deals them to the SMs
  // breadth-first, so CTA j lands on SM j mod n_sm: sort the tiles by work (rounds, then constraints),
  // heaviest first, and reverse every other layer of n_sm so that the per-SM sums even out.
  {
    TP.launch_order.resize(n_tiles);
    std::iota(TP.launch_order.begin(), TP.launch_order.end(), 0u);
    auto work = [&](uint32_t t) {
      return ((uint64_t)(TP.rounds[t].y + TP.rounds[t].z) << 32) | (uint32_t)(TP.ent_off[t + 1] - TP.ent_off[t]);
    };
    std::stable_sort(TP.launch_order.begin(), TP.launch_order.end(), [&](uint32_t a, uint32_t b) { return work(a) > work(b); });
    const uint32_t layer = (uint32_t)std::max(1, n_sm);
    for (uint32_t lo = layer; lo < n_tiles; lo += 2 * layer)
      std::reverse(TP.launch_order.begin() + lo, TP.launch_order.begin() + std::min(n_tiles, lo + layer));
  }
  return "";
}

// Next-level parts for the vertices of the cut constraints `cut`: a vertex joins
// the group keyed by the unordered pair {its part, the other part it shares most
// cut constraints with}; groups are packed (in key order) into tiles of <= cap.
uint32_t next_parts(const Plan &P, const DevTopo &D, const std::vector<int32_t> &cut, std::vector<int32_t> &part,
                    uint32_t cap, const std::vector<float> &dev_pos) {
  std::vector<uint64_t> votes;
  votes.reserve(cut.size() * 4);
  for (int32_t ent : cut) {
    int32_t vs[5];
    int n = ent_verts(D, ent, vs);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++)
        if (part[vs[j]] != part[vs[i]]) votes.push_back(((uint64_t)(uint32_t)vs[i] << 32) | (uint32_t)part[vs[j]]);
  }
  std::sort(votes.begin(), votes.end());
  struct KV {
    uint64_t key;
    uint32_t v;
  };
  std::vector<KV> kv;
  for (size_t i = 0; i < votes.size();) {
    uint32_t v = (uint32_t)(votes[i] >> 32);
    uint32_t best_q = 0, best_n = 0;
    while (i < votes.size() && (uint32_t)(votes[i] >> 32) == v) {
      uint32_t q = (uint32_t)votes[i];
      size_t j = i;
      while (j < votes.size() && votes[j] == votes[i]) j++;
      if ((uint32_t)(j - i) > best_n) {
        best_n = (uint32_t)(j - i);
        best_q = q;
      }
      i = j;
    }
    uint32_t p = (uint32_t)part[v];
    kv.push_back({((uint64_t)std::min(p, best_q) << 32) | std::max(p, best_q), v});
  }
  votes.clear();
  votes.shrink_to_fit();
  std::sort(kv.begin(), kv.end(), [](const KV &a, const KV &b) { return a.key < b.key || (a.key == b.key && a.v < b.v); });
  std::fill(part.begin(), part.end(), -1);
  uint32_t n_tiles = 0, open = 0;
  for (size_t i = 0; i < kv.size();) {
    size_t j = i;
    while (j < kv.size() && kv[j].key == kv[i].key) j++;
    uint32_t n = (uint32_t)(j - i);
    if (n > cap) {
      if (open) { n_tiles++; open = 0; }
      std::vector<uint32_t> idx(n);
      for (uint32_t k = 0; k < n; k++) idx[k] = kv[i + k].v;
      std::vector<size_t> ends;
      rcb(idx, 0, n, (n + cap - 1) / cap, dev_pos.data(), ends);
      size_t prev = 0;
      for (size_t e : ends) {
        for (size_t k = prev; k < e; k++) part[idx[k]] = (int32_t)n_tiles;
        n_tiles++;
        prev = e;
      }
    } else {
      if (open + n > cap) { n_tiles++; open = 0; }
      for (size_t k = i; k < j; k++) part[kv[k].v] = (int32_t)n_tiles;
      open += n;
    }
    i = j;
  }
  if (open) n_tiles++;
  (void)P;
  return n_tiles;
}

std::string global_colouring(Plan &P, const DevTopo &D, const std::vector<int32_t> &rest_in) {
  if (rest_in.empty()) return "";
  // a tet that ends up here brings its attached edges along as plain edges, and its mate (bi-tets) as a plain tet
  std::vector<int32_t> rest, tets_in;
  rest.reserve(rest_in.size());
  for (int32_t ent : rest_in)
    if (ent >= 0) rest.push_back(ent);
  for (int32_t ent : rest_in)
    if (ent < 0) {
      tets_in.push_back(ent);
      const int32_t mate = P.tet_mate.empty() ? -1 : P.tet_mate[ent & 0x7fffffff];
      if (mate >= 0) tets_in.push_back((int32_t)(0x80000000u | (uint32_t)mate));
    }
  for (int32_t ent : tets_in) {
    const int32_t id = ent & 0x7fffffff;
    if (P.tet_e01[id] >= 0) rest.push_back(P.tet_e01[id]);
    if (P.tet_e23[id] >= 0) rest.push_back(P.tet_e23[id]);
  }
  for (int32_t ent : tets_in) rest.push_back(ent);
  std::vector<Mask128> me(P.V), mt(P.V);
  std::vector<uint8_t> col(rest.size());
  std::vector<uint32_t> ecount, tcount;
  for (size_t i = 0; i < rest.size(); i++) {
    int32_t vs[5];
    int n = ent_verts(D, rest[i], vs, false);
    bool tet = rest[i] < 0;
    std::vector<Mask128> &M = tet ? mt : me;
    Mask128 u;
    for (int k = 0; k < n; k++) {
      u.lo |= M[vs[k]].lo;
      u.hi |= M[vs[k]].hi;
    }
    int c = first_free(u);
    if (c < 0) return "vertex valence needs more than 128 colours";
    for (int k = 0; k < n; k++) set_bit(M[vs[k]], c);
    col[i] = (uint8_t)c;
    std::vector<uint32_t> &cnt = tet ? tcount : ecount;
    if ((size_t)c >= cnt.size()) cnt.resize(c + 1, 0);
    cnt[c]++;
  }
  std::vector<uint32_t> eoff(ecount.size() + 1, 0), tof(tcount.size() + 1, 0);
  for (size_t c = 0; c < ecount.size(); c++) eoff[c + 1] = eoff[c] + ecount[c];
  for (size_t c = 0; c < tcount.size(); c++) tof[c + 1] = tof[c] + tcount[c];
  const uint32_t e0 = (uint32_t)P.g_edges.size(), t0 = (uint32_t)P.g_tets.size(); // a second group appends
  for (auto &o : eoff) o += e0;
  for (auto &o : tof) o += t0;
  P.g_edges.resize(eoff.back());
  P.g_elen.resize(eoff.back());
  P.g_eid.resize(eoff.back());
  P.g_tets.resize(tof.back());
  P.g_trest.resize(tof.back());
  P.g_tid.resize(tof.back());
  for (size_t c = 0; c < ecount.size(); c++) P.gbatches.push_back({false, eoff[c], ecount[c], 0});
  for (size_t c = 0; c < tcount.size(); c++) P.gbatches.push_back({true, tof[c], tcount[c], 0});
  std::vector<uint32_t> ecur(eoff.begin(), eoff.end() - 1), tcur(tof.begin(), tof.end() - 1);
  for (size_t i = 0; i < rest.size(); i++) {
    int32_t vs[5];
    ent_verts(D, rest[i], vs, false);
    if (rest[i] >= 0) {
      uint32_t k = ecur[col[i]]++;
      P.g_edges[k] = {vs[0], vs[1]};
      P.g_elen[k] = P.rest_len[rest[i]];
      P.g_eid[k] = rest[i];
    } else {
      int32_t id = rest[i] & 0x7fffffff;
      uint32_t k = tcur[col[i]]++;
      P.g_tets[k] = {vs[0], vs[1], vs[2], vs[3]};
      P.g_trest[k] = P.rest_vol6[id];
      P.g_tid[k] = id;
    }
  }
  return "";
}

} // namespace

void lumped_inv_mass_into(const float *pos, uint32_t V, const int32_t *tets, uint32_t T, float density, float *out) {
  Plan tmp;
  tmp.V = V;
  tmp.T = T;
  tmp.pos.assign(pos, pos + 3 * (size_t)V);
  tmp.tets.assign(tets, tets + 4 * (size_t)T);
  lumped_inv_mass(tmp, density);
  std::copy(tmp.inv_mass.begin(), tmp.inv_mass.end(), out);
}

void Plan::export_schedule(std::vector<int32_t> &order, std::vector<int64_t> &batch_off, bool reverse) const {
  order.clear();
  batch_off.clear();
  batch_off.push_back(0);
  auto close = [&]() {
    if ((int64_t)order.size() > batch_off.back()) batch_off.push_back((int64_t)order.size());
  };
  for (int group = 0; group < 2; group++) {
  for (size_t pi = 0; pi < passes.size(); pi++) {
    const TilePass &TP = passes[reverse ? passes.size() - 1 - pi : pi];
    if (TP.group != group) continue;
    const uint32_t nt = TP.n_tiles();
    // start of every colour of every tile inside TP.ents
    std::vector<uint64_t> cstart(TP.col_cnt.size() + 1, 0);
    for (uint32_t t = 0; t < nt; t++) {
      uint64_t at = TP.ent_off[t];
      for (uint32_t j = TP.col_off[t]; j < TP.col_off[t + 1]; j++) {
        cstart[j] = at;
        at += TP.col_cnt[j];
      }
    }
    for (int kind = 0; kind < 2; kind++) {
      const uint32_t maxc = kind ? TP.max_tcol : TP.max_ecol;
      for (uint32_t c = 0; c < maxc; c++) {
        for (uint32_t t = 0; t < nt; t++) {
          const uint32_t nk = kind ? TP.col_off[t + 1] - TP.col_off[t] - TP.n_ecol[t] : TP.n_ecol[t];
          if (c >= nk) continue;
          const uint32_t j = TP.col_off[t] + (kind ? TP.n_ecol[t] : 0) + c;
          for (uint32_t k = 0; k < TP.col_cnt[j]; k++) order.push_back(TP.ents[cstart[j] + k]);
        }
        close();
        if (!kind) continue;
        // the edges attached to this colour's tets: a thread projects tet, edge (0,1), edge (2,3) in a row -- and then
        // the tet's mate and its two edges (bi-tets); records of one colour are vertex-disjoint, so "all tets, all
        // (0,1) edges, all (2,3) edges, all mates, ..." is the same computation written as independent batches
        for (int half = 0; half < 2; half++) {
          if (half) { // the mates
            for (uint32_t t = 0; t < nt; t++) {
              const uint32_t nk = TP.col_off[t + 1] - TP.col_off[t] - TP.n_ecol[t];
              if (c >= nk) continue;
              const uint32_t j = TP.col_off[t] + TP.n_ecol[t] + c;
              for (uint32_t k = 0; k < TP.col_cnt[j]; k++) {
                const int32_t m = TP.width == 2 ? tet_mate[TP.ents[cstart[j] + k] & 0x7fffffff] : -1;
                if (m >= 0) order.push_back((int32_t)(0x80000000u | (uint32_t)m));
              }
            }
            close();
          }
          for (int which = 0; which < 2; which++) {
            const std::vector<int32_t> &att = which ? tet_e23 : tet_e01;
            for (uint32_t t = 0; t < nt; t++) {
              const uint32_t nk = TP.col_off[t + 1] - TP.col_off[t] - TP.n_ecol[t];
              if (c >= nk) continue;
              const uint32_t j = TP.col_off[t] + TP.n_ecol[t] + c;
              for (uint32_t k = 0; k < TP.col_cnt[j]; k++) {
                int32_t id = TP.ents[cstart[j] + k] & 0x7fffffff;
                if (half) id = TP.width == 2 ? tet_mate[id] : -1;
                const int32_t e = id >= 0 ? att[id] : -1;
                if (e >= 0) order.push_back(e);
              }
            }
            close();
          }
        }
      }
    }
  }
  for (const GlobalBatch &b : gbatches) {
    if (b.group != group) continue;
    for (uint32_t k = 0; k < b.cnt; k++)
      order.push_back(b.tet ? (int32_t)(0x80000000u | (uint32_t)g_tid[b.off + k]) : g_eid[b.off + k]);
    close();
  }
  }
}

std::string build_plan(const MeshInput &in, const PlanOptions &opt, Plan &P) {
  auto t0 = std::chrono::steady_clock::now();
  const bool timing = getenv("SB_PLAN_TIMING") != nullptr;
  auto lap = [&](const char *what) {
    if (timing) fprintf(stderr, "[plan] %-28s %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  };
  if (!in.pos_xyz || in.n_verts == 0) return "no vertices";
  if (in.n_tets && !in.tets) return "tets is NULL";
  if (in.n_tris && !in.surf_tris) return "surf_tris is NULL";
  if (in.n_verts >= 0x7fffffffu || in.n_tets >= 0x7fffffffu) return "mesh too large for 31-bit ids";
  int threads = resolve_threads(opt.threads);
  P = Plan();
  P.V = in.n_verts;
  P.T = in.n_tets;
  P.F = in.n_tris;
  if (in.n_ghost > in.n_verts) return "n_ghost_verts exceeds n_verts";
  P.n_ghost = in.n_ghost;
  P.pos.assign(in.pos_xyz, in.pos_xyz + 3 * (size_t)P.V);
  for (float c : P.pos)
    if (!std::isfinite(c)) return "non-finite rest position";
  P.tets.assign(in.tets, in.tets + 4 * (size_t)P.T);
  if (P.F) P.tris.assign(in.surf_tris, in.surf_tris + 3 * (size_t)P.F);
  std::string err;
  if (in.edges) {
    // explicit edge set: canonical (a < b), sorted, unique; tets are still validated
    for (size_t i = 0; i < 4 * (size_t)P.T; i++)
      if (P.tets[i] < 0 || (uint32_t)P.tets[i] >= P.V) return "tet vertex index out of range";
    std::vector<uint64_t> keys(in.n_edges);
    for (uint32_t e = 0; e < in.n_edges; e++) {
      int32_t a = in.edges[2 * (size_t)e], b = in.edges[2 * (size_t)e + 1];
      if (a < 0 || b < 0 || (uint32_t)a >= P.V || (uint32_t)b >= P.V || a == b) return "edge vertex index out of range or repeated";
      if (a > b) std::swap(a, b);
      keys[e] = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    P.E = (uint32_t)keys.size();
    P.edges.resize(2 * (size_t)P.E);
    for (uint32_t e = 0; e < P.E; e++) {
      P.edges[2 * (size_t)e] = (int32_t)(keys[e] >> 32);
      P.edges[2 * (size_t)e + 1] = (int32_t)(keys[e] & 0xffffffffu);
    }
  } else {
    err = build_edges(P, threads);
    if (!err.empty()) return err;
  }
  if (in.inv_mass) {
    P.inv_mass.assign(in.inv_mass, in.inv_mass + P.V);
    for (float w : P.inv_mass)
      if (!(w >= 0) || !std::isfinite(w)) return "inv_mass must be finite and >= 0";
  } else {
    if (!(in.density > 0)) return "density must be > 0 when inv_mass is NULL";
    lumped_inv_mass(P, in.density);
  }
  lap("edges, masses");
  // round_width 2 (the default for a mesh that is not one rank of a partition): tets in face-sharing pairs
  if (opt.round_width < 0 || opt.round_width > 2) return "round_width must be 1 or 2";
  static const int auto_width = getenv("SB_ROUND_WIDTH") ? atoi(getenv("SB_ROUND_WIDTH")) : 1; // (A/B switch for benches)
  const bool bitets = !P.n_ghost && (opt.round_width == 2 || (opt.round_width == 0 && auto_width == 2 && opt.compounds != 0));
  if (bitets) pair_and_attach(P, threads, opt.compounds != 0);
  else attach_edges(P, opt.compounds != 0);
  lap("attach edges");
  rest_values(P, threads);
  err = build_surface(P);
  lap("rest values, surface");
  if (!err.empty()) return err;

  // ---- tiling ---------------------------------------------------------------
  // default tile: 1024 vertices; a mesh big enough to fill every SM with several of them gets boxes of up to
  // 1728 (four warps per tile at the same rounds per tile: more warps in flight per SM)
  uint32_t cap = opt.tile_cap > 0 ? (uint32_t)opt.tile_cap : (P.V >= 300000u ? 1728u : 1024u);
  cap = std::min(cap, 65536u);
  P.tile_cap = cap;
  const uint32_t bt_opt = opt.block_threads > 0 ? (uint32_t)opt.block_threads : 0u;
  if (bt_opt && bt_opt != 32 && bt_opt != 64 && bt_opt != 128 && bt_opt != 160 && bt_opt != 192 && bt_opt != 256)
    return "block_threads must be 32, 64, 128, 160, 192 or 256";
  const uint32_t width = bitets ? 2u : 1u;
  P.round_width = width;
  int max_passes = opt.max_tile_passes < 0 ? 6 : std::min(opt.max_tile_passes, 8);

  // largest connected component decides the layout: many small bodies are packed whole
  // into tiles (one pass, nothing cut); a big mesh gets the balanced shifted tilings
  int n_tilings = opt.tilings;
  std::vector<Tiling> tilings;
  if (max_passes == 0) n_tilings = 1;
  if (n_tilings <= 0) {
    Uf uf(P.V);
    for (uint32_t t = 0; t < P.T; t++) {
      const int32_t *q = &P.tets[4 * (size_t)t];
      uf.unite(q[0], q[1]); uf.unite(q[0], q[2]); uf.unite(q[0], q[3]);
    }
    std::vector<uint32_t> cs(P.V, 0);
    uint32_t biggest = 0;
    for (uint32_t v = 0; v < P.V; v++) biggest = std::max(biggest, ++cs[uf.find(v)]);
    // bodies that fit one tile are kept whole (one pass, nothing cut): raise the tile size to the largest
    // body when that still leaves shared memory for several tiles per SM
    if (biggest > cap && biggest <= 3072 && opt.tile_cap <= 0) {
      cap = biggest;
      P.tile_cap = cap;
    }
    n_tilings = biggest > cap ? 4 : 1;
  }
  n_tilings = std::min(n_tilings, 8);
  std::vector<uint32_t> atom_key;
  if (n_tilings >= 2 && !grid_tilings(P, cap, n_tilings, tilings, atom_key, opt.dist_ranks)) n_tilings = 1;
  P.n_tilings = (uint32_t)n_tilings;

  std::vector<uint32_t> tile_off;
  std::vector<uint32_t> block_first_tile; // dist_ranks + 1 when the boxes were numbered block by block
  if (n_tilings >= 2 && opt.dist_ranks >= 2 && tilings[0].n_tiles >= (uint32_t)opt.dist_ranks) {
    // One mesh over several GPUs: cut the boxes of the unshifted tiling into dist_ranks compact blocks by recursive
    // coordinate bisection of their centroids (weights = vertex counts; the cut runs along the longest axis of the
    // part, ties ordered by the other two axes so that a cut through a layer of boxes stays one piece) and number
    // the boxes block by block: every rank then owns one range of device ids, and the surface between ranks -- the
    // tiles that straddle two GPUs -- is that of a block, not of a slab of the default x-fastest order.
    Tiling &T0 = tilings[0];
    const uint32_t nt = T0.n_tiles, K = (uint32_t)opt.dist_ranks;
    std::vector<double> cen(3 * (size_t)nt, 0.0);
    std::vector<uint64_t> wgt(nt, 0);
    for (uint32_t v = 0; v < P.V; v++) {
      const uint32_t t = (uint32_t)T0.part[v];
      for (int k = 0; k < 3; k++) cen[3 * (size_t)t + k] += P.pos[3 * (size_t)v + k];
      wgt[t]++;
    }
    for (uint32_t t = 0; t < nt; t++)
      for (int k = 0; k < 3; k++) cen[3 * (size_t)t + k] /= (double)std::max<uint64_t>(1, wgt[t]);
    std::vector<uint32_t> idx(nt), block(nt, 0);
    std::iota(idx.begin(), idx.end(), 0u);
    std::function<void(size_t, size_t, uint32_t, uint32_t)> split = [&](size_t lo, size_t hi, uint32_t k, uint32_t first) {
      if (k <= 1) {
        for (size_t i = lo; i < hi; i++) block[idx[i]] = first;
        return;
      }
      double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
      for (size_t i = lo; i < hi; i++)
        for (int c = 0; c < 3; c++) {
          mn[c] = std::min(mn[c], cen[3 * (size_t)idx[i] + c]);
          mx[c] = std::max(mx[c], cen[3 * (size_t)idx[i] + c]);
        }
      int ax[3] = {2, 1, 0}; // ties between equal extents: z first (the slowest axis of the box order)
      std::stable_sort(ax, ax + 3, [&](int a, int b) { return mx[a] - mn[a] > mx[b] - mn[b]; });
      // box centroids of one layer differ by the jitter of their vertices: compare them on a coarse grid
      const double q = std::max(1e-30, (mx[ax[0]] - mn[ax[0]]) * 1e-4);
      auto key = [&](uint32_t t, int c) { return std::floor((cen[3 * (size_t)t + c] - mn[c]) / q); };
      std::sort(idx.begin() + (ptrdiff_t)lo, idx.begin() + (ptrdiff_t)hi, [&](uint32_t a, uint32_t b) {
        for (int j = 0; j < 3; j++) {
          const double ka = key(a, ax[j]), kb = key(b, ax[j]);
          if (ka != kb) return ka < kb;
        }
        return a < b;
      });
      const uint32_t kl = k / 2;
      uint64_t total = 0, acc = 0;
      for (size_t i = lo; i < hi; i++) total += wgt[idx[i]];
      size_t cut = lo;
      while (cut < hi && (acc + wgt[idx[cut]] / 2) * k <= total * kl) acc += wgt[idx[cut++]];
      cut = std::min(std::max(cut, lo + kl), hi - (k - kl)); // at least one box per block
      split(lo, cut, kl, first);
      split(cut, hi, k - kl, first + kl);
    };
    split(0, nt, K, 0);
    // new box ids: by block, old order inside a block
    std::vector<uint32_t> order(nt), newid(nt);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return block[a] < block[b]; });
    block_first_tile.assign((size_t)K + 1, nt);
    for (uint32_t i = 0; i < nt; i++) {
      newid[order[i]] = i;
      block_first_tile[block[order[i]]] = std::min(block_first_tile[block[order[i]]], i);
    }
    for (uint32_t r = K; r-- > 0;) block_first_tile[r] = std::min(block_first_tile[r], block_first_tile[r + 1]);
    for (uint32_t v = 0; v < P.V; v++) T0.part[v] = (int32_t)newid[(uint32_t)T0.part[v]];
  }
  if (n_tilings >= 2) {
    // device numbering: tiles of tiling 0 are contiguous, ascending caller id inside a tile
    const Tiling &T0 = tilings[0];
    std::vector<uint32_t> cnt((size_t)T0.n_tiles + 1, 0);
    for (uint32_t v = 0; v < P.V; v++) cnt[(size_t)T0.part[v] + 1]++;
    for (uint32_t t = 0; t < T0.n_tiles; t++) cnt[t + 1] += cnt[t];
    tile_off.assign(cnt.begin(), cnt.end());
    P.perm.resize(P.V);
    for (uint32_t v = 0; v < P.V; v++) P.perm[cnt[T0.part[v]]++] = v;
    parallel_for(T0.n_tiles, threads, 1, [&](size_t t, int) {
      std::sort(P.perm.begin() + tile_off[t], P.perm.begin() + tile_off[t + 1], [&](uint32_t a, uint32_t b) {
        return atom_key[a] < atom_key[b] || (atom_key[a] == atom_key[b] && a < b);
      });
    });
    P.inv.resize(P.V);
    for (uint32_t d = 0; d < P.V; d++) P.inv[P.perm[d]] = d;
    if (!block_first_tile.empty()) {
      P.dist_ranks = (uint32_t)opt.dist_ranks;
      P.dist_slab_lo.resize(block_first_tile.size());
      for (size_t r = 0; r < block_first_tile.size(); r++) P.dist_slab_lo[r] = tile_off[block_first_tile[r]];
    }
  } else if (max_passes > 0) {
    tile_off = tile_vertices(P, cap, opt.n_sm);
  } else {
    P.perm.resize(P.V);
    std::iota(P.perm.begin(), P.perm.end(), 0u);
    P.inv = P.perm;
  }
  lap("tilings, numbering");
  DevTopo D;
  D.edges.resize(2 * (size_t)P.E);
  D.tets.resize(4 * (size_t)P.T);
  parallel_for(2 * (size_t)P.E, threads, 1 << 18, [&](size_t i, int) { D.edges[i] = (int32_t)P.inv[P.edges[i]]; });
  parallel_for(4 * (size_t)P.T, threads, 1 << 18, [&](size_t i, int) { D.tets[i] = (int32_t)P.inv[P.tet_roles[i]]; });
  if (P.tets_paired && n_tilings >= 2) {
    // a bi-tet that no tiling holds in one box (its five vertices span more than an element does: small boxes) is
    // split back into its two tets -- their roles and attached edges are valid on their own -- rather than left over
    uint64_t split = 0;
    for (uint32_t t = 0; t < P.T; t++) {
      if (!P.tet_lead[t] || P.tet_mate[t] < 0) continue;
      const int32_t u = P.tet_mate[t];
      const int32_t five[5] = {P.tet_roles[4 * (size_t)t], P.tet_roles[4 * (size_t)t + 1], P.tet_roles[4 * (size_t)t + 2],
                               P.tet_roles[4 * (size_t)t + 3], P.tet_roles[4 * (size_t)u]};
      bool somewhere = false;
      for (int s2 = 0; s2 < n_tilings && !somewhere; s2++) {
        bool in = true;
        for (int k = 1; k < 5; k++) in &= tilings[s2].part[five[k]] == tilings[s2].part[five[0]];
        somewhere = in;
      }
      if (!somewhere) {
        P.tet_mate[t] = P.tet_mate[u] = -1;
        P.tet_lead[u] = 1;
        split += 2;
      }
    }
    P.tets_paired -= split;
  }
  if (P.tets_paired) {
    D.apex.assign(P.T, -1);
    for (uint32_t t = 0; t < P.T; t++)
      if (P.tet_lead[t] && P.tet_mate[t] >= 0) D.apex[t] = (int32_t)P.inv[P.tet_roles[4 * (size_t)P.tet_mate[t]]]; // role 0 of B is its apex
  }
  std::vector<float> dev_pos(3 * (size_t)P.V);
  for (uint32_t d = 0; d < P.V; d++)
    for (int k = 0; k < 3; k++) dev_pos[3 * (size_t)d + k] = P.pos[3 * (size_t)P.perm[d] + k];

  auto plan_group = [&](const std::vector<int32_t> &cons, int group) -> std::string {
  const size_t first_pass = P.passes.size(), first_gb = P.gbatches.size();
  std::vector<int32_t> work, next;
  std::vector<int32_t> part(P.V);
  uint32_t n_tiles = 0;
  bool have_parts = false;

  if (n_tilings >= 2) {
    // parts in device numbering
    std::vector<std::vector<int32_t>> dpart(n_tilings, std::vector<int32_t>(P.V));
    for (int s = 0; s < n_tilings; s++)
      parallel_for(P.V, threads, 1 << 18, [&](size_t d, int) { dpart[s][d] = tilings[s].part[P.perm[d]]; });
    // deal every constraint to the least-loaded tiling it is interior to
    std::vector<std::vector<int32_t>> assigned(n_tilings);
    std::vector<uint64_t> load(n_tilings, 0);
    std::vector<uint8_t> mask(cons.size());
    parallel_for(mask.size(), threads, 1 << 16, [&](size_t i, int) {
      const int32_t ent = cons[i];
      int32_t vs[5];
      const int n = ent_verts(D, ent, vs);
      uint8_t m = 0;
      for (int s = 0; s < n_tilings; s++) {
        const int32_t p = dpart[s][vs[0]];
        bool in = true;
        for (int k = 1; k < n; k++) in &= dpart[s][vs[k]] == p;
        if (in) m |= (uint8_t)(1u << s);
      }
      mask[i] = m;
    });
    // Colours needed in a tile ~ the largest number of same-kind constraints meeting at one of
    // its vertices, so deal each constraint to the eligible tiling where its vertices carry
    // the fewest so far (ties: the globally least-loaded tiling).
    std::vector<std::vector<uint8_t>> deg(2 * (size_t)n_tilings, std::vector<uint8_t>(P.V, 0));
    // most constrained first: constraints with one eligible tiling have no choice, so they are placed
    // before the flexible ones, which then fill in around them
    std::vector<uint32_t> deal_order(mask.size());
    {
      uint32_t cnt[9] = {0};
      for (size_t i = 0; i < mask.size(); i++) cnt[__builtin_popcount(mask[i])]++;
      uint32_t start[9], acc = 0;
      for (int b = 0; b < 9; b++) { start[b] = acc; acc += cnt[b]; }
      for (size_t i = 0; i < mask.size(); i++) deal_order[start[__builtin_popcount(mask[i])]++] = (uint32_t)i;
    }
    std::vector<int8_t> chosen(mask.size(), -1);
    for (size_t oi = 0; oi < mask.size(); oi++) {
      const size_t i = deal_order[oi];
      const int32_t ent = cons[i];
      const int kind = ent < 0;
      // tets weigh more than edges in the kernels, and carry their attached edges
      uint64_t wgt = 2;
      if (kind) {
        const int32_t ta = ent & 0x7fffffff, tb = P.tet_mate[ta];
        wgt = 5 + 2 * ((P.tet_e01[ta] >= 0) + (P.tet_e23[ta] >= 0));
        if (tb >= 0) wgt += 5 + 2 * ((P.tet_e01[tb] >= 0) + (P.tet_e23[tb] >= 0));
      }
      int32_t vs[5];
      const int n = ent_verts(D, ent, vs);
      int best = -1;
      uint32_t best_deg = 0, best_sum = 0;
      // cut constraints (group 1) are few and their tiles nearly empty: there the cost is the number of
      // passes, not the rounds per tile, so they all go to the first tiling that can take them
      if (group == 1 && mask[i]) best = __builtin_ctz(mask[i]);
      for (int s = 0; s < n_tilings && group != 1; s++) {
        if (!(mask[i] >> s & 1)) continue;
        const std::vector<uint8_t> &dg = deg[2 * (size_t)s + kind];
        uint32_t md = 0, sd = 0;
        for (int k = 0; k < n; k++) {
          md = std::max<uint32_t>(md, dg[vs[k]]);
          sd += dg[vs[k]];
        }
        if (best < 0 || md < best_deg || (md == best_deg && (sd < best_sum || (sd == best_sum && load[s] < load[best])))) {
          best = s;
          best_deg = md;
          best_sum = sd;
        }
      }
      if (best < 0) continue;
      chosen[i] = (int8_t)best;
      load[best] += wgt;
      std::vector<uint8_t> &dg = deg[2 * (size_t)best + kind];
      for (int k = 0; k < n; k++)
        if (dg[vs[k]] < 255) dg[vs[k]]++;
    }
    lap("  deal to tilings");
    // keep ascending constraint id inside every tiling (the tile builder relies on a stable order)
    for (size_t i = 0; i < mask.size(); i++) {
      if (chosen[i] < 0) work.push_back(cons[i]);
      else assigned[chosen[i]].push_back(cons[i]);
    }
    deg.clear();
    mask.clear();
    mask.shrink_to_fit();
    bool wide = width == 1;
    for (int s = 0; s < n_tilings; s++) { // (a mesh spread over several GPUs: a rank runs its share of the tiles, a few more on some ranks)
      const uint64_t per_rank = opt.dist_ranks >= 2 ? ((uint64_t)tilings[s].n_tiles * 27 / 25 + opt.dist_ranks - 1) / opt.dist_ranks : tilings[s].n_tiles;
      wide = wide && per_rank <= 6ull * (uint64_t)std::max(1, opt.n_sm);
    }
    for (int s = 0; s < n_tilings; s++) {
      TilePass TP;
      // every tiling runs with the CTA width chosen for the unshifted one (whole boxes only)
      err = build_pass(P, D, dpart[s], tilings[s].n_tiles, s == 0 ? &tile_off : nullptr, assigned[s], next, TP, threads,
                       s == 0 ? bt_opt : P.passes[P.passes.size() - (size_t)s].bt, width, opt.n_sm, s == 0 ? nullptr : &dpart[0], wide);
      if (!err.empty()) return err;
      work.insert(work.end(), next.begin(), next.end()); // empty by construction
      P.passes.push_back(std::move(TP));
      assigned[s] = std::vector<int32_t>();
      lap("  tile pass built");
    }
    // dependencies between consecutive tilings (cyclic): tile t of tiling s must wait for every tile of
    // tiling s-1 that shares a vertex with it, by full box membership so that chains through vertices a
    // pass does not touch still order passes two apart
    if (group == 0) {
      const size_t base_pass = P.passes.size() - (size_t)n_tilings;
      for (int s = 0; s < n_tilings; s++) {
        const int q = (s + n_tilings - 1) % n_tilings;
        std::vector<uint64_t> pairs(P.V);
        parallel_for(P.V, threads, 1 << 18, [&](size_t d, int) { pairs[d] = ((uint64_t)(uint32_t)dpart[s][d] << 32) | (uint32_t)dpart[q][d]; });
        std::sort(pairs.begin(), pairs.end());
        pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
        TilePass &TP = P.passes[base_pass + s];
        TP.dep_off.assign((size_t)tilings[s].n_tiles + 1, 0);
        for (uint64_t pr : pairs) TP.dep_off[(size_t)(pr >> 32) + 1]++;
        for (uint32_t t = 0; t < tilings[s].n_tiles; t++) TP.dep_off[t + 1] += TP.dep_off[t];
        TP.dep_list.resize(pairs.size());
        for (size_t k = 0; k < pairs.size(); k++) TP.dep_list[k] = (uint32_t)pairs[k]; // sorted by (t, q): in place
      }
    }
    lap("  tile dependencies");
    part = dpart[0];
    n_tiles = tilings[0].n_tiles;
    have_parts = true;
    std::sort(work.begin(), work.end(), [](int32_t a, int32_t b) { return (uint32_t)a < (uint32_t)b; });
  } else {
    work = cons;
    if (max_passes > 0) {
      n_tiles = (uint32_t)tile_off.size() - 1;
      for (uint32_t t = 0; t < n_tiles; t++)
        for (uint32_t d = tile_off[t]; d < tile_off[t + 1]; d++) part[d] = (int32_t)t;
      have_parts = true;
    }
  }

  // hierarchical passes: constraints interior to a tile are taken, the cut ones go to a
  // next level whose tiles straddle the previous level's boundaries (the only scheme when
  // n_tilings == 1; the fallback for whatever the tilings left over otherwise)
  if (have_parts && !work.empty()) {
    auto level_cap = [&]() {
      uint32_t lcap = opt.later_cap > 0 ? (uint32_t)opt.later_cap : 0;
      if (!lcap) {
        std::vector<uint8_t> seen(P.V, 0);
        size_t nvr = 0;
        for (int32_t ent : work) {
          int32_t vs[5];
          int n = ent_verts(D, ent, vs);
          for (int j = 0; j < n; j++)
            if (!seen[vs[j]]) { seen[vs[j]] = 1; nvr++; }
        }
        lcap = (uint32_t)std::min<size_t>(cap, std::max<size_t>(2048, (nvr + 2 * opt.n_sm - 1) / (2 * opt.n_sm)));
      }
      return std::min(lcap, cap);
    };
    bool first_level = n_tilings < 2; // the tilings already took everything interior to tiling 0
    if (!first_level) n_tiles = next_parts(P, D, work, part, level_cap(), dev_pos);
    const int levels = std::min(max_passes, 8 - (int)P.passes.size()); // sb_info reports at most 8 passes
    for (int k = 0; k < levels && !work.empty(); k++) {
      TilePass TP;
      const bool contig = first_level && k == 0;
      err = build_pass(P, D, part, n_tiles, contig ? &tile_off : nullptr, work, next, TP, threads, bt_opt, width, opt.n_sm);
      if (!err.empty()) return err;
      size_t consumed = work.size() - next.size();
      work.swap(next);
      if (consumed > 0 || (contig && group == 0)) P.passes.push_back(std::move(TP));
      if (work.empty() || k + 1 == levels) break;
      if (consumed == 0 && !contig) break; // no progress: hand the rest to the global colours
      n_tiles = next_parts(P, D, work, part, level_cap(), dev_pos);
    }
  }
  err = global_colouring(P, D, work);
  if (!err.empty()) return err;
  for (size_t k = first_pass; k < P.passes.size(); k++) P.passes[k].group = group;
  for (size_t k = first_gb; k < P.gbatches.size(); k++) P.gbatches[k].group = group;
  return "";
  };

  // constraint groups: 0 = no ghost vertex (interior), 1 = touches a ghost (owned cut constraint);
  // constraints among ghosts only belong to another rank and are dropped
  const uint32_t first_ghost = P.V - P.n_ghost;
  std::vector<int32_t> cons0, cons1;
  cons0.reserve((size_t)P.E + P.T);
  for (size_t i = 0; i < (size_t)P.E + P.T; i++) {
    const int32_t ent = i < P.E ? (int32_t)i : (int32_t)(0x80000000u | (uint32_t)(i - P.E));
    if (i < P.E && P.edge_owner[i] >= 0) continue; // rides with its tet
    if (i >= P.E && !P.tet_lead[i - P.E]) continue;  // rides with its mate
    int ng = 0, n;
    if (ent >= 0) {
      n = 2;
      ng = ((uint32_t)P.edges[2 * (size_t)ent] >= first_ghost) + ((uint32_t)P.edges[2 * (size_t)ent + 1] >= first_ghost);
    } else {
      n = 4;
      const int32_t *q = &P.tets[4 * (size_t)(ent & 0x7fffffff)];
      for (int k = 0; k < 4; k++) ng += (uint32_t)q[k] >= first_ghost;
    }
    if (ng == 0) cons0.push_back(ent);
    else if (ng < n) cons1.push_back(ent);
  }
  err = plan_group(cons0, 0);
  if (!err.empty()) return err;
  lap("passes (deal, colour, order)");
  P.dag_ok = n_tilings >= 2 && !P.n_ghost && P.passes.size() == (size_t)n_tilings && P.gbatches.empty();
  if (P.n_ghost) {
    err = plan_group(cons1, 1);
    if (!err.empty()) return err;
  }
  P.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return "";
}

} // namespace sb
