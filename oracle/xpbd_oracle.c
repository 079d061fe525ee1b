/*
 * xpbd_oracle.c -- CPU oracle for the soft-body substep path.
 *
 * TEST INFRASTRUCTURE ONLY: nothing under softbodyunity_b200/ may include, link
 * or call this file; only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline do.
 *
 * PARITY UNPINNED.  The reference mount is /root/reference/README.md:1
 * ("# SoftbodyUnity") and nothing else: no C# solver, no tests, no golden
 * vectors.  This file therefore restates the
 * PUBLISHED algorithm the task names (XPBD with distance and tet-volume
 * constraints, BASELINE.json:5,8) and is pinned only by analytic known-answer
 * tests (tests/test_oracle_kat.py) and by an independent numpy restatement
 * (tests/np_xpbd.py).  Parity with the upstream C# solver is unverified.
 *
 * The arithmetic contract is in xpbd_oracle_impl.h.  Build: see oracle/Makefile
 * (gcc -O2 -ffp-contract=off -fopenmp; FMA only where written).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_FLAG_NO_GROUND 1

/* Same field order and meaning as sb_params in include/softbody_b200.h. */
typedef struct {
  float dt;
  int32_t substeps;
  int32_t iterations;
  float stiffness_distance; /* N/m; +inf = rigid (compliance 0); <= 0 = constraint family off */
  float stiffness_volume;   /* N/m^5-ish (1/compliance); same convention */
  float damping;            /* 1/s, v *= max(0, 1 - h*damping) */
  float friction;           /* 0..1, fraction of tangential motion removed on ground contact */
  float gravity[3];
  float ground_y;
  int32_t flags;
} orc_params;

/* stiffness -> compliance [SPEC]: 1/k for finite k > 0; 0 for +inf; -1 (off) otherwise */
static inline float orc_compliance(float k) {
  if (isinf(k) && k > 0) return 0.0f;
  if (k > 0) return 1.0f / k;
  return -1.0f;
}

int orc_abi_sizeof_params(void) { return (int)sizeof(orc_params); }

/* ---- topology: canonical edge list ------------------------------------- */

static int cmp_u64(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x < y ? -1 : x > y;
}

/*
 * Unique undirected edges of the tets, each as (a,b) with a < b, sorted by
 * (a, then b).  Call with edges == NULL to get the count.  Returns E or < 0.
 */
int64_t orc_build_edges(int32_t V, int32_t T, const int32_t *tets, int32_t *edges) {
  static const int pr[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
  size_t n = 6 * (size_t)T;
  uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (n ? n : 1));
  if (!keys) return -3;
  for (int32_t t = 0; t < T; t++)
    for (int k = 0; k < 6; k++) {
      int32_t a = tets[4 * (size_t)t + pr[k][0]], b = tets[4 * (size_t)t + pr[k][1]];
      if (a < 0 || b < 0 || a >= V || b >= V || a == b) { free(keys); return -2; }
      if (a > b) { int32_t s = a; a = b; b = s; }
      keys[6 * (size_t)t + k] = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
    }
  qsort(keys, n, sizeof(uint64_t), cmp_u64);
  int64_t E = 0;
  for (size_t i = 0; i < n; i++)
    if (i == 0 || keys[i] != keys[i - 1]) {
      if (edges) { edges[2 * E] = (int32_t)(keys[i] >> 32); edges[2 * E + 1] = (int32_t)(keys[i] & 0xffffffffu); }
      E++;
    }
  free(keys);
  return E;
}

/*
 * Lumped masses: m_i = sum over tets containing i, ascending tet index, of
 * (density * |det|/6) * 0.25, accumulated in double from the float positions;
 * inv_mass_i = (float)(1/m_i), or 0 when m_i == 0.
 */
void orc_lumped_inv_mass(int32_t V, const float *pos_xyz, int32_t T, const int32_t *tets, float density,
                         float *inv_mass) {
  double *m = (double *)calloc((size_t)(V > 0 ? V : 1), sizeof(double));
  for (int32_t t = 0; t < T; t++) {
    const int32_t *q = tets + 4 * (size_t)t;
    double p[4][3];
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 3; k++) p[j][k] = (double)pos_xyz[3 * (size_t)q[j] + k];
    double e1[3], e2[3], e3[3];
    for (int k = 0; k < 3; k++) { e1[k] = p[1][k] - p[0][k]; e2[k] = p[2][k] - p[0][k]; e3[k] = p[3][k] - p[0][k]; }
    double cx = e2[1] * e3[2] - e2[2] * e3[1];
    double cy = e2[2] * e3[0] - e2[0] * e3[2];
    double cz = e2[0] * e3[1] - e2[1] * e3[0];
    double det = e1[0] * cx + e1[1] * cy + e1[2] * cz;
    double share = ((double)density * (fabs(det) / 6.0)) * 0.25;
    for (int j = 0; j < 4; j++) m[q[j]] += share;
  }
  for (int32_t i = 0; i < V; i++) inv_mass[i] = m[i] > 0 ? (float)(1.0 / m[i]) : 0.0f;
  free(m);
}

/*
 * Diagnostics, accumulated in double in ascending index order:
 *  out[0] kinetic energy, out[1] gravitational potential (-sum m g.x),
 *  out[2] total volume (sum det6/6), out[3..5] mean position,
 *  out[6..8] linear momentum, out[9..11] angular momentum about the origin,
 *  out[12] max |len-L0|/L0, out[13] rms of the same, out[14] count of non-finite
 *  position components, out[15] min y.
 */
void orc_diagnostics(int32_t V, const float *x4, const float *v4, int32_t E, const int32_t *edges,
                     const float *rest_len, int32_t T, const int32_t *tets, const float *gravity,
                     double *out) {
  for (int k = 0; k < 16; k++) out[k] = 0;
  double miny = INFINITY;
  for (int32_t i = 0; i < V; i++) {
    const float *x = x4 + 4 * (size_t)i, *v = v4 + 4 * (size_t)i;
    for (int k = 0; k < 3; k++) {
      if (!isfinite(x[k])) out[14] += 1;
      out[3 + k] += x[k];
    }
    if (x[1] < miny) miny = x[1];
    if (x[3] > 0) {
      double m = 1.0 / (double)x[3];
      double vx = v[0], vy = v[1], vz = v[2];
      out[0] += 0.5 * m * (vx * vx + vy * vy + vz * vz);
      out[1] -= m * ((double)gravity[0] * x[0] + (double)gravity[1] * x[1] + (double)gravity[2] * x[2]);
      out[6] += m * vx; out[7] += m * vy; out[8] += m * vz;
      out[9] += m * ((double)x[1] * vz - (double)x[2] * vy);
      out[10] += m * ((double)x[2] * vx - (double)x[0] * vz);
      out[11] += m * ((double)x[0] * vy - (double)x[1] * vx);
    }
  }
  if (V > 0) { out[3] /= V; out[4] /= V; out[5] /= V; }
  out[15] = miny;
  for (int32_t t = 0; t < T; t++) {
    const int32_t *q = tets + 4 * (size_t)t;
    double e[3][3];
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < 3; k++) e[j][k] = (double)x4[4 * (size_t)q[j + 1] + k] - (double)x4[4 * (size_t)q[0] + k];
    double cx = e[1][1] * e[2][2] - e[1][2] * e[2][1];
    double cy = e[1][2] * e[2][0] - e[1][0] * e[2][2];
    double cz = e[1][0] * e[2][1] - e[1][1] * e[2][0];
    out[2] += (e[0][0] * cx + e[0][1] * cy + e[0][2] * cz) / 6.0;
  }
  double ss = 0;
  for (int32_t e = 0; e < E; e++) {
    const float *a = x4 + 4 * (size_t)edges[2 * (size_t)e], *b = x4 + 4 * (size_t)edges[2 * (size_t)e + 1];
    double dx = (double)a[0] - b[0], dy = (double)a[1] - b[1], dz = (double)a[2] - b[2];
    double r = fabs(sqrt(dx * dx + dy * dy + dz * dz) - (double)rest_len[e]) / (double)rest_len[e];
    if (r > out[12]) out[12] = r;
    ss += r * r;
  }
  out[13] = E > 0 ? sqrt(ss / E) : 0;
}

/* ---- the solver, fp32 then fp64 ----------------------------------------- */

#define REAL float
#define SUFFIX _f32
#define FMA fmaf
#define SQRT sqrtf
#include "xpbd_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef FMA
#undef SQRT

#define REAL double
#define SUFFIX _f64
#define FMA fma
#define SQRT sqrt
#include "xpbd_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef FMA
#undef SQRT
