/*
 * xpbd_oracle_impl.h -- body of the CPU oracle, included twice by xpbd_oracle.c
 * (once with REAL=float, once with REAL=double).
 *
 * TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: the reference mount holds only
 * /root/reference/README.md:1 ("# SoftbodyUnity"); there is no upstream solver,
 * test or golden vector to pin this against.  What is restated here is the
 * published XPBD soft-body formulation (Macklin, Mueller, Chentanez 2016,
 * "XPBD"; Macklin et al. 2019, "Small Steps"; distance + tet-volume constraints
 * as in Mueller's "Ten Minute Physics" soft-body demo) in the stage order that
 * BASELINE.json:5 names: predict -> project (I iterations) -> ground/collider
 * -> velocity update.  Every formula is [SPEC].
 *
 * ARITHMETIC CONTRACT (the CUDA kernels in exact mode follow it operation for
 * operation, so fp32 results are bit-identical; this file is compiled with
 * -ffp-contract=off so the only fused operations are the explicit FMA() calls):
 *
 *   h      = dt / (REAL)substeps
 *   inv_h  = 1 / h
 *   a_d    = compliance_distance / (h*h)            (alpha-tilde, distance)
 *   a_v36  = 36 * (compliance_volume / (h*h))       (alpha-tilde, volume, x36:
 *                                                    gradients are kept x6)
 *   damp   = max(0, 1 - h*damping)
 *   keep   = 1 - friction
 *
 *   predict(i), w_i > 0:  v = FMA(h, g, v);  x_prev = x;  x = FMA(h, v, x)
 *            w_i == 0:    x_prev = x
 *   distance(a,b):  d = x_a - x_b;  len2 = FMA(dz,dz, FMA(dy,dy, dx*dx))
 *                   skip unless w_a + w_b > 0 and 2^-101 <= len2 <= FLT_MAX
 *                   len = sqrt(len2);  C = len - L0;  den = (w_a + w_b + a_d) * len
 *                   skip unless 2^-126 <= den < 2^126
 *                   s = -C * RCP(den)                          RCP(x) = correctly rounded 1/x
 *                   x_a = FMA(s*w_a, d, x_a);  x_b = FMA(-(s*w_b), d, x_b)
 *   volume(p0..p3): e_k = x_pk - x_p0 (k=1..3)
 *                   G1 = e2 x e3, G2 = e3 x e1, G3 = e1 x e2, G0 = -((G1+G2)+G3)
 *                   cross(a,b).x = FMA(a.y, b.z, -(a.z*b.y))   (cyclic)
 *                   det = FMA(e1.z,G1.z, FMA(e1.y,G1.y, e1.x*G1.x))
 *                   n_k = FMA(G.z,G.z, FMA(G.y,G.y, G.x*G.x))
 *                   den = FMA(w3,n3, FMA(w2,n2, FMA(w1,n1, w0*n0))) + a_v36
 *                   skip unless 2^-126 <= den < 2^126
 *                   s = -(det - R6) * RCP(den);   x_pk = FMA(s*w_k, G_k, x_pk)
 *   (a reciprocal and a multiply instead of one division: the GPU's IEEE division almost always
 *    leaves its fast path for these operands.  The operand windows are those in which the GPU's
 *    branch-free sqrt / reciprocal sequences -- MUFU seed + FMA correction -- are correctly
 *    rounded; outside them (lengths below 1e-15 or non-finite state) a projection is skipped
 *    instead of taking a slow path, here and in the kernels alike.)
 *   finish(i), w_i > 0:   if x.y < ground_y: x.y = ground_y,
 *                             x.xz = FMA(keep, x.xz - x_prev.xz, x_prev.xz)
 *                         colliders: see COLLIDERS below
 *                         v = ((x - x_prev) * inv_h) * damp
 */

#ifndef REAL
#error "include from xpbd_oracle.c"
#endif

/* operand windows of the arithmetic contract (same numbers for float and double) */
#ifndef ORC_SQRT_MIN
#define ORC_SQRT_MIN ((REAL)0x1p-101)
#define ORC_SQRT_MAX ((REAL)3.40282346638528859812e+38) /* FLT_MAX */
#define ORC_RCP_MIN ((REAL)0x1p-126)
#define ORC_RCP_MAX ((REAL)0x1p126)
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

typedef struct {
  REAL h, inv_h, a_d, a_v36, damp, keep, gx, gy, gz, ground_y;
  int use_d, use_v, use_ground;
} FN(step_consts);

static FN(step_consts) FN(make_consts)(const orc_params *p) {
  FN(step_consts) c;
  c.h = (REAL)p->dt / (REAL)p->substeps;
  c.inv_h = (REAL)1 / c.h;
  REAL hh = c.h * c.h;
  REAL cd = orc_compliance(p->stiffness_distance);
  REAL cv = orc_compliance(p->stiffness_volume);
  c.use_d = cd >= 0;
  c.use_v = cv >= 0;
  c.a_d = c.use_d ? cd / hh : 0;
  c.a_v36 = c.use_v ? (REAL)36 * (cv / hh) : 0;
  REAL dm = (REAL)1 - c.h * (REAL)p->damping;
  c.damp = dm > 0 ? dm : 0;
  c.keep = (REAL)1 - (REAL)p->friction;
  c.gx = p->gravity[0]; c.gy = p->gravity[1]; c.gz = p->gravity[2];
  c.ground_y = p->ground_y;
  c.use_ground = !(p->flags & ORC_FLAG_NO_GROUND);
  return c;
}

static inline void FN(predict_one)(REAL *x, REAL *xp, REAL *v, const FN(step_consts) *c) {
  if (x[3] > 0) {
    v[0] = FMA(c->h, c->gx, v[0]);
    v[1] = FMA(c->h, c->gy, v[1]);
    v[2] = FMA(c->h, c->gz, v[2]);
    xp[0] = x[0]; xp[1] = x[1]; xp[2] = x[2];
    x[0] = FMA(c->h, v[0], x[0]);
    x[1] = FMA(c->h, v[1], x[1]);
    x[2] = FMA(c->h, v[2], x[2]);
  } else {
    xp[0] = x[0]; xp[1] = x[1]; xp[2] = x[2];
  }
}

static inline void FN(distance_one)(REAL *xa, REAL *xb, REAL L0, REAL a_d) {
  REAL wa = xa[3], wb = xb[3];
  REAL wsum = wa + wb;
  REAL dx = xa[0] - xb[0], dy = xa[1] - xb[1], dz = xa[2] - xb[2];
  REAL len2 = FMA(dz, dz, FMA(dy, dy, dx * dx));
  if (!(wsum > 0) || !(len2 >= ORC_SQRT_MIN && len2 <= ORC_SQRT_MAX)) return;
  REAL len = SQRT(len2);
  REAL C = len - L0;
  REAL den = (wsum + a_d) * len;
  if (!(den >= ORC_RCP_MIN && den < ORC_RCP_MAX)) return;
  REAL s = -C * ((REAL)1 / den);
  REAL sa = s * wa, sb = -(s * wb);
  xa[0] = FMA(sa, dx, xa[0]); xa[1] = FMA(sa, dy, xa[1]); xa[2] = FMA(sa, dz, xa[2]);
  xb[0] = FMA(sb, dx, xb[0]); xb[1] = FMA(sb, dy, xb[1]); xb[2] = FMA(sb, dz, xb[2]);
}

#define CROSS(o, a, b)                                  \
  do {                                                  \
    (o)[0] = FMA((a)[1], (b)[2], -((a)[2] * (b)[1]));   \
    (o)[1] = FMA((a)[2], (b)[0], -((a)[0] * (b)[2]));   \
    (o)[2] = FMA((a)[0], (b)[1], -((a)[1] * (b)[0]));   \
  } while (0)
#define DOT3(a, b) FMA((a)[2], (b)[2], FMA((a)[1], (b)[1], (a)[0] * (b)[0]))

/* six times the signed volume of (p0,p1,p2,p3) with the contract's operation order */
static inline REAL FN(det6)(const REAL *p0, const REAL *p1, const REAL *p2, const REAL *p3) {
  REAL e1[3], e2[3], e3[3], G1[3];
  for (int k = 0; k < 3; k++) { e1[k] = p1[k] - p0[k]; e2[k] = p2[k] - p0[k]; e3[k] = p3[k] - p0[k]; }
  CROSS(G1, e2, e3);
  return DOT3(e1, G1);
}

static inline void FN(volume_one)(REAL *p0, REAL *p1, REAL *p2, REAL *p3, REAL R6, REAL a_v36) {
  REAL e1[3], e2[3], e3[3], G0[3], G1[3], G2[3], G3[3];
  for (int k = 0; k < 3; k++) { e1[k] = p1[k] - p0[k]; e2[k] = p2[k] - p0[k]; e3[k] = p3[k] - p0[k]; }
  CROSS(G1, e2, e3);
  CROSS(G2, e3, e1);
  CROSS(G3, e1, e2);
  for (int k = 0; k < 3; k++) G0[k] = -((G1[k] + G2[k]) + G3[k]);
  REAL det = DOT3(e1, G1);
  REAL n0 = DOT3(G0, G0), n1 = DOT3(G1, G1), n2 = DOT3(G2, G2), n3 = DOT3(G3, G3);
  REAL den = FMA(p3[3], n3, FMA(p2[3], n2, FMA(p1[3], n1, p0[3] * n0))) + a_v36;
  if (!(den >= ORC_RCP_MIN && den < ORC_RCP_MAX)) return;
  REAL s = -(det - R6) * ((REAL)1 / den);
  REAL s0 = s * p0[3], s1 = s * p1[3], s2 = s * p2[3], s3 = s * p3[3];
  for (int k = 0; k < 3; k++) {
    p0[k] = FMA(s0, G0[k], p0[k]);
    p1[k] = FMA(s1, G1[k], p1[k]);
    p2[k] = FMA(s2, G2[k], p2[k]);
    p3[k] = FMA(s3, G3[k], p3[k]);
  }
}

/*
 * COLLIDERS, after the ground plane, in list order [SPEC] (the mount has no collider code; the kinds are Unity's
 * SphereCollider / CapsuleCollider / BoxCollider).  xp = position at the start of the substep.
 *
 *   SPHERE(c, r):   d = x - c;  l2 = FMA(dz,dz, FMA(dy,dy, dx*dx));  hit iff 0 < l2 < r*r:
 *                   rinv = RCP(sqrt(l2));  q = r * rinv;  x = FMA(q, d, c);  n = d * rinv
 *   CAPSULE(A, B, r): ab = B - A, il2 = 1 / |ab|^2 (0 if A == B), both prepared in float;
 *                   t = dot(x - A, ab) * il2 clamped to [0, 1] (t > 0 ? t : 0, then t < 1 ? t : 1);
 *                   then SPHERE(FMA(t, ab, A), r)               dot(a,b) = FMA(a.z,b.z, FMA(a.y,b.y, a.x*b.x))
 *   BOX(c, half, axes R_k): d = x - c;  l_k = dot(R_k, d);  p_k = half_k - |l_k|;  hit iff all p_k > 0:
 *                   k = first axis of least p_k;  x = FMA(l_k >= 0 ? p_k : -p_k, R_k, x);  n = R_k
 *   friction f > 0 on a hit:  m = x - xp;  mn = -dot(m, n);  x = FMA(-f, FMA(mn, n, m), x)
 *
 * Box axes: column k of the rotation matrix of the quaternion (x,y,z,w), evaluated in double as written in
 * orc_prepare_collider and rounded to float once; a zero quaternion is the identity.
 */
#ifndef ORC_COLLIDER_DEFINED
#define ORC_COLLIDER_DEFINED
typedef struct {
  int32_t kind; /* 0 sphere, 1 capsule, 2 box: same layout as sb_collider */
  float friction;
  float p[10];
} orc_collider;
typedef struct {
  int kind;
  float friction;
  float a[4], b[4], c[4], d[4];
} orc_prepared;
static void orc_prepare_collider(const orc_collider *in, orc_prepared *o) {
  const float *p = in->p;
  memset(o, 0, sizeof *o);
  o->kind = in->kind;
  o->friction = in->friction;
  o->a[0] = p[0]; o->a[1] = p[1]; o->a[2] = p[2]; o->a[3] = p[3];
  if (in->kind == 1) {
    for (int k = 0; k < 3; k++) o->b[k] = p[4 + k] - p[k];
    float l2 = fmaf(o->b[2], o->b[2], fmaf(o->b[1], o->b[1], o->b[0] * o->b[0]));
    o->b[3] = l2 > 0 ? 1.0f / l2 : 0.0f;
  } else if (in->kind == 2) {
    double x = p[6], y = p[7], z = p[8], w = p[9];
    double n2 = ((x * x + y * y) + z * z) + w * w;
    double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (n2 > 0) {
      double s = 2.0 / n2;
      R[0][0] = 1.0 - s * (y * y + z * z); R[0][1] = s * (x * y - z * w); R[0][2] = s * (x * z + y * w);
      R[1][0] = s * (x * y + z * w); R[1][1] = 1.0 - s * (x * x + z * z); R[1][2] = s * (y * z - x * w);
      R[2][0] = s * (x * z - y * w); R[2][1] = s * (y * z + x * w); R[2][2] = 1.0 - s * (x * x + y * y);
    }
    float *ax[3] = {o->b, o->c, o->d};
    for (int k = 0; k < 3; k++) {
      for (int j = 0; j < 3; j++) ax[k][j] = (float)R[j][k];
      ax[k][3] = p[3 + k];
    }
    o->a[3] = 0;
  }
}
#endif

static inline int FN(sphere_one)(REAL *x, REAL cx, REAL cy, REAL cz, REAL r, REAL *n) {
  REAL dx = x[0] - cx, dy = x[1] - cy, dz = x[2] - cz;
  REAL l2 = FMA(dz, dz, FMA(dy, dy, dx * dx));
  if (!(l2 > 0 && l2 < r * r)) return 0;
  REAL rinv = (REAL)1 / SQRT(l2);
  REAL q = r * rinv;
  x[0] = FMA(q, dx, cx); x[1] = FMA(q, dy, cy); x[2] = FMA(q, dz, cz);
  n[0] = dx * rinv; n[1] = dy * rinv; n[2] = dz * rinv;
  return 1;
}

static inline void FN(finish_one)(REAL *x, const REAL *xp, REAL *v, const FN(step_consts) *c,
                                  int n_col, const orc_prepared *cols) {
  if (!(x[3] > 0)) return;
  if (c->use_ground && x[1] < c->ground_y) {
    x[1] = c->ground_y;
    x[0] = FMA(c->keep, x[0] - xp[0], xp[0]);
    x[2] = FMA(c->keep, x[2] - xp[2], xp[2]);
  }
  for (int s = 0; s < n_col; s++) {
    const orc_prepared *o = cols + s;
    REAL n[3] = {0, 0, 0};
    int hit;
    if (o->kind == 0) {
      hit = FN(sphere_one)(x, o->a[0], o->a[1], o->a[2], o->a[3], n);
    } else if (o->kind == 1) {
      REAL abx = o->b[0], aby = o->b[1], abz = o->b[2];
      REAL t = FMA(x[2] - (REAL)o->a[2], abz, FMA(x[1] - (REAL)o->a[1], aby, (x[0] - (REAL)o->a[0]) * abx)) * (REAL)o->b[3];
      t = t > 0 ? t : 0;
      t = t < 1 ? t : 1;
      hit = FN(sphere_one)(x, FMA(t, abx, (REAL)o->a[0]), FMA(t, aby, (REAL)o->a[1]), FMA(t, abz, (REAL)o->a[2]), o->a[3], n);
    } else {
      const float *R[3] = {o->b, o->c, o->d};
      REAL dx = x[0] - (REAL)o->a[0], dy = x[1] - (REAL)o->a[1], dz = x[2] - (REAL)o->a[2];
      REAL l[3], pk[3];
      for (int k = 0; k < 3; k++) {
        l[k] = FMA((REAL)R[k][2], dz, FMA((REAL)R[k][1], dy, (REAL)R[k][0] * dx));
        pk[k] = (REAL)R[k][3] - (l[k] < 0 ? -l[k] : l[k]);
      }
      hit = pk[0] > 0 && pk[1] > 0 && pk[2] > 0;
      if (hit) {
        int km = 0;
        if (pk[1] < pk[km]) km = 1;
        if (pk[2] < pk[km]) km = 2;
        REAL dl = l[km] >= 0 ? pk[km] : -pk[km];
        for (int k = 0; k < 3; k++) { n[k] = R[km][k]; x[k] = FMA(dl, n[k], x[k]); }
      }
    }
    if (hit && o->friction > 0) {
      REAL m[3] = {x[0] - xp[0], x[1] - xp[1], x[2] - xp[2]};
      REAL mn = -FMA(m[2], n[2], FMA(m[1], n[1], m[0] * n[0]));
      REAL f = -(REAL)o->friction;
      for (int k = 0; k < 3; k++) x[k] = FMA(f, FMA(mn, n[k], m[k]), x[k]);
    }
  }
  v[0] = ((x[0] - xp[0]) * c->inv_h) * c->damp;
  v[1] = ((x[1] - xp[1]) * c->inv_h) * c->damp;
  v[2] = ((x[2] - xp[2]) * c->inv_h) * c->damp;
}

static inline void FN(apply_entry)(REAL *x4, int32_t ent, const int32_t *edges, const REAL *rest_len,
                                   const int32_t *tets, const REAL *rest_vol6,
                                   const FN(step_consts) *c) {
  int32_t id = ent & 0x7fffffff;
  if (ent >= 0) {
    if (c->use_d)
      FN(distance_one)(x4 + 4 * (size_t)edges[2 * (size_t)id], x4 + 4 * (size_t)edges[2 * (size_t)id + 1],
                       rest_len[id], c->a_d);
  } else {
    if (c->use_v) {
      const int32_t *t = tets + 4 * (size_t)id;
      FN(volume_one)(x4 + 4 * (size_t)t[0], x4 + 4 * (size_t)t[1], x4 + 4 * (size_t)t[2],
                     x4 + 4 * (size_t)t[3], rest_vol6[id], c->a_v36);
    }
  }
}

/*
 * Advance `n_frames` frames.  `order` is the Gauss-Seidel processing order for
 * ONE iteration (entry >= 0: edge id; entry < 0: tet id in the low 31 bits).
 * `batch_off` (n_batches+1 offsets into order) marks runs of vertex-disjoint
 * constraints; with threads > 1 each run is an OpenMP parallel-for (the result
 * does not depend on the thread count because runs are independent sets).
 * Pass n_batches = 0 for plain sequential order.
 *
 * orc_simulate2: iterations 0, 2, 4 ... of every substep follow `order`, iterations
 * 1, 3, 5 ... follow `order_odd` (NULL: `order` again).  The product sweeps its tile
 * passes forwards and backwards alternately (a symmetric Gauss-Seidel sweep), so that the
 * last pass of one iteration and the first of the next work on the same tiles.
 */
static void FN(run_order)(REAL *x4, int64_t n_order, const int32_t *order, int32_t n_batches, const int64_t *batch_off,
                          int par, int threads, const int32_t *edges, const REAL *rest_len, const int32_t *tets,
                          const REAL *rest_vol6, const FN(step_consts) *c) {
  (void)threads;
  if (!par) {
    for (int64_t k = 0; k < n_order; k++) FN(apply_entry)(x4, order[k], edges, rest_len, tets, rest_vol6, c);
  } else {
    for (int32_t b = 0; b < n_batches; b++) {
      int64_t lo = batch_off[b], hi = batch_off[b + 1];
#pragma omp parallel for schedule(static) num_threads(threads) if (hi - lo > 2048)
      for (int64_t k = lo; k < hi; k++) FN(apply_entry)(x4, order[k], edges, rest_len, tets, rest_vol6, c);
    }
  }
}

int FN(orc_simulate2)(int32_t V, REAL *x4, REAL *v4, int32_t E, const int32_t *edges,
                      const REAL *rest_len, int32_t T, const int32_t *tets, const REAL *rest_vol6,
                      const orc_params *p, int64_t n_order, const int32_t *order, int32_t n_batches,
                      const int64_t *batch_off, int64_t n_order_odd, const int32_t *order_odd, int32_t n_batches_odd,
                      const int64_t *batch_off_odd, int32_t n_col, const orc_collider *colliders,
                      int32_t n_frames, int32_t threads) {
  if (V < 0 || p->substeps <= 0 || p->iterations < 0 || !(p->dt > 0) || n_col < 0 || n_col > 16) return -1;
  orc_prepared cols[16];
  for (int s = 0; s < n_col; s++) {
    if (colliders[s].kind < 0 || colliders[s].kind > 2) return -1;
    orc_prepare_collider(colliders + s, cols + s);
  }
  if (!order_odd) { order_odd = order; n_order_odd = n_order; n_batches_odd = n_batches; batch_off_odd = batch_off; }
  for (int64_t k = 0; k < n_order; k++) {
    int32_t id = order[k] & 0x7fffffff;
    if (order[k] >= 0 ? id >= E : id >= T) return -2;
  }
  for (int64_t k = 0; k < n_order_odd; k++) {
    int32_t id = order_odd[k] & 0x7fffffff;
    if (order_odd[k] >= 0 ? id >= E : id >= T) return -2;
  }
  FN(step_consts) c = FN(make_consts)(p);
  REAL *xp = (REAL *)malloc(sizeof(REAL) * 3 * (size_t)(V > 0 ? V : 1));
  if (!xp) return -3;
  const int par = threads > 1 && n_batches > 0, par_odd = threads > 1 && n_batches_odd > 0;
  for (int f = 0; f < n_frames; f++) {
    for (int s = 0; s < p->substeps; s++) {
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
      for (int32_t i = 0; i < V; i++) FN(predict_one)(x4 + 4 * (size_t)i, xp + 3 * (size_t)i, v4 + 4 * (size_t)i, &c);
      for (int it = 0; it < p->iterations; it++) {
        if (it & 1) FN(run_order)(x4, n_order_odd, order_odd, n_batches_odd, batch_off_odd, par_odd, threads, edges, rest_len, tets, rest_vol6, &c);
        else FN(run_order)(x4, n_order, order, n_batches, batch_off, par, threads, edges, rest_len, tets, rest_vol6, &c);
      }
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
      for (int32_t i = 0; i < V; i++)
        FN(finish_one)(x4 + 4 * (size_t)i, xp + 3 * (size_t)i, v4 + 4 * (size_t)i, &c, n_col, cols);
    }
  }
  free(xp);
  return 0;
}

int FN(orc_simulate)(int32_t V, REAL *x4, REAL *v4, int32_t E, const int32_t *edges,
                     const REAL *rest_len, int32_t T, const int32_t *tets, const REAL *rest_vol6,
                     const orc_params *p, int64_t n_order, const int32_t *order, int32_t n_batches,
                     const int64_t *batch_off, int32_t n_col, const orc_collider *colliders,
                     int32_t n_frames, int32_t threads) {
  return FN(orc_simulate2)(V, x4, v4, E, edges, rest_len, T, tets, rest_vol6, p, n_order, order, n_batches, batch_off, 0, NULL,
                           0, NULL, n_col, colliders, n_frames, threads);
}

/*
 * Derived rest data.  rest_len[e] = |x_a - x_b| and rest_vol6[t] = det6 with the
 * contract's operation order, evaluated on the rest positions.
 */
void FN(orc_rest_values)(const REAL *x4, int32_t E, const int32_t *edges, REAL *rest_len, int32_t T,
                         const int32_t *tets, REAL *rest_vol6) {
  for (int32_t e = 0; e < E; e++) {
    const REAL *a = x4 + 4 * (size_t)edges[2 * (size_t)e], *b = x4 + 4 * (size_t)edges[2 * (size_t)e + 1];
    REAL dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    rest_len[e] = SQRT(FMA(dz, dz, FMA(dy, dy, dx * dx)));
  }
  for (int32_t t = 0; t < T; t++) {
    const int32_t *q = tets + 4 * (size_t)t;
    rest_vol6[t] = FN(det6)(x4 + 4 * (size_t)q[0], x4 + 4 * (size_t)q[1], x4 + 4 * (size_t)q[2], x4 + 4 * (size_t)q[3]);
  }
}

/*
 * Area-weighted vertex normals over the surface triangles: for vertex i the sum,
 * in ascending triangle index, of (p1-p0) x (p2-p0) over the triangles that
 * contain i, then n = sum * (1/sqrt(|sum|^2)) (zero if the sum is zero).
 * `normals` is float[3*V]-shaped REAL; vertices on no triangle get (0,0,0).
 */
void FN(orc_normals)(int32_t V, const REAL *x4, int32_t F, const int32_t *tris, REAL *normals) {
  for (size_t i = 0; i < 3 * (size_t)V; i++) normals[i] = 0;
  for (int32_t f = 0; f < F; f++) {
    const int32_t *t = tris + 3 * (size_t)f;
    const REAL *p0 = x4 + 4 * (size_t)t[0], *p1 = x4 + 4 * (size_t)t[1], *p2 = x4 + 4 * (size_t)t[2];
    REAL a[3], b[3], n[3];
    for (int k = 0; k < 3; k++) { a[k] = p1[k] - p0[k]; b[k] = p2[k] - p0[k]; }
    CROSS(n, a, b);
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < 3; k++) normals[3 * (size_t)t[j] + k] += n[k];
  }
  for (int32_t i = 0; i < V; i++) {
    REAL *n = normals + 3 * (size_t)i;
    REAL l2 = DOT3(n, n);
    if (l2 > 0) {
      REAL q = (REAL)1 / SQRT(l2);
      n[0] *= q; n[1] *= q; n[2] *= q;
    }
  }
}

/*
 * Render mesh bound to the tets: out4[i] = sum_k w_k * x[tets[4*tet_of[i] + k]] with the four barycentric weights
 * bary4[4i..] of render vertex i, as FMA(w3,x3, FMA(w2,x2, FMA(w1,x1, w0*x0))) per component; out4 is xyz0-strided
 * so that orc_normals can run over it with the render triangles.
 */
void FN(orc_skin)(const REAL *x4, const int32_t *tets, int32_t n, const int32_t *tet_of, const float *bary4, REAL *out4) {
  for (int32_t i = 0; i < n; i++) {
    const int32_t *q = tets + 4 * (size_t)tet_of[i];
    const float *w = bary4 + 4 * (size_t)i;
    for (int k = 0; k < 3; k++)
      out4[4 * (size_t)i + k] = FMA((REAL)w[3], x4[4 * (size_t)q[3] + k],
                                    FMA((REAL)w[2], x4[4 * (size_t)q[2] + k],
                                        FMA((REAL)w[1], x4[4 * (size_t)q[1] + k], (REAL)w[0] * x4[4 * (size_t)q[0] + k])));
    out4[4 * (size_t)i + 3] = 0;
  }
}

#undef CROSS
#undef DOT3
