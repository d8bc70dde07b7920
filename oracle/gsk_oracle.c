/*
 * gsk_oracle.c — CPU restatement of GeoStatsSolvers.jl's Kriging estimation path (and of the IDW / LWR
 * per-location bodies that share its searcher: src/estimation/idw.jl:112-142, src/estimation/lwr.jl:113-146).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. libgskrige.so never links
 * or calls it; the product has no CPU path.
 *
 * PARITY STATUS: *partially pinned*. The control flow restated here is the reference's
 * (src/estimation/krig.jl:166-234, src/ui.jl:11-50) and is checked against every
 * known-answer test the reference holds for this path (test/estimation/krig.jl:35-37,
 * 50-52,70-72; test/ui.jl:7-37) in tests/test_oracle_reference_checks.py. The arithmetic
 * itself lives in un-vendored Julia dependencies that are absent from /root/reference and
 * cannot be run here (no julia binary): GeoStatsModels 0.2, Variography 0.22, Meshes 0.37
 * (+NearestNeighbors), per ref Project.toml:25-43. Their published algorithms are restated
 * below; each [3P] behaviour is a run-time switch in gsk_problem (support offsets, Gaussian
 * nugget epsilon, variance clamp, sqrt round trip). Beyond the reference's nine atol=1e-3
 * checks, numerical parity with the Julia stack is UNPINNED.
 * A second, independent restatement (oracle/numpy_twin.py, LAPACK dsytrf/dpotrf as Julia's
 * bunchkaufman/cholesky would call) cross-checks this file.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off; no FMA contraction so that
 * squared distances are bit-identical to the CUDA search kernel's __dmul_rn/__dadd_rn chain).
 */
#include "gsk_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * variogram models — Variography 0.22 [3P], SURVEY §8a a14
 *   Gaussian     γ(h) = (s−n')(1 − exp(−3 (h/r)²)) + (h>0) n',  n' = n + eps (eps = 1e-6)
 *   Spherical    γ(h) = (s−n)(1.5 (h/r) − 0.5 (h/r)³) for h<r, (s−n) otherwise, + (h>0) n
 *   Exponential  γ(h) = (s−n)(1 − exp(−3 h/r)) + (h>0) n
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int kind;
  double range, sill, nugget; /* nugget already includes the Gaussian epsilon */
} vario_t;

static vario_t vario_from(const gsk_problem *p) {
  vario_t v;
  v.kind = p->vario_kind;
  v.range = p->vario_range;
  v.sill = p->vario_sill;
  v.nugget = p->vario_nugget + (p->vario_kind == GSK_VARIO_GAUSSIAN ? p->gaussian_nugget_eps : 0.0);
  return v;
}

static double vario_eval(const vario_t *v, double h) {
  double s = v->sill, n = v->nugget, r = v->range;
  double g;
  switch (v->kind) {
  case GSK_VARIO_GAUSSIAN: {
    double t = h / r;
    g = (s - n) * (1.0 - exp(-3.0 * (t * t)));
    break;
  }
  case GSK_VARIO_SPHERICAL: {
    double t = h / r;
    g = (h < r) ? (s - n) * (1.5 * t - 0.5 * (t * t * t)) : (s - n);
    break;
  }
  default: { /* exponential */
    g = (s - n) * (1.0 - exp(-3.0 * (h / r)));
    break;
  }
  }
  return g + ((h > 0.0) ? n : 0.0);
}

/* Squared Euclidean distance, summed left to right without FMA. The CUDA search kernel
 * evaluates exactly this chain so that neighbour sets are bit-identical (north_star). */
static inline double dist2(int dim, const double *a, const double *b) {
  double d0 = a[0] - b[0];
  double s = d0 * d0;
  if (dim > 1) { double d1 = a[1] - b[1]; s = s + d1 * d1; }
  if (dim > 2) { double d2 = a[2] - b[2]; s = s + d2 * d2; }
  return s;
}

/* ------------------------------------------------------------------------------------------
 * targets — Meshes `centroid(grid, ind)` [3P], SURVEY V9: origin + (ijk − ½)·spacing with a
 * column-major (x fastest) linear index (pinned by ref test/estimation/krig.jl:34-37,69-72).
 * Stated here with a 0-based index: origin + (i + 0.5)·spacing.
 * ---------------------------------------------------------------------------------------- */
int64_t gsk_oracle_num_targets(const gsk_problem *p) {
  if (p->grid_dims[0] > 0) {
    int64_t t = 1;
    for (int d = 0; d < p->dim; d++) t *= p->grid_dims[d];
    return t;
  }
  return p->n_points;
}

static void target_center(const gsk_problem *p, int64_t lin, double *c) {
  c[0] = c[1] = c[2] = 0.0;
  if (p->target_order) lin = p->target_order[lin]; /* traverse(pdomain, path): the j-th visited target (krig.jl:179,204) */
  if (p->grid_dims[0] > 0) {
    int64_t rem = lin;
    for (int d = 0; d < p->dim; d++) {
      int64_t i = rem % p->grid_dims[d];
      rem /= p->grid_dims[d];
      c[d] = p->grid_origin[d] + ((double)i + 0.5) * p->grid_spacing[d];
    }
  } else {
    for (int d = 0; d < p->dim; d++) c[d] = p->point_coords[d][lin];
  }
}

/* ------------------------------------------------------------------------------------------
 * neighbour search — Meshes KNearestSearch / KBallSearch `search!` (ref krig.jl:210) [3P]:
 * the k nearest samples sorted ascending by distance; ball variant keeps dist <= radius.
 * Ties are ordered by sample index (north_star).
 * ---------------------------------------------------------------------------------------- */
typedef struct { double d2; int32_t idx; } cand_t;

static inline int cand_less(double d2a, int32_t ia, double d2b, int32_t ib) {
  return d2a < d2b || (d2a == d2b && ia < ib);
}

/* insert into an ascending list of at most k */
static inline void topk_insert(cand_t *best, int *cnt, int k, double d2, int32_t idx) {
  int n = *cnt;
  if (n == k && !cand_less(d2, idx, best[n - 1].d2, best[n - 1].idx)) return;
  int pos = (n < k) ? n : k - 1;
  while (pos > 0 && cand_less(d2, idx, best[pos - 1].d2, best[pos - 1].idx)) {
    best[pos] = best[pos - 1];
    pos--;
  }
  best[pos].d2 = d2;
  best[pos].idx = idx;
  if (n < k) *cnt = n + 1;
}

static int knn_brute(const gsk_problem *p, const double *xyz /* n×3 AoS */, const double *c, int k, cand_t *best) {
  int cnt = 0;
  for (int64_t i = 0; i < p->n_samples; i++) topk_insert(best, &cnt, k, dist2(p->dim, c, xyz + 3 * i), (int32_t)i);
  return cnt;
}

/* KD-tree: median split on the widest axis, leaves of <= LEAF points. Stands in for
 * NearestNeighbors.jl's KDTree [3P]; result is identical to brute force by construction
 * (subtrees are pruned only when strictly farther than the current k-th candidate). */
#define KD_LEAF 12
typedef struct {
  int32_t lo, hi;  /* range in perm */
  int32_t left, right; /* children, −1 for leaf */
  int32_t axis;
  double split;
} kdnode_t;
typedef struct {
  kdnode_t *nodes;
  int32_t nnodes;
  int32_t *perm;
  const double *xyz;
  int dim;
} kdtree_t;

static const double *g_sort_xyz;
static int g_sort_axis;
static int cmp_axis(const void *a, const void *b) {
  double xa = g_sort_xyz[3 * (int64_t)(*(const int32_t *)a) + g_sort_axis];
  double xb = g_sort_xyz[3 * (int64_t)(*(const int32_t *)b) + g_sort_axis];
  return (xa > xb) - (xa < xb);
}

static int32_t kd_build(kdtree_t *t, int32_t lo, int32_t hi) {
  int32_t id = t->nnodes++;
  kdnode_t *nd = &t->nodes[id];
  nd->lo = lo; nd->hi = hi; nd->left = nd->right = -1; nd->axis = 0; nd->split = 0;
  if (hi - lo <= KD_LEAF) return id;
  double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int32_t i = lo; i < hi; i++)
    for (int d = 0; d < t->dim; d++) {
      double v = t->xyz[3 * (int64_t)t->perm[i] + d];
      if (v < mn[d]) mn[d] = v;
      if (v > mx[d]) mx[d] = v;
    }
  int ax = 0;
  for (int d = 1; d < t->dim; d++) if (mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
  if (!(mx[ax] > mn[ax])) return id; /* all points coincide: keep as a (large) leaf */
  g_sort_xyz = t->xyz; g_sort_axis = ax;
  qsort(t->perm + lo, (size_t)(hi - lo), sizeof(int32_t), cmp_axis);
  int32_t mid = lo + (hi - lo) / 2;
  double split = t->xyz[3 * (int64_t)t->perm[mid] + ax];
  /* move mid so that everything left is strictly < split where possible */
  while (mid > lo && t->xyz[3 * (int64_t)t->perm[mid - 1] + ax] == split) mid--;
  if (mid == lo) { /* fall back: first index with value > split */
    mid = lo + (hi - lo) / 2;
    while (mid < hi && t->xyz[3 * (int64_t)t->perm[mid] + ax] == split) mid++;
    if (mid == hi) return id;
    split = t->xyz[3 * (int64_t)t->perm[mid] + ax];
  }
  int32_t l = kd_build(t, lo, mid);
  int32_t r = kd_build(t, mid, hi);
  nd = &t->nodes[id]; /* nodes array is preallocated, pointer stays valid */
  nd->axis = ax; nd->split = split; nd->left = l; nd->right = r;
  return id;
}

static kdtree_t *kd_create(const double *xyz, int64_t n, int dim) {
  kdtree_t *t = (kdtree_t *)calloc(1, sizeof(kdtree_t));
  t->xyz = xyz; t->dim = dim;
  t->perm = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; i++) t->perm[i] = (int32_t)i;
  t->nodes = (kdnode_t *)malloc(sizeof(kdnode_t) * (size_t)(2 * n + 2));
  t->nnodes = 0;
  kd_build(t, 0, (int32_t)n);
  return t;
}
static void kd_free(kdtree_t *t) { if (t) { free(t->perm); free(t->nodes); free(t); } }

static void kd_query(const kdtree_t *t, int32_t id, const double *c, int k, cand_t *best, int *cnt) {
  const kdnode_t *nd = &t->nodes[id];
  if (nd->left < 0) {
    for (int32_t i = nd->lo; i < nd->hi; i++) {
      int32_t s = t->perm[i];
      topk_insert(best, cnt, k, dist2(t->dim, c, t->xyz + 3 * (int64_t)s), s);
    }
    return;
  }
  double diff = c[nd->axis] - nd->split;
  int32_t near = diff < 0 ? nd->left : nd->right;
  int32_t far = diff < 0 ? nd->right : nd->left;
  kd_query(t, near, c, k, best, cnt);
  /* (c−split)² is a lower bound of the computed d² of every point on the far side
   * (rounded subtraction and addition are monotone); prune only when strictly farther */
  double pd = diff * diff;
  if (*cnt < k || !(pd > best[*cnt - 1].d2)) kd_query(t, far, c, k, best, cnt);
}

/* ------------------------------------------------------------------------------------------
 * Universal Kriging monomials — GeoStatsModels UKexps [3P], SURVEY V4: all exponent vectors
 * of total degree 0..degree (per degree in lexicographically descending order, as
 * Combinatorics.multiexponents yields them), stably sorted by descending max exponent.
 * degree 1 → [x, y, (z), 1];  degree 2 in 2-D → [x², y², x, y, xy, 1].
 * ---------------------------------------------------------------------------------------- */
int gsk_oracle_uk_exponents(int degree, int dim, int32_t *out, int cap) {
  if (degree < 0 || degree > 2 || dim < 1 || dim > 3) return -1;
  int32_t tmp[16][3];
  int n = 0;
  for (int d = 0; d <= degree; d++) {
    /* multiexponents(dim, d): descending lexicographic */
    for (int a = d; a >= 0; a--) {
      if (dim == 1) { if (a == d) { tmp[n][0] = a; tmp[n][1] = tmp[n][2] = 0; n++; } continue; }
      for (int b = d - a; b >= 0; b--) {
        int c = d - a - b;
        if (dim == 2) { if (c == 0) { tmp[n][0] = a; tmp[n][1] = b; tmp[n][2] = 0; n++; } continue; }
        tmp[n][0] = a; tmp[n][1] = b; tmp[n][2] = c; n++;
      }
    }
  }
  if (n > cap) return -1;
  int k = 0;
  for (int mx = degree; mx >= 0; mx--)
    for (int i = 0; i < n; i++) {
      int m = tmp[i][0];
      if (tmp[i][1] > m) m = tmp[i][1];
      if (tmp[i][2] > m) m = tmp[i][2];
      if (m == mx) { for (int d = 0; d < dim; d++) out[k * dim + d] = tmp[i][d]; k++; }
    }
  return n;
}

static double monomial(int dim, const int32_t *e, const double *x) {
  double v = 1.0;
  for (int d = 0; d < dim; d++)
    for (int q = 0; q < e[d]; q++) v *= x[d];
  return v;
}

/* ------------------------------------------------------------------------------------------
 * dense factorisations (column-major, leading dimension m)
 *   lu_factor / lu_solve     : partial-pivot LU. Stands in for Julia's bunchkaufman(Symmetric(LHS),
 *                              check=false) on the OK/UK systems [3P, V3] — same solution up to rounding
 *                              (numpy_twin.py does call LAPACK dsytrf and pins the difference).
 *   chol_factor / chol_solve : Cholesky for Simple Kriging (Julia: cholesky(Symmetric(LHS), check=false)).
 * A zero pivot is not an error (check=false): the division produces Inf/NaN that flow to the output.
 * ---------------------------------------------------------------------------------------- */
static void lu_factor(int m, double *A, int32_t *piv) {
  for (int j = 0; j < m; j++) {
    int p = j;
    double best = fabs(A[j + (size_t)j * m]);
    for (int i = j + 1; i < m; i++) {
      double v = fabs(A[i + (size_t)j * m]);
      if (v > best) { best = v; p = i; }
    }
    piv[j] = p;
    if (p != j)
      for (int c = 0; c < m; c++) {
        double t = A[j + (size_t)c * m];
        A[j + (size_t)c * m] = A[p + (size_t)c * m];
        A[p + (size_t)c * m] = t;
      }
    double d = A[j + (size_t)j * m];
    for (int i = j + 1; i < m; i++) A[i + (size_t)j * m] /= d;
    for (int c = j + 1; c < m; c++) {
      double u = A[j + (size_t)c * m];
      if (u != 0.0)
        for (int i = j + 1; i < m; i++) A[i + (size_t)c * m] -= A[i + (size_t)j * m] * u;
    }
  }
}
static void lu_solve(int m, const double *A, const int32_t *piv, double *b) {
  for (int j = 0; j < m; j++) {
    int p = piv[j];
    if (p != j) { double t = b[j]; b[j] = b[p]; b[p] = t; }
  }
  for (int j = 0; j < m; j++) {
    double v = b[j];
    if (v != 0.0)
      for (int i = j + 1; i < m; i++) b[i] -= A[i + (size_t)j * m] * v;
  }
  for (int j = m - 1; j >= 0; j--) {
    b[j] /= A[j + (size_t)j * m];
    double v = b[j];
    for (int i = 0; i < j; i++) b[i] -= A[i + (size_t)j * m] * v;
  }
}
static void chol_factor(int m, double *A) { /* lower */
  for (int j = 0; j < m; j++) {
    double d = A[j + (size_t)j * m];
    for (int p = 0; p < j; p++) d -= A[j + (size_t)p * m] * A[j + (size_t)p * m];
    d = sqrt(d);
    A[j + (size_t)j * m] = d;
    for (int i = j + 1; i < m; i++) {
      double s = A[i + (size_t)j * m];
      for (int p = 0; p < j; p++) s -= A[i + (size_t)p * m] * A[j + (size_t)p * m];
      A[i + (size_t)j * m] = s / d;
    }
  }
}
static void chol_solve(int m, const double *A, double *b) {
  for (int j = 0; j < m; j++) {
    double s = b[j];
    for (int p = 0; p < j; p++) s -= A[j + (size_t)p * m] * b[p];
    b[j] = s / A[j + (size_t)j * m];
  }
  for (int j = m - 1; j >= 0; j--) {
    double s = b[j];
    for (int p = j + 1; p < m; p++) s -= A[p + (size_t)j * m] * b[p];
    b[j] = s / A[j + (size_t)j * m];
  }
}

/* ------------------------------------------------------------------------------------------
 * GeoStatsModels.fit [3P] (ref krig.jl:176,223), SURVEY §8a a13: LHS of the kriging system for
 * the k samples `nb` in covariance form (sill − γ), plus constraint blocks, then factorise.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int k, c, m;
  double *A;     /* m×m factor */
  int32_t *piv;  /* LU pivots (OK/UK) */
} fitted_t;

static int n_constraints(const gsk_problem *p, int32_t *exps) {
  if (p->estimator == GSK_EST_SIMPLE) return 0;
  if (p->estimator == GSK_EST_ORDINARY) return 1;
  return gsk_oracle_uk_exponents(p->uk_degree, p->dim, exps, GSK_MAX_DRIFT_TERMS);
}

static void fit_system(const gsk_problem *p, const vario_t *v, const double *xyz, const int32_t *nb, int k,
                       int c, const int32_t *exps, fitted_t *f) {
  int m = k + c;
  f->k = k; f->c = c; f->m = m;
  double *A = f->A;
  /* Variography.pairwise!: Γij = γ(‖xi − xj‖), diagonal γ(0) = 0; stationary γ → C = sill − Γ */
  for (int j = 0; j < k; j++) {
    A[j + (size_t)j * m] = v->sill - vario_eval(v, 0.0);
    for (int i = j + 1; i < k; i++) {
      double h = sqrt(dist2(p->dim, xyz + 3 * (int64_t)nb[i], xyz + 3 * (int64_t)nb[j]));
      double cij = v->sill - vario_eval(v, h);
      A[i + (size_t)j * m] = cij;
      A[j + (size_t)i * m] = cij;
    }
  }
  if (p->estimator == GSK_EST_ORDINARY) {
    for (int i = 0; i < k; i++) { A[k + (size_t)i * m] = 1.0; A[i + (size_t)k * m] = 1.0; }
    A[k + (size_t)k * m] = 0.0;
  } else if (p->estimator == GSK_EST_UNIVERSAL) {
    for (int i = 0; i < k; i++)
      for (int t = 0; t < c; t++) {
        double f_it = monomial(p->dim, exps + t * p->dim, xyz + 3 * (int64_t)nb[i]);
        A[(k + t) + (size_t)i * m] = f_it;
        A[i + (size_t)(k + t) * m] = f_it;
      }
    for (int a = 0; a < c; a++)
      for (int b = 0; b < c; b++) A[(k + a) + (size_t)(k + b) * m] = 0.0;
  }
  if (p->estimator == GSK_EST_SIMPLE) chol_factor(m, A);
  else lu_factor(m, A, f->piv);
}

/* GeoStatsModels.predictprob [3P] (ref krig.jl:180,226), SURVEY §8a a15-a16 */
static void predict(const gsk_problem *p, const vario_t *v, const double *xyz, const int32_t *nb,
                    const fitted_t *f, const int32_t *exps, const double *center, double *rhs, double *sol,
                    double *mean_out, double *var_out) {
  int k = f->k, c = f->c, m = f->m;
  int q = p->n_support;
  for (int j = 0; j < k; j++) {
    /* γ(U, xj) with U the target geometry: arithmetic mean over its sub-sample points */
    double acc = 0.0;
    for (int s = 0; s < q; s++) {
      double u[3] = {center[0], center[1], center[2]};
      for (int d = 0; d < p->dim; d++) u[d] = center[d] + (p->support_offsets[d] ? p->support_offsets[d][s] : 0.0);
      double h = sqrt(dist2(p->dim, u, xyz + 3 * (int64_t)nb[j]));
      acc += vario_eval(v, h);
    }
    rhs[j] = v->sill - acc / (double)q;
  }
  if (p->estimator == GSK_EST_ORDINARY) rhs[k] = 1.0;
  else if (p->estimator == GSK_EST_UNIVERSAL)
    for (int t = 0; t < c; t++) rhs[k + t] = monomial(p->dim, exps + t * p->dim, center);
  memcpy(sol, rhs, sizeof(double) * (size_t)m);
  if (p->estimator == GSK_EST_SIMPLE) chol_solve(m, f->A, sol);
  else lu_solve(m, f->A, f->piv, sol);
  /* mean */
  double mu = 0.0;
  if (p->estimator == GSK_EST_SIMPLE) {
    for (int i = 0; i < k; i++) mu += sol[i] * (p->values[nb[i]] - p->sk_mean);
    mu = p->sk_mean + mu;
  } else {
    for (int i = 0; i < k; i++) mu += sol[i] * p->values[nb[i]];
  }
  /* variance: sill − b·[λ;ν], clamped, then the Normal(μ,√σ²) → var() round trip */
  double c1 = 0.0, c2 = 0.0;
  for (int i = 0; i < k; i++) c1 += rhs[i] * sol[i];
  for (int i = k; i < m; i++) c2 += rhs[i] * sol[i];
  double s2 = v->sill - (c1 + c2);
  if (p->flags & GSK_FLAG_CLAMP_VARIANCE) s2 = (s2 > 0.0 || s2 != s2) ? s2 : 0.0;
  if (p->flags & GSK_FLAG_SQRT_ROUNDTRIP) { double sd = sqrt(s2); s2 = sd * sd; }
  *mean_out = mu;
  *var_out = s2;
}

/* ------------------------------------------------------------------------------------------
 * IDW (ref: src/estimation/idw.jl:118-140) and LWR (ref: src/estimation/lwr.jl:119-145) bodies for one target, given
 * the neighbours `nb` (sorted ascending by distance, as searchdists! returns them) and their squared distances.
 * ---------------------------------------------------------------------------------------- */
static double idw_pow(double d, double e) { /* Julia: ds .^ exponent — exact products for the small integer exponents */
  if (e == 1.0) return d;
  if (e == 2.0) return d * d;
  if (e == 3.0) return d * d * d;
  return pow(d, e);
}

static void idw_predict(const gsk_problem *p, const int32_t *nb, const double *d2, int nn, double *mu, double *sig) {
  double sw = 0.0;
  for (int i = 0; i < nn; i++) sw += 1.0 / idw_pow(sqrt(d2[i]), p->idw_exponent); /* ws = 1 ./ ds .^ exponent; Σw = sum(ws) */
  if (isinf(sw)) { /* some distance is zero: j = findfirst(iszero, ds) */
    for (int i = 0; i < nn; i++)
      if (d2[i] == 0.0) { *mu = p->values[nb[i]]; break; }
    *sig = 0.0;
    return;
  }
  double acc = 0.0;
  for (int i = 0; i < nn; i++) acc += ((1.0 / idw_pow(sqrt(d2[i]), p->idw_exponent)) / sw) * p->values[nb[i]];
  *mu = acc;
  double dmin = sqrt(d2[0]);
  for (int i = 1; i < nn; i++) dmin = fmin(dmin, sqrt(d2[i]));
  *sig = dmin; /* σ = minimum(ds) */
}

/* m×m (m <= 4) LU with partial pivoting, two right-hand sides — Julia's `A \ b` for a square matrix */
static int lu_solve2(int m, double A[4][4], double *b, double *b2) {
  for (int c = 0; c < m; c++) {
    int piv = c;
    double best = fabs(A[c][c]);
    for (int r = c + 1; r < m; r++)
      if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
    if (!(best > 0.0)) return 0;
    if (piv != c) {
      for (int j = 0; j < m; j++) { double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
      double t = b[c]; b[c] = b[piv]; b[piv] = t;
      t = b2[c]; b2[c] = b2[piv]; b2[piv] = t;
    }
    for (int r = c + 1; r < m; r++) {
      double f = A[r][c] / A[c][c];
      for (int j = c + 1; j < m; j++) A[r][j] -= f * A[c][j];
      b[r] -= f * b[c];
      b2[r] -= f * b2[c];
    }
  }
  for (int r = m - 1; r >= 0; r--) {
    double s = b[r], s2 = b2[r];
    for (int j = r + 1; j < m; j++) { s -= A[r][j] * b[j]; s2 -= A[r][j] * b2[j]; }
    b[r] = s / A[r][r];
    b2[r] = s2 / A[r][r];
  }
  return 1;
}

static void lwr_predict(const gsk_problem *p, const double *xyz, const int32_t *nb, const double *d2, int nn,
                        const double *ctr, double *mu, double *sig) {
  int dim = p->dim, m = dim + 1;
  double dmax = 0.0;
  for (int i = 0; i < nn; i++) dmax = fmax(dmax, sqrt(d2[i])); /* δs = ds ./ maximum(ds) */
  double A[4][4] = {{0}}, bz[4] = {0, 0, 0, 0}, x0[4] = {1.0, ctr[0], ctr[1], ctr[2]};
  for (int i = 0; i < nn; i++) {
    double h = sqrt(d2[i]) / dmax;
    double w = exp(-3.0 * (h * h)); /* weightfun default, lwr.jl:58 */
    const double *q = xyz + 3 * (int64_t)nb[i];
    double x[4] = {1.0, q[0], q[1], q[2]};
    for (int a = 0; a < m; a++) {
      double wx = x[a] * w;
      for (int b = 0; b < m; b++) A[a][b] += wx * x[b]; /* Xₗ' Wₗ Xₗ */
      bz[a] += wx * p->values[nb[i]];                   /* Xₗ' Wₗ zₗ */
    }
  }
  if (!lu_solve2(m, A, bz, x0)) { *mu = NAN; *sig = NAN; return; } /* the reference throws SingularException */
  double c0[4] = {1.0, ctr[0], ctr[1], ctr[2]};
  double zhat = 0.0;
  for (int a = 0; a < m; a++) zhat += bz[a] * c0[a]; /* ẑₒ = θₗ ⋅ xₒ */
  double rr = 0.0;
  for (int i = 0; i < nn; i++) {
    double h = sqrt(d2[i]) / dmax;
    double w = exp(-3.0 * (h * h));
    const double *q = xyz + 3 * (int64_t)nb[i];
    double x[4] = {1.0, q[0], q[1], q[2]};
    double xu = 0.0;
    for (int a = 0; a < m; a++) xu += x[a] * x0[a];
    double ri = w * xu; /* rₗ = Wₗ Xₗ (Xₗ'WₗXₗ \ xₒ) */
    rr += ri * ri;
  }
  *mu = zhat;
  *sig = sqrt(rr); /* r̂ₒ = norm(rₗ) */
}

/* the searcher of the last local call (see gsk_oracle_krige) */
static struct { double *xyz; kdtree_t *tree; int64_t n; int dim; uint64_t key; } g_prep;

static uint64_t prep_key(const gsk_problem *p) {
  uint64_t h = 0x9E3779B97F4A7C15ull;
  for (int d = 0; d < p->dim; d++)
    for (int64_t i = 0; i < p->n_samples; i++) {
      uint64_t b;
      memcpy(&b, &p->coords[d][i], 8);
      h = (h ^ b) * 0x100000001B3ull;
      h ^= h >> 29;
    }
  return h;
}

void gsk_oracle_clear_cache(void) {
  if (g_prep.tree) kd_free(g_prep.tree);
  free(g_prep.xyz);
  memset(&g_prep, 0, sizeof(g_prep));
}

/* ------------------------------------------------------------------------------------------
 * entry point: exactsolve (krig.jl:166-186) when max_neighbors == 0, approxsolve
 * (krig.jl:188-234) otherwise, over the slab [target_first, target_first+target_count).
 * ---------------------------------------------------------------------------------------- */
int gsk_oracle_krige(const gsk_problem *p, double *mean_out, double *var_out, int32_t *nneigh_out,
                     int32_t *neigh_idx_out, int search_kind, int nthreads) {
  if (!p || p->dim < 1 || p->dim > 3 || p->n_samples < 1 || p->n_support < 1) return GSK_ERR_INVALID;
  int64_t T = gsk_oracle_num_targets(p);
  int64_t first = p->target_first;
  int64_t count = p->target_count < 0 ? T - first : p->target_count;
  if (first < 0 || first + count > T) return GSK_ERR_INVALID;
  int64_t n = p->n_samples;
  int dim = p->dim;
  vario_t v = vario_from(p);
  int32_t exps[3 * GSK_MAX_DRIFT_TERMS];
  int c = n_constraints(p, exps);
  if (c < 0) return GSK_ERR_INVALID;

  /* preprocess (krig.jl:76-128) builds the searcher ONCE per solve; a benchmark that times bounded samples of a
   * grid calls this function many times with the same samples, so the packed coordinates and the KD-tree of the
   * last local call are kept (key: coordinate bits, n, dim) instead of being rebuilt per sample */
  const int cacheable = p->max_neighbors > 0 && search_kind == GSK_ORACLE_SEARCH_KDTREE;
  const uint64_t key = cacheable ? prep_key(p) : 0;
  const int hit = cacheable && g_prep.tree && g_prep.n == n && g_prep.dim == dim && g_prep.key == key;
  double *xyz;
  if (hit) {
    xyz = g_prep.xyz;
  } else {
    xyz = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    for (int64_t i = 0; i < n; i++)
      for (int d = 0; d < 3; d++) xyz[3 * i + d] = (d < dim) ? p->coords[d][i] : 0.0;
  }

  /* thread count of THIS call only (nthreads <= 0: all cores); the process-wide OpenMP setting is left alone */
#ifdef _OPENMP
  const int nt_call = nthreads > 0 ? nthreads : omp_get_max_threads();
#else
  const int nt_call = 1;
  (void)nthreads;
#endif
  (void)nt_call;

  if (p->solver != GSK_SOLVER_KRIGING && p->max_neighbors == 0) {
    /* ---- IDW / LWR with every sample as a neighbour (maxneighbors === nothing: nmax = n, idw.jl:93, lwr.jl:95) ---- */
#pragma omp parallel num_threads(nt_call)
    {
      cand_t *best = (cand_t *)malloc(sizeof(cand_t) * (size_t)n);
      int32_t *nb = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
      double *dd = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp for schedule(dynamic, 16)
      for (int64_t t = 0; t < count; t++) {
        double ctr[3];
        target_center(p, first + t, ctr);
        int cnt = knn_brute(p, xyz, ctr, (int)n, best);
        for (int i = 0; i < cnt; i++) { nb[i] = best[i].idx; dd[i] = best[i].d2; }
        if (nneigh_out) nneigh_out[t] = cnt;
        if (cnt < p->min_neighbors || cnt == 0) { mean_out[t] = NAN; var_out[t] = NAN; continue; }
        if (p->solver == GSK_SOLVER_IDW) idw_predict(p, nb, dd, cnt, &mean_out[t], &var_out[t]);
        else lwr_predict(p, xyz, nb, dd, cnt, ctr, &mean_out[t], &var_out[t]);
      }
      free(best); free(nb); free(dd);
    }
    free(xyz);
    return GSK_OK;
  }
  if (p->max_neighbors == 0) {
    /* ---- exactsolve: fit once on all samples, predict everywhere (krig.jl:176-180) ---- */
    int m = (int)n + c;
    fitted_t f;
    f.A = (double *)malloc(sizeof(double) * (size_t)m * m);
    f.piv = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);
    int32_t *nb = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int64_t i = 0; i < n; i++) nb[i] = (int32_t)i;
    fit_system(p, &v, xyz, nb, (int)n, c, exps, &f);
#pragma omp parallel num_threads(nt_call)
    {
      double *rhs = (double *)malloc(sizeof(double) * (size_t)m);
      double *sol = (double *)malloc(sizeof(double) * (size_t)m);
#pragma omp for schedule(dynamic, 16)
      for (int64_t t = 0; t < count; t++) {
        double ctr[3];
        target_center(p, first + t, ctr);
        predict(p, &v, xyz, nb, &f, exps, ctr, rhs, sol, &mean_out[t], &var_out[t]);
        if (nneigh_out) nneigh_out[t] = (int32_t)n;
      }
      free(rhs); free(sol);
    }
    free(nb); free(f.A); free(f.piv);
    free(xyz);
    return GSK_OK;
  }

  /* ---- approxsolve: per target search → fit → predict (krig.jl:205-228) ---- */
  int k = p->max_neighbors;
  if (k < 1 || k > n) { if (!hit) free(xyz); return GSK_ERR_INVALID; } /* host clamps first (ui.jl:16-23) */
  kdtree_t *tree = NULL;
  if (hit) {
    tree = g_prep.tree;
  } else if (search_kind == GSK_ORACLE_SEARCH_KDTREE) {
    tree = kd_create(xyz, n, dim);
    gsk_oracle_clear_cache();
    g_prep.xyz = xyz; g_prep.tree = tree; g_prep.n = n; g_prep.dim = dim; g_prep.key = key;
  }
  int use_ball = !(p->ball_radius != p->ball_radius);
  int mmax = k + c;
#pragma omp parallel num_threads(nt_call)
  {
    cand_t *best = (cand_t *)malloc(sizeof(cand_t) * (size_t)k);
    int32_t *nb = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
    fitted_t f;
    f.A = (double *)malloc(sizeof(double) * (size_t)mmax * mmax);
    f.piv = (int32_t *)malloc(sizeof(int32_t) * (size_t)mmax);
    double *rhs = (double *)malloc(sizeof(double) * (size_t)mmax);
    double *sol = (double *)malloc(sizeof(double) * (size_t)mmax);
    double *dd = (double *)malloc(sizeof(double) * (size_t)k);
#pragma omp for schedule(dynamic, 64)
    for (int64_t t = 0; t < count; t++) {
      double ctr[3];
      target_center(p, first + t, ctr); /* centroid(pdomain, ind), krig.jl:207 */
      int cnt = 0;
      if (tree) kd_query(tree, 0, ctr, k, best, &cnt);
      else cnt = knn_brute(p, xyz, ctr, k, best);
      int nn = cnt;
      if (use_ball) { /* KBallSearch: keep dists .<= radius (inclusive) */
        nn = 0;
        while (nn < cnt && sqrt(best[nn].d2) <= p->ball_radius) nn++;
      }
      for (int i = 0; i < nn; i++) nb[i] = best[i].idx;
      if (nneigh_out) nneigh_out[t] = nn;
      if (neigh_idx_out)
        for (int i = 0; i < k; i++) neigh_idx_out[t * (int64_t)k + i] = (i < nn) ? nb[i] : -1;
      if (nn < p->min_neighbors || nn == 0) { /* krig.jl:213-214 → (missing, missing) */
        mean_out[t] = NAN;
        var_out[t] = NAN;
        continue;
      }
      if (p->solver != GSK_SOLVER_KRIGING) { /* idw.jl:118-140 / lwr.jl:119-145 on the same neighbour list */
        for (int i = 0; i < nn; i++) dd[i] = best[i].d2;
        if (p->solver == GSK_SOLVER_IDW) idw_predict(p, nb, dd, nn, &mean_out[t], &var_out[t]);
        else lwr_predict(p, xyz, nb, dd, nn, ctr, &mean_out[t], &var_out[t]);
        continue;
      }
      fit_system(p, &v, xyz, nb, nn, c, exps, &f);
      predict(p, &v, xyz, nb, &f, exps, ctr, rhs, sol, &mean_out[t], &var_out[t]);
    }
    free(best); free(nb); free(f.A); free(f.piv); free(rhs); free(sol); free(dd);
  }
  if (!tree) free(xyz); /* brute-force search: nothing is kept; otherwise g_prep owns xyz and the tree */
  return GSK_OK;
}

/* search only (for neighbour-set parity tests) */
int gsk_oracle_search(const gsk_problem *p, int32_t *nneigh_out, int32_t *neigh_idx_out, double *d2_out,
                      int search_kind, int nthreads) {
  if (!p || p->max_neighbors < 1 || p->max_neighbors > p->n_samples) return GSK_ERR_INVALID;
  int64_t T = gsk_oracle_num_targets(p);
  int64_t first = p->target_first;
  int64_t count = p->target_count < 0 ? T - first : p->target_count;
  if (first < 0 || first + count > T) return GSK_ERR_INVALID;
  int64_t n = p->n_samples;
  int k = p->max_neighbors;
  double *xyz = (double *)malloc(sizeof(double) * 3 * (size_t)n);
  for (int64_t i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) xyz[3 * i + d] = (d < p->dim) ? p->coords[d][i] : 0.0;
  kdtree_t *tree = (search_kind == GSK_ORACLE_SEARCH_KDTREE) ? kd_create(xyz, n, p->dim) : NULL;
  int use_ball = !(p->ball_radius != p->ball_radius);
  /* thread count of THIS call only (nthreads <= 0: all cores); the process-wide OpenMP setting is left alone */
#ifdef _OPENMP
  const int nt_call = nthreads > 0 ? nthreads : omp_get_max_threads();
#else
  const int nt_call = 1;
  (void)nthreads;
#endif
  (void)nt_call;
#pragma omp parallel num_threads(nt_call)
  {
    cand_t *best = (cand_t *)malloc(sizeof(cand_t) * (size_t)k);
#pragma omp for schedule(dynamic, 64)
    for (int64_t t = 0; t < count; t++) {
      double ctr[3];
      target_center(p, first + t, ctr);
      int cnt = 0;
      if (tree) kd_query(tree, 0, ctr, k, best, &cnt);
      else cnt = knn_brute(p, xyz, ctr, k, best);
      int nn = cnt;
      if (use_ball) { nn = 0; while (nn < cnt && sqrt(best[nn].d2) <= p->ball_radius) nn++; }
      if (nneigh_out) nneigh_out[t] = nn;
      for (int i = 0; i < k; i++) {
        if (neigh_idx_out) neigh_idx_out[t * (int64_t)k + i] = (i < nn) ? best[i].idx : -1;
        if (d2_out) d2_out[t * (int64_t)k + i] = (i < nn) ? best[i].d2 : NAN;
      }
    }
    free(best);
  }
  kd_free(tree);
  free(xyz);
  return GSK_OK;
}

/* ------------------------------------------------------------------------------------------
 * Sequential Gaussian simulation: the loop of solvesingle(problem, covars, ::SeqSim, preproc)
 * (ref: src/simulation/seq.jl:102-135) with the estimator and marginal SGS sets up
 * (ref: src/simulation/sgs.jl:62-69: SimpleKriging(variogram, mean), Normal(mean, √sill)).
 * coords: the centroids of the domain (seq.jl:91); rank[i] < 0: data (mask true after initbuff,
 * seq.jl:88), else the position of i in traverse(pdomain, path) among the others. z[i]: the
 * standard normal draw used at element i (rand(rng, Normal(μ,σ)) = μ + σ·randn(rng)).
 * Optional outputs per element: neighbour count, indices (n × k), weights (n × k), σ.
 * ---------------------------------------------------------------------------------------- */
int gsk_oracle_sgs(int dim, int64_t n, const double *const *coords, const int64_t *rank, int vario_kind,
                   double vario_range, double vario_sill, double vario_nugget, double gaussian_nugget_eps, double mean,
                   int min_neighbors, int max_neighbors, double ball_radius, const double *values, const double *z,
                   double *out, int32_t *nneigh_out, int32_t *neigh_idx_out, double *weights_out, double *sigma_out) {
  if (dim < 1 || dim > 3 || n < 1 || !coords || !rank || !z || !out || max_neighbors < 1) return GSK_ERR_INVALID;
  int k = max_neighbors < n ? max_neighbors : (int)n;
  gsk_problem pr;
  memset(&pr, 0, sizeof(pr));
  pr.dim = dim;
  pr.n_samples = n;
  pr.vario_kind = vario_kind; pr.vario_range = vario_range; pr.vario_sill = vario_sill; pr.vario_nugget = vario_nugget;
  pr.gaussian_nugget_eps = gaussian_nugget_eps;
  pr.estimator = GSK_EST_SIMPLE;
  pr.sk_mean = mean;
  pr.n_support = 1; /* predictprob at the point pset[ind] (seq.jl:124) */
  pr.flags = GSK_FLAG_CLAMP_VARIANCE;
  pr.values = out; /* the realisation buffer doubles as the neighbours' values (seq.jl:116) */
  vario_t v = vario_from(&pr);
  double *xyz = (double *)malloc(sizeof(double) * 3 * (size_t)n);
  for (int64_t i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) xyz[3 * i + d] = (d < dim) ? coords[d][i] : 0.0;
  int64_t m = 0;
  for (int64_t i = 0; i < n; i++) if (rank[i] >= 0) m++;
  int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m > 0 ? m : 1));
  unsigned char *simulated = (unsigned char *)calloc((size_t)n, 1);
  for (int64_t i = 0; i < n; i++) {
    if (rank[i] >= 0) { if (rank[i] >= m) { free(xyz); free(order); free(simulated); return GSK_ERR_INVALID; } order[rank[i]] = i; }
    else { simulated[i] = 1; out[i] = values ? values[i] : 0.0; }
  }
  int use_ball = !(ball_radius != ball_radius);
  cand_t *best = (cand_t *)malloc(sizeof(cand_t) * (size_t)k);
  int32_t *nb = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
  fitted_t f;
  f.A = (double *)malloc(sizeof(double) * (size_t)k * k);
  f.piv = NULL;
  double *rhs = (double *)malloc(sizeof(double) * (size_t)k), *sol = (double *)malloc(sizeof(double) * (size_t)k);
  for (int64_t p = 0; p < m; p++) { /* for ind in traverse(pdomain, path); if !simulated[ind] */
    int64_t ind = order[p];
    const double *c = xyz + 3 * ind;
    /* search!(neighbors, pset[ind], searcher, mask=simulated) */
    int cnt = 0;
    for (int64_t j = 0; j < n; j++)
      if (simulated[j]) topk_insert(best, &cnt, k, dist2(dim, c, xyz + 3 * j), (int32_t)j);
    int nn = cnt;
    if (use_ball) { nn = 0; while (nn < cnt && sqrt(best[nn].d2) <= ball_radius) nn++; }
    double mu = mean, sd = sqrt(vario_sill); /* marginal */
    int fitted = 0;
    if (nn >= min_neighbors && nn >= 1) {
      for (int i = 0; i < nn; i++) nb[i] = best[i].idx;
      fit_system(&pr, &v, xyz, nb, nn, 0, NULL, &f);
      double mu_c, s2;
      predict(&pr, &v, xyz, nb, &f, NULL, c, rhs, sol, &mu_c, &s2);
      if (mu_c == mu_c && s2 == s2) { mu = mu_c; sd = sqrt(s2); fitted = 1; } /* status(fitted) */
    }
    out[ind] = mu + sd * z[ind];
    simulated[ind] = 1;
    if (nneigh_out) nneigh_out[ind] = fitted ? nn : 0;
    if (sigma_out) sigma_out[ind] = sd;
    for (int i = 0; i < k; i++) {
      if (neigh_idx_out) neigh_idx_out[ind * (int64_t)k + i] = (fitted && i < nn) ? nb[i] : -1;
      if (weights_out) weights_out[ind * (int64_t)k + i] = (fitted && i < nn) ? sol[i] : 0.0;
    }
  }
  free(best); free(nb); free(f.A); free(rhs); free(sol); free(xyz); free(order); free(simulated);
  return GSK_OK;
}

int gsk_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
