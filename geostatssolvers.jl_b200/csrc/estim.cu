// estim.cu — the per-location bodies of the two distance-weighting estimation solvers that share the Kriging
// solver's searcher, traversal and centroid step:
//   IDWSolver  ref: src/estimation/idw.jl:112-142   w = 1/d^exponent, μ = Σ (w/Σw) z, second output = nearest distance;
//                                                   a zero distance returns that sample's value and distance 0
//   LWRSolver  ref: src/estimation/lwr.jl:113-146   δ = d / max d, W = diag(weightfun(δ)), X = [1 x], θ = (XᵀWX)⁻¹XᵀWz,
//                                                   ẑ = θ·[1 x₀], second output = ‖W X (XᵀWX)⁻¹ [1 x₀]‖
// Local form (maxneighbors given): one thread per target walks the neighbour list K2 (search.cu) wrote — sorted by
// (d², index), so max d is the last entry and min d the first. Distances are re-formed from the records with the
// search kernel's own FMA-free chain, i.e. they are the values `searchdists!` returns. Global form (maxneighbors ===
// nothing: every sample is a neighbour, idw.jl:93 / lwr.jl:95): the same thread loops over all samples, staged
// through shared memory in tiles.
#include <math.h>

#include "gsk_internal.cuh"

namespace {

struct SimpleArgs {
  GskTargets tg;
  const double4 *rec;  // samples in original order {x, y, z, value}
  long long n;         // samples
  int k;               // clamped max neighbours; 0 = all samples
  int min_neighbors;
  int solver;          // GSK_SOLVER_IDW / GSK_SOLVER_LWR
  double exponent;
  int weightfun;
  long long first, count;
  const int *nn;
  const int *nbr;
  int *nn_out;         // optional (global form): neighbours used = n
  GskOut out;
};

// d^e as Julia evaluates `ds .^ exponent` (idw.jl:123): integer exponents by (compensated) repeated multiplication —
// exact products for 1, 2, 3 — otherwise pow
__device__ __forceinline__ double idw_pow(double d, double e) {
  if (e == 1.0) return d;
  if (e == 2.0) return d * d;
  if (e == 3.0) return d * d * d;
  return pow(d, e);
}

__device__ __forceinline__ double lwr_weight(int kind, double h) {
  (void)kind;  // GSK_LWR_WEIGHT_EXP3H2, the default of lwr.jl:58
  return exp(-3.0 * (h * h));
}

// solve the m×m system A x = b (m <= 4) by LU with partial pivoting (what Julia's `\` does for a square matrix);
// returns false when a pivot vanishes (the reference throws SingularException there)
__device__ bool small_lu_solve(int m, double (&A)[4][4], double (&b)[4], double (&b2)[4]) {
  for (int c = 0; c < m; ++c) {
    int piv = c;
    double best = fabs(A[c][c]);
    for (int r = c + 1; r < m; ++r)
      if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
    if (!(best > 0.0)) return false;
    if (piv != c) {
      for (int j = 0; j < m; ++j) { const double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
      double t = b[c]; b[c] = b[piv]; b[piv] = t;
      t = b2[c]; b2[c] = b2[piv]; b2[piv] = t;
    }
    for (int r = c + 1; r < m; ++r) {
      const double f = A[r][c] / A[c][c];
      for (int j = c + 1; j < m; ++j) A[r][j] -= f * A[c][j];
      b[r] -= f * b[c];
      b2[r] -= f * b2[c];
    }
  }
  for (int r = m - 1; r >= 0; --r) {
    double s = b[r], s2 = b2[r];
    for (int j = r + 1; j < m; ++j) { s -= A[r][j] * b[j]; s2 -= A[r][j] * b2[j]; }
    b[r] = s / A[r][r];
    b2[r] = s2 / A[r][r];
  }
  return true;
}

// ---- local form: neighbour lists ----
__global__ void __launch_bounds__(128) simple_local_kernel(const SimpleArgs a) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.count) return;
  double tc[3];
  gsk_target_center(a.tg, a.first + t, tc);
  const int dim = a.tg.dim;
  const int nn = a.nn[t];
  double mu = NAN, s2 = NAN;
  if (nn >= a.min_neighbors && nn > 0) {
    const int *nb = a.nbr + t * a.k;
    if (a.solver == GSK_SOLVER_IDW) {
      double sw = 0.0;
      for (int i = 0; i < nn; ++i) sw += 1.0 / idw_pow(sqrt(gsk_dist2_exact(dim, tc, a.rec[nb[i]])), a.exponent);
      if (isinf(sw)) {  // some distance is zero: the first one in the sorted list (idw.jl:127-130)
        for (int i = 0; i < nn; ++i) {
          const double4 r = a.rec[nb[i]];
          if (gsk_dist2_exact(dim, tc, r) == 0.0) { mu = r.w; break; }
        }
        s2 = 0.0;
      } else {
        double acc = 0.0;
        for (int i = 0; i < nn; ++i) {
          const double4 r = a.rec[nb[i]];
          const double w = (1.0 / idw_pow(sqrt(gsk_dist2_exact(dim, tc, r)), a.exponent)) / sw;
          acc += w * r.w;
        }
        mu = acc;
        s2 = sqrt(gsk_dist2_exact(dim, tc, a.rec[nb[0]]));  // minimum(ds): the list is sorted ascending
      }
    } else {
      const int m = dim + 1;
      const double dmax = sqrt(gsk_dist2_exact(dim, tc, a.rec[nb[nn - 1]]));  // maximum(ds)
      double A[4][4] = {}, bz[4] = {}, x0[4] = {1.0, tc[0], tc[1], tc[2]};
      for (int i = 0; i < nn; ++i) {
        const double4 r = a.rec[nb[i]];
        const double w = lwr_weight(a.weightfun, sqrt(gsk_dist2_exact(dim, tc, r)) / dmax);
        const double x[4] = {1.0, r.x, r.y, r.z};
        for (int p = 0; p < m; ++p) {
          const double wx = x[p] * w;
          for (int q = 0; q < m; ++q) A[p][q] += wx * x[q];
          bz[p] += wx * r.w;
        }
      }
      if (small_lu_solve(m, A, bz, x0)) {  // bz ← θ, x0 ← (XᵀWX)⁻¹ [1 x₀]
        const double c0[4] = {1.0, tc[0], tc[1], tc[2]};
        double zhat = 0.0;
        for (int p = 0; p < m; ++p) zhat += bz[p] * c0[p];
        double rr = 0.0;
        for (int i = 0; i < nn; ++i) {
          const double4 r = a.rec[nb[i]];
          const double w = lwr_weight(a.weightfun, sqrt(gsk_dist2_exact(dim, tc, r)) / dmax);
          const double x[4] = {1.0, r.x, r.y, r.z};
          double xu = 0.0;
          for (int p = 0; p < m; ++p) xu += x[p] * x0[p];
          const double ri = w * xu;
          rr += ri * ri;
        }
        mu = zhat;
        s2 = sqrt(rr);
      }
    }
  }
  gsk_store_result(a.out, t, mu, s2);
}

// ---- global form: all samples are neighbours; tiles of samples staged through shared memory ----
constexpr int GT_TILE = 256;
__global__ void __launch_bounds__(128) simple_global_kernel(const SimpleArgs a) {
  __shared__ double4 tile[GT_TILE];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < a.count;
  double tc[3] = {0.0, 0.0, 0.0};
  if (live) gsk_target_center(a.tg, a.first + t, tc);
  const int dim = a.tg.dim, m = dim + 1;
  // pass 1: IDW sums / LWR maximum distance
  double sw = 0.0, swz = 0.0, dmin = INFINITY, dmax = 0.0, zval = NAN;
  bool zero_seen = false;
  for (long long base = 0; base < a.n; base += GT_TILE) {
    const int cnt = (int)min((long long)GT_TILE, a.n - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) tile[i] = a.rec[base + i];
    __syncthreads();
    for (int i = 0; i < cnt; ++i) {
      const double4 r = tile[i];
      const double d = sqrt(gsk_dist2_exact(dim, tc, r));
      dmin = fmin(dmin, d);
      dmax = fmax(dmax, d);
      if (a.solver == GSK_SOLVER_IDW) {
        if (d == 0.0 && !zero_seen) { zero_seen = true; zval = r.w; }
        const double w = 1.0 / idw_pow(d, a.exponent);
        sw += w;
        swz += w * r.w;
      }
    }
  }
  double mu = NAN, s2 = NAN;
  const bool enough = a.n >= a.min_neighbors;
  if (a.solver == GSK_SOLVER_IDW) {
    if (enough) {
      if (isinf(sw)) { mu = zval; s2 = 0.0; }
      else { mu = swz / sw; s2 = dmin; }
    }
  } else {
    // pass 2: normal equations; pass 3: ‖W X u‖
    double A[4][4] = {}, bz[4] = {}, x0[4] = {1.0, tc[0], tc[1], tc[2]};
    for (long long base = 0; base < a.n; base += GT_TILE) {
      const int cnt = (int)min((long long)GT_TILE, a.n - base);
      __syncthreads();
      for (int i = threadIdx.x; i < cnt; i += blockDim.x) tile[i] = a.rec[base + i];
      __syncthreads();
      for (int i = 0; i < cnt; ++i) {
        const double4 r = tile[i];
        const double w = lwr_weight(a.weightfun, sqrt(gsk_dist2_exact(dim, tc, r)) / dmax);
        const double x[4] = {1.0, r.x, r.y, r.z};
        for (int p = 0; p < m; ++p) {
          const double wx = x[p] * w;
          for (int q = 0; q < m; ++q) A[p][q] += wx * x[q];
          bz[p] += wx * r.w;
        }
      }
    }
    const bool ok = small_lu_solve(m, A, bz, x0);
    double rr = 0.0;
    for (long long base = 0; base < a.n; base += GT_TILE) {
      const int cnt = (int)min((long long)GT_TILE, a.n - base);
      __syncthreads();
      for (int i = threadIdx.x; i < cnt; i += blockDim.x) tile[i] = a.rec[base + i];
      __syncthreads();
      for (int i = 0; i < cnt; ++i) {
        const double4 r = tile[i];
        const double w = lwr_weight(a.weightfun, sqrt(gsk_dist2_exact(dim, tc, r)) / dmax);
        const double x[4] = {1.0, r.x, r.y, r.z};
        double xu = 0.0;
        for (int p = 0; p < m; ++p) xu += x[p] * x0[p];
        const double ri = w * xu;
        rr += ri * ri;
      }
    }
    if (ok && enough) {
      const double c0[4] = {1.0, tc[0], tc[1], tc[2]};
      double zhat = 0.0;
      for (int p = 0; p < m; ++p) zhat += bz[p] * c0[p];
      mu = zhat;
      s2 = sqrt(rr);
    }
  }
  if (live) {
    gsk_store_result(a.out, t, mu, s2);
    if (a.nn_out) a.nn_out[t] = (int)a.n;
  }
}

}  // namespace

int gsk_launch_simple_solver(gsk_ctx *ctx, cudaStream_t st, long long first, long long count, const int *d_nn,
                             const int *d_nbr, long long out_off, int *d_nn_out, int *launches) {
  SimpleArgs a{};
  a.tg = ctx->tg;
  a.rec = ctx->d_rec_orig;
  a.n = ctx->prob.n_samples;
  a.k = ctx->prob.max_neighbors;
  a.min_neighbors = ctx->prob.min_neighbors;
  a.solver = ctx->prob.solver;
  a.exponent = ctx->prob.idw_exponent;
  a.weightfun = ctx->prob.lwr_weightfun;
  a.first = first;
  a.count = count;
  a.nn = d_nn;
  a.nbr = d_nbr;
  a.nn_out = d_nn_out;
  a.out = ctx->out;
  for (int p = 0; p < a.out.n; ++p) {
    a.out.mean[p] += out_off;
    a.out.var[p] += out_off;
  }
  if (count <= 0) return GSK_OK;
  const unsigned grid = (unsigned)((count + 127) / 128);
  if (a.k > 0) simple_local_kernel<<<grid, 128, 0, st>>>(a);
  else simple_global_kernel<<<grid, 128, 0, st>>>(a);
  GSK_CUDA_CHECK(ctx, cudaGetLastError());
  if (launches) *launches += 1;
  return GSK_OK;
}
