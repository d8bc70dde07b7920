// sgs.cu — sequential Gaussian simulation (ref: src/simulation/sgs.jl:56-89 sets the Simple Kriging estimator and the
// Normal(mean, √sill) marginal; src/simulation/seq.jl:102-135 is the loop: for every location of the path, search the
// neighbours among the already simulated locations, fit, draw from the conditional Normal).
//
// What the reference does one location at a time splits into a part that does not depend on the simulated VALUES and a
// part that does:
//   * which locations are the k nearest already-simulated ones of location i, their Simple Kriging weights λ_i and
//     the conditional standard deviation σ_i depend on the geometry, the variogram and the path only. gsk_sgs_plan
//     computes them for ALL locations at once: a rank-masked search (search.cu, RANKED: a record is a candidate only
//     if its rank in the path is lower than the target's; data locations have the lowest rank) and one small SPD
//     solve per location (sgs_weights_kernel, one thread per location).
//   * the values follow the recurrence v_i = μ + Σ_j λ_ij (v_n(i,j) − μ) + σ_i z_i in path order
//     (sgs_recurrence_kernel: one warp per realisation, lanes over the neighbours; the realisations of an ensemble run
//     on different SMs at the same time).
// The draws z come from the caller's generator (rand(rng, Normal(μ, σ)) = μ + σ·randn(rng), one per location in path
// order), so the library holds no random state.
#include <math.h>

#include <cstring>
#include <vector>

#include "gsk_internal.cuh"

struct SgsPlan {
  long long n = 0, m = 0;  // locations, of which m are simulated (rank >= 0)
  int k = 0, dim = 0;
  double mean = 0.0, sd_marginal = 1.0;
  int *nn = nullptr;      // n: neighbours of location i (0: draw from the marginal)
  int *nbr = nullptr;     // n × k location indices
  double *lam = nullptr;  // n × k Simple Kriging weights
  double *sig = nullptr;  // n conditional standard deviations
  int *order = nullptr;   // m: location visited at path position p
  int *isdata = nullptr;  // n: 1 where the value is given
  double *vals = nullptr, *z = nullptr, *out = nullptr;  // sample buffers (grown on demand)
  size_t cap_real = 0;
};

namespace {

__global__ void sgs_trank_kernel(const int *__restrict__ perm, const int *__restrict__ rankp1, long long n,
                                 int *__restrict__ trank) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) trank[s] = rankp1[perm[s]];
}

// One thread per location (bin-sorted order s, results stored at the location's own index). The k×k covariance
// matrix of the neighbours is held as a packed lower triangle in `work` with element-major layout
// work[e · stride + thread] (shared memory: stride = blockDim.x, bank-conflict free; global scratch for large k).
__global__ void sgs_weights_kernel(GskTargets tg, GskVario vg, const double4 *__restrict__ rec, const int *__restrict__ perm,
                                   const int *__restrict__ nn_s, const int *__restrict__ nbr_s, long long n, int k,
                                   int min_neighbors, double sd_marginal, double *__restrict__ gwork, int *__restrict__ nn_out,
                                   int *__restrict__ nbr_out, double *__restrict__ lam_out, double *__restrict__ sig_out) {
  extern __shared__ double swork[];
  double *work;
  long long stride;
  if (gwork) {
    stride = (long long)gridDim.x * blockDim.x;
    work = gwork + ((long long)blockIdx.x * blockDim.x + threadIdx.x);
  } else {
    stride = blockDim.x;
    work = swork + threadIdx.x;
  }
  const int ntri = k * (k + 1) / 2;
  double *A = work;                      // packed lower triangle, row i starts at i(i+1)/2
  double *b = work + (size_t)ntri * stride;  // right-hand side, then the weights
#define AT(i, j) A[(size_t)((i) * ((i) + 1) / 2 + (j)) * stride]
#define BV(i) b[(size_t)(i) * stride]
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (long long)gridDim.x * blockDim.x) {
    const long long o = perm[s];
    int nn = nn_s[s];
    const int *nb = nbr_s + s * k;
    double sig = sd_marginal;
    bool ok = nn >= min_neighbors && nn >= 1;
    if (ok) {
      double tc[3] = {tg.pts[0][s], tg.dim > 1 ? tg.pts[1][s] : 0.0, tg.dim > 2 ? tg.pts[2][s] : 0.0};
      // covariances among the neighbours and to the location (point support: the domain's centroids, seq.jl:91)
      for (int i = 0; i < nn; ++i) {
        const double4 ri = rec[nb[i]];
        const double ci[3] = {ri.x, ri.y, ri.z};
        for (int j = 0; j < i; ++j) AT(i, j) = gsk_cov_rt(vg, gsk_dist2_exact(tg.dim, ci, rec[nb[j]]));
        AT(i, i) = vg.sill;
        BV(i) = gsk_cov_rt(vg, gsk_dist2_exact(tg.dim, tc, ri));
      }
      // Cholesky (status(fitted) is the success of this factorisation, seq.jl:123)
      for (int j = 0; j < nn && ok; ++j) {
        double d = AT(j, j);
        for (int p = 0; p < j; ++p) { const double l = AT(j, p); d = fma(-l, l, d); }
        if (!(d > 0.0)) { ok = false; break; }
        const double ld = sqrt(d), inv = 1.0 / ld;
        AT(j, j) = ld;
        for (int i = j + 1; i < nn; ++i) {
          double v = AT(i, j);
          for (int p = 0; p < j; ++p) v = fma(-AT(i, p), AT(j, p), v);
          AT(i, j) = v * inv;
        }
      }
    }
    if (ok) {
      // σ² = sill − bᵀC⁻¹b = sill − ‖L⁻¹b‖² ; λ = L⁻ᵀ L⁻¹ b
      double q = 0.0;
      for (int i = 0; i < nn; ++i) {
        double v = BV(i);
        for (int p = 0; p < i; ++p) v = fma(-AT(i, p), BV(p), v);
        v /= AT(i, i);
        BV(i) = v;
        q = fma(v, v, q);
      }
      for (int i = nn - 1; i >= 0; --i) {
        double v = BV(i);
        for (int p = i + 1; p < nn; ++p) v = fma(-AT(p, i), BV(p), v);
        BV(i) = v / AT(i, i);
      }
      double s2 = vg.sill - q;
      s2 = (s2 > 0.0) ? s2 : 0.0;  // predictvar clamps at zero
      sig = sqrt(s2);
      for (int i = 0; i < k; ++i) {
        nbr_out[o * k + i] = (i < nn) ? nb[i] : -1;
        lam_out[o * k + i] = (i < nn) ? BV(i) : 0.0;
      }
    } else {
      nn = 0;  // fewer than min_neighbors, or the factorisation failed: the marginal (seq.jl:108-110,126-128)
      for (int i = 0; i < k; ++i) { nbr_out[o * k + i] = -1; lam_out[o * k + i] = 0.0; }
    }
    nn_out[o] = nn;
    sig_out[o] = sig;
  }
#undef AT
#undef BV
}

__global__ void sgs_init_kernel(const double *__restrict__ vals, const int *__restrict__ isdata, long long n, int nreal,
                                double *__restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * nreal) return;
  const long long loc = i % n;
  out[i] = isdata[loc] ? vals[loc] : nan("");
}

// One warp per realisation walks the path. The neighbour lists, weights, σ and the draw of position p + 1 are loaded
// while position p is reduced (they do not depend on simulated values); the only dependent chain per location is
// {load the k neighbour values from L2, 5 shuffles, store}. Values are read and written past L1 (ld.cg / st.cg): a
// location simulated a few iterations ago by lane 0 must be seen by the other lanes.
__global__ void __launch_bounds__(32) sgs_recurrence_kernel(const int *__restrict__ order, long long m, long long n, int k,
                                                            const int *__restrict__ nnv, const int *__restrict__ nbr,
                                                            const double *__restrict__ lam, const double *__restrict__ sig,
                                                            const double *__restrict__ z, double mean, double *__restrict__ out) {
  const int lane = threadIdx.x;
  const long long r = blockIdx.x;
  const double *zr = z + r * n;
  double *vr = out + r * n;
  struct Slot { int loc, nn, i0, i1; double l0, l1, sg, zz; };
  auto fetch = [&](long long p, Slot &s) {
    s.loc = -1; s.nn = 0; s.i0 = s.i1 = -1; s.l0 = s.l1 = 0.0; s.sg = 0.0; s.zz = 0.0;
    if (p >= m) return;
    const int loc = order[p];
    s.loc = loc;
    s.nn = nnv[loc];
    const long long base = (long long)loc * k;
    if (lane < s.nn) { s.i0 = nbr[base + lane]; s.l0 = lam[base + lane]; }
    if (lane + 32 < s.nn) { s.i1 = nbr[base + lane + 32]; s.l1 = lam[base + lane + 32]; }
    s.sg = sig[loc];
    s.zz = zr[loc];
  };
  Slot cur, nxt;
  fetch(0, cur);
  for (long long p = 0; p < m; ++p) {
    fetch(p + 1, nxt);
    double acc = 0.0;
    if (cur.i0 >= 0) acc = cur.l0 * (__ldcg(vr + cur.i0) - mean);
    if (cur.i1 >= 0) acc = fma(cur.l1, __ldcg(vr + cur.i1) - mean, acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) __stcg(vr + cur.loc, (mean + acc) + cur.sg * cur.zz);
    __syncwarp();
    cur = nxt;
  }
}

}  // namespace

void gsk_sgs_free(gsk_ctx *ctx) {
  SgsPlan *s = ctx->sgs;
  if (!s) return;
  cudaFree(s->nn); cudaFree(s->nbr); cudaFree(s->lam); cudaFree(s->sig); cudaFree(s->order); cudaFree(s->isdata);
  cudaFree(s->vals); cudaFree(s->z); cudaFree(s->out);
  delete s;
  ctx->sgs = nullptr;
}

int gsk_sgs_plan_impl(gsk_ctx *ctx, int dim, long long n, const double *const *coords, const long long *rank,
                      const GskVario &vg, double mean, int min_neighbors, int k, double ball_radius) {
  cudaStream_t st = ctx->stream;
  // ---- the path: order[p] = location of rank p; ranks must be a permutation of 0..m−1 over the locations without data
  long long m = 0;
  for (long long i = 0; i < n; ++i) if (rank[i] >= 0) ++m;
  if (m == 0) { ctx->err = "gsk_sgs_plan: every location holds data, nothing to simulate"; return GSK_ERR_INVALID; }
  std::vector<int> order((size_t)m, -1), rankp1((size_t)n), isdata((size_t)n);
  for (long long i = 0; i < n; ++i) {
    const long long r = rank[i];
    if (r >= m || (r >= 0 && order[(size_t)r] != -1)) {
      ctx->err = "gsk_sgs_plan: rank must be -1 (data) or a permutation of 0..m-1 over the other locations";
      return GSK_ERR_INVALID;
    }
    if (r >= 0) order[(size_t)r] = (int)i;
    rankp1[(size_t)i] = (r >= 0) ? (int)r + 1 : 0;
    isdata[(size_t)i] = r < 0;
  }
  // ---- bins over all locations; the record's w carries (rank + 1) << 32 | index ----
  {
    std::vector<double> hv((size_t)n);
    for (long long i = 0; i < n; ++i) {
      const long long bits = (long long)rankp1[(size_t)i] << 32;
      memcpy(&hv[(size_t)i], &bits, sizeof(double));
    }
    int rc = gsk_build_bins(ctx, coords[0], dim > 1 ? coords[1] : nullptr, dim > 2 ? coords[2] : nullptr, hv.data(), n, dim,
                            k, true);
    if (rc != GSK_OK) return rc;
  }
  // ---- the locations as explicit targets ----
  GskTargets &tg = ctx->tg;
  memset(&tg, 0, sizeof(tg));
  tg.is_grid = 0;
  tg.dim = dim;
  tg.npts = n;
  int rc;
  {
    double *hp = nullptr;
    if ((rc = gsk_host_stage(ctx, sizeof(double) * (size_t)n * 3, (void **)&hp)) != GSK_OK) return rc;
    for (int d = 0; d < dim; ++d) {
      if ((rc = gsk_buf(ctx, (GskBufId)(BUF_PTS0 + d), sizeof(double) * (size_t)n, (void **)&ctx->d_pts[d])) != GSK_OK) return rc;
      memcpy(hp + (size_t)d * n, coords[d], sizeof(double) * (size_t)n);
      GSK_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_pts[d], hp + (size_t)d * n, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
      tg.pts[d] = ctx->d_pts[d];
    }
  }
  ctx->prob = gsk_problem{};
  ctx->prob.dim = dim;
  ctx->prob.n_samples = n;
  ctx->prob.max_neighbors = k;
  ctx->prob.min_neighbors = min_neighbors;
  ctx->prob.ball_radius = ball_radius;
  ctx->vg = vg;

  SgsPlan *s = new SgsPlan();
  ctx->sgs = s;
  s->n = n; s->m = m; s->k = k; s->dim = dim; s->mean = mean; s->sd_marginal = sqrt(vg.sill);
  GSK_CUDA_CHECK(ctx, cudaMalloc(&s->nn, sizeof(int) * (size_t)n));
  GSK_CUDA_CHECK(ctx, cudaMalloc(&s->nbr, sizeof(int) * (size_t)n * k));
  GSK_CUDA_CHECK(ctx, cudaMalloc(&s->lam, sizeof(double) * (size_t)n * k));
  GSK_CUDA_CHECK(ctx, cudaMalloc(&s->sig, sizeof(double) * (size_t)n));
  GSK_CUDA_CHECK(ctx, cudaMalloc(&s->order, sizeof(int) * (size_t)m));
  GSK_CUDA_CHECK(ctx, cudaMalloc(&s->isdata, sizeof(int) * (size_t)n));
  int *d_rankp1 = nullptr, *d_trank = nullptr;
  if ((rc = gsk_buf(ctx, BUF_PT_MEAN, sizeof(int) * (size_t)n, (void **)&d_rankp1)) != GSK_OK) return rc;
  if ((rc = gsk_buf(ctx, BUF_PT_VAR, sizeof(int) * (size_t)n, (void **)&d_trank)) != GSK_OK) return rc;
  GSK_CUDA_CHECK(ctx, cudaMemcpyAsync(s->order, order.data(), sizeof(int) * (size_t)m, cudaMemcpyHostToDevice, st));
  GSK_CUDA_CHECK(ctx, cudaMemcpyAsync(s->isdata, isdata.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
  GSK_CUDA_CHECK(ctx, cudaMemcpyAsync(d_rankp1, rankp1.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));

  // ---- bin-sort the locations, search with the rank mask, solve ----
  int *perm = nullptr, *nn_s = nullptr, *nbr_s = nullptr;
  double *sx = nullptr, *sy = nullptr, *sz = nullptr;
  if ((rc = gsk_points_sort(ctx, 0, n, &perm, &sx, &sy, &sz)) != GSK_OK) return rc;
  if ((rc = gsk_buf(ctx, BUF_PT_NN, sizeof(int) * (size_t)n, (void **)&nn_s)) != GSK_OK) return rc;
  if ((rc = gsk_buf(ctx, BUF_PT_NBR, sizeof(int) * (size_t)n * k, (void **)&nbr_s)) != GSK_OK) return rc;
  const unsigned g = (unsigned)((n + 255) / 256);
  sgs_trank_kernel<<<g, 256, 0, st>>>(perm, d_rankp1, n, d_trank);
  GskTargets tgs = tg;
  tgs.pts[0] = sx; tgs.pts[1] = sy; tgs.pts[2] = sz;
  const GskTargets tg_saved = ctx->tg;
  ctx->tg = tgs;
  int launches = 0;
  rc = gsk_launch_search(ctx, st, 0, n, nn_s, nbr_s, &launches, d_trank);
  ctx->tg = tg_saved;
  if (rc != GSK_OK) return rc;

  const size_t per_thread = sizeof(double) * (size_t)(k * (k + 1) / 2 + k);
  int tpb = 128;
  while (tpb > 32 && per_thread * tpb > (size_t)96 * 1024) tpb >>= 1;
  double *gwork = nullptr;
  size_t smem = per_thread * tpb;
  unsigned grid = (unsigned)((n + tpb - 1) / tpb);
  if (smem > (size_t)200 * 1024) {  // large k: the triangles live in global scratch (L2-resident, coalesced by thread)
    tpb = 128;
    grid = (unsigned)std::min<long long>((n + tpb - 1) / tpb, (long long)ctx->sm_count * 8);
    if ((rc = gsk_buf(ctx, BUF_G_TMP, per_thread * (size_t)grid * tpb, (void **)&gwork)) != GSK_OK) return rc;
    smem = 0;
  } else {
    GSK_CUDA_CHECK(ctx, cudaFuncSetAttribute(sgs_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  sgs_weights_kernel<<<grid, tpb, smem, st>>>(tgs, vg, ctx->d_rec_orig, perm, nn_s, nbr_s, n, k, min_neighbors,
                                             s->sd_marginal, gwork, s->nn, s->nbr, s->lam, s->sig);
  GSK_CUDA_CHECK(ctx, cudaGetLastError());
  GSK_CUDA_CHECK(ctx, cudaStreamSynchronize(st));  // host vectors above go out of scope
  ctx->timing.launches = launches + 6 + 4 + 2;
  return GSK_OK;
}

int gsk_sgs_sample_impl(gsk_ctx *ctx, int nreal, const double *values, const double *z, double *out) {
  SgsPlan *s = ctx->sgs;
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)s->n;
  if (s->cap_real < (size_t)nreal) {
    cudaFree(s->z); cudaFree(s->out); cudaFree(s->vals);
    s->z = s->out = s->vals = nullptr;
    s->cap_real = 0;
    GSK_CUDA_CHECK(ctx, cudaMalloc(&s->vals, sizeof(double) * n));
    GSK_CUDA_CHECK(ctx, cudaMalloc(&s->z, sizeof(double) * n * nreal));
    GSK_CUDA_CHECK(ctx, cudaMalloc(&s->out, sizeof(double) * n * nreal));
    s->cap_real = (size_t)nreal;
  }
  if (values) GSK_CUDA_CHECK(ctx, cudaMemcpyAsync(s->vals, values, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  else GSK_CUDA_CHECK(ctx, cudaMemsetAsync(s->vals, 0, sizeof(double) * n, st));
  GSK_CUDA_CHECK(ctx, cudaMemcpyAsync(s->z, z, sizeof(double) * n * nreal, cudaMemcpyHostToDevice, st));
  const long long tot = (long long)n * nreal;
  sgs_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(s->vals, s->isdata, s->n, nreal, s->out);
  sgs_recurrence_kernel<<<(unsigned)nreal, 32, 0, st>>>(s->order, s->m, s->n, s->k, s->nn, s->nbr, s->lam, s->sig, s->z,
                                                        s->mean, s->out);
  GSK_CUDA_CHECK(ctx, cudaGetLastError());
  GSK_CUDA_CHECK(ctx, cudaMemcpyAsync(out, s->out, sizeof(double) * n * nreal, cudaMemcpyDeviceToHost, st));
  GSK_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  ctx->timing.launches = 2;
  return GSK_OK;
}

int gsk_sgs_weights_impl(gsk_ctx *ctx, int *nn_out, int *nbr_out, double *lam_out, double *sig_out) {
  SgsPlan *s = ctx->sgs;
  const size_t n = (size_t)s->n, k = (size_t)s->k;
  if (nn_out) GSK_CUDA_CHECK(ctx, cudaMemcpy(nn_out, s->nn, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (nbr_out) GSK_CUDA_CHECK(ctx, cudaMemcpy(nbr_out, s->nbr, sizeof(int) * n * k, cudaMemcpyDeviceToHost));
  if (lam_out) GSK_CUDA_CHECK(ctx, cudaMemcpy(lam_out, s->lam, sizeof(double) * n * k, cudaMemcpyDeviceToHost));
  if (sig_out) GSK_CUDA_CHECK(ctx, cudaMemcpy(sig_out, s->sig, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return GSK_OK;
}
