// api.cu — the C ABI of libgskrige.so (include/gskrige.h): context, plan/execute orchestration, the
// one-shot host-buffer entry point that replaces exactsolve/approxsolve (ref: src/estimation/krig.jl:
// 166-186, 188-234), and the host helpers every binding shares. No CPU compute path exists here:
// without an sm_100 device gsk_create fails and nothing else can be called.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "gsk_internal.cuh"

static thread_local std::string g_static_err;

static int fail(gsk_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg;
  else g_static_err = msg;
  return code;
}

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
extern "C" GSK_API int gsk_abi_version(void) { return GSK_ABI_VERSION; }

extern "C" GSK_API int64_t gsk_num_targets(const gsk_problem *p) {
  if (!p) return -1;
  if (p->grid_dims[0] > 0) {
    int64_t t = 1;
    for (int d = 0; d < p->dim && d < 3; ++d) t *= p->grid_dims[d];
    return t;
  }
  return p->n_points;
}

// GeoStatsModels' UKexps [3P]: exponent vectors of total degree 0..degree (each degree in descending
// lexicographic order), stably sorted by descending max exponent → degree 1: x, y, (z), 1.
extern "C" GSK_API int gsk_uk_exponents(int degree, int dim, int32_t *out, int cap) {
  if (degree < 0 || degree > 2 || dim < 1 || dim > 3 || !out) return GSK_ERR_INVALID;
  int tmp[16][3];
  int n = 0;
  for (int deg = 0; deg <= degree; ++deg)
    for (int a = deg; a >= 0; --a)
      for (int b = deg - a; b >= 0; --b) {
        int c = deg - a - b;
        if (dim == 1 && (b != 0 || c != 0)) continue;
        if (dim == 2 && c != 0) continue;
        tmp[n][0] = a; tmp[n][1] = b; tmp[n][2] = c;
        ++n;
      }
  if (n > cap) return GSK_ERR_INVALID;
  int w = 0;
  for (int mx = degree; mx >= 0; --mx)
    for (int i = 0; i < n; ++i)
      if (std::max(tmp[i][0], std::max(tmp[i][1], tmp[i][2])) == mx) {
        for (int d = 0; d < dim; ++d) out[w * dim + d] = tmp[i][d];
        ++w;
      }
  return n;
}

// Variography's geometry sub-sampling for γ(cell, point) [3P, SURVEY V1]: per axis
// n = ceil(side / (min(range, min side)/3)) points at parametric positions j/(n+1), j = 1..n.
extern "C" GSK_API int gsk_default_support(int dim, const double *spacing, double vario_range, double *ox, double *oy,
                                   double *oz, int cap) {
  if (dim < 1 || dim > 3 || !spacing || !ox) return GSK_ERR_INVALID;
  double lmin = INFINITY;
  for (int d = 0; d < dim; ++d)
    if (spacing[d] > 0) lmin = std::min(lmin, spacing[d]);
  if (!(lmin < INFINITY)) return GSK_ERR_INVALID;
  double step = ((vario_range > 0) ? std::min(vario_range, lmin) : lmin) / 3.0;
  int n[3] = {1, 1, 1};
  long long tot = 1;
  for (int d = 0; d < dim; ++d) {
    n[d] = std::max(1, (int)ceil(spacing[d] / step - 1e-12));
    tot *= n[d];
  }
  if (tot > cap) return GSK_ERR_INVALID;
  double *o[3] = {ox, oy, oz};
  int w = 0;
  for (int kz = 0; kz < n[2]; ++kz)
    for (int ky = 0; ky < n[1]; ++ky)
      for (int kx = 0; kx < n[0]; ++kx) {
        int kk[3] = {kx, ky, kz};
        for (int d = 0; d < dim; ++d)
          if (o[d]) o[d][w] = ((double)(kk[d] + 1) / (double)(n[d] + 1) - 0.5) * spacing[d];
        ++w;
      }
  return (int)tot;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
extern "C" GSK_API int gsk_create(gsk_ctx **out, int device_id) try {
  if (!out) return fail(nullptr, GSK_ERR_INVALID, "gsk_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, GSK_ERR_CUDA,
                std::string("no CUDA device visible (libgskrige has no CPU fallback): ") + cudaGetErrorString(e));
  if (device_id < 0 || device_id >= ndev) return fail(nullptr, GSK_ERR_INVALID, "gsk_create: device_id out of range");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device_id);
  if (e != cudaSuccess) return fail(nullptr, GSK_ERR_CUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, GSK_ERR_CUDA, "device is not sm_100 (libgskrige is built for B200 / sm_100a only)");
  gsk_ctx *ctx = new (std::nothrow) gsk_ctx();
  if (!ctx) return fail(nullptr, GSK_ERR_NOMEM, "gsk_create: out of host memory");
  ctx->device = device_id;
  ctx->sm_count = prop.multiProcessorCount;
  e = cudaSetDevice(device_id);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  ctx->own_stream = true;
  for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&ctx->ev_search[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_solve[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    std::string m = cudaGetErrorString(e);
    delete ctx;
    return fail(nullptr, GSK_ERR_CUDA, m);
  }
  *out = ctx;
  return GSK_OK;
} catch (const std::bad_alloc &) {  // no exception may cross the C ABI
  return fail(nullptr, GSK_ERR_NOMEM, "gsk_create: out of host memory");
} catch (const std::exception &e) {
  return fail(nullptr, GSK_ERR_STATE, std::string("gsk_create: ") + e.what());
}

int gsk_buf(gsk_ctx *ctx, GskBufId id, size_t bytes, void **out) {
  if (bytes == 0) bytes = 16;
  if (ctx->bufcap[id] < bytes) {
    cudaFree(ctx->bufp[id]);
    ctx->bufp[id] = nullptr;
    ctx->bufcap[id] = 0;
    cudaError_t e = cudaMalloc(&ctx->bufp[id], bytes);
    if (e != cudaSuccess) {
      ctx->err = std::string("device allocation failed: ") + cudaGetErrorString(e);
      return GSK_ERR_NOMEM;
    }
    ctx->bufcap[id] = bytes;
  }
  *out = ctx->bufp[id];
  return GSK_OK;
}

int gsk_host_stage(gsk_ctx *ctx, size_t bytes, void **out) {
  if (ctx->h_stage_cap < bytes) {
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr;
    ctx->h_stage_cap = 0;
    cudaError_t e = cudaHostAlloc(&ctx->h_stage, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
      ctx->err = std::string("pinned host allocation failed: ") + cudaGetErrorString(e);
      return GSK_ERR_NOMEM;
    }
    ctx->h_stage_cap = bytes;
  }
  *out = ctx->h_stage;
  return GSK_OK;
}

static void free_plan(gsk_ctx *ctx) {
  // device buffers stay cached in ctx->bufp; only the plan state is dropped
  ctx->d_rec_orig = ctx->d_rec_sorted = nullptr;
  ctx->d_cell_start = nullptr;
  ctx->d_sup = nullptr;
  for (int d = 0; d < 3; ++d) ctx->d_pts[d] = nullptr;
  gsk_global_free(ctx);
  ctx->planned = false;
}

extern "C" GSK_API void gsk_destroy(gsk_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  free_plan(ctx);
  for (int i = 0; i < BUF_COUNT; ++i) cudaFree(ctx->bufp[i]);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  cudaFree(ctx->d_nn);
  cudaFree(ctx->d_nbr);
  cudaFree(ctx->d_mean);
  cudaFree(ctx->d_var);
  for (int i = 0; i < 6; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_search[i]) cudaEventDestroy(ctx->ev_search[i]);
    if (ctx->ev_solve[i]) cudaEventDestroy(ctx->ev_solve[i]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  delete ctx;
}

extern "C" GSK_API const char *gsk_last_error(const gsk_ctx *ctx) { return ctx ? ctx->err.c_str() : g_static_err.c_str(); }

extern "C" GSK_API int gsk_set_stream(gsk_ctx *ctx, void *cuda_stream) {
  if (!ctx) return GSK_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return GSK_OK;
}

extern "C" GSK_API int gsk_synchronize(gsk_ctx *ctx) {
  if (!ctx) return GSK_ERR_INVALID;
  GSK_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  GSK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return GSK_OK;
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
static int validate(gsk_ctx *ctx, const gsk_problem *p) {
  if (!p) return fail(ctx, GSK_ERR_INVALID, "problem is NULL");
  if (p->abi_version != GSK_ABI_VERSION) return fail(ctx, GSK_ERR_INVALID, "gsk_problem.abi_version mismatch");
  if (p->dim < 1 || p->dim > 3) return fail(ctx, GSK_ERR_UNSUPPORTED, "dim must be 1, 2 or 3");
  if (p->n_samples < 1) return fail(ctx, GSK_ERR_INVALID, "n_samples must be >= 1");
  if (p->n_samples > 0x7fffffffLL) return fail(ctx, GSK_ERR_UNSUPPORTED, "n_samples must fit in int32");
  if (!p->values) return fail(ctx, GSK_ERR_INVALID, "values is NULL");
  for (int d = 0; d < p->dim; ++d)
    if (!p->coords[d]) return fail(ctx, GSK_ERR_INVALID, "coords[d] is NULL for d < dim");
  if (p->grid_dims[0] > 0) {
    for (int d = 0; d < p->dim; ++d) {
      if (p->grid_dims[d] < 1) return fail(ctx, GSK_ERR_INVALID, "grid_dims must be >= 1");
      if (!(p->grid_spacing[d] > 0.0) || !std::isfinite(p->grid_spacing[d]) || !std::isfinite(p->grid_origin[d]))
        return fail(ctx, GSK_ERR_INVALID, "grid_spacing must be finite and > 0, grid_origin finite");
    }
  } else {
    if (p->n_points < 0) return fail(ctx, GSK_ERR_INVALID, "n_points must be >= 0");
    for (int d = 0; d < p->dim; ++d)
      if (p->n_points > 0 && !p->point_coords[d]) return fail(ctx, GSK_ERR_INVALID, "point_coords[d] is NULL");
  }
  if (p->n_support < 1 || p->n_support > GSK_MAX_SUPPORT)
    return fail(ctx, GSK_ERR_INVALID, "n_support must be in [1, GSK_MAX_SUPPORT]");
  if (p->vario_kind < 0 || p->vario_kind > 2) return fail(ctx, GSK_ERR_UNSUPPORTED, "unknown variogram kind");
  if (!(p->vario_range > 0.0)) return fail(ctx, GSK_ERR_INVALID, "vario_range must be > 0");
  if (!(p->vario_sill > 0.0)) return fail(ctx, GSK_ERR_INVALID, "vario_sill must be > 0");
  if (!(p->vario_nugget >= 0.0) || !(p->gaussian_nugget_eps >= 0.0))
    return fail(ctx, GSK_ERR_INVALID, "vario_nugget and gaussian_nugget_eps must be >= 0");
  if (!(p->vario_sill - p->vario_nugget - (p->vario_kind == GSK_VARIO_GAUSSIAN ? p->gaussian_nugget_eps : 0.0) > 0.0) ||
      !std::isfinite(p->vario_sill) || !std::isfinite(p->vario_range))
    return fail(ctx, GSK_ERR_INVALID, "the nugget must be below a finite sill, the range finite");
  if (p->estimator == GSK_EST_SIMPLE && !std::isfinite(p->sk_mean))
    return fail(ctx, GSK_ERR_INVALID, "sk_mean must be finite");
  if (p->min_neighbors < 0) return fail(ctx, GSK_ERR_INVALID, "min_neighbors must be >= 0");
  if (p->estimator < 0 || p->estimator > 2) return fail(ctx, GSK_ERR_UNSUPPORTED, "unknown estimator");
  if (p->estimator == GSK_EST_UNIVERSAL && (p->uk_degree < 0 || p->uk_degree > 2))
    return fail(ctx, GSK_ERR_UNSUPPORTED, "uk_degree must be 0, 1 or 2");
  if (p->max_neighbors < 0) return fail(ctx, GSK_ERR_INVALID, "max_neighbors must be >= 0");
  if (p->max_neighbors > GSK_MAX_NEIGHBORS)
    return fail(ctx, GSK_ERR_UNSUPPORTED, "max_neighbors exceeds GSK_MAX_NEIGHBORS on the local path");
  if (p->max_neighbors > p->n_samples)
    return fail(ctx, GSK_ERR_INVALID, "max_neighbors must be clamped to n_samples by the host (ui.jl:16-23)");
  if (p->max_neighbors > 0 && !(p->ball_radius != p->ball_radius) && !(p->ball_radius > 0.0))
    return fail(ctx, GSK_ERR_INVALID, "ball_radius must be > 0 or NaN");
  int64_t T = gsk_num_targets(p);
  int64_t first = p->target_first, count = p->target_count < 0 ? T - first : p->target_count;
  if (first < 0 || count < 0 || first + count > T) return fail(ctx, GSK_ERR_INVALID, "target slab out of range");
  return GSK_OK;
}

extern "C" GSK_API int gsk_plan(gsk_ctx *ctx, const gsk_problem *p) try {
  if (!ctx) return GSK_ERR_INVALID;
  int rc = validate(ctx, p);
  if (rc != GSK_OK) return rc;
  GSK_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  GSK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  free_plan(ctx);
  GSK_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  ctx->prob = *p;
  ctx->n_targets = gsk_num_targets(p);
  const int dim = p->dim;

  // variogram constants
  GskVario &v = ctx->vg;
  v.kind = p->vario_kind;
  v.sill = p->vario_sill;
  double nug = p->vario_nugget + (p->vario_kind == GSK_VARIO_GAUSSIAN ? p->gaussian_nugget_eps : 0.0);
  v.cs = p->vario_sill - nug;
  v.range = p->vario_range;
  v.inv_r = 1.0 / p->vario_range;
  v.inv_r2 = v.inv_r * v.inv_r;
  v.hcs = 0.5 * v.cs;
  v.m15cs = -1.5 * v.cs;
  v.m3ir2 = -3.0 * v.inv_r2;
  v.m3ir = -3.0 * v.inv_r;

  // estimator
  GskEstimator &es = ctx->es;
  memset(&es, 0, sizeof(es));
  es.kind = p->estimator;
  es.sk_mean = p->sk_mean;
  es.nterms = 0;
  if (p->estimator == GSK_EST_ORDINARY) es.nterms = 1;
  if (p->estimator == GSK_EST_UNIVERSAL) {
    int32_t ex[3 * GSK_MAX_DRIFT_TERMS];
    int c = gsk_uk_exponents(p->uk_degree, dim, ex, GSK_MAX_DRIFT_TERMS);
    if (c < 0) return fail(ctx, GSK_ERR_UNSUPPORTED, "unsupported Universal Kriging degree");
    es.nterms = c;
    for (int t = 0; t < c; ++t)
      for (int d = 0; d < dim; ++d) es.exps[t][d] = ex[t * dim + d];
  }
  ctx->nterms = es.nterms;

  // targets
  GskTargets &tg = ctx->tg;
  memset(&tg, 0, sizeof(tg));
  tg.dim = dim;
  tg.is_grid = p->grid_dims[0] > 0;
  for (int d = 0; d < 3; ++d) {
    tg.gdim[d] = (tg.is_grid && d < dim) ? p->grid_dims[d] : 1;
    tg.gorg[d] = (d < dim) ? p->grid_origin[d] : 0.0;
    tg.gsp[d] = (d < dim) ? p->grid_spacing[d] : 1.0;
  }
  if (!tg.is_grid) {
    tg.npts = p->n_points;
    for (int d = 0; d < dim; ++d) {
      rc = gsk_buf(ctx, (GskBufId)(BUF_PTS0 + d), sizeof(double) * (size_t)std::max<int64_t>(1, p->n_points), (void **)&ctx->d_pts[d]);
      if (rc != GSK_OK) return rc;
      GSK_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_pts[d], p->point_coords[d], sizeof(double) * (size_t)p->n_points,
                                          cudaMemcpyHostToDevice, ctx->stream));
      tg.pts[d] = ctx->d_pts[d];
    }
  }

  // block support offsets, [3][nsup]
  {
    std::vector<double> sup(3 * (size_t)p->n_support, 0.0);
    for (int d = 0; d < dim; ++d)
      if (p->support_offsets[d])
        for (int q = 0; q < p->n_support; ++q) sup[(size_t)d * p->n_support + q] = p->support_offsets[d][q];
    rc = gsk_buf(ctx, BUF_SUP, sizeof(double) * sup.size(), (void **)&ctx->d_sup);
    if (rc != GSK_OK) return rc;
    GSK_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_sup, sup.data(), sizeof(double) * sup.size(), cudaMemcpyHostToDevice,
                                        ctx->stream));
    // block-support RHS shortcut for the exponential model (local_solve.cuh: rhs_block_support)
    double dmax2 = 0.0;
    for (int q = 0; q < p->n_support; ++q) {
      double d2 = 0.0;
      for (int d = 0; d < 3; ++d) d2 += sup[(size_t)d * p->n_support + q] * sup[(size_t)d * p->n_support + q];
      dmax2 = std::max(dmax2, d2);
    }
    ctx->sup_rmax = sqrt(dmax2);
    // tensor-grid support with 3 offsets per axis, x fastest (what gsk_default_support produces for cells no larger
    // than the range): the solve kernels then form the squared distances to the support points from per-axis squares
    {
      const int q = p->n_support;
      int want = 1;
      for (int d = 0; d < dim; ++d) want *= 3;
      bool ok = (q == want);
      for (int d = 0; d < 3 && ok; ++d) {
        const int stride = (d == 0) ? 1 : (d == 1 ? 3 : 9);
        for (int a = 0; a < 3; ++a) ctx->sup_ax[d][a] = (d < dim) ? sup[(size_t)d * q + (size_t)a * stride] : 0.0;
        for (int i = 0; i < q && ok; ++i) {
          const double expect = (d < dim) ? ctx->sup_ax[d][(i / stride) % 3] : 0.0;
          ok = (sup[(size_t)d * q + i] == expect);
        }
      }
      ctx->sup_tensor3 = ok ? 1 : 0;
    }
    ctx->rhs_taylor = (p->vario_kind == GSK_VARIO_EXPONENTIAL && p->n_support > 1 &&
                       3.0 * sqrt(dmax2) / p->vario_range <= 0.06 && !getenv("GSK_NO_RHS_TAYLOR")) ? 1 : 0;
    GSK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  }

  if (p->max_neighbors > 0) {
    rc = gsk_build_bins(ctx, p->coords[0], dim > 1 ? p->coords[1] : nullptr, dim > 2 ? p->coords[2] : nullptr,
                        p->values, p->n_samples, dim, p->max_neighbors);
  } else {
    rc = gsk_global_plan(ctx, p->coords[0], dim > 1 ? p->coords[1] : nullptr, dim > 2 ? p->coords[2] : nullptr,
                         p->values);
  }
  if (rc != GSK_OK) return rc;
  GSK_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  GSK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
  ctx->timing = gsk_timing{};
  ctx->timing.ms_plan = ms;
  // the host pointers of the problem are not kept
  for (int d = 0; d < 3; ++d) { ctx->prob.coords[d] = nullptr; ctx->prob.point_coords[d] = nullptr; ctx->prob.support_offsets[d] = nullptr; }
  ctx->prob.values = nullptr;
  ctx->planned = true;
  return GSK_OK;
} catch (const std::bad_alloc &) {  // no exception may cross the C ABI
  return fail(ctx, GSK_ERR_NOMEM, "gsk_plan: out of host memory");
} catch (const std::exception &e) {
  return fail(ctx, GSK_ERR_STATE, std::string("gsk_plan: ") + e.what());
}

// ---------------------------------------------------------------------------------------------
// execute
// ---------------------------------------------------------------------------------------------
static int ensure(gsk_ctx *ctx, void **buf, size_t *cap, size_t bytes) {
  if (*cap >= bytes) return GSK_OK;
  cudaFree(*buf);
  *buf = nullptr;
  *cap = 0;
  cudaError_t e = cudaMalloc(buf, bytes);
  if (e != cudaSuccess) return fail(ctx, GSK_ERR_NOMEM, std::string("device allocation failed: ") + cudaGetErrorString(e));
  *cap = bytes;
  return GSK_OK;
}

static long long local_chunk_targets() {
  static long long v = 0;
  if (v == 0) {
    const char *s = getenv("GSK_CHUNK_TARGETS");
    v = s ? std::max<long long>(1024, atoll(s)) : -1;  // -1: automatic
  }
  return v;
}

static int execute_impl(gsk_ctx *ctx, int64_t first, int64_t count, int32_t *d_nneigh, int32_t *d_neigh_idx);

extern "C" GSK_API int gsk_execute(gsk_ctx *ctx, int64_t first, int64_t count, double *d_mean, double *d_var,
                           int32_t *d_nneigh, int32_t *d_neigh_idx) try {
  if (!ctx) return GSK_ERR_INVALID;
  if (!d_mean || !d_var) return fail(ctx, GSK_ERR_INVALID, "output buffers are NULL");
  ctx->out = GskOut{};
  ctx->out.n = 1;
  ctx->out.mean[0] = d_mean;
  ctx->out.var[0] = d_var;
  return execute_impl(ctx, first, count, d_nneigh, d_neigh_idx);
} catch (const std::bad_alloc &) {  // no exception may cross the C ABI
  return fail(ctx, GSK_ERR_NOMEM, "gsk_execute: out of host memory");
} catch (const std::exception &e) {
  return fail(ctx, GSK_ERR_STATE, std::string("gsk_execute: ") + e.what());
}

extern "C" GSK_API int gsk_execute_peers(gsk_ctx *ctx, int64_t first, int64_t count, int n_peers,
                                 double *const *d_mean_peers, double *const *d_var_peers, int64_t out_offset,
                                 int multicast, int32_t *d_nneigh, int32_t *d_neigh_idx) try {
  if (!ctx) return GSK_ERR_INVALID;
  if (n_peers < 1 || n_peers > GSK_MAX_PEERS || !d_mean_peers || !d_var_peers)
    return fail(ctx, GSK_ERR_INVALID, "n_peers must be in [1, 8] with non-NULL pointer lists");
  if (multicast && n_peers != 1) return fail(ctx, GSK_ERR_INVALID, "multicast takes exactly one (multicast) address per field");
  ctx->out = GskOut{};
  ctx->out.n = n_peers;
  ctx->out.multicast = multicast ? 1 : 0;
  for (int p = 0; p < n_peers; ++p) {
    if (!d_mean_peers[p] || !d_var_peers[p]) return fail(ctx, GSK_ERR_INVALID, "peer output buffer is NULL");
    ctx->out.mean[p] = d_mean_peers[p] + out_offset;
    ctx->out.var[p] = d_var_peers[p] + out_offset;
  }
  return execute_impl(ctx, first, count, d_nneigh, d_neigh_idx);
} catch (const std::bad_alloc &) {  // no exception may cross the C ABI
  return fail(ctx, GSK_ERR_NOMEM, "gsk_execute_peers: out of host memory");
} catch (const std::exception &e) {
  return fail(ctx, GSK_ERR_STATE, std::string("gsk_execute_peers: ") + e.what());
}

static int execute_impl(gsk_ctx *ctx, int64_t first, int64_t count, int32_t *d_nneigh, int32_t *d_neigh_idx) {
  if (!ctx->planned) return fail(ctx, GSK_ERR_STATE, "gsk_execute called before gsk_plan");
  if (first < 0 || count < 0 || first + count > ctx->n_targets) return fail(ctx, GSK_ERR_INVALID, "target range out of bounds");
  GSK_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const bool phase_timing = ctx->phase_timing;
  int launches = 0;
  double ms_search = 0.0, ms_solve = 0.0;
  GSK_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  int rc = GSK_OK;
  if (ctx->prob.max_neighbors == 0) {
    rc = gsk_global_execute(ctx, first, count, d_nneigh, &launches);
    if (rc != GSK_OK) return rc;
  } else if (!ctx->tg.is_grid && count > 0) {
    // explicit points: process the slab in bin-sorted order (spatially coherent CTAs), then scatter back
    const int k = ctx->prob.max_neighbors;
    int *perm = nullptr, *nn_s = nullptr, *nbr_s = nullptr;
    double *sx = nullptr, *sy = nullptr, *sz = nullptr, *ms = nullptr, *vs = nullptr;
    if ((rc = gsk_points_sort(ctx, first, count, &perm, &sx, &sy, &sz)) != GSK_OK) return rc;
    if ((rc = gsk_buf(ctx, BUF_PT_MEAN, sizeof(double) * (size_t)count, (void **)&ms)) != GSK_OK) return rc;
    if ((rc = gsk_buf(ctx, BUF_PT_VAR, sizeof(double) * (size_t)count, (void **)&vs)) != GSK_OK) return rc;
    if ((rc = gsk_buf(ctx, BUF_PT_NN, sizeof(int) * (size_t)count, (void **)&nn_s)) != GSK_OK) return rc;
    if ((rc = gsk_buf(ctx, BUF_PT_NBR, sizeof(int) * (size_t)count * k, (void **)&nbr_s)) != GSK_OK) return rc;
    const GskTargets tg_saved = ctx->tg;
    const GskOut out_saved = ctx->out;
    ctx->tg.pts[0] = sx; ctx->tg.pts[1] = sy; ctx->tg.pts[2] = sz;
    ctx->tg.npts = count;
    ctx->out = GskOut{};
    ctx->out.n = 1;
    ctx->out.mean[0] = ms;
    ctx->out.var[0] = vs;
    if (phase_timing) cudaEventRecord(ctx->ev[3], ctx->stream);
    rc = gsk_launch_search(ctx, ctx->stream, 0, count, nn_s, nbr_s, &launches);
    if (phase_timing) cudaEventRecord(ctx->ev[4], ctx->stream);
    if (rc == GSK_OK) rc = gsk_launch_local_solve(ctx, ctx->stream, 0, count, nn_s, nbr_s, 0, &launches);
    ctx->tg = tg_saved;
    ctx->out = out_saved;
    if (rc != GSK_OK) return rc;
    if (phase_timing) {
      cudaEventRecord(ctx->ev[5], ctx->stream);
      cudaEventSynchronize(ctx->ev[5]);
      float a = 0.f, b2 = 0.f;
      cudaEventElapsedTime(&a, ctx->ev[3], ctx->ev[4]);
      cudaEventElapsedTime(&b2, ctx->ev[4], ctx->ev[5]);
      ms_search += a;
      ms_solve += b2;
    }
    rc = gsk_points_unscatter(ctx, perm, count, ms, vs, ctx->out, nn_s, d_nneigh, nbr_s, d_neigh_idx, k);
    if (rc != GSK_OK) return rc;
    launches += 4;
  } else {
    const int k = ctx->prob.max_neighbors;
    // chunks of ~1M targets bound the neighbour-list scratch (4k+4 B per target); chunk edges are aligned
    // to whole tile layers of the grid where that is cheap. GSK_OVERLAP=1 runs the search of chunk c+1 on
    // a side stream while chunk c is solved (measured on B200: no gain — both kernels already fill the
    // SMs and small chunks add tail effects — so it is off by default).
    static const bool want_overlap = getenv("GSK_OVERLAP") != nullptr;
    long long chunk = local_chunk_targets();
    if (chunk <= 0) {
      chunk = 1ll << 20;
      if (ctx->tg.is_grid) {
        const int dim = ctx->tg.dim;
        long long unit = (dim == 3) ? ctx->tg.gdim[0] * ctx->tg.gdim[1] * 4 : (dim == 2 ? ctx->tg.gdim[0] * 8 : 128);
        if (unit <= (1ll << 21)) chunk = (chunk + unit - 1) / unit * unit;
      }
    }
    chunk = std::min<long long>(chunk, std::max<long long>(count, 1));
    const bool overlap = want_overlap && !phase_timing && count > chunk;
    const int nbuf = overlap ? 2 : 1;
    if (!d_nneigh) {
      rc = ensure(ctx, (void **)&ctx->d_nn, &ctx->cap_nn, sizeof(int) * (size_t)chunk * nbuf);
      if (rc != GSK_OK) return rc;
    }
    if (!d_neigh_idx) {
      rc = ensure(ctx, (void **)&ctx->d_nbr, &ctx->cap_nbr, sizeof(int) * (size_t)chunk * k * nbuf);
      if (rc != GSK_OK) return rc;
    }
    if (overlap) {
      GSK_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
      GSK_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    }
    long long c = 0;
    for (long long off = 0; off < count; off += chunk, ++c) {
      const long long cnt = std::min<long long>(chunk, count - off);
      const int b = (int)(c % nbuf);
      int *nn = d_nneigh ? d_nneigh + off : ctx->d_nn + (size_t)b * chunk;
      int *nbr = d_neigh_idx ? d_neigh_idx + off * k : ctx->d_nbr + (size_t)b * chunk * k;
      if (overlap) {
        // the scratch of parity b is free once the solve of chunk c-2 has read it
        if (c >= 2) GSK_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_solve[b], 0));
        rc = gsk_launch_search(ctx, ctx->stream2, first + off, cnt, nn, nbr, &launches);
        if (rc != GSK_OK) return rc;
        GSK_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_search[b], ctx->stream2));
        GSK_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_search[b], 0));
        rc = gsk_launch_local_solve(ctx, ctx->stream, first + off, cnt, nn, nbr, off, &launches);
        if (rc != GSK_OK) return rc;
        GSK_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_solve[b], ctx->stream));
        continue;
      }
      if (phase_timing) cudaEventRecord(ctx->ev[3], ctx->stream);
      rc = gsk_launch_search(ctx, ctx->stream, first + off, cnt, nn, nbr, &launches);
      if (rc != GSK_OK) return rc;
      if (phase_timing) cudaEventRecord(ctx->ev[4], ctx->stream);
      rc = gsk_launch_local_solve(ctx, ctx->stream, first + off, cnt, nn, nbr, off, &launches);
      if (rc != GSK_OK) return rc;
      if (phase_timing) {
        cudaEventRecord(ctx->ev[5], ctx->stream);
        cudaEventSynchronize(ctx->ev[5]);
        float a = 0.f, b2 = 0.f;
        cudaEventElapsedTime(&a, ctx->ev[3], ctx->ev[4]);
        cudaEventElapsedTime(&b2, ctx->ev[4], ctx->ev[5]);
        ms_search += a;
        ms_solve += b2;
      }
    }
  }
  GSK_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  ctx->timing.ms_search = ms_search;
  ctx->timing.ms_solve = ms_solve;
  ctx->timing.launches = launches;
  ctx->timing.targets = count;
  ctx->timing_pending = true;
  return GSK_OK;
}

extern "C" GSK_API int gsk_set_phase_timing(gsk_ctx *ctx, int on) {
  if (!ctx) return GSK_ERR_INVALID;
  ctx->phase_timing = on != 0;
  return GSK_OK;
}

extern "C" GSK_API int gsk_get_timing(const gsk_ctx *cctx, gsk_timing *out) {
  gsk_ctx *ctx = const_cast<gsk_ctx *>(cctx);
  if (!ctx || !out) return GSK_ERR_INVALID;
  if (ctx->timing_pending) {
    cudaSetDevice(ctx->device);
    if (cudaEventSynchronize(ctx->ev[1]) == cudaSuccess) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[1]) == cudaSuccess) ctx->timing.ms_total = ms;
    }
    ctx->timing_pending = false;
  }
  *out = ctx->timing;
  return GSK_OK;
}

// ---------------------------------------------------------------------------------------------
// one-shot host-buffer call: plan + execute + copies
// ---------------------------------------------------------------------------------------------
extern "C" GSK_API int gsk_krige(gsk_ctx *ctx, const gsk_problem *p, double *mean_out, double *var_out, int32_t *nneigh_out,
                         int32_t *neigh_idx_out) try {
  if (!ctx) return GSK_ERR_INVALID;
  if (!mean_out || !var_out) return fail(ctx, GSK_ERR_INVALID, "mean_out / var_out are NULL");
  int rc = gsk_plan(ctx, p);
  if (rc != GSK_OK) return rc;
  const int64_t T = ctx->n_targets;
  const int64_t first = p->target_first;
  const int64_t count = p->target_count < 0 ? T - first : p->target_count;
  if (count == 0) return GSK_OK;
  const int k = p->max_neighbors;
  rc = ensure(ctx, (void **)&ctx->d_mean, &ctx->cap_out, sizeof(double) * 2 * (size_t)count);
  if (rc != GSK_OK) return rc;
  double *d_mean = ctx->d_mean, *d_var = ctx->d_mean + count;
  int *d_nn = nullptr, *d_idx = nullptr;
  if (nneigh_out) {
    rc = ensure(ctx, (void **)&ctx->d_nn, &ctx->cap_nn, sizeof(int) * (size_t)count);
    if (rc != GSK_OK) return rc;
    d_nn = ctx->d_nn;
  }
  if (neigh_idx_out && k > 0) {
    rc = ensure(ctx, (void **)&ctx->d_nbr, &ctx->cap_nbr, sizeof(int) * (size_t)count * k);
    if (rc != GSK_OK) return rc;
    d_idx = ctx->d_nbr;
  }
  // The slab is computed in a few pieces; the device→host copies of a finished piece run on the side stream
  // while the next piece is being computed (they overlap only when the host buffers are page-locked).
  {
    // piece boundaries: 40 / 30 / 20 / 10 % of the slab (rounded to whole rows or planes of a grid) — only the
    // last piece's copies are exposed after the compute has finished, so it is the smallest
    std::vector<long long> bounds{0, (long long)count};
    if (count >= (1ll << 19)) {
      long long unit = 128;
      if (ctx->tg.is_grid) {
        const int dim = ctx->tg.dim;
        unit = (dim == 3) ? ctx->tg.gdim[0] * ctx->tg.gdim[1] * 4 : (dim == 2 ? ctx->tg.gdim[0] * 8 : 128);
      }
      bounds.clear();
      bounds.push_back(0);
      const double cum[3] = {0.4, 0.7, 0.9};
      for (int i = 0; i < 3; ++i) {
        long long b = (long long)(cum[i] * (double)count);
        if (unit * 8 <= count) b = (b + unit - 1) / unit * unit;
        b = std::min<long long>(b, count);
        if (b > bounds.back()) bounds.push_back(b);
      }
      if (count > bounds.back()) bounds.push_back(count);
    }
    for (size_t pi = 0; pi + 1 < bounds.size(); ++pi) {
      const long long off = bounds[pi], cnt = bounds[pi + 1] - bounds[pi];
      rc = gsk_execute(ctx, first + off, cnt, d_mean + off, d_var + off, d_nn ? d_nn + off : nullptr,
                       d_idx ? d_idx + off * k : nullptr);
      if (rc != GSK_OK) return rc;
      const int eb = (int)(pi & 1);
      GSK_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_search[eb], ctx->stream));  // reused as a "piece done" marker
      GSK_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_search[eb], 0));
      GSK_CUDA_CHECK(ctx, cudaMemcpyAsync(mean_out + off, d_mean + off, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost, ctx->stream2));
      GSK_CUDA_CHECK(ctx, cudaMemcpyAsync(var_out + off, d_var + off, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost, ctx->stream2));
      if (d_nn)
        GSK_CUDA_CHECK(ctx, cudaMemcpyAsync(nneigh_out + off, d_nn + off, sizeof(int) * (size_t)cnt, cudaMemcpyDeviceToHost, ctx->stream2));
      if (d_idx)
        GSK_CUDA_CHECK(ctx, cudaMemcpyAsync(neigh_idx_out + off * k, d_idx + off * k, sizeof(int) * (size_t)cnt * k, cudaMemcpyDeviceToHost, ctx->stream2));
    }
    ctx->timing.targets = count;
  }
  GSK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream2));
  GSK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  GSK_CUDA_CHECK(ctx, cudaGetLastError());
  return GSK_OK;
} catch (const std::bad_alloc &) {  // no exception may cross the C ABI
  return fail(ctx, GSK_ERR_NOMEM, "gsk_krige: out of host memory");
} catch (const std::exception &e) {
  return fail(ctx, GSK_ERR_STATE, std::string("gsk_krige: ") + e.what());
}

extern "C" GSK_API int gsk_measure_fp64_peak(gsk_ctx *ctx, double *dfma_tflops, double *dmma_tflops) {
  if (!ctx || !dfma_tflops || !dmma_tflops) return GSK_ERR_INVALID;
  GSK_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  return gsk_peak_measure(ctx, dfma_tflops, dmma_tflops);
}

// ---------------------------------------------------------------------------------------------
// single-process multi-GPU: one host thread + one cached context per piece
// ---------------------------------------------------------------------------------------------
namespace {
std::mutex g_multi_mutex;
std::vector<gsk_ctx *> g_multi_ctx;   // slot i serves piece i (re-created if the device id changes)
}  // namespace

extern "C" GSK_API int gsk_krige_multi(const int *device_ids, int n_devices, const gsk_problem *p, double *mean_out,
                                       double *var_out, int32_t *nneigh_out, int32_t *neigh_idx_out, char *errbuf,
                                       int errbuf_len) try {
  auto set_err = [&](const std::string &m) {
    if (errbuf && errbuf_len > 0) {
      strncpy(errbuf, m.c_str(), (size_t)errbuf_len - 1);
      errbuf[errbuf_len - 1] = 0;
    }
  };
  if (!device_ids || n_devices < 1 || n_devices > 64 || !p || !mean_out || !var_out) {
    set_err("gsk_krige_multi: invalid arguments");
    return GSK_ERR_INVALID;
  }
  std::lock_guard<std::mutex> lock(g_multi_mutex);  // one multi-GPU call at a time per process
  if ((int)g_multi_ctx.size() < n_devices) g_multi_ctx.resize(n_devices, nullptr);
  for (int i = 0; i < n_devices; ++i) {
    if (g_multi_ctx[i] && g_multi_ctx[i]->device != device_ids[i]) {
      gsk_destroy(g_multi_ctx[i]);
      g_multi_ctx[i] = nullptr;
    }
    if (!g_multi_ctx[i]) {
      int rc = gsk_create(&g_multi_ctx[i], device_ids[i]);
      if (rc != GSK_OK) {
        set_err(gsk_last_error(nullptr));
        return rc;
      }
    }
  }
  const int64_t T = gsk_num_targets(p);
  const int64_t first = p->target_first;
  const int64_t count = p->target_count < 0 ? T - first : p->target_count;
  if (first < 0 || count < 0 || first + count > T) {
    set_err("gsk_krige_multi: target slab out of range");
    return GSK_ERR_INVALID;
  }
  const int k = p->max_neighbors;
  std::vector<int> rcs(n_devices, GSK_OK);
  std::vector<std::thread> workers;
  for (int i = 0; i < n_devices; ++i) {
    workers.emplace_back([&, i]() {
      const int64_t lo = count * i / n_devices, hi = count * (i + 1) / n_devices;
      gsk_problem pi = *p;
      pi.target_first = first + lo;
      pi.target_count = hi - lo;
      rcs[i] = gsk_krige(g_multi_ctx[i], &pi, mean_out + lo, var_out + lo, nneigh_out ? nneigh_out + lo : nullptr,
                         (neigh_idx_out && k > 0) ? neigh_idx_out + lo * k : nullptr);
    });
  }
  for (auto &w : workers) w.join();
  for (int i = 0; i < n_devices; ++i)
    if (rcs[i] != GSK_OK) {
      set_err(gsk_last_error(g_multi_ctx[i]));
      return rcs[i];
    }
  return GSK_OK;
} catch (const std::exception &e) {  // no exception may cross the C ABI (std::thread, allocations)
  if (errbuf && errbuf_len > 0) {
    strncpy(errbuf, e.what(), (size_t)errbuf_len - 1);
    errbuf[errbuf_len - 1] = 0;
  }
  return GSK_ERR_NOMEM;
}
