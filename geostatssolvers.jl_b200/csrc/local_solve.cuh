// local_solve.cuh — K3: per-target kriging system assembly + factorisation + solve + mean/variance,
// replacing `GeoStatsModels.fit(estimator, samples)` and `predictprob(krig, var, pdomain[ind])`
// inside approxsolve's loop (ref: src/estimation/krig.jl:217-226).
//
// Formulation (FP64 throughout). For k neighbours with covariance block C (k×k, SPD) and the
// "extra" rows E = [b; z; f_1 … f_c] (b = block-support RHS, z = values (− μ for SK), f = drift
// monomials: OK c=1 → ones, UK → UKexps order), one Cholesky sweep over the augmented matrix
//        [ C  ]           [ L ]
//        [ E  ]    →      [ Y ]      with  Y = E L^-T  (forward substitution comes for free),
// and the Schur complement  Gm = Y Yᵀ = E C⁻¹ Eᵀ  (e×e) holds every quantity kriging needs:
//   SK :  mean = μ + Gm[b,z]                     var = sill − Gm[b,b]
//   OK/UK:  ν = Gm[f,f]⁻¹ (Gm[f,b] − f₀)          (Lagrange multipliers; f₀ = drift at the target)
//           mean = Gm[b,z] − Gm[f,z]·ν            var = sill − (Gm[b,b] − Gm[f,b]·ν + f₀·ν)
// which is the reference's  s = LHS \ RHS;  μ̂ = Σλᵢzᵢ;  σ² = sill − RHS·s  solved by block
// elimination (C first, then the c×c Schur system) — no back substitution, no pivot search.
//
// Mapping: G lanes cooperate on one target (32/G targets per warp). Lane l owns rows l, l+G, …
// (R register slots); the factor lives packed in shared memory (column p keeps rows ≥ p&~3) and
// is built in place, W columns at a time: left-looking update from finished columns (own entries
// + W broadcast entries per finished column → R·W independent DFMAs), then the W×W panel.
#pragma once
#include <math.h>

#include "gsk_internal.cuh"

namespace gsk_local {

constexpr int CTA_THREADS = 128;

struct Layout {
  int KC, EP, RT;    // padded neighbour columns, padded extra rows, total rows
  int e;             // live extra rows
  int stor;          // doubles of packed factor storage per target
  int gsz;           // doubles per target group (all per-target shared memory)
  int off_nb, off_v, off_b, off_s, off_gm, off_piv;
};

__host__ __device__ inline int col_start_row(int p) { return p & ~3; }

template <int DIM>
__host__ inline Layout make_layout(int k, int e, int W) {
  Layout L;
  L.KC = (k + W - 1) / W * W;
  L.EP = (e + W - 1) / W * W;
  L.RT = L.KC + L.EP;
  L.e = e;
  int stor = 0;
  for (int p = 0; p < L.KC; ++p) stor += L.RT - col_start_row(p);
  L.stor = stor;
  int o = 0;
  L.off_nb = o; o += DIM * L.KC;
  L.off_v = o;  o += L.KC;
  L.off_b = o;  o += L.KC;
  L.off_s = o;  o += stor;       // all of the above are multiples of 4 doubles → 32-B aligned
  L.off_gm = o; o += L.EP * L.EP;
  L.off_piv = o; o += 4;
  // spread the groups of a warp over the banks: make the stride ≡ 4 (mod 8) doubles
  while ((o & 7) != 4) o += 4;
  L.gsz = o;
  return L;
}

template <int G, int R, int W, int DIM, int VK>
__global__ void __launch_bounds__(CTA_THREADS) local_solve_kernel(const GskLocalArgs a, const Layout L) {
  constexpr int TPW = 32 / G;                   // targets per warp
  constexpr int TPC = TPW * (CTA_THREADS / 32);  // targets per CTA
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  // CTA-wide tables
  double *sup = sm;                                     // [3][nsup]
  int nsup_pad = (3 * a.nsup + 3) & ~3;
  int *colOff = reinterpret_cast<int *>(sm + nsup_pad);  // [KC]
  int coff_pad = ((L.KC + 1) / 2 + 3) & ~3;             // in doubles
  double *groups = sm + nsup_pad + coff_pad;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int l = lane % G;                 // lane within the target group
  const int grp = tid / G;                // group within the CTA
  const int KC = L.KC, RT = L.RT, EP = L.EP, e = L.e;

  for (int i = tid; i < 3 * a.nsup; i += CTA_THREADS) sup[i] = a.sup[i];
  if (tid == 0) {
    int o = 0;
    for (int p = 0; p < KC; ++p) { colOff[p] = o; o += RT - col_start_row(p); }
  }
  __syncthreads();

  double *gs = groups + (size_t)grp * L.gsz;
  double *nbX = gs + L.off_nb;
  double *nbY = nbX + KC;
  double *nbZ = (DIM == 3) ? nbY + KC : nbY;
  double *V = gs + L.off_v;
  double *B = gs + L.off_b;
  double *S = gs + L.off_s;
  double *GM = gs + L.off_gm;
  double *PIV = gs + L.off_piv;

  const long long t = (long long)blockIdx.x * TPC + grp;  // slab-local target
  const bool live = t < a.count;
  const GskVario vg = a.vg;

  // ---- target centroid ----
  double tc[3] = {0.0, 0.0, 0.0};
  int nn = 0;
  if (live) {
    long long lin = a.first + t;
    if (a.tg.is_grid) {
      long long rem = lin;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (d < a.tg.dim) {
          long long c = rem % a.tg.gdim[d];
          rem /= a.tg.gdim[d];
          tc[d] = gsk_cell_center(a.tg.gorg[d], a.tg.gsp[d], c);
        }
      }
    } else {
      for (int d = 0; d < a.tg.dim; ++d) tc[d] = a.tg.pts[d][lin];
    }
    nn = a.nn[t];
  }
  const bool estimate = live && nn >= a.min_neighbors && nn > 0;
  if (!estimate) nn = 0;

  // ---- phase 1: gather neighbours (coordinates + value in one 32-byte record) ----
  for (int j = l; j < KC; j += G) {
    double4 rc = make_double4(0.0, 0.0, 0.0, 0.0);
    if (j < nn) {
      int idx = a.nbr[t * a.k + j];
      rc = a.rec_orig[idx];
    }
    nbX[j] = rc.x;
    nbY[j] = rc.y;
    if (DIM == 3) nbZ[j] = rc.z;
    V[j] = (j < nn) ? ((a.es.kind == GSK_EST_SIMPLE) ? rc.w - a.es.sk_mean : rc.w) : 0.0;
  }
  __syncwarp();

  // ---- phase 2: block-support right-hand side  b_j = mean_q C(‖t + δ_q − x_j‖) ----
  {
    const double inv_q = 1.0 / (double)a.nsup;
    for (int j = l; j < KC; j += G) {
      double acc = 0.0;
      if (j < nn) {
        const double xj = nbX[j], yj = nbY[j], zj = (DIM == 3) ? nbZ[j] : 0.0;
        for (int q = 0; q < a.nsup; ++q) {
          double dx = (tc[0] + sup[q]) - xj;
          double dy = (tc[1] + sup[a.nsup + q]) - yj;
          double d2 = fma(dy, dy, dx * dx);
          if (DIM == 3) {
            double dz = (tc[2] + sup[2 * a.nsup + q]) - zj;
            d2 = fma(dz, dz, d2);
          }
          acc += gsk_cov<VK>(vg, d2);
        }
      }
      B[j] = acc * inv_q;
    }
  }
  __syncwarp();

  // ---- phase 3: fill the augmented matrix in place (column p, rows >= p&~3) ----
  for (int p = 0; p < KC; ++p) {
    const bool valid_p = p < nn;
    const double xp = nbX[p], yp = nbY[p], zp = (DIM == 3) ? nbZ[p] : 0.0;
    const int sp = col_start_row(p);
    double *col = S + colOff[p] - sp;
    for (int i = sp + l; i < RT; i += G) {
      double v = 0.0;
      if (i < KC) {
        if (i == p) {
          v = valid_p ? vg.sill : 1.0;
        } else if (valid_p && i < nn) {
          double dx = nbX[i] - xp, dy = nbY[i] - yp;
          double d2 = fma(dy, dy, dx * dx);
          if (DIM == 3) {
            double dz = nbZ[i] - zp;
            d2 = fma(dz, dz, d2);
          }
          v = gsk_cov<VK>(vg, d2);
        }
      } else if (valid_p) {
        const int r = i - KC;
        if (r == 0) v = B[p];
        else if (r == 1) v = V[p];
        else if (r < e) {
          if (a.es.kind == GSK_EST_ORDINARY) v = 1.0;
          else {
            const int *ex = a.es.exps[r - 2];
            v = gsk_ipow(xp, ex[0]) * gsk_ipow(yp, ex[1]);
            if (DIM == 3) v *= gsk_ipow(zp, ex[2]);
          }
        }
      }
      col[i] = v;
    }
  }
  __syncwarp();

  // ---- phase 4: blocked in-place Cholesky of the augmented matrix ----
  double acc[R][W];
  for (int c0 = 0; c0 < KC; c0 += W) {
    const int rmin = c0 / G;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = r * G + l;
#pragma unroll
      for (int jj = 0; jj < W; ++jj) {
        const int j = c0 + jj;
        const int sj = col_start_row(j);
        acc[r][jj] = (r >= rmin && i >= sj && i < RT) ? S[colOff[j] - sj + i] : 0.0;
      }
    }
    // left-looking update from the finished columns p < c0
    for (int p = 0; p < c0; ++p) {
      const double *col = S + colOff[p] - col_start_row(p);
      double piv[W];
#pragma unroll
      for (int jj = 0; jj < W; jj += 2) {
        double2 t2 = *reinterpret_cast<const double2 *>(col + c0 + jj);
        piv[jj] = t2.x;
        piv[jj + 1] = t2.y;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int i = r * G + l;
        if (r >= rmin) {
          const double own = (i < RT) ? col[i] : 0.0;
#pragma unroll
          for (int jj = 0; jj < W; ++jj) acc[r][jj] = fma(-own, piv[jj], acc[r][jj]);
        }
      }
    }
    // the W×W panel
#pragma unroll
    for (int jj = 0; jj < W; ++jj) {
      const int j = c0 + jj;
      const int sj = col_start_row(j);
      double *colj = S + colOff[j] - sj;
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r >= rmin && r * G + l == j) PIV[0] = acc[r][jj];
      __syncwarp();
      const double rinv = rsqrt(PIV[0]);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int i = r * G + l;
        if (r >= rmin) {
          acc[r][jj] *= rinv;
          if (i >= j && i < RT) colj[i] = acc[r][jj];
        }
      }
      __syncwarp();
#pragma unroll
      for (int j2 = jj + 1; j2 < W; ++j2) {
        const double lj = colj[c0 + j2];
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (r >= rmin) acc[r][j2] = fma(-acc[r][jj], lj, acc[r][j2]);
      }
    }
  }

  // ---- phase 5: Schur complement of the extra rows: Gm = Y Yᵀ ----
  {
    const int rmin = KC / G;
    for (int cc = 0; cc < EP; cc += W) {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int jj = 0; jj < W; ++jj) acc[r][jj] = 0.0;
      for (int p = 0; p < KC; ++p) {
        const double *col = S + colOff[p] - col_start_row(p);
        double piv[W];
#pragma unroll
        for (int jj = 0; jj < W; jj += 2) {
          double2 t2 = *reinterpret_cast<const double2 *>(col + KC + cc + jj);
          piv[jj] = t2.x;
          piv[jj + 1] = t2.y;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = r * G + l;
          if (r >= rmin) {
            const double own = (i >= KC && i < RT) ? col[i] : 0.0;
#pragma unroll
            for (int jj = 0; jj < W; ++jj) acc[r][jj] = fma(own, piv[jj], acc[r][jj]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int i = r * G + l;
        if (r >= rmin && i >= KC && i < RT) {
#pragma unroll
          for (int jj = 0; jj < W; ++jj) GM[(i - KC) * EP + cc + jj] = acc[r][jj];
        }
      }
    }
  }
  __syncwarp();

  // ---- phase 6: e×e algebra, one lane per target ----
  if (l == 0 && live) {
    double mean = NAN, var = NAN;
    if (estimate) {
      const int c = a.es.nterms;
      const double gbb = GM[0], gbz = GM[1];
      if (c == 0) {
        mean = a.es.sk_mean + gbz;
        var = vg.sill - gbb;
      } else {
        // solve Gff ν = gfb − f0 (c×c SPD) by Cholesky in place on GM
        double f0[GSK_MAX_DRIFT_TERMS], nu[GSK_MAX_DRIFT_TERMS];
        for (int t2 = 0; t2 < c; ++t2) {
          if (a.es.kind == GSK_EST_ORDINARY) f0[t2] = 1.0;
          else {
            const int *ex = a.es.exps[t2];
            double m = gsk_ipow(tc[0], ex[0]) * gsk_ipow(tc[1], ex[1]);
            if (DIM == 3) m *= gsk_ipow(tc[2], ex[2]);
            f0[t2] = m;
          }
        }
#define GFF(i, j) GM[(2 + (i)) * EP + 2 + (j)]
        for (int j = 0; j < c; ++j) {
          double d = GFF(j, j);
          for (int p = 0; p < j; ++p) d -= GFF(j, p) * GFF(j, p);
          d = sqrt(d);
          GFF(j, j) = d;
          for (int i = j + 1; i < c; ++i) {
            double s = GFF(i, j);
            for (int p = 0; p < j; ++p) s -= GFF(i, p) * GFF(j, p);
            GFF(i, j) = s / d;
          }
        }
        for (int j = 0; j < c; ++j) {
          double s = GM[(2 + j) * EP + 0] - f0[j];
          for (int p = 0; p < j; ++p) s -= GFF(j, p) * nu[p];
          nu[j] = s / GFF(j, j);
        }
        for (int j = c - 1; j >= 0; --j) {
          double s = nu[j];
          for (int p = j + 1; p < c; ++p) s -= GFF(p, j) * nu[p];
          nu[j] = s / GFF(j, j);
        }
#undef GFF
        double mz = 0.0, mb = 0.0, mf = 0.0;
        for (int j = 0; j < c; ++j) {
          mz += GM[(2 + j) * EP + 1] * nu[j];
          mb += GM[(2 + j) * EP + 0] * nu[j];
          mf += f0[j] * nu[j];
        }
        mean = gbz - mz;
        var = vg.sill - (gbb - mb + mf);
      }
      if (a.flags & GSK_FLAG_CLAMP_VARIANCE) var = (var > 0.0 || var != var) ? var : 0.0;
      if (a.flags & GSK_FLAG_SQRT_ROUNDTRIP) { double sd = sqrt(var); var = sd * sd; }
    }
    a.mean[t] = mean;
    a.var[t] = var;
  }
}

template <int G, int R, int W, int DIM, int VK>
inline cudaError_t launch_one(const GskLocalArgs &a, int e, cudaStream_t st) {
  constexpr int TPW = 32 / G;
  constexpr int TPC = TPW * (CTA_THREADS / 32);
  Layout L = make_layout<DIM>(a.k, e, W);
  int nsup_pad = (3 * a.nsup + 3) & ~3;
  int coff_pad = ((L.KC + 1) / 2 + 3) & ~3;
  size_t smem = sizeof(double) * ((size_t)nsup_pad + coff_pad + (size_t)TPC * L.gsz);
  auto kern = local_solve_kernel<G, R, W, DIM, VK>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  unsigned grid = (unsigned)((a.count + TPC - 1) / TPC);
  kern<<<grid, CTA_THREADS, smem, st>>>(a, L);
  return cudaGetLastError();
}

template <int G, int R, int W>
inline cudaError_t launch_cfg(const GskLocalArgs &a, int e, cudaStream_t st) {
  const bool d3 = a.tg.dim == 3;
  switch (a.vg.kind) {
    case GSK_VARIO_GAUSSIAN:
      return d3 ? launch_one<G, R, W, 3, GSK_VARIO_GAUSSIAN>(a, e, st) : launch_one<G, R, W, 2, GSK_VARIO_GAUSSIAN>(a, e, st);
    case GSK_VARIO_SPHERICAL:
      return d3 ? launch_one<G, R, W, 3, GSK_VARIO_SPHERICAL>(a, e, st) : launch_one<G, R, W, 2, GSK_VARIO_SPHERICAL>(a, e, st);
    default:
      return d3 ? launch_one<G, R, W, 3, GSK_VARIO_EXPONENTIAL>(a, e, st) : launch_one<G, R, W, 2, GSK_VARIO_EXPONENTIAL>(a, e, st);
  }
}

}  // namespace gsk_local

// one translation unit per register/lanes configuration (compiled in parallel)
cudaError_t gsk_local_launch_A(const GskLocalArgs &a, int e, cudaStream_t st);  // G=4  R=3 W=4  (≤ 12 rows)
cudaError_t gsk_local_launch_B(const GskLocalArgs &a, int e, cudaStream_t st);  // G=4  R=6 W=4  (≤ 24 rows)
cudaError_t gsk_local_launch_C(const GskLocalArgs &a, int e, cudaStream_t st);  // G=8  R=5 W=8  (≤ 40 rows)
cudaError_t gsk_local_launch_D(const GskLocalArgs &a, int e, cudaStream_t st);  // G=16 R=5 W=8  (≤ 80 rows)
cudaError_t gsk_local_launch_E(const GskLocalArgs &a, int e, cudaStream_t st);  // G=32 R=4 W=8  (≤ 128 rows)
