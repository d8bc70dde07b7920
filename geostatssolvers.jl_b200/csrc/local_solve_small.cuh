// local_solve_small.cuh — K3 specialised for k <= 20 neighbours with Simple / Ordinary Kriging (e <= 3
// extra rows): the configuration BASELINE.json's metric is quoted on (C2: k = 20, OK). Same
// formulation as local_solve.cuh (augmented in-place Cholesky + Schur/Gram algebra), but
//   * 4 lanes per target, 5 register slots of neighbour rows, panels of 4 columns, all loops unrolled:
//     every shared-memory address is base + immediate and every register index is static;
//   * the extra rows (b, z, ones) never touch shared memory: lane l keeps extra row l in registers
//     (yreg[20]); they are updated alongside the neighbour rows and their Gram products are formed with
//     shuffles — shared memory per target is just the tightly packed factor (212 doubles = 1696 B), so 512
//     threads (128 targets, two 256-thread CTAs) are resident per SM, the limit the 128 registers set too;
//   * pivot column coordinates, pivots and panel rows all travel by warp shuffles inside the 4-lane group.
#pragma once
#include "local_solve.cuh"

namespace gsk_local {

constexpr int SK_KMAX = 20;   // neighbour columns (multiple of 4)
constexpr int SK_R = SK_KMAX / 4;
constexpr int SK_STOR = col_off<SK_KMAX, 1>(SK_KMAX);  // tightly packed factor: column p keeps rows >= p
constexpr int sk_pad(int g) { return ((g & 15) == 4 || (g & 15) == 12) ? g : sk_pad(g + 1); }
constexpr int SK_GSZ = sk_pad(SK_STOR);  // ≡ 4 or 12 (mod 16) doubles: the groups of a warp spread over the banks

// NTH threads per CTA (NTH/4 targets); 512 threads per SM either way — 256 measured 1 % ahead of 64/128, 512 8 % behind
// FULL: k = 20, the column count is a compile-time constant and the per-panel / per-column guards fold away
// NUG0: the variogram has no nugget (sill == cs; never true for the Gaussian model, which carries its 1e-6)
template <int DIM, int VK, bool FULL, bool NUG0 = false, int NTH = 256>
__global__ void __launch_bounds__(NTH, 512 / NTH) local_solve_small_kernel(const GskLocalArgs a, const int KC_arg) {
  const int KC = FULL ? SK_KMAX : KC_arg;
  constexpr int G = 4, R = SK_R, W = 4, RT = SK_KMAX, A = 1, KM = SK_KMAX;
  constexpr int TPC = NTH / 4;  // targets per CTA
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  double *sup = sm;
  const int nsup_pad = a.sup_smem ? ((3 * a.nsup + 3) & ~3) : 0;
  const int tid = threadIdx.x, lane = tid & 31, l = lane & 3, grp = tid >> 2, gbase = lane & ~3;
  const long long t = (long long)blockIdx.x * TPC + grp;
  const bool live = t < a.count;
  // issue the neighbour-index loads first: their latency (and that of the dependent record loads below) then
  // overlaps the support staging, the barrier and the centroid arithmetic
  int nidx[SK_R];
#pragma unroll
  for (int jj = 0; jj < SK_R; ++jj) {
    const int j = jj * 4 + l;
    nidx[jj] = (live && j < a.k) ? a.nbr[t * a.k + j] : -1;
  }
  const int nn_in = live ? a.nn[t] : 0;
  double4 nrec[SK_R];
#pragma unroll
  for (int jj = 0; jj < SK_R; ++jj) {
    nrec[jj] = make_double4(0.0, 0.0, 0.0, 0.0);
    if (nidx[jj] >= 0) nrec[jj] = a.rec_orig[nidx[jj]];
  }
  // spherical model: lengths in units of the range, so that d² is the polynomial's argument (one DMUL less per
  // evaluation; no drift terms here that would need the coordinates themselves)
  constexpr bool UNIT = (VK == GSK_VARIO_SPHERICAL);
  const double cscale = UNIT ? a.vg.inv_r : 1.0;
  if (a.sup_smem)
    for (int i = tid; i < 3 * a.nsup; i += NTH) sup[i] = UNIT ? a.sup[i] * cscale : a.sup[i];
  __syncthreads();

  double *res = sm + nsup_pad;                      // [2][TPC]: the CTA's results, staged for coalesced stores
  double *S = sm + nsup_pad + 2 * TPC + (size_t)grp * SK_GSZ;
  double *Sl = S + l;
  const GskVario vg = a.vg;

  // ---- target centroid ----
  double tc[3] = {0.0, 0.0, 0.0};
  int nn = 0;
  if (live) {
    const long long lin = a.first + t;
    if (a.tg.is_grid) {
      if (a.tg.gdim[0] * a.tg.gdim[1] * a.tg.gdim[2] < 0x7fffffffLL) {
        unsigned rem = (unsigned)lin;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          if (d < a.tg.dim) {
            const unsigned gd = (unsigned)a.tg.gdim[d];
            const unsigned qd = rem / gd;
            tc[d] = gsk_cell_center(a.tg.gorg[d], a.tg.gsp[d], (long long)(rem - qd * gd));
            rem = qd;
          }
        }
      } else {
        long long rem = lin;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          if (d < a.tg.dim) {
            long long c = rem % a.tg.gdim[d];
            rem /= a.tg.gdim[d];
            tc[d] = gsk_cell_center(a.tg.gorg[d], a.tg.gsp[d], c);
          }
        }
      }
    } else {
      for (int d = 0; d < a.tg.dim; ++d) tc[d] = a.tg.pts[d][lin];
    }
    nn = nn_in;
  }
  const bool estimate = live && nn >= a.min_neighbors && nn > 0;
  if (!estimate) nn = 0;

  // ---- phase 1: gather my neighbours j = 4·jj + l ----
  double nx[R], ny[R], nz[R], nv[R], bacc[R];
#pragma unroll
  for (int jj = 0; jj < R; ++jj) {
    const double4 rc = nrec[jj];
    // (relative to the target centroid before scaling: scaling absolute coordinates would lose digits in the
    // differences taken below)
    nx[jj] = UNIT ? (rc.x - tc[0]) * cscale : rc.x;
    ny[jj] = UNIT ? (rc.y - tc[1]) * cscale : rc.y;
    nz[jj] = UNIT ? (rc.z - tc[2]) * cscale : rc.z;
    nv[jj] = (a.es.kind == GSK_EST_SIMPLE) ? rc.w - a.es.sk_mean : rc.w;
    // Unused slots (j >= nn: k not a multiple of 4, a ball that kept fewer, or a target that is not estimated)
    // become samples of value 0 parked far away, each at its own place: every covariance with them evaluates to
    // zero (1e-304 for the exp models) and their diagonal to C(0), so the rows decouple by themselves and neither
    // the evaluation loops nor the fill need a validity select. 1e100 dwarfs any coordinate and its square is finite.
    const int j = jj * G + l;
    if (j >= nn) {
      nx[jj] = 1e100 * (double)(j + 1);
      ny[jj] = 0.0;
      nz[jj] = 0.0;
      nv[jj] = 0.0;
    }
    bacc[jj] = 0.0;
  }
  // ---- phase 2: block-support RHS, q outermost (5 independent chains per lane) ----
  if (UNIT) tc[0] = tc[1] = tc[2] = 0.0;  // the centroid is the local origin
  if (!a.sup_smem) {
    // supports above GSK_MAX_SUPPORT points: read from global memory (uniform addresses), in the same frame
    rhs_block_support<VK, DIM, R, UNIT, NUG0>(a, vg, UNIT ? a.sup_unit : a.sup, tc, nx, ny, nz, bacc);
  } else if (VK == GSK_VARIO_SPHERICAL) {
    // warp-uniform fast path: every support point of every neighbour of the warp's targets lies inside the range
    // (the coordinates are centroid-relative in range units here), so the range select is not needed
    bool inr = true;
#pragma unroll
    for (int jj = 0; jj < R; ++jj) {
      double d2c = fma(ny[jj], ny[jj], nx[jj] * nx[jj]);
      if (DIM == 3) d2c = fma(nz[jj], nz[jj], d2c);
      inr = inr && (d2c < a.rhs_inr_lim2);
    }
    if (__all_sync(0xffffffffu, inr)) rhs_block_support<VK, DIM, R, UNIT, NUG0, true>(a, vg, sup, tc, nx, ny, nz, bacc);
    else rhs_block_support<VK, DIM, R, UNIT, NUG0, false>(a, vg, sup, tc, nx, ny, nz, bacc);
  } else {
    rhs_block_support<VK, DIM, R, UNIT, NUG0>(a, vg, sup, tc, nx, ny, nz, bacc);
  }
  // ---- phase 3: extra rows into registers: lane 0 ← b, lane 1 ← z, lane 2 ← ones (OK), lane 3 ← (unused).
  //      Transposed through shared memory — the factor storage S is not written before phase 4: every lane stores
  //      its 5 entries of the rows b, z, ones; lane l then reads row min(l, 2) with 10 LDS.128. (The shuffle
  //      version cost 80 SHFL and as many selects per warp.) ----
  double yreg[KM];
  {
    const double inv_q = 1.0 / (double)a.nsup;
    const bool ok_row = a.es.kind != GSK_EST_SIMPLE;
#pragma unroll
    for (int jj = 0; jj < R; ++jj) {
      const int j = jj * G + l;
      Sl[0 * KM + jj * G] = bacc[jj] * inv_q;            // zero in the unused slots by construction
      Sl[1 * KM + jj * G] = nv[jj];                      // ditto
      Sl[2 * KM + jj * G] = (ok_row && j < nn) ? 1.0 : 0.0;
    }
    __syncwarp();
    const double2 *row = reinterpret_cast<const double2 *>(S + (l < 2 ? l : 2) * KM);
#pragma unroll
    for (int i = 0; i < KM / 2; ++i) {
      const double2 v = row[i];
      yreg[2 * i] = v.x;
      yreg[2 * i + 1] = v.y;
    }
    __syncwarp();  // phase 4 overwrites S
  }
  // ---- phase 4: covariance block into shared memory, column p rows >= p & ~3 (all static) ----
#pragma unroll
  for (int p = 0; p < KM; ++p) {
    if (p < KC) {
      const int sb = p & ~3, rlo = sb / G;
      const double xp = __shfl_sync(0xffffffffu, nx[p / G], gbase + (p % G));
      const double yp = __shfl_sync(0xffffffffu, ny[p / G], gbase + (p % G));
      const double zp = (DIM == 3) ? __shfl_sync(0xffffffffu, nz[p / G], gbase + (p % G)) : 0.0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r >= rlo) {
          const int i = r * G + l;
          const double dx = nx[r] - xp, dy = ny[r] - yp;
          double d2 = fma(dy, dy, dx * dx);
          if (DIM == 3) {
            const double dz = nz[r] - zp;
            d2 = fma(dz, dz, d2);
          }
          // no selects: the diagonal is d2 == 0 → C(0) = sill, unused slots evaluate to zero by construction,
          // and rows above the diagonal are neither stored here nor read by the factorisation
          const double v = cov_fast<VK, UNIT, NUG0>(vg, d2);
          if (r * G >= p || i >= p) Sl[col_off<RT, A>(p) - p + r * G] = v;
        }
      }
    }
  }
  __syncwarp();

  // ---- phase 5: blocked in-place Cholesky; extra rows ride along in registers ----
  double acc[R][W], accx[W];
#pragma unroll
  for (int c0 = 0; c0 < KM; c0 += W) {
    if (c0 < KC) {
      const int rmin = c0 / G;
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r >= rmin)
#pragma unroll
          for (int jj = 0; jj < W; ++jj)
            acc[r][jj] = (r * G >= c0 + jj || r * G + l >= c0 + jj) ? Sl[col_off<RT, A>(c0 + jj) - (c0 + jj) + r * G] : 0.0;
#pragma unroll
      for (int jj = 0; jj < W; ++jj) accx[jj] = yreg[c0 + jj];
      // left-looking update from the finished columns
#pragma unroll
      for (int p = 0; p < c0; ++p) {
        const double *col = S + col_off<RT, A>(p) - p;
        const double piv[W] = {col[c0], col[c0 + 1], col[c0 + 2], col[c0 + 3]};
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r >= rmin) {
            const double own = col[r * G + l];
#pragma unroll
            for (int jj = 0; jj < W; ++jj) acc[r][jj] = fma(-own, piv[jj], acc[r][jj]);
          }
        }
#pragma unroll
        for (int jj = 0; jj < W; ++jj) accx[jj] = fma(-yreg[p], piv[jj], accx[jj]);
      }
      // the 4×4 panel
#pragma unroll
      for (int jj = 0; jj < W; ++jj) {
        const int j = c0 + jj;
        const double d = __shfl_sync(0xffffffffu, acc[j / G][jj], gbase + (j % G));
        const double rinv = gsk_rsqrt(d);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r >= rmin) {
            acc[r][jj] *= rinv;
            if (r * G >= j || r * G + l >= j) Sl[col_off<RT, A>(j) - j + r * G] = acc[r][jj];
          }
        }
        accx[jj] *= rinv;
        yreg[j] = accx[jj];
#pragma unroll
        for (int j2 = jj + 1; j2 < W; ++j2) {
          const double lj = __shfl_sync(0xffffffffu, acc[(c0 + j2) / G][jj], gbase + ((c0 + j2) % G));
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (r >= rmin) acc[r][j2] = fma(-acc[r][jj], lj, acc[r][j2]);
          accx[j2] = fma(-accx[jj], lj, accx[j2]);
        }
      }
      __syncwarp();
    }
  }

  // ---- phase 6: Gram products of the extra rows by shuffles. Lane l forms <row l, row l> and
  //      <row l, row (l+1)%3>: lane 0 → bb, bz; lane 1 → zz, zf; lane 2 → ff, fb ----
  double g_self = 0.0, g_next = 0.0;
  {
    const int src = gbase + ((l + 1) % 3);
#pragma unroll
    for (int p = 0; p < KM; ++p) {
      const double o = __shfl_sync(0xffffffffu, yreg[p], src);
      g_self = fma(yreg[p], yreg[p], g_self);
      g_next = fma(yreg[p], o, g_next);
    }
  }
  const double gbb = __shfl_sync(0xffffffffu, g_self, gbase + 0);
  const double gbz = __shfl_sync(0xffffffffu, g_next, gbase + 0);
  const double gzf = __shfl_sync(0xffffffffu, g_next, gbase + 1);
  const double gff = __shfl_sync(0xffffffffu, g_self, gbase + 2);
  const double gfb = __shfl_sync(0xffffffffu, g_next, gbase + 2);

  if (l == 0 && live) {
    double mean = NAN, var = NAN;
    if (estimate) {
      if (a.es.kind == GSK_EST_SIMPLE) {
        mean = a.es.sk_mean + gbz;
        var = vg.sill - gbb;
      } else {
        const double nu = (gfb - 1.0) / gff;
        mean = gbz - gzf * nu;
        var = vg.sill - (gbb - gfb * nu + nu);
      }
      if (a.flags & GSK_FLAG_CLAMP_VARIANCE) var = (var > 0.0 || var != var) ? var : 0.0;
      if (a.flags & GSK_FLAG_SQRT_ROUNDTRIP) { double sd = sqrt(var); var = sd * sd; }
    }
    res[grp] = mean;
    res[TPC + grp] = var;
  }
  // coalesced result stores: thread i writes the mean of the CTA's i-th target, thread TPC + i its variance
  __syncthreads();
  if (tid < 2 * TPC) {
    const int i = tid % TPC, field = tid / TPC;
    const long long ti = (long long)blockIdx.x * TPC + i;
    if (ti < a.count) gsk_store_field(a.out, field, ti, res[tid]);
  }
}

template <int DIM, int VK, bool FULL, bool NUG0 = false, int NTH = 256>
inline cudaError_t launch_small_full(const GskLocalArgs &a, cudaStream_t st) {
  const int KC = (a.k + 3) / 4 * 4;
  GskLocalArgs b = a;
  b.sup_smem = (a.nsup <= GSK_MAX_SUPPORT) ? 1 : 0;
  const int nsup_pad = b.sup_smem ? ((3 * a.nsup + 3) & ~3) : 0;
  const size_t smem = sizeof(double) * ((size_t)nsup_pad + 2 * (NTH / 4) + (NTH / 4) * (size_t)SK_GSZ);
  auto kern = local_solve_small_kernel<DIM, VK, FULL, NUG0, NTH>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  const unsigned grid = (unsigned)((a.count + NTH / 4 - 1) / (NTH / 4));
  kern<<<grid, NTH, smem, st>>>(b, KC);
  return cudaGetLastError();
}

template <int DIM, int VK>
inline cudaError_t launch_small_one(const GskLocalArgs &a, cudaStream_t st) {
  if ((a.k + 3) / 4 * 4 != SK_KMAX) return launch_small_full<DIM, VK, false>(a, st);
  if (VK != GSK_VARIO_GAUSSIAN && a.vg.sill == a.vg.cs) return launch_small_full<DIM, VK, true, VK != GSK_VARIO_GAUSSIAN>(a, st);
  return launch_small_full<DIM, VK, true>(a, st);
}

}  // namespace gsk_local

// k <= 20, Simple / Ordinary Kriging (UK degree 0 is OK): the small-system fast path
cudaError_t gsk_local_launch_small(const GskLocalArgs &a, cudaStream_t st);
