/*
 * krige_c.c — a host written in plain C99 against include/gskrige.h, the same calls the Julia shim's
 * ccall makes (julia/GSKrige.jl). No Python, no torch: libgskrige.so is self-contained.
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude examples/krige_c.c \
 *       -Lgeostatssolvers.jl_b200/csrc -lgskrige -Wl,-rpath,$PWD/geostatssolvers.jl_b200/csrc -lm -o krige_c
 *   ./krige_c in.bin out.bin
 *
 * in.bin  : int64 n, gx, gy, k ; double range ; then x[n], y[n], value[n]        (little endian)
 * out.bin : double mean[gx*gy], var[gx*gy] ; int32 nneigh[gx*gy]
 *
 * The problem is the reference's local kriging call (ref: test/estimation/krig.jl:43-52):
 * OrdinaryKriging with a SphericalVariogram(range), maxneighbors = k, on a unit-spaced CartesianGrid
 * whose cells carry the default 3x3 block support.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gskrige.h"

static int fail(const char *what, gsk_ctx *ctx) {
  fprintf(stderr, "krige_c: %s: %s\n", what, gsk_last_error(ctx));
  if (ctx) gsk_destroy(ctx);
  return 1;
}

int main(int argc, char **argv) {
  if (argc != 3) {
    fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]);
    return 2;
  }
  FILE *fi = fopen(argv[1], "rb");
  if (!fi) return fail("cannot open input", NULL);
  int64_t hdr[4];
  double range;
  if (fread(hdr, sizeof(int64_t), 4, fi) != 4 || fread(&range, sizeof(double), 1, fi) != 1) return fail("short header", NULL);
  const int64_t n = hdr[0], gx = hdr[1], gy = hdr[2], k = hdr[3], T = gx * gy;
  double *buf = (double *)malloc(sizeof(double) * 3 * (size_t)n);
  if (!buf || fread(buf, sizeof(double), 3 * (size_t)n, fi) != 3 * (size_t)n) return fail("short sample arrays", NULL);
  fclose(fi);

  double sup_x[GSK_MAX_SUPPORT], sup_y[GSK_MAX_SUPPORT];
  const double spacing[3] = {1.0, 1.0, 1.0};
  const int q = gsk_default_support(2, spacing, range, sup_x, sup_y, NULL, GSK_MAX_SUPPORT);
  if (q <= 0) return fail("gsk_default_support", NULL);

  gsk_problem p;
  memset(&p, 0, sizeof(p));
  p.abi_version = GSK_ABI_VERSION;
  p.dim = 2;
  p.n_samples = n;
  p.coords[0] = buf;
  p.coords[1] = buf + n;
  p.values = buf + 2 * n;
  p.grid_dims[0] = gx;
  p.grid_dims[1] = gy;
  p.grid_dims[2] = 1;
  p.grid_spacing[0] = p.grid_spacing[1] = p.grid_spacing[2] = 1.0;
  p.target_first = 0;
  p.target_count = -1;
  p.n_support = q;
  p.support_offsets[0] = sup_x;
  p.support_offsets[1] = sup_y;
  p.vario_kind = GSK_VARIO_SPHERICAL;
  p.vario_range = range;
  p.vario_sill = 1.0;
  p.vario_nugget = 0.0;
  p.gaussian_nugget_eps = 1e-6;
  p.estimator = GSK_EST_ORDINARY;
  p.min_neighbors = 1;
  p.max_neighbors = (int32_t)k;
  p.ball_radius = NAN;
  p.flags = GSK_FLAGS_DEFAULT;
  if (gsk_num_targets(&p) != T) return fail("gsk_num_targets disagrees", NULL);

  gsk_ctx *ctx = NULL;
  if (gsk_create(&ctx, 0) != GSK_OK) return fail("gsk_create", ctx);
  double *mean = (double *)malloc(sizeof(double) * (size_t)T), *var = (double *)malloc(sizeof(double) * (size_t)T);
  int32_t *nn = (int32_t *)malloc(sizeof(int32_t) * (size_t)T);
  if (!mean || !var || !nn) return fail("out of host memory", ctx);
  if (gsk_krige(ctx, &p, mean, var, nn, NULL) != GSK_OK) return fail("gsk_krige", ctx);
  gsk_timing tm;
  if (gsk_get_timing(ctx, &tm) == GSK_OK)
    printf("krige_c: %lld targets, %lld kernel launches, %.3f ms on the device\n", (long long)tm.targets,
           (long long)tm.launches, tm.ms_total);
  gsk_destroy(ctx);

  FILE *fo = fopen(argv[2], "wb");
  if (!fo) return fail("cannot open output", NULL);
  fwrite(mean, sizeof(double), (size_t)T, fo);
  fwrite(var, sizeof(double), (size_t)T, fo);
  fwrite(nn, sizeof(int32_t), (size_t)T, fo);
  fclose(fo);
  free(buf);
  free(mean);
  free(var);
  free(nn);
  return 0;
}
