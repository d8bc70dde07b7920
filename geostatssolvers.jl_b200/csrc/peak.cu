// peak.cu — measured FP64 roofline denominators for this device (MEASURED_PEAKS.json has none):
// a dependency-free DFMA loop (vector FP64 pipe) and an mma.sync f64 loop (DMMA, FP64 tensor path).
// tcgen05 has no FP64 kind, so mma.sync is the only tensor route for doubles on sm_100a.
#include "gsk_internal.cuh"

namespace {

constexpr int CHAINS = 8;

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = (double)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double a, double b) {
  // m16n8k16: A 16×16 (8 regs), B 16×8 (4 regs), C/D 16×8 (4 regs) per thread
  double c[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
  double ra[8], rb[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) ra[i] = a + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) rb[i] = b + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
          "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
          : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
          : "d"(ra[0]), "d"(ra[1]), "d"(ra[2]), "d"(ra[3]), "d"(ra[4]), "d"(ra[5]), "d"(ra[6]), "d"(ra[7]),
            "d"(rb[0]), "d"(rb[1]), "d"(rb[2]), "d"(rb[3]));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[j][i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

int gsk_peak_measure(gsk_ctx *ctx, double *dfma, double *dmma) {
  double *d_out = nullptr;
  const int blocks = ctx->sm_count * 8, threads = 256;
  int rc = gsk_buf(ctx, BUF_PEAK, sizeof(double) * blocks * threads, (void **)&d_out);
  if (rc != GSK_OK) return rc;
  cudaEvent_t e0 = ctx->ev[3], e1 = ctx->ev[4];
  auto time_best = [&](auto launch) -> double {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0, ctx->stream);
      launch();
      cudaEventRecord(e1, ctx->stream);
      cudaEventSynchronize(e1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    return (double)best;
  };
  const int it1 = 20000;
  double ms1 = time_best([&] { dfma_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, it1, 1.0000001, 1e-9); });
  *dfma = (double)blocks * threads * (double)it1 * CHAINS * 2.0 / (ms1 * 1e-3) / 1e12;
  const int it2 = 4000;
  double ms2 = time_best([&] { dmma_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, it2, 1.0000001, 1e-9); });
  // per warp and mma: 16·8·16 FMAs
  *dmma = (double)blocks * (threads / 32) * (double)it2 * 4.0 * (16.0 * 8.0 * 16.0 * 2.0) / (ms2 * 1e-3) / 1e12;
  GSK_CUDA_CHECK(ctx, cudaGetLastError());
  return GSK_OK;
}
