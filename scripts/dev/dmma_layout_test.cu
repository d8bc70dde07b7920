// Verifies the register layout assumed for mma.sync.m16n8k16 f64 fragments (dev aid for global.cu).
#include <cstdio>
#include <cmath>
__global__ void k(const double* A, const double* B, double* C) {  // A 16x16 row-major, B 16x8 row-major (k x n), C 16x8
  int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double a[8], b[4], c[4] = {0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) a[i] = A[(g + 8 * (i & 1)) * 16 + t + 4 * (i >> 1)];
  for (int i = 0; i < 4; ++i) b[i] = B[(t + 4 * i) * 8 + g];
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
    : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
    : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  C[g * 8 + 2 * t] = c[0]; C[g * 8 + 2 * t + 1] = c[1]; C[(g + 8) * 8 + 2 * t] = c[2]; C[(g + 8) * 8 + 2 * t + 1] = c[3];
}
int main() {
  double hA[256], hB[128], hC[128], ref[128];
  for (int i = 0; i < 256; ++i) hA[i] = sin(i * 0.37) + 0.01 * i;
  for (int i = 0; i < 128; ++i) hB[i] = cos(i * 0.11) - 0.02 * i;
  for (int m = 0; m < 16; ++m) for (int n = 0; n < 8; ++n) { double s = 0; for (int kk = 0; kk < 16; ++kk) s += hA[m * 16 + kk] * hB[kk * 8 + n]; ref[m * 8 + n] = s; }
  double *dA, *dB, *dC; cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dC, sizeof hC);
  cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
  k<<<1, 32>>>(dA, dB, dC); cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost);
  double e = 0; for (int i = 0; i < 128; ++i) e = fmax(e, fabs(hC[i] - ref[i]));
  printf("dmma m16n8k16 layout max err = %.3e (%s)\n", e, e < 1e-10 ? "OK" : "MISMATCH");
  return e < 1e-10 ? 0 : 1;
}
