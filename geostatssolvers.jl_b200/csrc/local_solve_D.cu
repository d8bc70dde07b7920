// local_solve_D.cu — instantiates K3 for G=16 lanes/target, R=5 row slots, W=8 panel columns (see local_solve.cuh)
#include "local_solve.cuh"
cudaError_t gsk_local_launch_D(const GskLocalArgs &a, int e, cudaStream_t st) {
  return gsk_local::launch_cfg<16, 5, 8>(a, e, st);
}
