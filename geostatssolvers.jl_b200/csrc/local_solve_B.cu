// local_solve_B.cu — instantiates K3 for G=4 lanes/target, R=6 row slots, W=4 panel columns (see local_solve.cuh)
#include "local_solve.cuh"
cudaError_t gsk_local_launch_B(const GskLocalArgs &a, int e, cudaStream_t st) {
  return gsk_local::launch_cfg<4, 6, 4>(a, e, st);
}
