// local_solve_A.cu — instantiates K3 for G=4 lanes/target, R=3 row slots, W=4 panel columns (see local_solve.cuh)
#include "local_solve.cuh"
cudaError_t gsk_local_launch_A(const GskLocalArgs &a, int e, cudaStream_t st) {
  return gsk_local::launch_cfg<4, 3, 4>(a, e, st);
}
