// local_solve_C.cu — instantiates K3 for G=8 lanes/target, R=5 row slots, W=8 panel columns (see local_solve.cuh)
#include "local_solve.cuh"
cudaError_t gsk_local_launch_C(const GskLocalArgs &a, int e, cudaStream_t st) {
  return gsk_local::launch_cfg<8, 5, 8>(a, e, st);
}
