// local_solve_E.cu — instantiates K3 for G=32 lanes/target, R=4 row slots, W=8 panel columns (see local_solve.cuh)
#include "local_solve.cuh"
cudaError_t gsk_local_launch_E(const GskLocalArgs &a, int e, cudaStream_t st) {
  return gsk_local::launch_cfg<32, 4, 8>(a, e, st);
}
