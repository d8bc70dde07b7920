// local_solve_C2.cu — instantiates K3 for <G, R, W, RS, NT> = <8, 5, 4, 36, 128> (see local_solve.cuh)
#include "local_solve.cuh"
cudaError_t gsk_local_launch_C2(const GskLocalArgs &a, int e, cudaStream_t st) {
  return gsk_local::launch_cfg<8, 5, 4, 36, 128>(a, e, st);
}
