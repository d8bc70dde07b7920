// local_solve_C4.cu — instantiates K3 for <G, R, W, RS, NT> = <8, 5, 8, 40, 96> (see local_solve.cuh)
#include "local_solve.cuh"
cudaError_t gsk_local_launch_C4(const GskLocalArgs &a, int e, cudaStream_t st) {
  return gsk_local::launch_cfg<8, 5, 8, 40, 96>(a, e, st);
}
