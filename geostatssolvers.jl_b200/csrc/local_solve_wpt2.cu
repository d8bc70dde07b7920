// local_solve_wpt2.cu — instantiates the block-pool kernel with two targets per warp (k <= 32; local_solve_wpt.cuh)
#include "local_solve_wpt.cuh"

cudaError_t gsk_local_launch_wpt2(const GskLocalArgs &a, int e, cudaStream_t st) { return gsk_wpt::launch_wpt_any<2>(a, e, st); }
