// local_solve_wpt.cu — instantiates the block-pool kernel with one warp per target (32 < k <= 64; local_solve_wpt.cuh)
#include "local_solve_wpt.cuh"

cudaError_t gsk_local_launch_wpt(const GskLocalArgs &a, int e, cudaStream_t st) { return gsk_wpt::launch_wpt_any<1>(a, e, st); }
