// local_solve_small.cu — instantiates the k <= 20 Simple/Ordinary-Kriging fast path (local_solve_small.cuh)
#include "local_solve_small.cuh"

cudaError_t gsk_local_launch_small(const GskLocalArgs &a, cudaStream_t st) {
  using namespace gsk_local;
  const bool d3 = a.tg.dim == 3;
  switch (a.vg.kind) {
    case GSK_VARIO_GAUSSIAN:
      return d3 ? launch_small_one<3, GSK_VARIO_GAUSSIAN>(a, st) : launch_small_one<2, GSK_VARIO_GAUSSIAN>(a, st);
    case GSK_VARIO_SPHERICAL:
      return d3 ? launch_small_one<3, GSK_VARIO_SPHERICAL>(a, st) : launch_small_one<2, GSK_VARIO_SPHERICAL>(a, st);
    default:
      return d3 ? launch_small_one<3, GSK_VARIO_EXPONENTIAL>(a, st) : launch_small_one<2, GSK_VARIO_EXPONENTIAL>(a, st);
  }
}
