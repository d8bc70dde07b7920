// local_solve_wpt.cu — instantiates the block-pool kernel for 32 < k <= 64 (local_solve_wpt.cuh)
#include "local_solve_wpt.cuh"

cudaError_t gsk_local_launch_wpt(const GskLocalArgs &a, int e, cudaStream_t st) {
  using namespace gsk_wpt;
  constexpr int NW = 8;
  const bool d3 = a.tg.dim == 3;
  switch (a.vg.kind) {
    case GSK_VARIO_GAUSSIAN:
      return d3 ? launch_wpt<3, GSK_VARIO_GAUSSIAN, NW>(a, e, st) : launch_wpt<2, GSK_VARIO_GAUSSIAN, NW>(a, e, st);
    case GSK_VARIO_SPHERICAL:
      return d3 ? launch_wpt<3, GSK_VARIO_SPHERICAL, NW>(a, e, st) : launch_wpt<2, GSK_VARIO_SPHERICAL, NW>(a, e, st);
    default:
      return d3 ? launch_wpt<3, GSK_VARIO_EXPONENTIAL, NW>(a, e, st) : launch_wpt<2, GSK_VARIO_EXPONENTIAL, NW>(a, e, st);
  }
}
