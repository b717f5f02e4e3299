// instantiations of the warp-per-pair fast EM kernel (em_warp.cuh): register depth R x warps per pair G
#include "em_warp.cuh"
namespace emwarp {
#define V(r, g)                                                                                             \
  {r, g, (const void *)em_warp_kernel<r, false, 2, g>, (const void *)em_warp_kernel<r, true, 1, g>, \
   (const void *)em_warp_kernel<r, false, 1, g>}
extern const WarpVariant warp_variants[] = {V(1, 1), V(2, 1), V(3, 1), V(4, 1), V(5, 1), V(6, 1), V(7, 1), V(8, 1),
                                            V(4, 2), V(5, 2), V(6, 2), V(4, 4), V(5, 4), V(6, 4)};
#undef V
extern const int warp_variants_count = 14;
}  // namespace emwarp
