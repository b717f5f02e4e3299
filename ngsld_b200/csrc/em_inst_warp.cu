// instantiations of the warp-per-pair fast EM kernel (em_warp.cuh), one per register depth R
#include "em_warp.cuh"
namespace emwarp {
#define V(r) {r, (const void *)em_warp_kernel<r, false, 2>, (const void *)em_warp_kernel<r, true, 1>, (const void *)em_warp_kernel<r, false, 1>}
extern const WarpVariant warp_variants[] = {V(1), V(2), V(3), V(4), V(5), V(6), V(7), V(8)};
#undef V
extern const int warp_variants_count = 8;
}  // namespace emwarp
