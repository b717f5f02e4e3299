// generated layout: instantiations of the fast EM kernels for groups of 64 lanes
#include "em_kernels.cuh"
namespace emfast {
extern const EmVariant em_variants_lpg64[] = {
    {5, 64, (const void *)em_list_kernel<5, 64>, (const void *)em_tile_kernel<5, 64>},
    {6, 64, (const void *)em_list_kernel<6, 64>, (const void *)em_tile_kernel<6, 64>},
    {7, 64, (const void *)em_list_kernel<7, 64>, (const void *)em_tile_kernel<7, 64>},
    {8, 64, (const void *)em_list_kernel<8, 64>, (const void *)em_tile_kernel<8, 64>},
};
extern const int em_variants_lpg64_count = 4;
}  // namespace emfast
