// generated layout: instantiations of the fast EM kernels for groups of 32 lanes
#include "em_kernels.cuh"
namespace emfast {
extern const EmVariant em_variants_lpg32[] = {
    {5, 32, (const void *)em_list_kernel<5, 32>, (const void *)em_tile_kernel<5, 32>},
    {6, 32, (const void *)em_list_kernel<6, 32>, (const void *)em_tile_kernel<6, 32>},
    {7, 32, (const void *)em_list_kernel<7, 32>, (const void *)em_tile_kernel<7, 32>},
    {8, 32, (const void *)em_list_kernel<8, 32>, (const void *)em_tile_kernel<8, 32>},
};
extern const int em_variants_lpg32_count = 4;
}  // namespace emfast
