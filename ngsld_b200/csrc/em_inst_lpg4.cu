// generated layout: instantiations of the fast EM kernels for groups of 4 lanes
#include "em_kernels.cuh"
namespace emfast {
extern const EmVariant em_variants_lpg4[] = {
    {1, 4, (const void *)em_list_kernel<1, 4>, (const void *)em_tile_kernel<1, 4>},
    {2, 4, (const void *)em_list_kernel<2, 4>, (const void *)em_tile_kernel<2, 4>},
    {3, 4, (const void *)em_list_kernel<3, 4>, (const void *)em_tile_kernel<3, 4>},
    {4, 4, (const void *)em_list_kernel<4, 4>, (const void *)em_tile_kernel<4, 4>},
    {5, 4, (const void *)em_list_kernel<5, 4>, (const void *)em_tile_kernel<5, 4>},
    {6, 4, (const void *)em_list_kernel<6, 4>, (const void *)em_tile_kernel<6, 4>},
    {7, 4, (const void *)em_list_kernel<7, 4>, (const void *)em_tile_kernel<7, 4>},
    {8, 4, (const void *)em_list_kernel<8, 4>, (const void *)em_tile_kernel<8, 4>},
};
extern const int em_variants_lpg4_count = 8;
}  // namespace emfast
