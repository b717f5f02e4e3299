// generated layout: instantiations of the fast EM kernels for groups of 128 lanes
#include "em_kernels.cuh"
namespace emfast {
extern const EmVariant em_variants_lpg128[] = {
    {5, 128, (const void *)em_list_kernel<5, 128>, (const void *)em_tile_kernel<5, 128>},
    {6, 128, (const void *)em_list_kernel<6, 128>, (const void *)em_tile_kernel<6, 128>},
    {7, 128, (const void *)em_list_kernel<7, 128>, (const void *)em_tile_kernel<7, 128>},
    {8, 128, (const void *)em_list_kernel<8, 128>, (const void *)em_tile_kernel<8, 128>},
};
extern const int em_variants_lpg128_count = 4;
}  // namespace emfast
