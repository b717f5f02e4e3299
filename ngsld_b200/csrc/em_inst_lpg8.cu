// generated layout: instantiations of the fast EM kernels for groups of 8 lanes
#include "em_kernels.cuh"
namespace emfast {
extern const EmVariant em_variants_lpg8[] = {
    {5, 8, (const void *)em_list_kernel<5, 8>, (const void *)em_tile_kernel<5, 8>},
    {6, 8, (const void *)em_list_kernel<6, 8>, (const void *)em_tile_kernel<6, 8>},
    {7, 8, (const void *)em_list_kernel<7, 8>, (const void *)em_tile_kernel<7, 8>},
    {8, 8, (const void *)em_list_kernel<8, 8>, (const void *)em_tile_kernel<8, 8>},
};
extern const int em_variants_lpg8_count = 4;
}  // namespace emfast
