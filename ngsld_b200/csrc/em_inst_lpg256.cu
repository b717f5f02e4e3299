// generated layout: instantiations of the fast EM kernels for groups of 256 lanes
#include "em_kernels.cuh"
namespace emfast {
extern const EmVariant em_variants_lpg256[] = {
    {5, 256, (const void *)em_list_kernel<5, 256>, (const void *)em_tile_kernel<5, 256>},
    {6, 256, (const void *)em_list_kernel<6, 256>, (const void *)em_tile_kernel<6, 256>},
    {7, 256, (const void *)em_list_kernel<7, 256>, (const void *)em_tile_kernel<7, 256>},
    {8, 256, (const void *)em_list_kernel<8, 256>, (const void *)em_tile_kernel<8, 256>},
};
extern const int em_variants_lpg256_count = 4;
}  // namespace emfast
