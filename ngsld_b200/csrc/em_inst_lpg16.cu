// generated layout: instantiations of the fast EM kernels for groups of 16 lanes
#include "em_kernels.cuh"
namespace emfast {
extern const EmVariant em_variants_lpg16[] = {
    {5, 16, (const void *)em_list_kernel<5, 16>, (const void *)em_tile_kernel<5, 16>},
    {6, 16, (const void *)em_list_kernel<6, 16>, (const void *)em_tile_kernel<6, 16>},
    {7, 16, (const void *)em_list_kernel<7, 16>, (const void *)em_tile_kernel<7, 16>},
    {8, 16, (const void *)em_list_kernel<8, 16>, (const void *)em_tile_kernel<8, 16>},
};
extern const int em_variants_lpg16_count = 4;
}  // namespace emfast
