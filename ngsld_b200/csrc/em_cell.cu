// Site palettes, the data-set statistics that decide for or against the class-compressed EM, and the instantiations of
// that kernel (em_cell.cuh).
#include "em_cell.cuh"

namespace emcell {

__global__ void __launch_bounds__(CTA_THREADS) build_palette_kernel(const double *gl, uint32_t n_sites, uint32_t n_ind,
                                                                    uint32_t n_pad, uint32_t n_cpad, uint8_t *cls, double *pal,
                                                                    uint8_t *pal_k, uint64_t *pal_miss, unsigned int *max_k) {
  __shared__ unsigned long long ps[WARPS_PER_CTA][3][NGSLD_KMAX];  // the palette being built, as bit patterns
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long(*p)[NGSLD_KMAX] = ps[warp];
  for (uint32_t s = blockIdx.x * WARPS_PER_CTA + warp; s < n_sites; s += gridDim.x * WARPS_PER_CTA) {
    const double *row = gl + (size_t)s * n_pad * 3;
    uint8_t *crow = cls + (size_t)s * n_cpad;
    uint32_t k = 0;
    bool over = false;
    for (uint32_t base = 0; base < n_ind && !over; base += 32u) {
      const uint32_t i = base + lane;
      const bool valid = i < n_ind;
      unsigned long long g0 = 0, g1 = 0, g2 = 0;
      if (valid) {
        g0 = (unsigned long long)__double_as_longlong(row[3 * (size_t)i]);
        g1 = (unsigned long long)__double_as_longlong(row[3 * (size_t)i + 1]);
        g2 = (unsigned long long)__double_as_longlong(row[3 * (size_t)i + 2]);
      }
      int id = -1;
      for (uint32_t c = 0; c < k; c++)
        if (id < 0 && p[0][c] == g0 && p[1][c] == g1 && p[2][c] == g2) id = (int)c;
      unsigned un = __ballot_sync(0xffffffffu, valid && id < 0);
      while (un) {  // warp-uniform: the lowest unmatched lane's triple becomes the next class
        if (k == NGSLD_KMAX) {
          over = true;
          break;
        }
        const int src = __ffs(un) - 1;
        const unsigned long long b0 = __shfl_sync(0xffffffffu, g0, src), b1 = __shfl_sync(0xffffffffu, g1, src),
                                 b2 = __shfl_sync(0xffffffffu, g2, src);
        if (lane == 0) {
          p[0][k] = b0;
          p[1][k] = b1;
          p[2][k] = b2;
        }
        if (valid && id < 0 && g0 == b0 && g1 == b1 && g2 == b2) id = (int)k;
        k++;
        un = __ballot_sync(0xffffffffu, valid && id < 0);
      }
      __syncwarp();  // new palette entries are compared against by every lane in the next round
      if (!over && valid) crow[i] = (uint8_t)id;
    }
    for (uint32_t i = n_ind + lane; i < n_cpad; i += 32u) crow[i] = 0;
    if (over) k = 0;
    uint64_t miss = 0;
#pragma unroll
    for (int h = 0; h < NGSLD_KMAX / 32; h++) {
      const uint32_t c = (uint32_t)lane + 32u * h;
      bool m = false;
      if (c < k) {
        const double a = __longlong_as_double((long long)p[0][c]), b = __longlong_as_double((long long)p[1][c]),
                     d = __longlong_as_double((long long)p[2][c]);
        double *o = pal + ((size_t)s * NGSLD_KMAX + c) * 3;
        o[0] = a;
        o[1] = b;
        o[2] = d;
        m = gl_missing(a, b, d);
      }
      miss |= (uint64_t)__ballot_sync(0xffffffffu, m) << (32 * h);
    }
    if (lane == 0) {
      pal_k[s] = (uint8_t)k;
      pal_miss[s] = miss;
      atomicMax(max_k, k);  // the largest palette sizes the joint-class tables of the EM kernel
    }
    __syncwarp();  // the next site overwrites the shared palette
  }
}

__device__ __forceinline__ uint32_t mix32(uint32_t x) {  // lowbias32 integer hash
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

__global__ void __launch_bounds__(CTA_THREADS) cell_stats_kernel(SiteTable T, uint32_t n_samples, int ignore_miss,
                                                                 unsigned long long *out, unsigned int *hist) {
  __shared__ __align__(16) uint16_t all_bins[WARPS_PER_CTA][NBINS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint16_t *bins = all_bins[warp];
  wipe_bins(bins, NGSLD_KMAX, lane);
  for (uint32_t j = blockIdx.x * WARPS_PER_CTA + warp; j < n_samples; j += gridDim.x * WARPS_PER_CTA) {
    const uint32_t s1 = mix32(2u * j + 1u) % T.n_sites;
    uint32_t s2 = mix32(2u * j + 2u) % T.n_sites;
    if (s2 == s1) s2 = (s2 + 1u) % T.n_sites;
    if (T.pal_k[s1] == 0 || T.pal_k[s2] == 0) {
      if (lane == 0) {
        atomicAdd(&out[1], 1ull);
        atomicAdd(&out[2], 1ull);
      }
      continue;
    }
    uint32_t used = 0;
    const uint32_t n = joint_classes(T, s1, s2, ignore_miss != 0, bins, NGSLD_KMAX, nullptr, 0, used, lane);
    wipe_bins(bins, NGSLD_KMAX, lane);
    if (lane == 0) {
      atomicAdd(&out[0], (unsigned long long)n);
      atomicAdd(&out[1], 1ull);
      atomicAdd(&hist[n / 32u < 128u ? n / 32u : 128u], 1u);
    }
  }
}

#define V(r, b) {r, b, (const void *)em_cell_kernel<r, false, b>, (const void *)em_cell_kernel<r, true, b>}
extern const CellVariant cell_variants[] = {V(2, 4), V(2, 5), V(4, 3), V(4, 4), V(4, 5), V(6, 3), V(6, 4)};
#undef V
extern const int cell_variants_count = 7;

}  // namespace emcell
