// Kernel bodies of the fast EM: pair-list variant (rows read straight from global/L2 into registers)
// and site-tile variant (rows of a TA x TB block of sites staged once into shared memory by TMA bulk
// copies and reused by every pair of the block).  Instantiated per (IPL, LPG) in em_inst_*.cu.
#pragma once
#include "em_fast.cuh"

namespace emfast {

// ------------------------------------------------------------------------------------------------
// pair-list fetcher: dynamic hand-out from a global counter
struct ListFetch {
  const uint32_t *s1, *s2;
  unsigned long long n;
  unsigned long long *counter;
  const double *gl;
  size_t row_doubles;
  __device__ __forceinline__ bool next(unsigned long long &idx, uint32_t &la, uint32_t &lb) {
    idx = atomicAdd(counter, 1ull);
    if (idx >= n) return false;
    la = s1[idx];
    lb = s2[idx];
    return true;
  }
  __device__ __forceinline__ PairRows resolve(uint32_t la, uint32_t lb) const {
    PairRows r;
    r.a = gl + (size_t)la * row_doubles;
    r.b = gl + (size_t)lb * row_doubles;
    return r;
  }
  __device__ __forceinline__ void sites(uint32_t la, uint32_t lb, uint32_t &a, uint32_t &b) const {
    a = la;
    b = lb;
  }
};

// Register cap: like the warp-per-pair kernel, 152 (128 for IPL <= 4) so that 12-16 warps share an SM.
template <int IPL, int LPG>
__global__ void __maxnreg__(IPL <= 4 ? 128 : 152) em_list_kernel(SiteTable T, PairChunk C, int ignore_miss,
                                                              DevCounters *ctr) {
  __shared__ GroupScratch<LPG> scr;
  ListFetch fetch;
  fetch.s1 = C.s1;
  fetch.s2 = C.s2;
  fetch.n = C.n_pairs;
  fetch.counter = &ctr->next_pair;
  fetch.gl = T.gl;
  fetch.row_doubles = (size_t)T.n_pad * 3;
  run_groups<IPL, LPG>(T, C.rows, ignore_miss != 0, fetch, scr, &ctr->em_passes);
}

// ------------------------------------------------------------------------------------------------
// site-tile variant
struct TileArgs {
  const uint2 *tiles;          // (a_lo, b_lo) in compact site numbering
  unsigned long long n_tiles;
  const uint32_t *cs;          // compact -> site index (NULL = identity)
  const uint32_t *cw_end;      // per compact first site: exclusive end of its partner window (compact)
  const unsigned long long *row_off;  // per compact first site: global output row of its first pair
  unsigned long long row_lo, row_hi;  // this chunk covers global rows [row_lo, row_hi)
  uint32_t n_compact;
  uint32_t TA, TB;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_row_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct TileFetch {
  const TileArgs *A;
  const double *rows_smem;      // [na + nb][row_doubles]
  const uint32_t *site_ids;     // [na + nb]
  unsigned int *cursor;         // shared-memory cursor inside the tile
  uint32_t a_lo, b_lo, na, nb;
  size_t row_doubles;
  __device__ __forceinline__ bool next(unsigned long long &idx, uint32_t &la, uint32_t &lb) {
    const uint32_t total = na * nb;
    for (;;) {
      const uint32_t k = atomicAdd(cursor, 1u);
      if (k >= total) return false;
      const uint32_t ai = k / nb, bi = k - ai * nb;
      const uint32_t c1 = a_lo + ai, c2 = b_lo + bi;
      if (c2 <= c1 || c2 >= A->cw_end[c1]) continue;
      const unsigned long long g = A->row_off[c1] + (c2 - c1 - 1);
      if (g < A->row_lo || g >= A->row_hi) continue;
      idx = g - A->row_lo;
      la = ai;
      lb = na + bi;
      return true;
    }
  }
  __device__ __forceinline__ PairRows resolve(uint32_t la, uint32_t lb) const {
    PairRows r;
    r.a = rows_smem + (size_t)la * row_doubles;
    r.b = rows_smem + (size_t)lb * row_doubles;
    return r;
  }
  __device__ __forceinline__ void sites(uint32_t la, uint32_t lb, uint32_t &a, uint32_t &b) const {
    a = site_ids[la];
    b = site_ids[lb];
  }
};

template <int IPL, int LPG>
__global__ void __maxnreg__(IPL <= 4 ? 128 : 152) em_tile_kernel(SiteTable T, ngsld_pair_row *rows_out, TileArgs A,
                                                              int ignore_miss, DevCounters *ctr) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ GroupScratch<LPG> scr;
  __shared__ __align__(8) uint64_t bar;
  __shared__ unsigned int cursor;
  __shared__ unsigned long long cur_tile;
  __shared__ uint32_t site_ids[128];

  const size_t row_doubles = (size_t)T.n_pad * 3;
  const uint32_t row_bytes = (uint32_t)(row_doubles * sizeof(double));
  double *rows_smem = reinterpret_cast<double *>(dyn_smem);
  const int tid = threadIdx.x;
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  uint32_t phase = 0;

  for (;;) {
    if (tid == 0) {
      cur_tile = atomicAdd(&ctr->next_tile, 1ull);
      cursor = 0;
    }
    __syncthreads();
    const unsigned long long t = cur_tile;
    if (t >= A.n_tiles) break;
    const uint2 td = A.tiles[t];
    const uint32_t na = min(A.TA, A.n_compact - td.x), nb = min(A.TB, A.n_compact - td.y);
    // stage the tile's site rows: one TMA bulk copy per row, all signalling one mbarrier
    if (tid == 0) mbar_expect_tx(&bar, (na + nb) * row_bytes);
    if (tid < (int)(na + nb)) {
      const uint32_t c = tid < (int)na ? td.x + tid : td.y + (tid - na);
      const uint32_t site = A.cs ? A.cs[c] : c;
      site_ids[tid] = site;
      tma_row_load(rows_smem + (size_t)tid * row_doubles, T.gl + (size_t)site * row_doubles, row_bytes, &bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    __syncthreads();  // site_ids visible

    TileFetch fetch;
    fetch.A = &A;
    fetch.rows_smem = rows_smem;
    fetch.site_ids = site_ids;
    fetch.cursor = &cursor;
    fetch.a_lo = td.x;
    fetch.b_lo = td.y;
    fetch.na = na;
    fetch.nb = nb;
    fetch.row_doubles = row_doubles;
    run_groups<IPL, LPG>(T, rows_out, ignore_miss != 0, fetch, scr, &ctr->em_passes);
    __syncthreads();  // everyone is done reading the tile before it is overwritten
  }
}

// ------------------------------------------------------------------------------------------------
// host-visible launch table (one entry per instantiated (IPL, LPG); filled in em_inst_*.cu)
struct EmVariant {
  int ipl, lpg;
  const void *list_fn;
  const void *tile_fn;
};

}  // namespace emfast
