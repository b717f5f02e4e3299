// r2_ExpG of one site pair: the pair-dependent half of gsl_stats_correlation's recurrence (reference ngsLD.cpp:365-367;
// GSL statistics/covariance_source.c) in emulated x87 arithmetic.  Per site, aux::site_terms_kernel already produced
// delta_i = x_i - mean_(i-1) (80-bit image) and q = sqrt((double)sum_sq); here
//     sum_cross = sum_{i>=1} fl80( fl80(da_i * db_i) * (long double)(i/(i+1.0)) )   (in order)
//     r = fl80( sum_cross / (long double)(qa*qb) ),  r2 = (double)r * (double)r.
// One THREAD per pair (the sum is inherently sequential); callers give the 32 lanes of a warp pairs with the same
// first site and consecutive second sites, so the loads from the individual-major tables coalesce.
#pragma once
#include "common.cuh"
#include "fp80.cuh"

// 1: the four terms of a block of individuals are formed first (their multiplications interleave), then added one after the
// other; 0: term, add, term, add ...
#ifndef NGSLD_X87_ILP
#define NGSLD_X87_ILP 1
#endif

namespace pearson {

__device__ __forceinline__ double pair_r2(const SiteTable &T, uint32_t s1, uint32_t s2) {
  // Blocks of four individuals (individual 0 and the padding behind the last one are stored as zero terms, which mac3
  // skips).  One register set, refilled on the fly: as soon as the four terms of a block are formed, the next block is
  // requested into the same registers, so every request has the four additions (~300 instructions) to come back from L2,
  // and nothing is ever copied.  The tables
  // have one spare block row behind the last one, so the last refill needs no guard.  Record (blk, s) has the index
  // blk * n_sites + s < 2^32 (a site table of that many 24-byte likelihood triples would not fit the memory).
  // The four ratios of a block are the same words for every lane and every pair: they sit in L1 and are fetched when
  // needed instead of travelling through prefetch registers.
  x87::acc96 acc3 = x87::acc96_zero();
  const ulonglong2 *SIG = reinterpret_cast<const ulonglong2 *>(T.dx_sig);  // two individuals per element
  const uint32_t *SE = reinterpret_cast<const uint32_t *>(T.dx_se);        // two individuals per element
  const ulonglong2 *pr = reinterpret_cast<const ulonglong2 *>(T.ratio);
  uint32_t ia = s1, ib = s2;
  ulonglong2 a01 = SIG[2 * (size_t)ia], a23 = SIG[2 * (size_t)ia + 1], b01 = SIG[2 * (size_t)ib], b23 = SIG[2 * (size_t)ib + 1];
  uint32_t ea01 = SE[2 * (size_t)ia], ea23 = SE[2 * (size_t)ia + 1], eb01 = SE[2 * (size_t)ib], eb23 = SE[2 * (size_t)ib + 1];
  for (uint32_t blk = 0; blk < T.n_blk; blk++) {
    const ulonglong2 r01 = __ldg(pr + 2 * blk), r23 = __ldg(pr + 2 * blk + 1);
    ia += T.n_sites;
    ib += T.n_sites;
#if NGSLD_X87_ILP
    // the four terms first (independent of each other and of the sum: the compiler interleaves their multiplications),
    // then the four additions, which are the sequential part
    const x87::term96 t0 = x87::term3(a01.x, ea01 & 0xffffu, b01.x, eb01 & 0xffffu, r01.x);
    const x87::term96 t1 = x87::term3(a01.y, ea01 >> 16, b01.y, eb01 >> 16, r01.y);
    const x87::term96 t2 = x87::term3(a23.x, ea23 & 0xffffu, b23.x, eb23 & 0xffffu, r23.x);
    const x87::term96 t3 = x87::term3(a23.y, ea23 >> 16, b23.y, eb23 >> 16, r23.y);
    a01 = SIG[2 * (size_t)ia]; b01 = SIG[2 * (size_t)ib];
    ea01 = SE[2 * (size_t)ia]; eb01 = SE[2 * (size_t)ib];
    a23 = SIG[2 * (size_t)ia + 1]; b23 = SIG[2 * (size_t)ib + 1];
    ea23 = SE[2 * (size_t)ia + 1]; eb23 = SE[2 * (size_t)ib + 1];
    x87::acc3(acc3, t0);
    x87::acc3(acc3, t1);
    x87::acc3(acc3, t2);
    x87::acc3(acc3, t3);
#else
    x87::mac3(acc3, a01.x, ea01 & 0xffffu, b01.x, eb01 & 0xffffu, r01.x);
    x87::mac3(acc3, a01.y, ea01 >> 16, b01.y, eb01 >> 16, r01.y);
    a01 = SIG[2 * (size_t)ia]; b01 = SIG[2 * (size_t)ib];
    ea01 = SE[2 * (size_t)ia]; eb01 = SE[2 * (size_t)ib];
    x87::mac3(acc3, a23.x, ea23 & 0xffffu, b23.x, eb23 & 0xffffu, r23.x);
    x87::mac3(acc3, a23.y, ea23 >> 16, b23.y, eb23 >> 16, r23.y);
    a23 = SIG[2 * (size_t)ia + 1]; b23 = SIG[2 * (size_t)ib + 1];
    ea23 = SE[2 * (size_t)ia + 1]; eb23 = SE[2 * (size_t)ib + 1];
#endif
  }
  const x87::ext acc = x87::acc96_to_ext(acc3);
  const double den = __dmul_rn(T.q[s1], T.q[s2]);
  double r;
  if (den == 0.0 || den != den) {
    // x87: 0/0 -> default NaN; finite/0 -> signed infinity
    if (acc.sig == 0 || den != den)
      r = __longlong_as_double(0xfff8000000000000ll);
    else
      r = __longlong_as_double(acc.neg ? 0xfff0000000000000ll : 0x7ff0000000000000ll);
  } else {
    r = x87::to_double(x87::div(acc, x87::from_double(den)));
  }
  return __dmul_rn(r, r);
}

}  // namespace pearson
