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

#ifndef NGSLD_PEARSON_UNROLL2
#define NGSLD_PEARSON_UNROLL2 0
#endif

namespace pearson {

__device__ __forceinline__ double pair_r2(const SiteTable &T, uint32_t s1, uint32_t s2) {
#if NGSLD_PEARSON_UNROLL2
  // Experiment (not the default: measured together with the straight-line accumulate step it was 3 % slower): two
  // individuals per trip with two register sets, so that no set is ever copied into the other; the tables have a spare
  // row behind the last individual, so the request for "the next one" needs no guard.
  x87::ext acc = x87::zero(0);
  const size_t stride = T.n_sites;
  const uint64_t *p1 = T.dx_sig + stride + s1, *p2 = T.dx_sig + stride + s2;  // row i = 1
  const uint16_t *q1 = T.dx_se + stride + s1, *q2 = T.dx_se + stride + s2;
  const uint64_t *pr = T.ratio + 1;
  uint64_t a0 = 0, b0 = 0, r0 = 0, a1, b1, r1;
  uint32_t ea0 = 0, eb0 = 0, ea1, eb1;
  if (T.n_ind > 1) {
    a0 = *p1; b0 = *p2; ea0 = *q1; eb0 = *q2; r0 = __ldg(pr);
  }
  for (uint32_t i = 1; i < T.n_ind; i += 2) {
    p1 += stride; p2 += stride; q1 += stride; q2 += stride; pr++;
    a1 = *p1; b1 = *p2; ea1 = *q1; eb1 = *q2; r1 = __ldg(pr);
    x87::mac_ratio(acc, a0, ea0, b0, eb0, r0);
    if (i + 1 >= T.n_ind) break;
    p1 += stride; p2 += stride; q1 += stride; q2 += stride; pr++;
    a0 = *p1; b0 = *p2; ea0 = *q1; eb0 = *q2; r0 = __ldg(pr);
    x87::mac_ratio(acc, a1, ea1, b1, eb1, r1);
  }
#else
  x87::ext acc = x87::zero(0);
  // software-pipelined by hand: the operands of individual i + 1 are requested before individual i is accumulated,
  // otherwise every iteration would wait out a full L2 round trip (the loop body is too branchy for the compiler
  // to hoist the loads itself)
  const uint64_t *sig = T.dx_sig + T.n_sites;  // row i = 1
  const uint16_t *se = T.dx_se + T.n_sites;
  uint64_t a_sig = 0, b_sig = 0, r_sig = 0;
  uint32_t a_se = 0, b_se = 0;
  if (T.n_ind > 1) {
    a_sig = sig[s1]; b_sig = sig[s2]; a_se = se[s1]; b_se = se[s2]; r_sig = __ldg(T.ratio + 1);
  }
  for (uint32_t i = 1; i < T.n_ind; i++) {
    uint64_t na_sig = 0, nb_sig = 0, nr_sig = 0;
    uint32_t na_se = 0, nb_se = 0;
    if (i + 1 < T.n_ind) {
      sig += T.n_sites;
      se += T.n_sites;
      na_sig = sig[s1]; nb_sig = sig[s2]; na_se = se[s1]; nb_se = se[s2]; nr_sig = __ldg(T.ratio + i + 1);
    }
    x87::mac_ratio(acc, a_sig, a_se, b_sig, b_se, r_sig);
    a_sig = na_sig; b_sig = nb_sig; a_se = na_se; b_se = nb_se; r_sig = nr_sig;
  }
#endif
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
