/* TEST INFRASTRUCTURE — not product code.
 *
 * Minimal stand-in for <gsl/gsl_rng.h> so the UNMODIFIED reference sources under
 * /root/reference compile in an image that has no GSL (oracle/Makefile, target _ref/ngsLD).
 * Only the four entry points the reference calls are provided
 * (reference call sites: ngsLD.cpp:69-70,165-166,216,357; shared/gen_func.cpp:117-119).
 *
 * Third-party algorithm restated: GSL `rng/taus.c`, generator `gsl_rng_taus`
 * (L'Ecuyer 1996 maximally-equidistributed three-component Tausworthe; the plain variant,
 * i.e. without the taus2 seed fix-ups).  GSL is not vendored in the reference and has no pinned
 * version (reference README.md:20 names "v1.15 tested"); the known answer used to pin this
 * restatement is GSL's own self-test value: seed 1 -> 10000th output 2733957125 (tests/test_oracle.py).
 */
#ifndef NGSLD_ORACLE_GSL_RNG_SHIM_H
#define NGSLD_ORACLE_GSL_RNG_SHIM_H
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int id; } gsl_rng_type;
typedef struct { unsigned long int s1, s2, s3; } gsl_rng;

static const gsl_rng_type gsl_rng_taus_shim_type = { 1 };
#define gsl_rng_taus (&gsl_rng_taus_shim_type)

#define NGSLD_SHIM_MASK 0xffffffffUL
#define NGSLD_SHIM_TAUS(s, a, b, c, d) \
  ((((s) & (c)) << (d)) & NGSLD_SHIM_MASK) ^ (((((s) << (a)) & NGSLD_SHIM_MASK) ^ (s)) >> (b))

static inline unsigned long int gsl_rng_get(gsl_rng *r) {
  r->s1 = NGSLD_SHIM_TAUS(r->s1, 13, 19, 4294967294UL, 12);
  r->s2 = NGSLD_SHIM_TAUS(r->s2, 2, 25, 4294967288UL, 4);
  r->s3 = NGSLD_SHIM_TAUS(r->s3, 3, 11, 4294967280UL, 17);
  return r->s1 ^ r->s2 ^ r->s3;
}

static inline void gsl_rng_set(gsl_rng *r, unsigned long int seed) {
  if (seed == 0) seed = 1;
  r->s1 = (69069UL * seed) & NGSLD_SHIM_MASK;
  r->s2 = (69069UL * r->s1) & NGSLD_SHIM_MASK;
  r->s3 = (69069UL * r->s2) & NGSLD_SHIM_MASK;
  for (int k = 0; k < 6; k++) gsl_rng_get(r); /* warm-up */
}

static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *t) {
  (void)t;
  gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
  if (r) gsl_rng_set(r, 0);
  return r;
}

static inline double gsl_rng_uniform(gsl_rng *r) { return gsl_rng_get(r) / 4294967296.0; }

static inline void gsl_rng_free(gsl_rng *r) { free(r); }

#ifdef __cplusplus
}
#endif
#endif
