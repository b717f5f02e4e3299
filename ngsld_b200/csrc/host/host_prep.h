// Host-side per-site preparation (see host_prep.cpp).
#pragma once
#include <stdint.h>

namespace hostprep {

struct PrepOptions {
  bool log_scale = false;       // file already holds log-likelihoods (--log_scale)
  bool from_log_cells = false;  // cells are already in log space (text input path)
  bool ignore_miss = false;
  bool call_geno = false;
  double n_thresh = 0, call_thresh = 0;
};

int prepare_sites(const double *raw, uint64_t n_sites, uint64_t n_ind, const PrepOptions &o, int n_threads, double *gl,
                  double *expg, double *maf);
void pearson_site_terms(const double *expg, uint64_t n_sites, uint64_t n_ind, uint64_t n_pad, int n_threads,
                        uint64_t *dx_sig, uint16_t *dx_se, double *q);
void site_seeds(uint64_t seed, uint64_t n_sites, uint64_t *out);

// gsl_rng_taus (GSL rng/taus.c) seeded like gsl_rng_set; uniform() = get()/2^32 as gsl_rng_uniform.
struct TausStream {
  uint32_t a, b, c;
  explicit TausStream(uint64_t seed);
  uint32_t get() {
    a = ((a & 4294967294u) << 12) ^ (((a << 13) ^ a) >> 19);
    b = ((b & 4294967288u) << 4) ^ (((b << 2) ^ b) >> 25);
    c = ((c & 4294967280u) << 17) ^ (((c << 3) ^ c) >> 11);
    return a ^ b ^ c;
  }
  double uniform() { return get() / 4294967296.0; }
};

// Jump-ahead tables of the generator: each of its three components is a linear map over GF(2), so the state after
// 512 * 2^j draws is a 32 x 32 bit-matrix product.  jump[j][c][b] = column b (image of bit b) of component c's matrix
// for 512 * 2^j steps, j < TAUS_JUMP_LEVELS.  Lets a kernel start anywhere in a site's stream (aux::taus_sample_kernel).
constexpr int TAUS_JUMP_LEVELS = 32;
constexpr int TAUS_SEGMENT = 512;
void taus_jump_tables(uint32_t *jump /* [TAUS_JUMP_LEVELS][3][32] */);

}  // namespace hostprep
