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

}  // namespace hostprep
