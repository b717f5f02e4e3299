// Declarations of the non-EM-fast kernels (definitions in aux_kernels.cu).
#pragma once
#include "common.cuh"

namespace aux {
__global__ void em_strict_kernel(SiteTable T, PairChunk C, int ignore_miss, DevCounters *ctr);
__global__ void pearson_kernel(SiteTable T, PairChunk C, DevCounters *ctr);
__global__ void site_terms_kernel(const double *expg, uint32_t n_sites, uint32_t n_ind, uint32_t n_blk, uint64_t *dx_sig,
                                  uint16_t *dx_se, double *q, uint64_t *ratio);
__global__ void expand_window_kernel(const unsigned long long *row_off, const uint32_t *cs, uint32_t n_compact,
                                     unsigned long long row_lo, unsigned long long n, uint32_t *s1, uint32_t *s2);
__global__ void fill_rows_kernel(SiteTable T, PairChunk C);
__global__ void taus_sample_kernel(const unsigned long long *site_seeds, const uint32_t *cs, const uint32_t *cw_end,
                                   uint32_t c_lo, uint32_t c_hi, uint32_t keep_max, int mode, unsigned long long *counts,
                                   const unsigned long long *row_off, unsigned long long row_base, unsigned long long row_cap,
                                   uint32_t *s1, uint32_t *s2, const uint32_t *jump);
__global__ void decay_bins_kernel(const ngsld_pair_row *rows, unsigned long long n, double bin_size,
                                  unsigned long long n_bins, ngsld_decay_bin *bins, unsigned long long *outside);
__global__ void prune_edges_kernel(const ngsld_pair_row *rows, unsigned long long n, ngsld_prune_params q, double precision,
                                   ngsld_edge *edges, unsigned long long *n_edges, unsigned char *seen);
__global__ void prep_sites_kernel(double *gl, uint32_t n_sites, uint32_t n_ind, uint32_t n_pad, int to_log, int ignore_miss,
                                  int call_geno, double n_thresh, double call_thresh, double *expg, double *maf, int *nan_flag);
__global__ void fp64_probe_kernel(double *out, int iters);
}  // namespace aux
