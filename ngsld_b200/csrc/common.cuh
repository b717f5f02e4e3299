// Internal types shared by the kernels and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ngsld_b200.h"

#define NGSLD_EPS 1e-5      // reference shared/gen_func.hpp:16
#define NGSLD_ITER_MAX 100  // reference shared/gen_func.hpp:18

// Device-resident site table (read-only during a scan).
struct SiteTable {
  const double *gl;        // [n_sites][n_pad][3] normal-space genotype likelihoods, rows 16-byte aligned
  const double *maf;       // [n_sites]
  const uint64_t *dx_sig;  // [n_blk][n_sites][4] x87 significand of the Pearson deviation x[i]-mean_(i-1) of individual
                           // i = 4 blk + j (zero for i = 0 and behind the last individual): the 32 pairs of a warp share
                           // s1 and have consecutive s2, so a block row is read as one coalesced 1 KB request
  const uint16_t *dx_se;   // [n_blk][n_sites][4] sign and exponent of the same, packed for mac3 (fp80.cuh, "se14")
  const double *q;         // [n_sites] sqrt((double)sum_xsq)
  const uint64_t *ratio;   // [4 n_blk] x87 significand of (long double)(i / (i + 1.0)) (exponent -1); entry 0 unused
  const double *cum;       // [n_sites] exact prefix sum of finite pos_dist (NULL: no positions)
  const uint32_t *seg;     // [n_sites] chromosome segment id (increments at each +inf pos_dist)
  // site palettes (em_cell.cuh): the DISTINCT genotype-likelihood triples of a site and, per individual, which one it has
  const uint8_t *cls;      // [n_sites][n_cpad] class of individual i = index of its triple in the site's palette
  const double *pal;       // [n_sites][NGSLD_KMAX][3] palette, classes in order of first appearance
  const uint8_t *pal_k;    // [n_sites] number of classes; 0 = more than NGSLD_KMAX distinct triples (site not coded)
  const uint64_t *pal_miss;  // [n_sites] bit c set: class c is "missing data" (flat triple, gl_missing)
  uint32_t n_sites, n_ind, n_pad, n_cpad;
  uint32_t n_blk;          // (n_ind + 3) / 4: blocks of four individuals in dx_sig / dx_se
};

#define NGSLD_KMAX 64  // classes per site palette (joint classes of a pair index a 64 x 64 table of 16-bit counts)

// One chunk of output rows: pair p (0-based inside the chunk) is (s1[p], s2[p]) -> rows[p].
struct PairChunk {
  const uint32_t *s1, *s2;
  ngsld_pair_row *rows;
  uint64_t n_pairs;
};

// Counters a scan accumulates on the device.  The first NGSLD_WORK_COUNTERS fields are reset per chunk.
#define NGSLD_WORK_COUNTERS 5
struct DevCounters {
  unsigned long long next_pair;     // dynamic work counter (list / warp / cell EM kernels)
  unsigned long long next_tile;     // dynamic work counter (tile kernel)
  unsigned long long next_pearson;  // dynamic work counter (r2_ExpG kernel)
  unsigned long long next_resid;    // dynamic work counter of the dense kernel over the cell kernel's left-over pairs
  unsigned long long n_resid;       // pairs of this chunk the cell kernel handed to the dense kernel
  unsigned long long em_passes;     // total EM passes executed
  unsigned long long cell_passes;   // sum over pairs of (weighted cells x passes) executed by the cell kernel
  unsigned long long cells;         // sum over pairs of weighted cells (cell kernel)
  unsigned long long cell_pairs;    // pairs the cell kernel computed itself
  unsigned long long resid_pairs;   // pairs handed to the dense kernel (whole scan)
};

// ---- arithmetic in the reference's operation order (no contraction) -----------------------------
__device__ __forceinline__ double ref_min(double a, double b) { return a <= b ? a : b; }  // gen_func.hpp:22

// D, D', r2, hap_maf, chi2 from converged haplotype frequencies: reference ngsLD.cpp:296-306,328-333.
// Every operation is an explicitly rounded intrinsic so the result is bit-identical to x86 SSE2
// given identical f[].
__device__ __forceinline__ void derive_and_store(ngsld_pair_row *row, const double f[4], uint32_t n_iter,
                                                 uint32_t n_used) {
  const double m0 = __dsub_rn(1.0, __dadd_rn(f[0], f[1]));
  const double m1 = __dsub_rn(1.0, __dadd_rn(f[0], f[2]));
  const double D = __dsub_rn(__dmul_rn(f[0], f[3]), __dmul_rn(f[1], f[2]));
  const double om0 = __dsub_rn(1.0, m0), om1 = __dsub_rn(1.0, m1);
  double den;
  if (D < 0)
    den = -ref_min(__dmul_rn(m0, m1), __dmul_rn(om0, om1));
  else
    den = ref_min(__dmul_rn(m0, om1), __dmul_rn(om0, m1));
  const double Dp = __ddiv_rn(D, den);
  const double prod = __dmul_rn(__dmul_rn(__dmul_rn(m0, m1), om0), om1);
  const double qq = __ddiv_rn(D, __dsqrt_rn(prod));
  const float fa = (float)__dadd_rn(f[0], f[1]);
  const float fb = (float)__dadd_rn(f[0], f[2]);
  const float ofa = __fsub_rn(1.0f, fa), ofb = __fsub_rn(1.0f, fb);
  const float e[4] = {__fmul_rn(fa, fb), __fmul_rn(fa, ofb), __fmul_rn(ofa, fb), __fmul_rn(ofa, ofb)};
  float chi2 = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const double d = __dsub_rn(f[k], (double)e[k]);
    const double term = __ddiv_rn(__dmul_rn(d, d), (double)e[k]);
    chi2 = (float)__dadd_rn((double)chi2, term);
  }
  row->D = D;
  row->Dp = Dp;
  row->r2 = __dmul_rn(qq, qq);
  row->hap[0] = f[0];
  row->hap[1] = f[1];
  row->hap[2] = f[2];
  row->hap[3] = f[3];
  row->hap_maf[0] = m0;
  row->hap_maf[1] = m1;
  row->chi2 = chi2;
  row->n_iter = n_iter;
  row->n_used = n_used;
}

// "all three likelihoods equal within 1e-5": reference shared/gen_func.cpp:862-868
__device__ __forceinline__ bool gl_missing(double g0, double g1, double g2) {
  double a = __dsub_rn(g0, g1), b = __dsub_rn(g1, g2);
  a = a >= 0 ? a : -a;
  b = b >= 0 ? b : -b;
  return a < NGSLD_EPS && b < NGSLD_EPS;
}
