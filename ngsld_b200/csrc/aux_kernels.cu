// Everything on the path that is not the fast EM: the bit-faithful EM, the x87-exact r2_ExpG,
// window -> pair-list expansion, per-site taus sampling, and an FP64 issue-rate probe.
#include "aux_kernels.cuh"
#include "fixed6.cuh"
#include "fp80.cuh"
#include "pearson.cuh"

namespace aux {

// ------------------------------------------------------------------------------------------------
// Bit-faithful EM: one thread per pair, individuals in order, every operation individually rounded
// in the reference's association order (shared/gen_func.cpp:1076-1119, 1027-1059).  Rounding-
// preserving hoists only: P[k][h] = f[k]*f[h] per pass, L[a][b] = p1[a]*p2[b] per individual (the
// reference's bracket p*q + p*q is exactly 2L).  The `sum` terms keep ((f_k f_h) p1) p2.
__global__ void __launch_bounds__(128) em_strict_kernel(SiteTable T, PairChunk C, int ignore_miss, DevCounters *ctr) {
  const size_t row_doubles = (size_t)T.n_pad * 3;
  unsigned long long passes = 0;
  for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < C.n_pairs;
       p += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t s1 = C.s1[p], s2 = C.s2[p];
    const double *ga = T.gl + (size_t)s1 * row_doubles, *gb = T.gl + (size_t)s2 * row_doubles;
    const double m1 = T.maf[s1], m2 = T.maf[s2];
    double f[4];
    f[0] = __dmul_rn(__dsub_rn(1.0, m1), __dsub_rn(1.0, m2));
    f[1] = __dmul_rn(__dsub_rn(1.0, m1), m2);
    f[2] = __dmul_rn(m1, __dsub_rn(1.0, m2));
    f[3] = __dmul_rn(m1, m2);
    uint32_t it, used = 0;
    for (it = 0; it < NGSLD_ITER_MAX; it++) {
      double P[4][4];
#pragma unroll
      for (int k = 0; k < 4; k++)
#pragma unroll
        for (int h = 0; h < 4; h++) P[k][h] = __dmul_rn(f[k], f[h]);
      double acc[4] = {0, 0, 0, 0};
      used = 0;
      for (uint32_t i = 0; i < T.n_ind; i++) {
        double pa[3], pb[3];
#pragma unroll
        for (int g = 0; g < 3; g++) {
          pa[g] = ga[3 * (size_t)i + g];
          pb[g] = gb[3 * (size_t)i + g];
        }
        if (ignore_miss && (gl_missing(pa[0], pa[1], pa[2]) || gl_missing(pb[0], pb[1], pb[2]))) continue;
        used++;
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
          for (int h = 0; h < 4; h++) {
            const int a = (k >> 1) + (h >> 1), b = (k & 1) + (h & 1);
            tot = __dadd_rn(tot, __dmul_rn(__dmul_rn(P[k][h], pa[a]), pb[b]));
          }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          double part = 0.0;
#pragma unroll
          for (int h = 0; h < 4; h++) {
            const int a = (k >> 1) + (h >> 1), b = (k & 1) + (h & 1);
            const double l = __dmul_rn(pa[a], pb[b]);
            part = __dadd_rn(part, __dmul_rn(P[k][h], __dadd_rn(l, l)));
          }
          acc[k] = __dadd_rn(acc[k], __ddiv_rn(part, tot));
        }
      }
      const double prev[4] = {f[0], f[1], f[2], f[3]};
      const double two_x = (double)(2ull * used);
#pragma unroll
      for (int k = 0; k < 4; k++) f[k] = __ddiv_rn(acc[k], two_x);
#pragma unroll
      for (int k = 0; k < 4; k++)  // sequential: later entries see the already-updated earlier ones
        f[k] = __ddiv_rn(f[k], __dadd_rn(__dadd_rn(__dadd_rn(f[0], f[1]), f[2]), f[3]));
      double eps = 0.0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const double d = fabs(__dsub_rn(f[k], prev[k]));
        if (d > eps) eps = d;
      }
      if (eps < NGSLD_EPS) break;
    }
    passes += it < NGSLD_ITER_MAX ? it + 1 : NGSLD_ITER_MAX;
    derive_and_store(C.rows + p, f, it, used);
  }
  // warp-aggregate the pass counter
  for (int o = 16; o > 0; o >>= 1) passes += __shfl_xor_sync(0xffffffffu, passes, o);
  if ((threadIdx.x & 31) == 0 && passes) atomicAdd(&ctr->em_passes, passes);
}

// ------------------------------------------------------------------------------------------------
// r2_ExpG (reference ngsLD.cpp:365-367): pearson::pair_r2 (pearson.cuh), one thread per pair.
__global__ void __launch_bounds__(128) pearson_kernel(SiteTable T, PairChunk C, DevCounters *ctr) {
  // Persistent: each warp repeatedly claims 32 consecutive pairs.  The launch is sized to ONE small CTA per SM when
  // the warp-per-pair EM kernel runs beside it: that kernel is bound by the FP64 pipe and leaves integer issue slots
  // and 10 K registers per SM free, so this integer-only kernel rides along for free instead of running before it.
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&ctr->next_pearson, 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= C.n_pairs) break;
    const unsigned long long p = base + lane;
    if (p >= C.n_pairs) continue;
    C.rows[p].r2_expg = pearson::pair_r2(T, C.s1[p], C.s2[p]);
  }
}

// ------------------------------------------------------------------------------------------------
// Per-site half of the same recurrence (gsl_stats_correlation, GSL statistics/covariance_source.c, called at
// reference ngsLD.cpp:366): for site s with expected genotypes x[0..n),
//     mean = x[0];  for i >= 1:  delta_i = x[i] - mean;  sum_sq += delta_i * delta_i * ratio_i;  mean += delta_i / (i + 1.0)
// all in x87 extended precision (ratio_i = i / (i + 1.0) is a double division widened).  Stores the 80-bit memory image
// of every delta_i and q = sqrt((double)sum_sq).  One thread per site; the recurrence is sequential in i.
__global__ void __launch_bounds__(128) site_terms_kernel(const double *expg, uint32_t n_sites, uint32_t n_ind, uint32_t n_blk,
                                                         uint64_t *dx_sig, uint16_t *dx_se, double *q, uint64_t *ratio) {
  // the ratio table shared by every pair: significand of (long double)(i / (i + 1.0))
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < 4u * n_blk; i += gridDim.x * blockDim.x)
    ratio[i] = i ? x87::ratio_sig(__ddiv_rn((double)i, __dadd_rn((double)i, 1.0))) : 0ull;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_sites; s += gridDim.x * blockDim.x) {
    const double *x = expg + (size_t)s * n_ind;
    // element i of site s lives at [(i / 4) * n_sites + s][i % 4]: four individuals are collected and stored together
    x87::ext mean = x87::from_double(x[0]);
    x87::ext ssq = x87::zero(0);
    for (uint32_t blk = 0; blk < n_blk; blk++) {
      uint64_t sig4[4];
      uint16_t se4[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t i = 4u * blk + j;
        sig4[j] = 0;
        se4[j] = 0;
        if (i == 0 || i >= n_ind) continue;
        const double ip1 = __dadd_rn((double)i, 1.0);
        const x87::ext ratio = x87::from_double(__ddiv_rn((double)i, ip1));
        x87::ext neg_mean = mean;
        neg_mean.neg ^= 1u;
        const x87::ext delta = x87::add(x87::from_double(x[i]), neg_mean);
        ssq = x87::add(ssq, x87::mul(x87::mul(delta, delta), ratio));
        mean = x87::add(mean, x87::div(delta, x87::from_double(ip1)));
        sig4[j] = delta.sig;
        se4[j] = x87::se14_pack(delta.neg, delta.exp, delta.sig);
      }
      const size_t at = (size_t)blk * n_sites + s;
      ulonglong2 *ps = reinterpret_cast<ulonglong2 *>(dx_sig) + 2 * at;
      ps[0] = make_ulonglong2(sig4[0], sig4[1]);
      ps[1] = make_ulonglong2(sig4[2], sig4[3]);
      reinterpret_cast<uint2 *>(dx_se)[at] =
          make_uint2((uint32_t)se4[0] | ((uint32_t)se4[1] << 16), (uint32_t)se4[2] | ((uint32_t)se4[3] << 16));
    }
    q[s] = __dsqrt_rn(x87::to_double(ssq));
  }
}

// ------------------------------------------------------------------------------------------------
// Window plan -> explicit pair list for output rows [row_lo, row_lo + n): row g belongs to the
// compact first site c1 with row_off[c1] <= g < row_off[c1+1]; its partner is c1 + 1 + (g - row_off[c1]).
__global__ void expand_window_kernel(const unsigned long long *row_off, const uint32_t *cs, uint32_t n_compact,
                                     unsigned long long row_lo, unsigned long long n, uint32_t *s1, uint32_t *s2) {
  for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < n;
       p += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long g = row_lo + p;
    uint32_t lo = 0, hi = n_compact;  // largest c with row_off[c] <= g
    while (hi - lo > 1) {
      const uint32_t mid = lo + (hi - lo) / 2;
      if (row_off[mid] <= g) lo = mid; else hi = mid;
    }
    const uint32_t c1 = lo, c2 = c1 + 1 + (uint32_t)(g - row_off[c1]);
    s1[p] = cs ? cs[c1] : c1;
    s2[p] = cs ? cs[c2] : c2;
  }
}

// site indices + accumulated distance (reference ngsLD.cpp:241) into the output rows
__global__ void fill_rows_kernel(SiteTable T, PairChunk C) {
  for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < C.n_pairs;
       p += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t a = C.s1[p], b = C.s2[p];
    double dist = __longlong_as_double(0x7ff0000000000000ll);
    if (T.cum != nullptr && T.seg[a] == T.seg[b]) dist = T.cum[b] - T.cum[a];
    ngsld_pair_row *r = C.rows + p;
    r->dist = dist;
    r->s1 = a;
    r->s2 = b;
    r->reserved = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// Random pair sampling (reference ngsLD.cpp:277 with the per-site gsl_rng_taus of ngsLD.cpp:165-166;
// generator restated from GSL rng/taus.c).  One thread walks one first site's candidate stream.
struct Taus {
  uint32_t a, b, c;
  __device__ __forceinline__ uint32_t get() {
    a = ((a & 4294967294u) << 12) ^ (((a << 13) ^ a) >> 19);
    b = ((b & 4294967288u) << 4) ^ (((b << 2) ^ b) >> 25);
    c = ((c & 4294967280u) << 17) ^ (((c << 3) ^ c) >> 11);
    return a ^ b ^ c;
  }
  __device__ __forceinline__ void set(unsigned long long seed) {
    if (seed == 0) seed = 1;
    a = (uint32_t)(69069ull * seed);
    b = (uint32_t)(69069ull * a);
    c = (uint32_t)(69069ull * b);
    for (int k = 0; k < 6; k++) get();
  }
};

// One CTA per first site; the site's candidate stream is cut into segments of TAUS_SEG draws and every thread takes one
// segment at a time, starting in the middle of the stream: each component of the generator is linear over GF(2), so the
// state after 512 * 2^j draws is a bit-matrix product with a precomputed table (hostprep::taus_jump_tables).  A site with
// a million candidates (BASELINE config 5) is thus walked by the whole CTA at once instead of by a single thread.
// A draw is kept iff !(get() / 2^32 > rnd_sample)  <=>  get() <= keep_max = floor(rnd_sample * 2^32) (both scalings by a
// power of two are exact).
// mode 0: counts[c1 - c_lo] = kept pairs; mode 1: write the kept partners at row_off[c1] - row_base.
constexpr int TAUS_SEG = 512;
constexpr int TAUS_CTA = 256;

__device__ __forceinline__ uint32_t gf2_apply(const uint32_t *cols, uint32_t x) {
  uint32_t y = 0;
#pragma unroll 8
  for (int b = 0; b < 32; b++) y ^= cols[b] & (0u - ((x >> b) & 1u));
  return y;
}

__global__ void __launch_bounds__(TAUS_CTA) taus_sample_kernel(const unsigned long long *site_seeds, const uint32_t *cs,
                                                               const uint32_t *cw_end, uint32_t c_lo, uint32_t c_hi,
                                                               uint32_t keep_max, int mode, unsigned long long *counts,
                                                               const unsigned long long *row_off, unsigned long long row_base,
                                                               unsigned long long row_cap, uint32_t *s1, uint32_t *s2,
                                                               const uint32_t *jump) {
  __shared__ unsigned int warp_tot[TAUS_CTA / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (uint32_t c1 = c_lo + blockIdx.x; c1 < c_hi; c1 += gridDim.x) {
    const uint32_t site = cs ? cs[c1] : c1;
    const uint32_t first = c1 + 1, end = cw_end[c1];
    const unsigned long long W = end > first ? end - first : 0u;  // candidate partners, in stream order
    Taus g0;
    g0.set(site_seeds[site]);
    const unsigned long long n_seg = (W + TAUS_SEG - 1) / TAUS_SEG;
    const unsigned long long base = mode ? row_off[c1] - row_base : 0;  // may wrap: rows before the chunk fail o < row_cap
    unsigned long long running = 0;  // kept draws of the segments before this round (the same in every thread)
    for (unsigned long long seg0 = 0; seg0 < n_seg; seg0 += TAUS_CTA) {
      const unsigned long long k = seg0 + tid;
      Taus g = g0;
      unsigned int cnt = 0;
      const unsigned long long lo = k * TAUS_SEG, hi = lo + TAUS_SEG < W ? lo + TAUS_SEG : W;
      if (k < n_seg) {
        for (int j = 0; (k >> j) != 0; j++)
          if ((k >> j) & 1ull) {
            g.a = gf2_apply(jump + (j * 3 + 0) * 32, g.a);
            g.b = gf2_apply(jump + (j * 3 + 1) * 32, g.b);
            g.c = gf2_apply(jump + (j * 3 + 2) * 32, g.c);
          }
        Taus w = g;
        for (unsigned long long i = lo; i < hi; i++) cnt += w.get() <= keep_max ? 1u : 0u;
      }
      // exclusive prefix of cnt over the CTA
      unsigned int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) warp_tot[warp] = incl;
      __syncthreads();
      unsigned int before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < TAUS_CTA / 32; w++) {
        const unsigned int t = warp_tot[w];
        if (w < warp) before += t;
        total += t;
      }
      __syncthreads();
      if (mode && k < n_seg && cnt) {
        unsigned long long o = base + running + before + (incl - cnt);
        for (unsigned long long i = lo; i < hi; i++)
          if (g.get() <= keep_max) {
            if (o < row_cap) {
              s1[o] = site;
              const uint32_t c2 = first + (uint32_t)i;
              s2[o] = cs ? cs[c2] : c2;
            }
            o++;
          }
      }
      running += total;
    }
    if (!mode && tid == 0) counts[c1 - c_lo] = running;
  }
}

// ------------------------------------------------------------------------------------------------
// LD-decay bins (the binning of the reference's scripts/fit_LDdecay.R:133-150, see ngsld_scan_decay): one thread
// per row; rows arrive sorted by (s1, s2), so neighbouring lanes mostly hit the same bin -> lanes with equal bins
// are combined in the warp (match.any) and one lane per distinct bin issues the atomics.
__global__ void __launch_bounds__(256) decay_bins_kernel(const ngsld_pair_row *rows, unsigned long long n, double bin_size,
                                                         unsigned long long n_bins, ngsld_decay_bin *bins,
                                                         unsigned long long *outside) {
  const int lane = threadIdx.x & 31;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long n_round = (n + 31ull) & ~31ull;  // whole warps stay together for the warp collectives
  for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < n_round; p += stride) {
    long long bin = -1;
    double v[4] = {0, 0, 0, 0};
    bool live = p < n;
    if (live) {
      const ngsld_pair_row r = rows[p];
      v[0] = r.r2_expg; v[1] = r.D; v[2] = r.Dp; v[3] = r.r2;
      if (isfinite(r.dist)) {
        bin = (long long)ceil(r.dist / bin_size) - 1;  // right-closed bins (k*b, (k+1)*b]
        if (bin < 0) bin = 0;
      }
      if (bin < 0 || (unsigned long long)bin >= n_bins) bin = -1;
    }
    const unsigned out_mask = __ballot_sync(0xffffffffu, live && bin < 0);
    if (lane == 0 && out_mask) atomicAdd(outside, (unsigned long long)__popc(out_mask));
    const unsigned peers = __match_any_sync(0xffffffffu, bin);  // lanes of this warp with the same bin
    const int leader = __ffs(peers) - 1;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const bool fin = bin >= 0 && isfinite(v[j]);
      double s = fin ? v[j] : 0.0;
      unsigned cnt = fin ? 1u : 0u;
      // sum over the peer group: every lane walks its own peer mask (groups are disjoint)
      double tot = 0.0;
      unsigned tc = 0;
      for (unsigned m = peers; m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        tot += __shfl_sync(peers, s, src);
        tc += __shfl_sync(peers, cnt, src);
      }
      if (lane == leader && bin >= 0 && tc) {
        atomicAdd(&bins[bin].sum[j], tot);
        atomicAdd((unsigned long long *)&bins[bin].n[j], (unsigned long long)tc);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LD pruning, device half (ngsld_scan_edges): the row filter of the reference's scripts/prune_graph.pl:118-137 applied
// right behind the EM.  The weight is what the script parses from the TSV: the value rounded to six decimals
// (fmt::fixed6, exactly as "%f" prints it) read back as a double.  Rows that pass are appended to `edges` (the lanes of
// a warp reserve their slots with one atomic); every site that occurs in a row is flagged in `seen`.
__global__ void __launch_bounds__(256) prune_edges_kernel(const ngsld_pair_row *rows, unsigned long long n, ngsld_prune_params q,
                                                          double precision, ngsld_edge *edges, unsigned long long *n_edges,
                                                          unsigned char *seen) {
  const int lane = threadIdx.x & 31;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long n_round = (n + 31ull) & ~31ull;
  for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < n_round; p += stride) {
    bool edge = false;
    ngsld_edge e;
    e.s1 = e.s2 = 0;
    e.label = 0;
    if (p < n) {
      const ngsld_pair_row r = rows[p];
      seen[r.s1] = 1;
      seen[r.s2] = 1;
      const double x = q.field == 4 ? r.r2_expg : q.field == 5 ? r.D : q.field == 6 ? r.Dp : r.r2;
      if (isfinite(x) && isfinite(r.dist) && !(r.dist > q.max_dist)) {
        unsigned long long N;
        double w = fmt::fixed6(x, N) ? __ddiv_rn((double)N, 1000000.0) : fabs(x);  // the printed decimal, parsed back
        if (x < 0) w = -w;
        if (q.weight_type == 'a') w = fabs(w);
        if (!(w < q.min_weight)) {
          if (q.weight_type == 'n') w = 1.0;
          edge = true;
          e.s1 = r.s1;
          e.s2 = r.s2;
          e.label = (int32_t)__dmul_rn(w, precision);  // int(): truncation towards zero
        }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, edge);
    if (m) {
      unsigned long long base = 0;
      if (lane == __ffs(m) - 1) base = atomicAdd(n_edges, (unsigned long long)__popc(m));
      base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
      if (edge) edges[base + __popc(m & ((1u << lane) - 1u))] = e;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K0 (opt-in): the per-cell preparation of the reference on the device -- read_geno()'s log +
// normalisation (shared/read_data.cpp:28-46 with conv_space / post_prob / logsum, gen_func.cpp:123-151, 920-932), the
// optional call_geno() pass (gen_func.cpp:886-914), est_maf() (gen_func.cpp:974-1009) and the exp + expected-genotype
// loop of main (ngsLD.cpp:107-114).  One warp per site, in place on the uploaded file cells.  CUDA's log/exp are not
// glibc's: likelihoods differ from the host path by a few ulp (and the allele frequency, summed as a tree instead of
// in individual order, by ~sqrt(n_ind) ulp), so results are NOT bit-identical to the reference with this path; the
// default stays ngsld_prepare_sites on the host.
__device__ __forceinline__ double log_norm3_dev(double a, double b, double c) {  // logsum(), gen_func.cpp:135-151
  double top = a;
  if (b >= top) top = b;
  if (c >= top) top = c;
  if (top == -INFINITY) return -INFINITY;
  double s = 0;
  s += exp(a - top);
  s += exp(b - top);
  s += exp(c - top);
  return log(s) + top;
}

__global__ void __launch_bounds__(128) prep_sites_kernel(double *gl, uint32_t n_sites, uint32_t n_ind, uint32_t n_pad, int to_log,
                                                         int ignore_miss, int call_geno, double n_thresh, double call_thresh,
                                                         double *expg, double *maf, int *nan_flag) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const double big = 1e15;
  for (uint32_t s = warp; s < n_sites; s += n_warps) {
    double *row = gl + (size_t)s * n_pad * 3;
    double num = 0, den = 0;
    for (uint32_t i = lane; i < n_ind; i += 32) {
      double v[3];
#pragma unroll
      for (int g = 0; g < 3; g++) {
        double x = row[3 * (size_t)i + g];
        if (to_log) {  // conv_space(log)
          x = log(x);
          if (x == -INFINITY) x = -big;
        }
        v[g] = x;
      }
      const double z = log_norm3_dev(v[0], v[1], v[2]);  // post_prob()
      v[0] -= z; v[1] -= z; v[2] -= z;
      if (v[0] != v[0] || v[1] != v[1] || v[2] != v[2]) *nan_flag = 1;  // read_data.cpp:42-45
      if (call_geno) {  // call_geno(geno, 3, log_scale = true, N_thresh, call_thresh, 0)
        int imax = 0, imin = 0;
        double vmax = -INFINITY, vmin = INFINITY;
#pragma unroll
        for (int g = 0; g < 3; g++) {
          if (v[g] > vmax) { vmax = v[g]; imax = g; }
          if (v[g] < vmin) { vmin = v[g]; imin = g; }
        }
        double best = exp(v[imax]);
        if (v[imin] == v[imax]) best = -1;
        if (best < n_thresh) v[0] = v[1] = v[2] = log(1.0 / 3.0);
        if (best >= call_thresh) {
          v[0] = v[1] = v[2] = -big;
          v[imax] = 0.0;
        }
      }
      // est_maf(): posterior under the uniform prior, accumulated over the individuals
      double a = v[0] - v[1], b = v[1] - v[2];
      if (!(a >= 0)) a = -a;
      if (!(b >= 0)) b = -b;
      const bool flat = a < NGSLD_EPS && b < NGSLD_EPS;
      if (!(flat && ignore_miss)) {
        const double z2 = log_norm3_dev(v[0], v[1], v[2]);
        const double p0 = exp(v[0] - z2), p1 = exp(v[1] - z2), p2 = exp(v[2] - z2);
        num += p1 + p2 * 2.0;
        den += 2.0 * p1 + (p0 + p2) * 2.0;
      }
      // normal space + expected genotype (ngsLD.cpp:107-114)
      const double e0 = exp(v[0]), e1 = exp(v[1]), e2 = exp(v[2]);
      row[3 * (size_t)i] = e0;
      row[3 * (size_t)i + 1] = e1;
      row[3 * (size_t)i + 2] = e2;
      expg[(size_t)s * n_ind + i] = e1 + 2.0 * e2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      num += __shfl_xor_sync(0xffffffffu, num, o);
      den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    if (lane == 0) maf[s] = num / den;
  }
}

// ------------------------------------------------------------------------------------------------
// FP64 FMA issue-rate probe: 8 independent chains per thread.
__global__ void fp64_probe_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
    a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace aux
