// Host-side per-site preparation.  O(n_sites * n_ind) work that must match the reference's glibc
// results bit for bit (log/exp) and its x87 long-double Pearson recurrence, so it stays on the
// host CPU, spread over threads by site.  Compiled with g++ -ffp-contract=off (see Makefile).
#include "host_prep.h"

#include <math.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

namespace hostprep {

static const double kBig = 1e15;   // reference shared/gen_func.hpp:15 (INF)
static const double kEps = 1e-5;   // reference shared/gen_func.hpp:16 (EPSILON)

// log-space normaliser of one genotype triple: reference logsum(), shared/gen_func.cpp:135-151
static inline double log_norm3(const double *v) {
  double top = v[0];
  if (v[1] >= top) top = v[1];
  if (v[2] >= top) top = v[2];
  if (top == -INFINITY) return -INFINITY;
  double s = 0;
  s += exp(v[0] - top);
  s += exp(v[1] - top);
  s += exp(v[2] - top);
  return log(s) + top;
}

static inline bool flat_triple(const double *v) {  // miss_data(), shared/gen_func.cpp:862-868
  double a = v[0] - v[1], b = v[1] - v[2];
  if (!(a >= 0)) a = -a;
  if (!(b >= 0)) b = -b;
  return a < kEps && b < kEps;
}

// call_geno(geno, 3, log_scale=true, N_thresh, call_thresh, miss_data=0): shared/gen_func.cpp:886-914
static inline void call_triple(double *v, double n_thresh, double call_thresh) {
  int imax = 0, imin = 0;
  double vmax = -INFINITY, vmin = INFINITY;
  for (int g = 0; g < 3; g++) {
    if (v[g] > vmax) { vmax = v[g]; imax = g; }
    if (v[g] < vmin) { vmin = v[g]; imin = g; }
  }
  double best = exp(v[imax]);
  if (v[imin] == v[imax]) best = -1;
  if (best < n_thresh) {
    const double u = log((double)1 / 3);
    v[0] = v[1] = v[2] = u;
  }
  if (best >= call_thresh) {
    v[0] = v[1] = v[2] = -kBig;
    v[imax] = log(1);
  }
}

// est_maf() with indF == NULL: shared/gen_func.cpp:974-1009.  The accumulators carry over between
// passes exactly as in the reference (they are declared outside its do-loop).
static double site_maf(const double *lg, uint64_t n_ind, bool ignore_miss) {
  double num = 0, den = 0, freq = 0.01, before;
  int guard = 0;
  for (;;) {
    before = freq;
    for (uint64_t i = 0; i < n_ind; i++) {
      const double *c = lg + 3 * i;
      if (flat_triple(c) && ignore_miss) continue;
      double post[3] = {c[0], c[1], c[2]};
      const double z = log_norm3(post);
      for (int g = 0; g < 3; g++) {
        post[g] = exp(post[g] - z);
        if (post[g] == -INFINITY) post[g] = -kBig;
      }
      const double F = 0;
      num += post[1] + post[2] * (2 - F);
      den += 2 * post[1] + (post[0] + post[2]) * (2 - F);
    }
    freq = num / den;
    double d = before - freq;
    if (!(d >= 0)) d = -d;
    if (!(d > kEps)) break;
    if (!(guard++ < 100)) break;
  }
  return freq;
}

static int prepare_range(const double *raw, uint64_t s_lo, uint64_t s_hi, uint64_t n_ind, const PrepOptions &o,
                         double *gl, double *expg, double *maf) {
  for (uint64_t s = s_lo; s < s_hi; s++) {
    double *row = gl + s * n_ind * 3;
    const double *src = raw + s * n_ind * 3;
    for (uint64_t i = 0; i < n_ind; i++) {
      double *c = row + 3 * i;
      for (int g = 0; g < 3; g++) {
        double v = src[3 * i + g];
        if (!o.from_log_cells && !o.log_scale) {  // conv_space(log), shared/gen_func.cpp:123-130
          v = log(v);
          if (v == -INFINITY) v = -kBig;
        }
        c[g] = v;
      }
      const double z = log_norm3(c);  // post_prob(), shared/gen_func.cpp:920-932
      c[0] -= z;
      c[1] -= z;
      c[2] -= z;
      if (isnan(c[0]) || isnan(c[1]) || isnan(c[2])) return -1;  // read_data.cpp:42-45
    }
    if (o.call_geno)
      for (uint64_t i = 0; i < n_ind; i++) call_triple(row + 3 * i, o.n_thresh, o.call_thresh);
    maf[s] = site_maf(row, n_ind, o.ignore_miss);
    for (uint64_t i = 0; i < n_ind; i++) {  // ngsLD.cpp:107-114
      double *c = row + 3 * i;
      for (int g = 0; g < 3; g++) {
        c[g] = exp(c[g]);
        if (c[g] == -INFINITY) c[g] = -kBig;
      }
      expg[s * n_ind + i] = c[1] + 2 * c[2];
    }
  }
  return 0;
}

template <class F>
static void parallel_sites(uint64_t n_sites, int n_threads, F body) {
  if (n_threads < 1) n_threads = 1;
  const uint64_t block = 64;
  std::atomic<uint64_t> next(0);
  auto worker = [&]() {
    for (;;) {
      const uint64_t lo = next.fetch_add(block);
      if (lo >= n_sites) break;
      body(lo, lo + block < n_sites ? lo + block : n_sites);
    }
  };
  if (n_threads == 1 || n_sites <= block) {
    worker();
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; t++) pool.emplace_back(worker);
  for (auto &t : pool) t.join();
}

int prepare_sites(const double *raw, uint64_t n_sites, uint64_t n_ind, const PrepOptions &o, int n_threads, double *gl,
                  double *expg, double *maf) {
  std::atomic<int> bad(0);
  parallel_sites(n_sites, n_threads, [&](uint64_t lo, uint64_t hi) {
    if (prepare_range(raw, lo, hi, n_ind, o, gl, expg, maf) != 0) bad.store(1);
  });
  return bad.load() ? -1 : 0;
}

// Per-site half of gsl_stats_correlation's recurrence (GSL statistics/covariance_source.c, called at
// reference ngsLD.cpp:366) on the host FPU in native long double: delta_i = x_i - mean_(i-1) (the
// 80-bit value, stored as its memory image), sum_sq += delta*delta*ratio, mean += delta/(i+1.0);
// q = sqrt((double)sum_sq).  n_pad >= n_ind is the device row pitch.
void pearson_site_terms(const double *expg, uint64_t n_sites, uint64_t n_ind, uint64_t n_pad, int n_threads,
                        uint64_t *dx_sig, uint16_t *dx_se, double *q) {
  static_assert(sizeof(long double) >= 10, "x87 long double required");
  parallel_sites(n_sites, n_threads, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t s = lo; s < hi; s++) {
      const double *x = expg + s * n_ind;
      uint64_t *sig = dx_sig + s * n_pad;
      uint16_t *se = dx_se + s * n_pad;
      long double mean = x[0], ssq = 0.0L;
      sig[0] = 0;
      se[0] = 0;
      for (uint64_t i = 1; i < n_ind; i++) {
        const long double ratio = i / (i + 1.0);
        const long double delta = x[i] - mean;
        ssq += delta * delta * ratio;
        mean += delta / (i + 1.0);
        memcpy(&sig[i], &delta, 8);
        memcpy(&se[i], (const char *)&delta + 8, 2);
      }
      for (uint64_t i = n_ind; i < n_pad; i++) {
        sig[i] = 0;
        se[i] = 0;
      }
      q[s] = sqrt((double)ssq);
    }
  });
}

TausStream::TausStream(uint64_t seed) {
  if (seed == 0) seed = 1;
  a = (uint32_t)(69069ull * seed);
  b = (uint32_t)(69069ull * a);
  c = (uint32_t)(69069ull * b);
  for (int k = 0; k < 6; k++) get();
}

// reference ngsLD.cpp:69-70,165-166: seed_s = (unsigned long)(0 + uniform(master) * (1e15 - 0))
void site_seeds(uint64_t seed, uint64_t n_sites, uint64_t *out) {
  TausStream master(seed);
  const uint64_t lo = 0, hi = (uint64_t)kBig;
  for (uint64_t s = 0; s < n_sites; s++) {
    const double u = master.get() / 4294967296.0;
    out[s] = (uint64_t)(lo + u * (hi - lo));
  }
}

void taus_jump_tables(uint32_t *jump) {
  // one step of each component as a function of its own 32-bit state (GSL rng/taus.c)
  auto step = [](int c, uint32_t x) -> uint32_t {
    if (c == 0) return ((x & 4294967294u) << 12) ^ (((x << 13) ^ x) >> 19);
    if (c == 1) return ((x & 4294967288u) << 4) ^ (((x << 2) ^ x) >> 25);
    return ((x & 4294967280u) << 17) ^ (((x << 3) ^ x) >> 11);
  };
  auto apply = [](const uint32_t *cols, uint32_t x) -> uint32_t {
    uint32_t y = 0;
    for (int b = 0; b < 32; b++)
      if ((x >> b) & 1u) y ^= cols[b];
    return y;
  };
  for (int c = 0; c < 3; c++) {
    uint32_t m[32], sq[32];
    for (int b = 0; b < 32; b++) m[b] = step(c, 1u << b);  // the one-step matrix
    int steps = 1;
    for (int j = 0; j < TAUS_JUMP_LEVELS; j++) {
      while (steps < TAUS_SEGMENT || j > 0) {  // square up to 512 steps for level 0, then once per level
        for (int b = 0; b < 32; b++) sq[b] = apply(m, m[b]);
        for (int b = 0; b < 32; b++) m[b] = sq[b];
        steps *= 2;
        if (j > 0) break;
      }
      for (int b = 0; b < 32; b++) jump[((size_t)j * 3 + c) * 32 + b] = m[b];
    }
  }
}

}  // namespace hostprep
