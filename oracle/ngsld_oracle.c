/* ============================================================================================
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement of the pairwise-LD hot path of fgvieira/ngsLD (reference @ 596bec1f, "1.2.1"),
 * on flat arrays, used ONLY as the parity checker by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  Nothing under ngsld_b200/ links, imports or
 * calls this file; the product path fails loudly when its CUDA library is missing.
 *
 * Parity status: PINNED for everything except the GSL boundary.  tests/test_oracle.py checks this
 * restatement md5-for-md5 against outputs of the unmodified reference binary (oracle/_ref/ngsLD,
 * built by oracle/Makefile from /root/reference with a two-header GSL stand-in) on every fixture in
 * tests/golden/.  The two GSL routines on the path (taus RNG, stats_correlation) are not in the
 * reference tree; they are restated from GSL's published algorithm.  The RNG is pinned by GSL's own
 * known answer (seed 1 -> 10000th output 2733957125); gsl_stats_correlation is "parity unpinned"
 * (no libgsl in this image) -- see DESIGN.md.
 *
 * Arithmetic contract (reference Makefile:9): -O3, SSE2 doubles, no FMA contraction, no fast-math,
 * x87 long double in the Pearson recurrence.  Build: oracle/Makefile (-ffp-contract=off).
 * ============================================================================================ */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EPS 1e-5      /* reference shared/gen_func.hpp:16 EPSILON  */
#define ORC_BIG 1e15      /* reference shared/gen_func.hpp:15 INF      */
#define ORC_ITER_MAX 100  /* reference shared/gen_func.hpp:18 ITER_MAX */

/* ---------------------------------------------------------------- taus RNG (GSL rng/taus.c) -- */
typedef struct { uint32_t a, b, c; } orc_taus;

uint32_t orc_taus_get(orc_taus *r) {
  r->a = ((r->a & 4294967294u) << 12) ^ (((r->a << 13) ^ r->a) >> 19);
  r->b = ((r->b & 4294967288u) << 4) ^ (((r->b << 2) ^ r->b) >> 25);
  r->c = ((r->c & 4294967280u) << 17) ^ (((r->c << 3) ^ r->c) >> 11);
  return r->a ^ r->b ^ r->c;
}

void orc_taus_set(orc_taus *r, uint64_t seed) {
  if (seed == 0) seed = 1;
  r->a = (uint32_t)(69069ull * seed);
  r->b = (uint32_t)(69069ull * r->a);
  r->c = (uint32_t)(69069ull * r->b);
  for (int k = 0; k < 6; k++) orc_taus_get(r);
}

/* gsl_rng_uniform; the reference wraps it as min + u*(max-min) (shared/gen_func.cpp:117-119). */
double orc_taus_uniform(orc_taus *r) { return orc_taus_get(r) / 4294967296.0; }

/* Per-site generator seeds, drawn serially from the master stream in s1 order
 * (reference ngsLD.cpp:69-70,165-166: seed_s1 = (unsigned long)(0 + u*(1e15 - 0))). */
void orc_site_seeds(uint64_t seed, uint64_t n_sites, uint64_t *out) {
  orc_taus m;
  orc_taus_set(&m, seed);
  for (uint64_t s = 0; s < n_sites; s++) {
    uint64_t lo = 0, hi = (uint64_t)ORC_BIG;
    out[s] = (uint64_t)(lo + orc_taus_uniform(&m) * (hi - lo));
  }
}

/* ------------------------------------------------------------------- per-cell preprocessing -- */
static double orc_logsum3(const double *a) { /* shared/gen_func.cpp:135-151 */
  double m = a[0];
  for (int k = 1; k < 3; k++) m = (a[k] >= m ? a[k] : m);
  if (m == -INFINITY) return -INFINITY;
  double acc = 0;
  for (int k = 0; k < 3; k++) acc += exp(a[k] - m);
  return log(acc) + m;
}

static int orc_missing(const double *g) { /* shared/gen_func.cpp:862-868 (abs is a macro there) */
  double d01 = g[0] - g[1], d12 = g[1] - g[2];
  d01 = d01 >= 0 ? d01 : -d01;
  d12 = d12 >= 0 ? d12 : -d12;
  return d01 < ORC_EPS && d12 < ORC_EPS;
}

/* shared/gen_func.cpp:886-914 with log_scale = true, miss_data = 0 (the only way ngsLD.cpp:98 calls it) */
static void orc_call_geno(double *g, double n_thresh, double call_thresh) {
  int hi = 0, lo = 0;
  double vmax = -INFINITY, vmin = INFINITY;
  for (int k = 0; k < 3; k++) {
    if (g[k] > vmax) { vmax = g[k]; hi = k; }
    if (g[k] < vmin) { vmin = g[k]; lo = k; }
  }
  double top = exp(g[hi]);
  if (g[lo] == g[hi]) top = -1;
  if (top < n_thresh)
    for (int k = 0; k < 3; k++) g[k] = log((double)1 / 3);
  if (top >= call_thresh) {
    for (int k = 0; k < 3; k++) g[k] = -ORC_BIG;
    g[hi] = log(1);
  }
}

/* shared/gen_func.cpp:974-1009, indF == NULL branch.  num/den deliberately NOT reset per pass. */
static double orc_est_maf(const double *lg /*[n_ind][3] log space*/, uint64_t n_ind, int ignore_miss) {
  int iters = 0;
  double num = 0, den = 0, prev, freq = 0.01;
  do {
    prev = freq;
    for (uint64_t i = 0; i < n_ind; i++) {
      const double *g = lg + 3 * i;
      if (orc_missing(g) && ignore_miss) continue;
      double pp[3] = {g[0], g[1], g[2]};
      double norm = orc_logsum3(pp);
      for (int k = 0; k < 3; k++) pp[k] -= norm;
      for (int k = 0; k < 3; k++) {
        pp[k] = exp(pp[k]);
        if (pp[k] == -INFINITY) pp[k] = -ORC_BIG;
      }
      double F = 0;
      num += pp[1] + pp[2] * (2 - F);
      den += 2 * pp[1] + (pp[0] + pp[2]) * (2 - F);
    }
    freq = num / den;
    double d = prev - freq;
    d = d >= 0 ? d : -d;
    if (!(d > ORC_EPS)) break;
  } while (iters++ < 100);
  return freq;
}

/* Binary-input preprocessing: reference shared/read_data.cpp:28-46 (log, clamp, normalise, NaN check),
 * ngsLD.cpp:92-98 (optional genotype calling), :103-104 (maf), :107-114 (exp, expected genotype).
 * raw: [n_sites][n_ind][3] as in the file.  Outputs: gl [n_sites][n_ind][3] normal space,
 * expg [n_sites][n_ind], maf [n_sites].  Returns 0, or -1 if a NaN appears (reference: fatal). */
int orc_preprocess(const double *raw, uint64_t n_sites, uint64_t n_ind, int log_scale, int ignore_miss,
                   int call_geno, double n_thresh, double call_thresh,
                   double *gl, double *expg, double *maf) {
  for (uint64_t s = 0; s < n_sites; s++) {
    double *row = gl + s * n_ind * 3;
    for (uint64_t i = 0; i < n_ind; i++) {
      double *g = row + 3 * i;
      const double *src = raw + (s * n_ind + i) * 3;
      for (int k = 0; k < 3; k++) {
        g[k] = src[k];
        if (!log_scale) {
          g[k] = log(g[k]);
          if (g[k] == -INFINITY) g[k] = -ORC_BIG;
        }
      }
      double norm = orc_logsum3(g);
      for (int k = 0; k < 3; k++) g[k] -= norm;
      if (isnan(g[0]) || isnan(g[1]) || isnan(g[2])) return -1;
    }
    if (call_geno)
      for (uint64_t i = 0; i < n_ind; i++) orc_call_geno(row + 3 * i, n_thresh, call_thresh);
    maf[s] = orc_est_maf(row, n_ind, ignore_miss);
    for (uint64_t i = 0; i < n_ind; i++) {
      double *g = row + 3 * i;
      for (int k = 0; k < 3; k++) {
        g[k] = exp(g[k]);
        if (g[k] == -INFINITY) g[k] = -ORC_BIG;
      }
      expg[s * n_ind + i] = g[1] + 2 * g[2];
    }
  }
  return 0;
}

/* Log-space normalised cells straight from parsed text input (read_data.cpp:83-97 already took the
 * log); continues with calling / maf / exp exactly as above.  lg is overwritten with normal space. */
int orc_preprocess_logcells(double *lg, uint64_t n_sites, uint64_t n_ind, int ignore_miss,
                            int call_geno, double n_thresh, double call_thresh,
                            double *expg, double *maf) {
  for (uint64_t s = 0; s < n_sites; s++) {
    double *row = lg + s * n_ind * 3;
    for (uint64_t i = 0; i < n_ind; i++) {
      double *g = row + 3 * i;
      double norm = orc_logsum3(g);
      for (int k = 0; k < 3; k++) g[k] -= norm;
    }
    if (call_geno)
      for (uint64_t i = 0; i < n_ind; i++) orc_call_geno(row + 3 * i, n_thresh, call_thresh);
    maf[s] = orc_est_maf(row, n_ind, ignore_miss);
    for (uint64_t i = 0; i < n_ind; i++) {
      double *g = row + 3 * i;
      for (int k = 0; k < 3; k++) {
        g[k] = exp(g[k]);
        if (g[k] == -INFINITY) g[k] = -ORC_BIG;
      }
      expg[s * n_ind + i] = g[1] + 2 * g[2];
    }
  }
  return 0;
}

/* ------------------------------------------------------ Pearson (GSL covariance_source.c) ---- */
double orc_pearson_r2(const double *x, const double *y, uint64_t n) { /* ngsLD.cpp:365-367 */
  long double sxx = 0.0L, syy = 0.0L, sxy = 0.0L, mx = x[0], my = y[0];
  for (uint64_t i = 1; i < n; i++) {
    long double ratio = i / (i + 1.0);
    long double dx = x[i] - mx, dy = y[i] - my;
    sxx += dx * dx * ratio;
    syy += dy * dy * ratio;
    sxy += dx * dy * ratio;
    mx += dx / (i + 1.0);
    my += dy / (i + 1.0);
  }
  long double r = sxy / (sqrt((double)sxx) * sqrt((double)syy));
  double rd = (double)r;
  return rd * rd; /* pow(r, 2) */
}

/* ------------------------------------------------------------- haplotype-frequency EM -------- */
/* One EM pass: shared/gen_func.cpp:1076-1119.  Haplotype index bit1 = allele at site 1, bit0 = allele
 * at site 2, so an ordered haplotype pair (k,h) implies genotypes ga = (k>>1)+(h>>1), gb = (k&1)+(h&1). */
static uint64_t orc_em_pass(double f[4], const double *ga, const double *gb, uint64_t n_ind, int ignore_miss) {
  double acc[4] = {0, 0, 0, 0};
  uint64_t used = 0;
  for (uint64_t i = 0; i < n_ind; i++) {
    const double *p = ga + 3 * i, *q = gb + 3 * i;
    if ((orc_missing(p) || orc_missing(q)) && ignore_miss) continue;
    used++;
    double tot = 0;
    for (int k = 0; k < 4; k++)
      for (int h = 0; h < 4; h++) {
        int a = (k >> 1) + (h >> 1), b = (k & 1) + (h & 1);
        tot += f[k] * f[h] * p[a] * q[b];
      }
    for (int k = 0; k < 4; k++) {
      double part = 0;
      for (int h = 0; h < 4; h++) {
        int a = (k >> 1) + (h >> 1), b = (k & 1) + (h & 1);
        part += f[k] * f[h] * (p[a] * q[b] + p[a] * q[b]);
      }
      acc[k] += part / tot;
    }
  }
  for (int k = 0; k < 4; k++) f[k] = acc[k] / (2 * used);
  for (int k = 0; k < 4; k++) f[k] /= f[0] + f[1] + f[2] + f[3]; /* sequential: sees updated entries */
  return used;
}

/* shared/gen_func.cpp:1027-1059.  Returns the 0-based index of the converging pass, 100 if none. */
uint64_t orc_haplo_freq(double f[4], uint64_t *n_used, const double *ga, const double *gb,
                        double maf_a, double maf_b, uint64_t n_ind, int ignore_miss) {
  f[0] = (1 - maf_a) * (1 - maf_b);
  f[1] = (1 - maf_a) * maf_b;
  f[2] = maf_a * (1 - maf_b);
  f[3] = maf_a * maf_b;
  uint64_t it;
  for (it = 0; it < ORC_ITER_MAX; it++) {
    double prev[4] = {f[0], f[1], f[2], f[3]}, eps = 0;
    *n_used = orc_em_pass(f, ga, gb, n_ind, ignore_miss);
    for (int j = 0; j < 4; j++) {
      double d = fabs(f[j] - prev[j]);
      if (d > eps) eps = d;
    }
    if (eps < ORC_EPS) break;
  }
  return it;
}

/* --------------------------------------------------------------------- one pair, all columns -- */
typedef struct {
  double r2pear, D, Dp, r2;
  double hap[4];
  double hmaf[2];
  float chi2;
  uint64_t n_used, n_iter;
} orc_pair_out;

#define ORC_MIN(a, b) ((a) <= (b) ? (a) : (b))

void orc_pair(const double *gl, const double *expg, const double *maf, uint64_t n_ind, uint64_t s1,
              uint64_t s2, int ignore_miss, orc_pair_out *o) { /* ngsLD.cpp:290-306,328-333 */
  o->r2pear = orc_pearson_r2(expg + s1 * n_ind, expg + s2 * n_ind, n_ind);
  o->n_used = 0;
  o->n_iter = orc_haplo_freq(o->hap, &o->n_used, gl + s1 * n_ind * 3, gl + s2 * n_ind * 3, maf[s1], maf[s2],
                             n_ind, ignore_miss);
  const double *f = o->hap;
  double m0 = 1 - (f[0] + f[1]), m1 = 1 - (f[0] + f[2]);
  o->hmaf[0] = m0;
  o->hmaf[1] = m1;
  o->D = f[0] * f[3] - f[1] * f[2];
  o->Dp = o->D / (o->D < 0 ? -ORC_MIN(m0 * m1, (1 - m0) * (1 - m1)) : ORC_MIN(m0 * (1 - m1), (1 - m0) * m1));
  double q = o->D / sqrt(m0 * m1 * (1 - m0) * (1 - m1));
  o->r2 = q * q;
  float chi2 = 0, fa = f[0] + f[1], fb = f[0] + f[2];
  float e[4] = {fa * fb, fa * (1 - fb), (1 - fa) * fb, (1 - fa) * (1 - fb)};
  for (int k = 0; k < 4; k++) {
    double d = f[k] - e[k];
    chi2 += d * d / e[k];
  }
  o->chi2 = chi2;
}

/* --------------------------------------------------------------------------- the s1 scan ----- */
typedef struct {
  const double *gl, *expg, *maf, *pos_dist;
  char *const *labels; /* NULL -> "(null)" like glibc prints a NULL %s */
  uint64_t n_sites, n_ind;
  uint64_t max_kb_dist, max_snp_dist;
  double min_maf, rnd_sample;
  uint64_t seed;
  int ignore_miss, extend_out;
} orc_job;

typedef struct {
  char *buf;
  size_t len, cap;
  uint64_t n_pairs, sum_iter;
} orc_rowbuf;

static void orc_emit(orc_rowbuf *rb, const char *txt, size_t n) {
  if (rb->len + n + 1 > rb->cap) {
    rb->cap = (rb->cap ? rb->cap * 2 : 4096) + n;
    rb->buf = (char *)realloc(rb->buf, rb->cap);
  }
  memcpy(rb->buf + rb->len, txt, n);
  rb->len += n;
}

/* ngsLD.cpp:229-359 for one first site.  want_text = 0 only counts (CPU-baseline timing). */
static void orc_scan_site(const orc_job *J, uint64_t s1, uint64_t site_seed, int want_text, orc_rowbuf *rb) {
  orc_taus rng;
  orc_taus_set(&rng, site_seed);
  double dist = 0;
  char line[1024];
  for (uint64_t s2 = s1 + 1; s2 < J->n_sites; s2++) {
    dist += J->pos_dist[s2];
    if (J->max_kb_dist > 0 && J->max_kb_dist * 1000 < dist) break;
    if (J->max_snp_dist > 0 && J->max_snp_dist < s2 - s1) break;
    if (J->maf[s1] < J->min_maf) break;
    if (J->maf[s2] < J->min_maf) continue;
    if (0 + orc_taus_uniform(&rng) * (1 - 0) > J->rnd_sample) continue;
    orc_pair_out o;
    orc_pair(J->gl, J->expg, J->maf, J->n_ind, s1, s2, J->ignore_miss, &o);
    rb->n_pairs++;
    rb->sum_iter += o.n_iter < ORC_ITER_MAX ? o.n_iter + 1 : ORC_ITER_MAX;
    if (!want_text) continue;
    int n = snprintf(line, sizeof line, "%s\t%s\t%.0f\t%f\t%f\t%f\t%f", J->labels ? J->labels[s1] : "(null)",
                     J->labels ? J->labels[s2] : "(null)", dist, o.r2pear, o.D, o.Dp, o.r2);
    if (J->extend_out)
      n += snprintf(line + n, sizeof line - n, "\t%lu\t%f\t%f\t%f\t%f\t%f\t%f\t%f\t%f\t%f\t%f\t%lu",
                    (unsigned long)o.n_used, J->maf[s1], J->maf[s2], o.hap[0], o.hap[1], o.hap[2], o.hap[3],
                    o.hmaf[0], o.hmaf[1], o.chi2, 0.0, (unsigned long)o.n_iter);
    line[n++] = '\n';
    orc_emit(rb, line, (size_t)n);
  }
}

typedef struct {
  const orc_job *J;
  const uint64_t *seeds;
  orc_rowbuf *rows; /* per s1 */
  uint64_t s1_lo, s1_hi;
  volatile uint64_t *next;
  int want_text;
} orc_worker_arg;

static void *orc_worker(void *vp) {
  orc_worker_arg *w = (orc_worker_arg *)vp;
  for (;;) {
    uint64_t s1 = __sync_fetch_and_add(w->next, 1);
    if (s1 >= w->s1_hi) break;
    orc_scan_site(w->J, s1, w->seeds[s1], w->want_text, &w->rows[s1 - w->s1_lo]);
  }
  return NULL;
}

/* Whole job for first sites [s1_lo, s1_hi): rows come out in (s1, s2) order == the reference with
 * --n_threads 1.  out_path NULL -> count only.  header: write the ngsLD.cpp:77 header first.
 * Returns pairs computed; *sum_iter_out = total EM passes executed. */
uint64_t orc_run(const double *gl, const double *expg, const double *maf, const double *pos_dist,
                 char *const *labels, uint64_t n_sites, uint64_t n_ind, uint64_t max_kb_dist,
                 uint64_t max_snp_dist, double min_maf, double rnd_sample, uint64_t seed, int ignore_miss,
                 int extend_out, uint64_t s1_lo, uint64_t s1_hi, int n_threads, const char *out_path,
                 int header, uint64_t *sum_iter_out) {
  orc_job J = {gl, expg, maf, pos_dist, labels, n_sites, n_ind, max_kb_dist, max_snp_dist,
               min_maf, rnd_sample, seed, ignore_miss, extend_out};
  if (s1_hi > n_sites) s1_hi = n_sites;
  uint64_t *seeds = (uint64_t *)malloc(sizeof(uint64_t) * (n_sites ? n_sites : 1));
  orc_site_seeds(seed, n_sites, seeds);
  uint64_t span = s1_hi > s1_lo ? s1_hi - s1_lo : 0;
  orc_rowbuf *rows = (orc_rowbuf *)calloc(span ? span : 1, sizeof(orc_rowbuf));
  volatile uint64_t next = s1_lo;
  if (n_threads < 1) n_threads = 1;
  orc_worker_arg arg = {&J, seeds, rows, s1_lo, s1_hi, &next, out_path != NULL};
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
  for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, orc_worker, &arg);
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  uint64_t pairs = 0, iters = 0;
  FILE *fh = out_path ? fopen(out_path, "w") : NULL;
  if (fh && header)
    fprintf(fh, "site1\tsite2\tdist\tr2_ExpG\tD\tDp\tr2%s\n",
            extend_out ? "\tsample_size\tmaf1\tmaf2\thap00\thap01\thap10\thap11\thap_maf1\thap_maf2\tchi2\tloglike\tnIter" : "");
  for (uint64_t k = 0; k < span; k++) {
    pairs += rows[k].n_pairs;
    iters += rows[k].sum_iter;
    if (fh && rows[k].len) fwrite(rows[k].buf, 1, rows[k].len, fh);
    free(rows[k].buf);
  }
  if (fh) fclose(fh);
  free(rows);
  free(seeds);
  free(th);
  if (sum_iter_out) *sum_iter_out = iters;
  return pairs;
}

/* Positions -> inter-site distances: shared/read_data.cpp:199-214 on already-split (chr, pos) text.
 * chr[s], pos[s] are the first two tab-separated fields of line s.  Returns 0 or -1 (distance < 1). */
int orc_pos_dist(char *const *chr, char *const *pos, uint64_t n_sites, double *out) {
  const char *prev_chr = NULL;
  uint64_t prev_pos = 0;
  for (uint64_t s = 0; s < n_sites; s++) {
    if (prev_chr == NULL) prev_chr = chr[s];
    if (strcmp(prev_chr, chr[s]) == 0) {
      out[s] = strtod(pos[s], NULL) - prev_pos;
      if (out[s] < 1) return -1;
    } else {
      out[s] = INFINITY;
      prev_chr = chr[s];
    }
    prev_pos = strtoul(pos[s], NULL, 0);
  }
  return 0;
}
