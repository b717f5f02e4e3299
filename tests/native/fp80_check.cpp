// Host self-check of ngsld_b200/csrc/fp80.cuh against the FPU's native long double (x87).
// Built and run by tests/test_fp80.py.  Exit code 0 = all identical.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <random>
#include <vector>
#include "../../ngsld_b200/csrc/fp80.cuh"

static x87::ext from_ld(long double v) {
  uint64_t sig;
  uint16_t se;
  memcpy(&sig, &v, 8);
  memcpy(&se, (char *)&v + 8, 2);
  return x87::from_bits(sig, se);
}
static bool same(const x87::ext &a, long double v) {
  x87::ext b = from_ld(v);
  if (a.sig == 0 && b.sig == 0) return a.neg == b.neg;
  return a.sig == b.sig && a.exp == b.exp && a.neg == b.neg;
}

int main(int argc, char **argv) {
  long n = argc > 1 ? atol(argv[1]) : 2000000;
  std::mt19937_64 g(12345);
  std::uniform_real_distribution<double> U(-2.0, 2.0);
  long bad = 0;
  auto rnd_ld = [&](int spread) {
    // random long double with full 64-bit significand and modest exponent spread
    long double v = (long double)U(g) + (long double)U(g) * 0x1p-40L + (long double)U(g) * 0x1p-62L;
    int sh = (int)(g() % (2 * spread + 1)) - spread;
    return ldexpl(v, sh);
  };
  for (long k = 0; k < n; k++) {
    int spread = (k % 5 == 0) ? 140 : (k % 5 == 1 ? 70 : 3);
    volatile long double a = rnd_ld(spread), b = rnd_ld(spread);
    if (k % 97 == 0) b = -a;                      // exact cancellation
    if (k % 101 == 0) b = -a * (1 + 0x1p-63L);    // massive cancellation
    if (k % 103 == 0) a = 0.0L;
    if (k % 107 == 0) b = ldexpl(b, -64);         // alignment distance exactly at the guard boundary
    if (k % 109 == 0) b = ldexpl(b, -65);
    volatile long double s = a + b, p = a * b;
    x87::ext ea = from_ld(a), eb = from_ld(b);
    if (!same(x87::add(ea, eb), s)) { if (bad++ < 5) printf("add mismatch %La %La\n", (long double)a, (long double)b); }
    if (!same(x87::mul(ea, eb), p)) { if (bad++ < 5) printf("mul mismatch %La %La\n", (long double)a, (long double)b); }
    if (b != 0.0L) {
      volatile long double q = a / b;
      if (!same(x87::div(ea, eb), q)) { if (bad++ < 5) printf("div mismatch %La %La\n", (long double)a, (long double)b); }
    }
    volatile double d = (double)a;
    double mine = x87::to_double(ea);
    if (memcmp(&mine, (const void *)&d, 8) != 0) { if (bad++ < 5) printf("to_double mismatch %La\n", (long double)a); }
    double dd = U(g) * std::ldexp(1.0, (int)(g() % 60) - 30);
    if (!same(x87::from_double(dd), (long double)dd)) { if (bad++ < 5) printf("from_double mismatch %a\n", dd); }
  }
  // narrowing into the subnormal / overflow ranges of double
  for (int e = -1090; e <= -1010; e++) {
    volatile long double a = ldexpl(1.0L + 0x1.123456789abcdp-1L + 0x1p-60L, e);
    volatile double d = (double)a;
    double mine = x87::to_double(from_ld(a));
    if (memcmp(&mine, (const void *)&d, 8) != 0) { if (bad++ < 5) printf("subnormal narrowing mismatch e=%d\n", e); }
  }
  // the fused accumulate step of the r2_ExpG inner loop against native long double, including long random-walk
  // sums (cancellations, tiny terms under a large accumulator, zero terms)
  const int mac_reps = argc > 2 ? atoi(argv[2]) : 4000;
  long mac3_steps = 0;
  for (int rep = 0; rep < mac_reps; rep++) {
    const int n_ind = 2 + (int)(g() % 700);
    volatile long double sum = 0.0L;
    x87::ext acc = x87::zero(0);
    x87::acc96 acc3 = x87::acc96_zero();  // the 32-bit-limb form the kernels use
    bool acc3_live = true;
    mac3_steps += 0;
    const int mode = rep % 8;
    for (int i = 1; i < n_ind; i++) {
      long double da = rnd_ld(mode == 0 ? 40 : 2), db = rnd_ld(mode == 1 ? 70 : 2);
      if (mode == 2 && i % 3 == 0) da = 0.0L;
      if (mode == 3 && i % 2 == 0) { da = 1.0L; db = -(long double)sum / ((long double)(i / (i + 1.0))); }  // near-exact cancellation
      if (mode == 4) { da = ldexpl(da, -(int)(g() % 140)); }
      if (mode == 5 && i == n_ind / 2) { da = 0x1p+60L; }
      if (mode == 6) { da = ldexpl(da, (int)(g() % 140) - 70); if (i % 4 == 1) db = -db; }                 // wide exponent scatter, both signs
      if (mode == 7 && i > 1) { da = 1.0L; db = -(long double)sum * (1.0L + ldexpl(1.0L, -(int)(g() % 66))) / ((long double)(i / (i + 1.0))); }  // cancellations that leave 1..66 low bits
      const long double ratio = i / (i + 1.0);
      sum += da * db * ratio;
      const x87::ext ea = from_ld(da), eb = from_ld(db);
      uint64_t as, bs; uint16_t ae, be;
      long double tda = da, tdb = db;
      memcpy(&as, &tda, 8); memcpy(&ae, (char *)&tda + 8, 2);
      memcpy(&bs, &tdb, 8); memcpy(&be, (char *)&tdb + 8, 2);
      x87::ext acc_ref = acc;  // the straight-line version of the same step must agree too
      x87::mac_ratio_flat(acc_ref, as, ae, bs, be, x87::ratio_sig((double)i / ((double)i + 1.0)));
      x87::mac_ratio(acc, as, ae, bs, be, x87::ratio_sig((double)i / ((double)i + 1.0)));
      if (acc_ref.sig != acc.sig || acc_ref.exp != acc.exp || acc_ref.neg != acc.neg) {
        if (bad++ < 5) printf("mac_ratio != mac_ratio_flat rep %d mode %d i %d\n", rep, mode, i);
        break;
      }
      // (the packed term format of mac3 holds exponents within +-8191 -- deviations of doubles lie within +-1200 -- while the
      // repeated cancellations of mode 7 drive this test's operands far below that: mac3 is followed as far as it is defined)
      auto in_range = [](uint64_t sig, uint16_t se) { const int e = (int)(se & 0x7fff) - 16383; return sig == 0 || (e > -8000 && e < 8000); };
      if (!in_range(as, ae) || !in_range(bs, be)) acc3_live = false;
      if (acc3_live) {
        const uint64_t rs = x87::ratio_sig((double)i / ((double)i + 1.0));
        x87::mac3(acc3, as, as ? x87::se14_from_x87(ae) : 0, bs, bs ? x87::se14_from_x87(be) : 0, rs);
        const x87::ext a3 = x87::acc96_to_ext(acc3);
        mac3_steps++;
        if (a3.sig != acc.sig || (acc.sig && (a3.exp != acc.exp || a3.neg != acc.neg))) {
          if (bad++ < 5) printf("mac3 != mac_ratio rep %d mode %d i %d: %016llx e%d n%u  vs  %016llx e%d n%u\n", rep, mode, i, (unsigned long long)a3.sig, a3.exp, a3.neg, (unsigned long long)acc.sig, acc.exp, acc.neg);
          break;
        }
      }
      if (!same(acc, sum)) {
        if (bad++ < 5) printf("mac_ratio mismatch rep %d mode %d i %d: %La * %La\n", rep, mode, i, da, db);
        break;
      }
      (void)ea; (void)eb;
    }
  }
  // the packed sign/exponent word of the kernels' term tables: from the 80-bit memory image and from (sign, exponent)
  for (int e = -8000; e <= 8000; e += 7)
    for (uint32_t neg = 0; neg < 2; neg++) {
      const uint16_t x87se = (uint16_t)((neg << 15) | (uint32_t)(e + 16383));
      const uint16_t a = x87::se14_from_x87(x87se), b = x87::se14_pack(neg, e, 1ull << 63);
      if (a != b || (a >> 15) != neg || (int)(a & 0x7fff) - 8192 != e || a == 0) { if (bad++ < 5) printf("se14 mismatch e=%d\n", e); }
    }
  if (x87::se14_pack(1, 5, 0) != 0) { if (bad++ < 5) printf("se14 of a zero term must be 0\n"); }
  // the per-site half of gsl_stats_correlation as aux::site_terms_kernel performs it, against native long double
  std::uniform_real_distribution<double> E(0.0, 2.0);
  for (int rep = 0; rep < 300; rep++) {
    const int n_ind = 2 + (int)(g() % 600);
    std::vector<double> x(n_ind);
    for (auto &v : x) v = (g() % 7 == 0) ? 0.0 : ((g() % 5 == 0) ? 1.0 : E(g));
    if (rep % 17 == 0) std::fill(x.begin(), x.end(), 0.75);  // monomorphic site: every delta is exactly zero
    volatile long double mean = x[0], ssq = 0.0L;
    x87::ext emean = x87::from_double(x[0]), essq = x87::zero(0);
    for (int i = 1; i < n_ind; i++) {
      const long double ratio = i / (i + 1.0);
      const long double delta = x[i] - mean;
      ssq += delta * delta * ratio;
      mean += delta / (i + 1.0);
      const double ip1 = (double)i + 1.0;
      const x87::ext eratio = x87::from_double((double)i / ip1);
      x87::ext neg_mean = emean;
      neg_mean.neg ^= 1u;
      const x87::ext edelta = x87::add(x87::from_double(x[i]), neg_mean);
      essq = x87::add(essq, x87::mul(x87::mul(edelta, edelta), eratio));
      emean = x87::add(emean, x87::div(edelta, x87::from_double(ip1)));
      if (!same(edelta, delta) || !same(emean, mean) || !same(essq, ssq)) {
        if (bad++ < 5) printf("site recurrence mismatch rep %d i %d\n", rep, i);
        break;
      }
    }
    volatile double qn = sqrt((double)ssq);
    double qe = sqrt(x87::to_double(essq));
    if (memcmp(&qe, (const void *)&qn, 8) != 0) { if (bad++ < 5) printf("q mismatch rep %d\n", rep); }
  }
  printf("checked %ld random operand pairs and %ld limb-form accumulate steps, %ld mismatches\n", n, mac3_steps, bad);
  return bad ? 1 : 0;
}
