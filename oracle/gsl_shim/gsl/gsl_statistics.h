/* TEST INFRASTRUCTURE — not product code.
 *
 * Minimal stand-in for <gsl/gsl_statistics.h>: only gsl_stats_correlation, the one statistics
 * routine the reference calls (ngsLD.cpp:366).
 *
 * Third-party algorithm restated: GSL >= 1.10 `statistics/covariance_source.c`
 * (FUNCTION(gsl_stats,correlation)): a single-pass recurrence whose accumulators are all
 * `long double` (x87 80-bit on x86-64 Linux); `ratio` and the divisor `i + 1.0` are formed in
 * double; the final square roots go through the C `sqrt(double)`, so each sum of squares is first
 * rounded to double.  PARITY UNPINNED at this boundary: no real libgsl exists in this image to
 * check against (see DESIGN.md "Oracle").  Compiled as C++ here, so ::sqrt is forced to the double
 * overload explicitly to keep the C semantics.
 */
#ifndef NGSLD_ORACLE_GSL_STATISTICS_SHIM_H
#define NGSLD_ORACLE_GSL_STATISTICS_SHIM_H
#include <math.h>
#include <stddef.h>

static inline double gsl_stats_correlation(const double data1[], const size_t stride1,
                                           const double data2[], const size_t stride2,
                                           const size_t n) {
  long double sum_xsq = 0.0L, sum_ysq = 0.0L, sum_cross = 0.0L;
  long double ratio, delta_x, delta_y, mean_x, mean_y, r;
  mean_x = data1[0 * stride1];
  mean_y = data2[0 * stride2];
  for (size_t i = 1; i < n; ++i) {
    ratio = i / (i + 1.0);
    delta_x = data1[i * stride1] - mean_x;
    delta_y = data2[i * stride2] - mean_y;
    sum_xsq += delta_x * delta_x * ratio;
    sum_ysq += delta_y * delta_y * ratio;
    sum_cross += delta_x * delta_y * ratio;
    mean_x += delta_x / (i + 1.0);
    mean_y += delta_y / (i + 1.0);
  }
  r = sum_cross / (sqrt((double)sum_xsq) * sqrt((double)sum_ysq));
  return (double)r;
}
#endif
