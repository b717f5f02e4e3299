// |v| * 10^6 rounded half-to-even on the EXACT value, for a finite double: the six decimals glibc's "%f" prints.
// A double is m * 2^e with an integer m < 2^53, so |v| * 10^6 = (m * 10^6) * 2^e is formed in 128-bit integer
// arithmetic.  Returns false for |v| >= 1e9 (the callers' fallback range); N is the rounded count of millionths.
#pragma once
#include <cuda_runtime.h>

namespace fmt {

__device__ __forceinline__ bool fixed6(double v, unsigned long long &N) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const int be = (int)((bits >> 52) & 0x7ff);
  const unsigned long long frac = bits & 0xfffffffffffffull;
  N = 0;
  if (be != 0 || frac != 0) {
    const unsigned long long m = be ? (frac | 0x10000000000000ull) : frac;
    const int e2 = (be ? be : 1) - 1075;
    if (fabs(v) >= 1e9) return false;
    // P = m * 10^6 (< 2^73) as hi:lo
    const unsigned long long lo = m * 1000000ull, hi = __umul64hi(m, 1000000ull);
    if (e2 >= 0) {
      N = lo << e2;  // |v| < 1e9 guarantees this fits
    } else {
      const int s = -e2;
      if (s >= 75) {
        N = 0;  // P < 2^73 <= half an ulp of the result
      } else {
        unsigned long long q, rem_hi, rem_lo, half_hi, half_lo;
        if (s < 64) {
          q = (lo >> s) | (s ? (hi << (64 - s)) : 0);  // hi < 2^9, q fits because |v|*1e6 < 2^64
          rem_hi = 0;
          rem_lo = lo & ((1ull << s) - 1);
          half_hi = 0;
          half_lo = 1ull << (s - 1);
        } else {
          q = (s == 64) ? hi : (hi >> (s - 64));
          rem_hi = (s == 64) ? 0 : (hi & ((1ull << (s - 64)) - 1));
          rem_lo = lo;
          half_hi = (s == 64) ? 0 : (1ull << (s - 65));
          half_lo = (s == 64) ? (1ull << 63) : 0;
        }
        const bool gt = rem_hi > half_hi || (rem_hi == half_hi && rem_lo > half_lo);
        const bool eq = rem_hi == half_hi && rem_lo == half_lo;
        if (gt || (eq && (q & 1))) q += 1;
        N = q;
      }
    }
  }
  return true;
}

}  // namespace fmt
