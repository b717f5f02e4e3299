// x87 extended-precision (64-bit significand) arithmetic in integer registers.
//
// The reference's r2_ExpG column is gsl_stats_correlation (reference ngsLD.cpp:365-367), whose
// accumulators are `long double`; on x86-64 that is the x87 80-bit format, round-to-nearest-even,
// 64-bit precision control.  To reproduce that column bit for bit on a GPU (no fp80 hardware) the
// per-pair part of the recurrence -- two multiplies and one add per individual, one divide per
// pair -- is carried out here on (sign, exponent, 64-bit significand) triples.
//
// Only zero and normal numbers are representable (no x87 denormals / inf / nan): the operands are
// differences and products of expected genotypes in [0,2], > 16000 binades away from either end
// of the x87 exponent range.  Compiles as plain C++ too (tests/test_fp80.py checks it against the
// host FPU's native long double).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define X87_HD __host__ __device__ __forceinline__
#else
#define X87_HD static inline
#endif

namespace x87 {

struct ext {       // value = (-1)^neg * sig * 2^(exp - 63);  sig == 0 -> zero, else bit 63 set
  uint64_t sig;
  int32_t exp;
  uint32_t neg;
};

X87_HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return __builtin_clzll(v);
#endif
}

X87_HD void mul64(uint64_t a, uint64_t b, uint64_t &hi, uint64_t &lo) {
#if defined(__CUDA_ARCH__)
  hi = __umul64hi(a, b);
  lo = a * b;
#else
  unsigned __int128 p = (unsigned __int128)a * b;
  hi = (uint64_t)(p >> 64);
  lo = (uint64_t)p;
#endif
}

// Round a 128-bit significand (hi:lo, hi normalised) plus a sticky flag to 64 bits, ties to even.
X87_HD ext round_pack(uint32_t neg, int32_t exp, uint64_t hi, uint64_t lo, bool sticky) {
  bool guard = (lo >> 63) != 0;
  bool rest = ((lo << 1) != 0) || sticky;
  if (guard && (rest || (hi & 1))) {
    hi += 1;
    if (hi == 0) {  // carried out of 64 bits
      hi = 0x8000000000000000ull;
      exp += 1;
    }
  }
  ext r;
  r.sig = hi;
  r.exp = exp;
  r.neg = neg;
  return r;
}

X87_HD ext zero(uint32_t neg = 0) {
  ext r;
  r.sig = 0;
  r.exp = 0;
  r.neg = neg;
  return r;
}

// Exact widening of an IEEE double (finite; subnormals handled) -- the x87 FLD m64.
X87_HD ext from_double(double d) {
#if defined(__CUDA_ARCH__)
  uint64_t b = (uint64_t)__double_as_longlong(d);
#else
  uint64_t b;
  __builtin_memcpy(&b, &d, 8);
#endif
  uint32_t neg = (uint32_t)(b >> 63);
  int32_t be = (int32_t)((b >> 52) & 0x7ff);
  uint64_t frac = b & 0xfffffffffffffull;
  if (be == 0) {
    if (frac == 0) return zero(neg);
    int sh = clz64(frac);
    ext r;
    r.sig = frac << sh;
    r.exp = -1022 - 52 + (63 - sh);
    r.neg = neg;
    return r;
  }
  ext r;
  r.sig = (frac | 0x10000000000000ull) << 11;
  r.exp = be - 1023;
  r.neg = neg;
  return r;
}

// The 10-byte memory image of a long double: sig = bytes 0-7, se = bytes 8-9 (sign | biased exp).
X87_HD ext from_bits(uint64_t sig, uint16_t se) {
  ext r;
  r.sig = sig;
  r.exp = (int32_t)(se & 0x7fff) - 16383;
  r.neg = (uint32_t)(se >> 15);
  if (sig == 0) r.exp = 0;
  return r;
}

X87_HD ext mul(const ext &a, const ext &b) {
  uint32_t neg = a.neg ^ b.neg;
  if (a.sig == 0 || b.sig == 0) return zero(neg);
  uint64_t hi, lo;
  mul64(a.sig, b.sig, hi, lo);
  int32_t e = a.exp + b.exp + 1;
  if (!(hi >> 63)) {  // product in [2^126, 2^127): renormalise
    hi = (hi << 1) | (lo >> 63);
    lo <<= 1;
    e -= 1;
  }
  return round_pack(neg, e, hi, lo, false);
}

X87_HD ext add(const ext &x, const ext &y) {
  if (y.sig == 0) {
    if (x.sig == 0) return zero(x.neg & y.neg);  // (+0)+(-0) = +0 under round-to-nearest
    return x;
  }
  if (x.sig == 0) return y;
  // a = operand of larger magnitude
  bool swap = (y.exp > x.exp) || (y.exp == x.exp && y.sig > x.sig);
  const ext &a = swap ? y : x;
  const ext &b = swap ? x : y;
  uint32_t d = (uint32_t)(a.exp - b.exp);
  // b's significand as a 128-bit value aligned under a (a = a.sig:0), jamming lost bits into `sticky`
  uint64_t bhi, blo;
  bool sticky = false;
  if (d == 0) {
    bhi = b.sig;
    blo = 0;
  } else if (d < 64) {
    bhi = b.sig >> d;
    blo = b.sig << (64 - d);
  } else if (d == 64) {
    bhi = 0;
    blo = b.sig;
  } else if (d < 128) {
    bhi = 0;
    blo = b.sig >> (d - 64);
    sticky = (b.sig << (128 - d)) != 0;
  } else {
    bhi = 0;
    blo = 0;
    sticky = true;
  }
  uint64_t hi, lo;
  int32_t e = a.exp;
  if (a.neg == b.neg) {
    lo = blo;
    hi = a.sig + bhi;
    if (hi < a.sig) {  // carry out of bit 127: shift right one
      sticky = sticky || (lo & 1);
      lo = (lo >> 1) | (hi << 63);
      hi = (hi >> 1) | 0x8000000000000000ull;
      e += 1;
    }
    return round_pack(a.neg, e, hi, lo, sticky);
  }
  // magnitude subtraction a - b - (sticky ? tiny : 0)
  uint64_t borrow_in = sticky ? 1 : 0;  // lost bits of b make the true difference slightly smaller
  uint64_t lo0 = 0 - blo;
  uint64_t br = (blo != 0) ? 1 : 0;
  lo = lo0 - borrow_in;
  if (lo0 < borrow_in) br = 1;
  hi = a.sig - bhi - br;
  // (with sticky set, lo now holds the floor of the difference and the remainder is still non-zero)
  if (hi == 0 && lo == 0) return zero(0);
  if (hi == 0) {
    hi = lo;
    lo = 0;
    e -= 64;
  }
  int sh = clz64(hi);
  if (sh) {
    hi = (hi << sh) | (lo >> (64 - sh));
    lo <<= sh;
    e -= sh;
  }
  return round_pack(a.neg, e, hi, lo, sticky);
}

// Correctly rounded quotient; b != 0.
X87_HD ext div(const ext &a, const ext &b) {
  uint32_t neg = a.neg ^ b.neg;
  if (a.sig == 0) return zero(neg);
  // Scale the numerator so the quotient of significands lies in [1, 2); then restoring division.
  uint64_t rem;
  int32_t e = a.exp - b.exp;
  if (a.sig >= b.sig) {
    rem = a.sig - b.sig;
  } else {
    rem = a.sig - (b.sig - a.sig);  // 2*a.sig - b.sig, which fits because a.sig < b.sig
    e -= 1;
  }
  uint64_t q = 1;
  for (int k = 0; k < 63; k++) {
    bool top = (rem >> 63) != 0;
    rem <<= 1;
    bool bit = top || rem >= b.sig;
    if (bit) rem -= b.sig;
    q = (q << 1) | (bit ? 1 : 0);
  }
  bool top = (rem >> 63) != 0;
  rem <<= 1;
  bool guard = top || rem >= b.sig;
  if (guard) rem -= b.sig;
  return round_pack(neg, e, q, guard ? 0x8000000000000000ull : 0, rem != 0);
}

// ---- fused fast path of the r2_ExpG inner loop ---------------------------------------------------------------
// One step of sum += fl80( fl80(a * b) * r ) with r = rsig * 2^-64 in [0.5, 1) (the ratio i/(i+1.0) of
// gsl_stats_correlation widened to 80 bits: its exponent is always -1).  Same results as
// add(acc, mul(mul(a, b), r)) above, with the case analysis the general routines need stripped down to what
// this loop can see: a, b, r are zero or normal, the accumulator is never -0 (it starts at +0, +0 + -0 = +0 and an
// exact cancellation gives +0 under round-to-nearest).

// 128-bit product of two normalised significands -> rounded 64-bit significand; returns the exponent increment
// relative to (ea + eb): 1 if the product had its top bit set, 0 otherwise, +1 if rounding carried out.
X87_HD int32_t mul_round(uint64_t a, uint64_t b, uint64_t &out) {
  uint64_t hi, lo;
  mul64(a, b, hi, lo);
  const uint32_t top = (uint32_t)(hi >> 63);
  if (!top) {
    hi = (hi << 1) | (lo >> 63);
    lo <<= 1;
  }
  const uint64_t inc = (lo >> 63) & (uint64_t)(((lo << 1) != 0) | (hi & 1));
  hi += inc;
  const uint32_t carry = hi == 0;  // only an all-ones significand can wrap
  out = hi | ((uint64_t)carry << 63);
  return (int32_t)(top + carry);
}

// Straight-line form of the same step: identical results to mac_ratio below (cross-checked in
// tests/native/fp80_check.cpp), organised so that a GPU could execute it as one predicated block instead of a dozen
// branches.  NOT used by the kernels: nvcc turns it into as many instructions as the case-by-case version (238 vs 243 per
// individual) and with 58 instead of 40 registers, and the fused kernel measured 3 % slower with it (round 2).
//   * both operands are aligned in one 128-bit window under the larger one; an exponent distance above 66 is clamped to
//     66 (the small operand then only leaves a sticky trace below the guard bit, which can never change the result);
//   * a subtraction is the addition of the two's complement (the bits that fell off the window turn the +1 into +0);
//   * one normalisation (one place to the right after a carry, or clz places to the left after a cancellation) and one
//     rounding serve both cases.
X87_HD void mac_ratio_flat(ext &acc, uint64_t asig, uint32_t ase, uint64_t bsig, uint32_t bse, uint64_t rsig) {
  if (asig == 0 || bsig == 0) return;  // a zero term leaves the (never -0) accumulator unchanged
  uint64_t psig, t;
  int32_t e = (int32_t)((ase & 0x7fffu) + (bse & 0x7fffu)) - 2 * 16383;
  e += mul_round(asig, bsig, psig);     // P = fl80(a * b)
  e += mul_round(psig, rsig, t) - 1;    // T = fl80(P * r), r.exp = -1
  const uint32_t tneg = ((ase ^ bse) >> 15) & 1u;
  if (acc.sig == 0) {  // first term
    acc.sig = t;
    acc.exp = e;
    acc.neg = tneg;
    return;
  }
  const int32_t dd = acc.exp - e;
  const bool swap = dd < 0 || (dd == 0 && t > acc.sig);
  const uint64_t big = swap ? t : acc.sig, small = swap ? acc.sig : t;
  const uint32_t nbig = swap ? tneg : acc.neg;
  const bool sub = (acc.neg ^ tneg) != 0;
  int32_t er = swap ? e : acc.exp;
  uint32_t d = (uint32_t)(dd < 0 ? -dd : dd);
  d = d > 66u ? 66u : d;
  // small under big (= big:0) as shi:slo, bits below the window in `sticky`
  const uint32_t k = d & 63u;
  const uint64_t xh = small >> k, xl = k ? small << (64u - k) : 0ull;
  const bool far = d >= 64u;
  const uint64_t shi = far ? 0ull : xh, slo = far ? xh : xl;
  uint32_t sticky = far ? (uint32_t)(xl != 0) : 0u;
  // big:0 + shi:slo, or big:0 - shi:slo - (sticky ? something below the window : 0) as a two's-complement addition
  const uint64_t m = sub ? ~0ull : 0ull;
  const uint64_t cin = sub ? (uint64_t)(1u - sticky) : 0ull;
  uint64_t lo = (slo ^ m) + cin;
  const uint64_t c1 = (uint64_t)(lo < cin);
  const uint64_t y = (shi ^ m) + c1;          // cannot wrap: shi ^ m == ~0 needs shi == 0, slo == 0 -> handled by c1 <= 1 below
  const uint64_t cy = (uint64_t)(y < c1);
  uint64_t hi = big + y;
  const bool carry_out = (hi < big) || cy;  // addition: a carry out of bit 127; subtraction: always set (no borrow)
  if (sub && (hi | lo) == 0) {  // exact cancellation -> +0
    acc.sig = 0;
    acc.exp = 0;
    acc.neg = 0;
    return;
  }
  if (!sub && carry_out) {  // one place to the right
    sticky |= (uint32_t)(lo & 1);
    lo = (lo >> 1) | (hi << 63);
    hi = (hi >> 1) | 0x8000000000000000ull;
    er += 1;
  }
  if (sub) {  // up to 127 places to the left
    if (hi == 0) {
      hi = lo;
      lo = 0;
      er -= 64;
    }
    const int sh = clz64(hi);
    if (sh) {
      hi = (hi << sh) | (lo >> (64 - sh));
      lo <<= sh;
      er -= sh;
    }
  }
  const uint64_t inc = (lo >> 63) & (uint64_t)((((lo << 1) != 0) | sticky) | (hi & 1));
  hi += inc;
  const uint32_t carry = hi == 0;
  acc.sig = hi | ((uint64_t)carry << 63);
  acc.exp = er + (int32_t)carry;
  acc.neg = nbig;
}

X87_HD void mac_ratio(ext &acc, uint64_t asig, uint32_t ase, uint64_t bsig, uint32_t bse, uint64_t rsig) {
  if (asig == 0 || bsig == 0) return;  // a zero term leaves the (never -0) accumulator unchanged
  uint64_t psig, tsig;
  int32_t e = (int32_t)(ase & 0x7fff) + (int32_t)(bse & 0x7fff) - 2 * 16383;
  e += mul_round(asig, bsig, psig);        // P = fl80(a * b)
  e += mul_round(psig, rsig, tsig) - 1;    // T = fl80(P * r), r.exp = -1
  const uint32_t tneg = ((ase ^ bse) >> 15) & 1u;
  if (acc.sig == 0) {
    acc.sig = tsig;
    acc.exp = e;
    acc.neg = tneg;
    return;
  }
  // big = operand of larger magnitude, small the other
  const bool swap = (e > acc.exp) || (e == acc.exp && tsig > acc.sig);
  const uint64_t big = swap ? tsig : acc.sig, small = swap ? acc.sig : tsig;
  const int32_t ebig = swap ? e : acc.exp;
  const uint32_t nbig = swap ? tneg : acc.neg, nsmall = swap ? acc.neg : tneg;
  const uint32_t d = (uint32_t)(ebig - (swap ? acc.exp : e));
  acc.neg = nbig;
  if (d > 65) {  // small < ulp(big)/4: it cannot move big, not even across a binade boundary when subtracting
    acc.sig = big;
    acc.exp = ebig;
    return;
  }
  // small aligned under big as 128 bits (big = big:0); bits that fall off the window only exist for d == 65
  uint64_t shi, slo;
  uint32_t sticky = 0;
  if (d < 64) {
    shi = small >> d;
    slo = (small << 1) << (63 - d);
  } else if (d == 64) {
    shi = 0;
    slo = small;
  } else {
    shi = 0;
    slo = small >> 1;
    sticky = (uint32_t)(small & 1);
  }
  uint64_t hi, lo;
  int32_t er = ebig;
  if (nbig == nsmall) {
    lo = slo;
    hi = big + shi;
    if (hi < big) {  // carry out of bit 127
      sticky |= (uint32_t)(lo & 1);
      lo = (lo >> 1) | (hi << 63);
      hi = (hi >> 1) | 0x8000000000000000ull;
      er += 1;
    }
  } else {
    // big:0 - shi:slo - (sticky ? something below the window : 0)
    const uint64_t l0 = 0 - slo;
    uint64_t borrow = slo != 0;
    lo = l0 - sticky;
    borrow |= (uint64_t)(l0 < sticky);
    hi = big - shi - borrow;
    if ((hi | lo) == 0) {  // exact cancellation -> +0
      acc.sig = 0;
      acc.exp = 0;
      acc.neg = 0;
      return;
    }
    if (hi == 0) {
      hi = lo;
      lo = 0;
      er -= 64;
    }
    const int sh = clz64(hi);
    if (sh) {
      hi = (hi << sh) | (lo >> (64 - sh));
      lo <<= sh;
      er -= sh;
    }
  }
  const uint64_t inc = (lo >> 63) & (uint64_t)((((lo << 1) != 0) | sticky) | (hi & 1));
  hi += inc;
  const uint32_t carry = hi == 0;
  acc.sig = hi | ((uint64_t)carry << 63);
  acc.exp = er + (int32_t)carry;
}

// significand of (long double)(i / (i + 1.0)) for i >= 1 (the value lies in [0.5, 1): exponent -1)
X87_HD uint64_t ratio_sig(double ratio) {
#if defined(__CUDA_ARCH__)
  const uint64_t b = (uint64_t)__double_as_longlong(ratio);
#else
  uint64_t b;
  __builtin_memcpy(&b, &ratio, 8);
#endif
  return ((b & 0xfffffffffffffull) | 0x10000000000000ull) << 11;
}

// Narrow to double with round-to-nearest-even (x87 FSTP m64); overflow -> inf, underflow -> subnormal/0.
X87_HD double to_double(const ext &a) {
  uint64_t bits;
  uint64_t sign = (uint64_t)a.neg << 63;
  if (a.sig == 0) {
    bits = sign;
  } else {
    int32_t e = a.exp;
    uint64_t m = a.sig;
    int drop = 11;
    if (e < -1022) drop += (-1022 - e);
    if (drop > 64) {
      bits = sign;  // far below the smallest subnormal
    } else {
      uint64_t kept, lost;
      bool half, rest;
      if (drop == 64) {
        kept = 0;
        lost = m;
      } else {
        kept = m >> drop;
        lost = m << (64 - drop);
      }
      half = (lost >> 63) != 0;
      rest = (lost << 1) != 0;
      if (half && (rest || (kept & 1))) kept += 1;
      if (e < -1022) {
        bits = sign | kept;  // subnormal (a carry into bit 52 lands on the smallest normal correctly)
      } else {
        if (kept >> 53) {
          kept >>= 1;
          e += 1;
        }
        if (e > 1023)
          bits = sign | 0x7ff0000000000000ull;
        else
          bits = sign | ((uint64_t)(e + 1023) << 52) | (kept & 0xfffffffffffffull);
      }
    }
  }
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)bits);
#else
  double d;
  __builtin_memcpy(&d, &bits, 8);
  return d;
#endif
}

}  // namespace x87
