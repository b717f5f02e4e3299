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

// ---- the same step on 32-bit limbs (what the kernels use) ----------------------------------------------------------
// mac3 is mac_ratio written for a machine whose registers are 32 bits wide and whose warps pay for every path any lane
// takes: the significands are two words, the aligned addition runs in a 96-bit window (64 bits + one guard word whose
// lowest bit collects everything that falls off: "jamming") with ONE bit of headroom on top, so that addition and
// subtraction are the same three-word two's-complement sum followed by the same count-leading-zeros normalisation, and
// only two cases leave the straight line: a term more than 2^31 times smaller (or larger) than the running sum, and a
// cancellation of 31 or more bits.  Results are identical to mac_ratio (tests/native/fp80_check.cpp).
//
// Terms come in a packed form of their own ("se14"): bit 15 = sign, bits 0-14 = exponent + 8192 (every deviation of an
// expected genotype lies within 2^+-1200), so that one integer addition of two such words yields the exponent sum AND
// (in bit 15: a one-bit sum without carry-in) the sign of the product.
struct acc96 {   // value = (-1)^neg * (h1:h0) * 2^(exp - ACC_BIAS - 63); zero: h1 == 0 (exp = ACC_ZERO_EXP, neg = 0)
  uint32_t h1, h0;
  int32_t exp;
  uint32_t neg;
};
constexpr int32_t ACC_BIAS = 16381;            // (ea + 8192) + (eb + 8192) - 1 (the ratio's exponent) - 2 (see mul_round3)
constexpr int32_t ACC_ZERO_EXP = -(1 << 28);   // far below every term: the first term then simply wins the comparison

X87_HD uint16_t se14_pack(uint32_t neg, int32_t exp, uint64_t sig) {  // zero <=> 0 (a non-zero term has exp + 8192 > 0)
  return sig ? (uint16_t)((neg << 15) | ((uint32_t)(exp + 8192) & 0x7fffu)) : (uint16_t)0;
}
X87_HD uint16_t se14_from_x87(uint16_t se) {  // from the sign | biased-exponent word of the 80-bit memory image
  return (uint16_t)((se & 0x8000u) | ((uint32_t)((int32_t)(se & 0x7fffu) - 16383 + 8192) & 0x7fffu));
}

X87_HD uint32_t fsh_l(uint32_t lo, uint32_t hi, uint32_t s) {  // upper word of (hi:lo) << s, 0 <= s <= 31
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, s);
#else
  return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}
X87_HD uint32_t fsh_rc(uint32_t lo, uint32_t hi, uint32_t s) {  // lower word of (hi:lo) >> min(s, 32)
#if defined(__CUDA_ARCH__)
  return __funnelshift_rc(lo, hi, s);
#else
  return s >= 32 ? hi : (s ? (lo >> s) | (hi << (32 - s)) : lo);
#endif
}
X87_HD uint32_t clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__clz((int)v);
#else
  return v ? (uint32_t)__builtin_clz(v) : 32u;
#endif
}
// r2:r1:r0 = (a2:a1:a0) + (b2:b1:b0), carries rippling upwards (the carry out of the top word is dropped)
X87_HD void add96(uint32_t a2, uint32_t a1, uint32_t a0, uint32_t b2, uint32_t b1, uint32_t b0, uint32_t &r2, uint32_t &r1,
                  uint32_t &r0) {
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %0, %3, %6;\n\taddc.cc.u32 %1, %4, %7;\n\taddc.u32 %2, %5, %8;"
      : "=r"(r0), "=r"(r1), "=r"(r2)
      : "r"(a0), "r"(a1), "r"(a2), "r"(b0), "r"(b1), "r"(b2));
#else
  const uint64_t s0 = (uint64_t)a0 + b0, s1 = (uint64_t)a1 + b1 + (s0 >> 32);
  r0 = (uint32_t)s0;
  r1 = (uint32_t)s1;
  r2 = a2 + b2 + (uint32_t)(s1 >> 32);
#endif
}

// Rounds the window x1:x0 | g (g = the 32 bits below, lowest bit jammed) to 64 bits, ties to even -> h1:h0 with the top
// bit set; nw = 0 if the rounding carried out of 64 bits (significand 2^63, exponent one up), else 1.
X87_HD void round64(uint32_t x1, uint32_t x0, uint32_t g, uint32_t &h1, uint32_t &h0, uint32_t &nw) {
  const uint32_t rest = (uint32_t)((g & 0x7fffffffu) != 0);
  const uint32_t inc = (g >> 31) & (rest | x0);
  const uint32_t y0 = x0 + inc;
  const uint32_t y1 = x1 + (uint32_t)(y0 < inc);
  nw = y1 >> 31;           // x1 has its top bit set: only an all-ones significand can wrap, and then to zero
  h1 = y1 | 0x80000000u;
  h0 = y0;
}

// a * b, both normalised, rounded to 64 bits -> h1:h0.  The exponent of the result is ea + eb + 2 - s - nw:
// s = 1 if the product lay in [2^126, 2^127) (moved one place to the left), nw = 0 if the rounding carried out.
X87_HD void mul_round3(uint64_t a, uint64_t b, uint32_t &h1, uint32_t &h0, uint32_t &s, uint32_t &nw) {
  uint64_t hi, lo;
  mul64(a, b, hi, lo);  // (nvcc shares the partial products of __umul64hi and a * b: seven instructions; spelling the
                        // schoolbook product out on 32-bit halves compiled to more)
  const uint32_t w0 = (uint32_t)lo, w1 = (uint32_t)(lo >> 32), w2 = (uint32_t)hi, w3 = (uint32_t)(hi >> 32);
  s = (~w3) >> 31;
  const uint32_t x1 = fsh_l(w2, w3, s), x0 = fsh_l(w1, w2, s);
  const uint32_t g = fsh_l(w0, w1, s) | (uint32_t)((w0 << s) != 0);  // the lowest word only matters as "something there"
  round64(x1, x0, g, h1, h0, nw);
}

// One term of the sum: T = fl80( fl80(a * b) * r ), r = rsig * 2^-64 in [0.5, 1);  ase / bse in se14 form.
// Straight-line code without a branch, so that the terms of several individuals (which do not depend on each other, only
// the additions do) can be scheduled into each other.  se == 0 marks a zero term (t1:t0 is then meaningless).
struct term96 {
  uint32_t t1, t0;
  uint32_t se;  // bit 15: sign, bits 0-14: exponent biased by ACC_BIAS; 0: the term is zero
};
X87_HD term96 term3(uint64_t asig, uint32_t ase, uint64_t bsig, uint32_t bse, uint64_t rsig) {
  uint32_t p1, p0, sa, na, sb, nb2;
  term96 t;
  mul_round3(asig, bsig, p1, p0, sa, na);                            // P = fl80(a * b)
  mul_round3(((uint64_t)p1 << 32) | p0, rsig, t.t1, t.t0, sb, nb2);  // T = fl80(P * r)
  const uint32_t se = ase + bse;  // exponent sum in bits 0-14 (no carry into bit 15: both fields are below 2^14), sign in bit 15
  const uint32_t e = (se & 0x7fffu) - (sa + na) - (sb + nb2);  // >= 2 * (8192 - 1200) - 4: far from wrapping
  t.se = (ase == 0 || bse == 0) ? 0u : ((se & 0x8000u) | e);
  return t;
}

// acc += T
X87_HD void acc3(acc96 &acc, const term96 &tm) {
  if (tm.se == 0) return;  // a zero term leaves the sum as it is
  const uint32_t t1 = tm.t1, t0 = tm.t0;
  const int32_t e = (int32_t)(tm.se & 0x7fffu);
  const uint32_t tneg = tm.se >> 15;
  // big = operand of larger magnitude
  const int32_t dd = acc.exp - e;
  const uint64_t A = ((uint64_t)acc.h1 << 32) | acc.h0, T = ((uint64_t)t1 << 32) | t0;
  const bool tb = (dd < 0) | ((dd == 0) & (T > A));
  const uint32_t B1 = tb ? t1 : acc.h1, B0 = tb ? t0 : acc.h0, S1 = tb ? acc.h1 : t1, S0 = tb ? acc.h0 : t0;
  int32_t eb = tb ? e : acc.exp;
  const uint32_t sub = acc.neg ^ tneg;
  acc.neg = tb ? tneg : acc.neg;
  uint32_t k = (uint32_t)(dd < 0 ? -dd : dd);
  k = (k > 66u ? 66u : k) + 1u;  // beyond 66 places the small operand cannot matter: any distance above is as good as 66
  // window of 96 bits, big at bits 94..31 (one bit of headroom), small k places further down
  uint32_t X1 = fsh_rc(S1, 0u, k), X0 = fsh_rc(S0, S1, k), G = fsh_rc(0u, S0, k);
  if (k > 32u) {  // (rare) the words above hold the shift by 32; the rest of the way, jamming what falls off
    const uint32_t k2 = k - 32u;  // 1..35
    const uint64_t v = ((uint64_t)S1 << 32) | S0;
    const uint64_t u = v >> k2;
    X1 = 0;
    X0 = (uint32_t)(u >> 32);
    G = (uint32_t)u | (uint32_t)((v << (64u - k2)) != 0);
  }
  // big + small, or big - small as big + ~small + 1 (the "+ 1" rides in the 31 empty low bits of big's guard word)
  const uint32_t m = 0u - sub;
  uint32_t R1, R0, g;
  add96(B1 >> 1, fsh_rc(B0, B1, 1u), (B0 << 31) | sub, X1 ^ m, X0 ^ m, G ^ m, R1, R0, g);
  if (R1 == 0) {  // (rare) 31 or more leading bits cancelled; the guard word then holds at most its two top bits
    if ((R0 | g) == 0) {  // exact cancellation -> +0
      acc.h1 = 0;
      acc.h0 = 0;
      acc.exp = ACC_ZERO_EXP;
      acc.neg = 0;
      return;
    }
    if (R0 == 0) {
      R1 = g;
      g = 0;
      eb -= 64;
    } else {
      R1 = R0;
      R0 = g;
      g = 0;
      eb -= 32;
    }
  }
  const uint32_t z = clz32(R1);
  uint32_t nw;
  round64(fsh_l(R0, R1, z), fsh_l(g, R0, z), g << z, acc.h1, acc.h0, nw);
  acc.exp = eb + 2 - (int32_t)z - (int32_t)nw;
}

// acc += fl80( fl80(a * b) * r )
X87_HD void mac3(acc96 &acc, uint64_t asig, uint32_t ase, uint64_t bsig, uint32_t bse, uint64_t rsig) {
  acc3(acc, term3(asig, ase, bsig, bse, rsig));
}

X87_HD acc96 acc96_zero() {
  acc96 a;
  a.h1 = 0;
  a.h0 = 0;
  a.exp = ACC_ZERO_EXP;
  a.neg = 0;
  return a;
}
X87_HD ext acc96_to_ext(const acc96 &a) {
  ext r;
  r.sig = ((uint64_t)a.h1 << 32) | a.h0;
  r.exp = r.sig ? a.exp - ACC_BIAS : 0;
  r.neg = r.sig ? a.neg : 0u;
  return r;
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
