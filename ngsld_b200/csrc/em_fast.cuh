// Fast haplotype-frequency EM (replaces haplo_freq + pair_freq_iter, reference
// shared/gen_func.cpp:1027-1119) for sm_100a.
//
// Work decomposition: one GROUP of LPG lanes owns one site pair; each lane keeps IPL individuals of
// that pair ENTIRELY IN REGISTERS for the whole EM (their six genotype likelihoods), so the <=100 EM
// passes touch no memory at all.  LPG < 32: several groups share a warp; LPG > 32: a group spans
// LPG/32 warps and combines through shared memory + a named barrier.  Per pass and individual the
// E-step is estep() below (27 FP64 + 1 MUFU; same arithmetic as the warp-per-pair kernel of
// em_warp.cuh), and the M-step is f_k <- f_k A_k / n_used with A_k the group-wide sum of i_k / s
// (f_k i_k is the reference's tmp_k / 2 and its M-step is ff/(2x)).
// The reference's sequential renormalisation divides by 1 +- 1ulp; it is skipped inside the pass loop and
// applied once to the frequencies that are written out.  Results agree with the bit-faithful kernel
// (aux_kernels.cu) to ~1e-15, far inside the 1e-9 contract, at equal nIter.
//
// The pass loop is flattened into a state machine (fetch / iterate) so that sub-warp groups of one
// warp that converge after different numbers of passes never wait for each other: a finished group
// stores its row and loads its next pair while its neighbours keep iterating.
#pragma once
#include "common.cuh"

namespace emfast {

constexpr int CTA_THREADS = 128;                         // four warps; groups wider than that get their own CTA size
template <int LPG>
struct Cta {
  static constexpr int THREADS = LPG > CTA_THREADS ? LPG : CTA_THREADS;
};

struct Ind {
  double p0, p1, p2, q0, q1, q2;
};

// One individual's E-step:  u0 = f0 q0 + f1 q1   u1 = f2 q0 + f3 q1   v0 = f0 q1 + f1 q2   v1 = f2 q1 + f3 q2
//                           i0 = p0 u0 + p1 u1   i1 = p0 v0 + p1 v1   i2 = p1 u0 + p2 u1   i3 = p1 v0 + p2 v1
//                           s  = f0 i0 + f1 i1 + f2 i2 + f3 i3   (== the reference's `sum`, gen_func.cpp:1094-1096)
__device__ __forceinline__ void estep(const double f0, const double f1, const double f2, const double f3, const Ind &g,
                                      double &i0, double &i1, double &i2, double &i3, double &s) {
  const double u0 = __fma_rn(f1, g.q1, f0 * g.q0);
  const double u1 = __fma_rn(f3, g.q1, f2 * g.q0);
  const double v0 = __fma_rn(f1, g.q2, f0 * g.q1);
  const double v1 = __fma_rn(f3, g.q2, f2 * g.q1);
  i0 = __fma_rn(g.p1, u1, g.p0 * u0);
  i1 = __fma_rn(g.p1, v1, g.p0 * v0);
  i2 = __fma_rn(g.p2, u1, g.p1 * u0);
  i3 = __fma_rn(g.p2, v1, g.p1 * v0);
  s = __fma_rn(f3, i3, __fma_rn(f2, i2, __fma_rn(f1, i1, f0 * i0)));
}

__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));  // MUFU.RCP64H seed (~2^-20)
  double e = __fma_rn(-x, y, 1.0);
  e = __fma_rn(e, e, e);        // e + e^2
  return __fma_rn(y, e, y);     // y (1 + e + e^2): relative error ~ e^3 + 1 ulp
}

// Sums a0..a3 over aligned groups of GL lanes (GL = 4, 8, 16, 32); every lane of a group receives the same four
// totals.  Transposing butterfly: the first round halves the number of values a lane carries (4 -> 2), the second
// halves it again (2 -> 1), the remaining rounds add single values, and four shuffles hand the totals back:
// log2(GL) + 1 additions instead of 4 log2(GL).
template <int GL>
__device__ __forceinline__ void group_sum4(double &a0, double &a1, double &a2, double &a3, int lane) {
  constexpr int H = GL / 2, Q = GL / 4;
  const bool hi = lane & H, lo = lane & Q;
  double x0 = hi ? a2 : a0, x1 = hi ? a3 : a1;
  const double y0 = hi ? a0 : a2, y1 = hi ? a1 : a3;
  x0 += __shfl_xor_sync(0xffffffffu, y0, H);
  x1 += __shfl_xor_sync(0xffffffffu, y1, H);
  double z = lo ? x1 : x0;
  const double w = lo ? x0 : x1;
  z += __shfl_xor_sync(0xffffffffu, w, Q);
#pragma unroll
  for (int o = Q / 2; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
  const int base = lane & ~(GL - 1);
  a0 = __shfl_sync(0xffffffffu, z, base);
  a1 = __shfl_sync(0xffffffffu, z, base + Q);
  a2 = __shfl_sync(0xffffffffu, z, base + H);
  a3 = __shfl_sync(0xffffffffu, z, base + H + Q);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Shared-memory scratch a CTA needs for multi-warp groups and pair hand-out.
template <int LPG>
struct GroupScratch {
  static constexpr int W = LPG > 32 ? LPG / 32 : 1;
  static constexpr int G = Cta<LPG>::THREADS / LPG;
  double red[G][2][W][4];          // per group, double-buffered by pass parity
  unsigned long long fetched[G][2];  // {output index, row locators} handed to a multi-warp group
  unsigned int n_used[G][W];
};

// Row source for a pair: generic pointers (shared or global) to [n_pad][3] doubles.
struct PairRows {
  const double *a, *b;
};

// The per-group engine.  FETCH hands out work: next(idx, la, lb) is called by a group's leader lane
// only (false = no more work) and yields the output row index plus two opaque row locators;
// resolve(la, lb) -> row pointers and sites(la, lb, s1, s2) -> site indices are called by every lane.
template <int IPL, int LPG, class FETCH>
__device__ __forceinline__ void run_groups(const SiteTable &T, ngsld_pair_row *rows_out, bool ignore_miss,
                                           FETCH &fetch, GroupScratch<LPG> &scr, unsigned long long *pass_counter) {
  constexpr int W = LPG > 32 ? LPG / 32 : 1;
  constexpr int GL = LPG > 32 ? 32 : LPG;  // lanes of the group inside one warp
  constexpr bool MULTI = LPG > 32;
  const int lane = threadIdx.x & 31;
  const int grp = threadIdx.x / LPG;
  const int glane = threadIdx.x % LPG;
  const int gwarp = (threadIdx.x >> 5) % W;  // warp index inside the group
  const unsigned gmask = (GL == 32) ? 0xffffffffu : (((1u << GL) - 1u) << (lane & ~(GL - 1)));
  const int leader = lane & ~(GL - 1);

  Ind gl6[IPL];
  double f[4] = {0, 0, 0, 0};
  double inv_x = 0.0;
  uint32_t vmask = 0, n_used = 0, it = 0, s1 = 0, s2 = 0;
  unsigned long long out_index = 0;
  bool need = true, done = false;
  unsigned long long my_passes = 0;

#pragma unroll
  for (int j = 0; j < IPL; j++) gl6[j].p0 = gl6[j].p1 = gl6[j].p2 = gl6[j].q0 = gl6[j].q1 = gl6[j].q2 = 0.0;

  for (;;) {
    if (need && !done) {
      // ---------------- fetch the next pair for this group ----------------
      unsigned long long idx = ~0ull;
      uint32_t la = 0, lb = 0;
      if (glane == 0) {
        if (!fetch.next(idx, la, lb)) idx = ~0ull;
      }
      if (MULTI) {
        // The reduction barrier of the previous pass orders earlier reads of these slots before this write.
        if (glane == 0) {
          scr.fetched[grp][0] = idx;
          scr.fetched[grp][1] = ((unsigned long long)la << 32) | lb;
        }
        named_bar_sync(1 + grp, LPG);
        idx = scr.fetched[grp][0];
        const unsigned long long v = scr.fetched[grp][1];
        la = (uint32_t)(v >> 32);
        lb = (uint32_t)v;
      } else {
        idx = __shfl_sync(gmask, idx, leader);
        la = __shfl_sync(gmask, la, leader);
        lb = __shfl_sync(gmask, lb, leader);
      }
      if (idx == ~0ull) {
        done = true;
        vmask = 0;
      } else {
        const PairRows pr = fetch.resolve(la, lb);
        fetch.sites(la, lb, s1, s2);
        out_index = idx;
        // ---------------- load this lane's individuals into registers ----------------
        vmask = 0;
#pragma unroll
        for (int j = 0; j < IPL; j++) {
          const uint32_t i = (uint32_t)glane + (uint32_t)LPG * j;
          double p0 = 0, p1 = 0, p2 = 0, q0 = 0, q1 = 0, q2 = 0;
          bool ok = i < T.n_ind;
          if (ok) {
            const double *pa = pr.a + 3 * (size_t)i, *pb = pr.b + 3 * (size_t)i;
            p0 = pa[0]; p1 = pa[1]; p2 = pa[2];
            q0 = pb[0]; q1 = pb[1]; q2 = pb[2];
            if (ignore_miss && (gl_missing(p0, p1, p2) || gl_missing(q0, q1, q2))) ok = false;
          }
          if (!ok) { p0 = p1 = p2 = 0.0; }
          vmask |= (ok ? 1u : 0u) << j;
          gl6[j].p0 = p0; gl6[j].p1 = p1; gl6[j].p2 = p2;
          gl6[j].q0 = q0; gl6[j].q1 = q1; gl6[j].q2 = q2;
        }
        // individuals used by the EM (reference: x in pair_freq_iter)
        uint32_t cnt = __popc(vmask);
#pragma unroll
        for (int o = GL / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(gmask, cnt, o);
        if (MULTI) {
          if (lane == 0) scr.n_used[grp][gwarp] = cnt;
          named_bar_sync(1 + grp, LPG);
          cnt = 0;
#pragma unroll
          for (int w = 0; w < W; w++) cnt += scr.n_used[grp][w];
        }
        n_used = cnt;
        inv_x = __ddiv_rn(1.0, (double)n_used);
        const double m1 = T.maf[s1], m2 = T.maf[s2];  // haplo_freq start point, gen_func.cpp:1034-1037
        f[0] = __dmul_rn(__dsub_rn(1.0, m1), __dsub_rn(1.0, m2));
        f[1] = __dmul_rn(__dsub_rn(1.0, m1), m2);
        f[2] = __dmul_rn(m1, __dsub_rn(1.0, m2));
        f[3] = __dmul_rn(m1, m2);
        it = 0;
        need = false;
      }
    }
    if (MULTI) {
      if (done) break;
    } else {
      if (__all_sync(0xffffffffu, done)) break;
    }

    // ---------------- one EM pass (uniform across the warp) ----------------
    double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
#pragma unroll
    for (int j = 0; j < IPL; j++) {
      double i0, i1, i2, i3, s;
      estep(f[0], f[1], f[2], f[3], gl6[j], i0, i1, i2, i3, s);
      double inv = rcp_fast(s);
      if (!((vmask >> j) & 1u)) inv = 0.0;
      acc0 = __fma_rn(i0, inv, acc0);
      acc1 = __fma_rn(i1, inv, acc1);
      acc2 = __fma_rn(i2, inv, acc2);
      acc3 = __fma_rn(i3, inv, acc3);
    }
    // group-wide sums (every lane of the group receives the same bits)
    group_sum4<GL>(acc0, acc1, acc2, acc3, lane);
    if (MULTI) {
      const int par = it & 1;
      if (lane == 0) {
        double *slot = scr.red[grp][par][gwarp];
        slot[0] = acc0; slot[1] = acc1; slot[2] = acc2; slot[3] = acc3;
      }
      named_bar_sync(1 + grp, LPG);
      acc0 = acc1 = acc2 = acc3 = 0.0;
#pragma unroll
      for (int w = 0; w < W; w++) {
        const double *slot = scr.red[grp][par][w];
        acc0 += slot[0]; acc1 += slot[1]; acc2 += slot[2]; acc3 += slot[3];
      }
    }
    // M-step and convergence test (reference gen_func.cpp:1049-1055: eps = max |f - f_last| < 1e-5)
    acc0 *= f[0]; acc1 *= f[1]; acc2 *= f[2]; acc3 *= f[3];
    const double n0 = acc0 * inv_x, n1 = acc1 * inv_x, n2 = acc2 * inv_x, n3 = acc3 * inv_x;
    double eps = 0.0, d;
    d = fabs(n0 - f[0]); if (d > eps) eps = d;
    d = fabs(n1 - f[1]); if (d > eps) eps = d;
    d = fabs(n2 - f[2]); if (d > eps) eps = d;
    d = fabs(n3 - f[3]); if (d > eps) eps = d;
    if (!done) {
      f[0] = n0; f[1] = n1; f[2] = n2; f[3] = n3;
      const bool conv = eps < NGSLD_EPS;
      if (conv || it == NGSLD_ITER_MAX - 1) {
        if (glane == 0) {
          // Output M-step in the reference's own arithmetic (gen_func.cpp:1108-1113): true divisions and the
          // sequential renormalisation, so that exactly-degenerate pairs (a frequency sum of exactly 1 or 0)
          // land on the same 0/0 -> NaN outcomes as the reference.  Once per pair, not per pass.
          const double xd = (double)n_used;
          double g[4] = {__ddiv_rn(acc0, xd), __ddiv_rn(acc1, xd), __ddiv_rn(acc2, xd), __ddiv_rn(acc3, xd)};
#pragma unroll
          for (int k = 0; k < 4; k++)
            g[k] = __ddiv_rn(g[k], __dadd_rn(__dadd_rn(__dadd_rn(g[0], g[1]), g[2]), g[3]));
          ngsld_pair_row *row = rows_out + out_index;
          derive_and_store(row, g, conv ? it : (uint32_t)NGSLD_ITER_MAX, n_used);
          my_passes += it + 1;
        }
        need = true;
      } else {
        it++;
      }
    }
  }
  if (my_passes) atomicAdd(pass_counter, my_passes);
}

}  // namespace emfast
