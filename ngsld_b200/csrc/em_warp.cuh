// Warp-per-pair fast EM (replaces haplo_freq + pair_freq_iter, reference shared/gen_func.cpp:1027-1119)
// for samples of 160 individuals and more (one warp per pair up to ~580, G = 2 or 4 warps per pair beyond).
//
// One warp owns one site pair for its whole EM; at G = 1 warps never synchronise with each other.  Each lane
// keeps R individuals of the pair in registers (their six genotype likelihoods p[3], q[3]); the rest of
// the two rows (individuals [32R, n_ind)) is staged ONCE per pair into a warp-private shared-memory
// slice by two TMA bulk copies and re-read from there on every pass with conflict-free 128-bit loads
// (a lane owns two adjacent individuals = 48 contiguous bytes of each row).  Registers + shared memory
// together hold the pair, so the <=100 passes never touch L2/HBM, and the footprint per warp (152 registers,
// ~15 KB of shared memory for 500 individuals at R = 6) leaves three warps per scheduler resident plus one CTA
// of the integer-only r2_ExpG kernel, which runs in the issue slots this FP64-bound kernel leaves idle.
//
// Per pass and individual (27 FP64 + 1 MUFU):
//     u0 = f0 q0 + f1 q1   u1 = f2 q0 + f3 q1   v0 = f0 q1 + f1 q2   v1 = f2 q1 + f3 q2        (8)
//     i0 = p0 u0 + p1 u1   i1 = p0 v0 + p1 v1   i2 = p1 u0 + p2 u1   i3 = p1 v0 + p2 v1        (8)
//     s  = f0 i0 + f1 i1 + f2 i2 + f3 i3   == the reference's `sum`                             (4)
//     a_k += i_k / s                                                                       (3 + 4)
// and per pass once: A_k = sum over lanes of a_k (transposing butterfly emfast::group_sum4: 6 adds instead of 20),
// f_k <- f_k A_k / n_used.  f_k i_k is the reference's tmp_k / 2 and its M-step is ff/(2x), so this is
// the same fixed-point iteration; results agree with the bit-faithful kernel to ~1e-15 at equal nIter.
#pragma once
#include "common.cuh"
#include "em_fast.cuh"     // rcp_fast
#include "em_kernels.cuh"  // mbarrier / TMA helpers

namespace emwarp {

constexpr int WARPS_PER_CTA = 4;
constexpr int CTA_THREADS = 32 * WARPS_PER_CTA;

using emfast::Ind;
using emfast::estep;

// How a row of n_pad individual slots is divided among the G warps of a pair (host and device agree on this).
struct WarpGeom {
  // slots per warp: even (16-byte aligned slices) and a multiple of 2
  __host__ __device__ static inline uint32_t slice(uint32_t n_pad, int g) {
    const uint32_t s = (n_pad + (uint32_t)g - 1u) / (uint32_t)g;
    return (s + 1u) & ~1u;
  }
  // shared-memory slots per warp and row: the largest tail among the group's warps (the first warp's)
  __host__ __device__ static inline uint32_t tail_slots(uint32_t n_pad, int g, int r) {
    const uint32_t s = slice(n_pad, g) < n_pad ? slice(n_pad, g) : n_pad;
    return s > 32u * (uint32_t)r ? s - 32u * (uint32_t)r : 0u;
  }
};

__device__ __forceinline__ void accum(double &a0, double &a1, double &a2, double &a3, double i0, double i1, double i2,
                                      double i3, double inv) {
  a0 = __fma_rn(i0, inv, a0);
  a1 = __fma_rn(i1, inv, a1);
  a2 = __fma_rn(i2, inv, a2);
  a3 = __fma_rn(i3, inv, a3);
}

__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// A lane's two adjacent individuals of one tail iteration: 48 contiguous bytes of each row slice.
struct TailVec {
  double2 a0, a1, a2, b0, b1, b2;
  __device__ __forceinline__ void load(uint32_t sa, uint32_t sb) {
    a0 = lds128(sa); a1 = lds128(sa + 16); a2 = lds128(sa + 32);
    b0 = lds128(sb); b1 = lds128(sb + 16); b2 = lds128(sb + 32);
  }
  template <bool MASKED>
  __device__ __forceinline__ void run(double f0, double f1, double f2, double f3, double &s0, double &s1, double &s2,
                                      double &s3, bool ok_a, bool ok_b) const {
    Ind ga, gb;
    ga.p0 = a0.x; ga.p1 = a0.y; ga.p2 = a1.x; gb.p0 = a1.y; gb.p1 = a2.x; gb.p2 = a2.y;
    ga.q0 = b0.x; ga.q1 = b0.y; ga.q2 = b1.x; gb.q0 = b1.y; gb.q1 = b2.x; gb.q2 = b2.y;
    double i0, i1, i2, i3, s, k0, k1, k2, k3, t;
    estep(f0, f1, f2, f3, ga, i0, i1, i2, i3, s);
    estep(f0, f1, f2, f3, gb, k0, k1, k2, k3, t);
    double inv_a = emfast::rcp_fast(s), inv_b = emfast::rcp_fast(t);
    if (MASKED) {
      if (!ok_a) inv_a = 0.0;
      if (!ok_b) inv_b = 0.0;
    }
    accum(s0, s1, s2, s3, i0, i1, i2, i3, inv_a);
    accum(s0, s1, s2, s3, k0, k1, k2, k3, inv_b);
  }
};

// R    individuals per lane held in registers (individuals lane + 32 r, r < R)
// IGN  --ignore_miss_data: individuals whose likelihoods are flat at either site are left out
// Dynamic shared memory: WARPS_PER_CTA * 2 * tail_bytes, tail_bytes = (n_pad - 32 R) * 24 (0 if n_pad <= 32 R).
// U    full tail iterations fused per loop trip (2U individuals in flight per lane)
// G    warps per pair (1, 2 or 4): warp w of a group owns the individuals [w*slice, (w+1)*slice) of both rows
//      (slice = WarpGeom::slice(n_pad, G)); the G partial sums of a pass meet in shared memory behind a named
//      barrier, and every warp of the group then performs the identical M-step.  G > 1 keeps the per-warp
//      footprint of a 500-individual pair for samples of 1000 (G = 2) or 2000 (G = 4) individuals.
// Register cap: 152 (R >= 5) keeps three CTAs (58 K registers) plus one CTA of the r2_ExpG kernel (5 K) resident per SM;
// 128 (R <= 4) allows four CTAs.
// sel / n_sel: when sel is not NULL the kernel works through the pairs sel[0 .. *n_sel) (indices into the chunk) instead of
//      all of them: the pairs the class-compressed kernel (em_cell.cuh) left over, counted on the device.
template <int R, bool IGN, int U, int G>
__global__ void __maxnreg__(R <= 4 ? 128 : 152) em_warp_kernel(SiteTable T, PairChunk C, DevCounters *ctr, const uint32_t *sel,
                                                               const unsigned long long *n_sel) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ __align__(8) uint64_t bars[WARPS_PER_CTA];
  __shared__ double red[WARPS_PER_CTA / G][2][G][4];            // per group, double-buffered by pass parity
  __shared__ unsigned long long fetched[WARPS_PER_CTA / G];     // pair index handed to the group's warps
  __shared__ unsigned int used_cnt[WARPS_PER_CTA / G][G];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = warp / G, gw = warp % G;                      // group in the CTA, warp in the group
  const uint32_t slice = WarpGeom::slice(T.n_pad, G);
  const uint32_t lo = (uint32_t)gw * slice;                     // first individual of this warp's slice
  const uint32_t n_ind = T.n_ind > lo ? (T.n_ind - lo < slice ? T.n_ind - lo : slice) : 0u;   // individuals in the slice
  const uint32_t n_padw = T.n_pad > lo ? (T.n_pad - lo < slice ? T.n_pad - lo : slice) : 0u;  // row slots in the slice
  const uint32_t reg_n = n_ind < 32u * R ? n_ind : 32u * R;       // individuals covered by registers
  const uint32_t tail_n = n_ind - reg_n;                          // individuals streamed from shared memory
  const uint32_t tail_pad = n_padw > 32u * R ? n_padw - 32u * R : 0u;
  const uint32_t tail_bytes = tail_pad * 24u;                     // this warp's copy size
  const uint32_t slot_bytes = WarpGeom::tail_slots(T.n_pad, G, R) * 24u;  // every warp's slot size (the largest tail)
  const size_t row_doubles = (size_t)T.n_pad * 3;
  const double *tail_a = reinterpret_cast<const double *>(dyn_smem + (size_t)warp * 2 * slot_bytes);
  const double *tail_b = reinterpret_cast<const double *>(dyn_smem + (size_t)warp * 2 * slot_bytes + slot_bytes);
  uint64_t *bar = &bars[warp];
  if (lane == 0) emfast::mbar_init(bar, 1);
  __syncwarp();
  uint32_t phase = 0;
  unsigned long long my_passes = 0;
  const uint32_t n_iter_tail = (tail_n + 63u) / 64u;  // each lane handles individuals 2*lane + 64*j (+1)
  const uint32_t n_full = tail_n / 64u;               // iterations in which every lane has two valid individuals
  const uint32_t sa = emfast::smem_u32(dyn_smem) + (uint32_t)warp * 2u * slot_bytes + (uint32_t)lane * 48u;
  const uint32_t sb = sa + slot_bytes;

  const unsigned long long n_work = sel ? *n_sel : C.n_pairs;
  unsigned long long *next = sel ? &ctr->next_resid : &ctr->next_pair;
  for (;;) {
    unsigned long long idx = 0;
    if (G == 1) {
      if (lane == 0) idx = atomicAdd(next, 1ull);
      idx = __shfl_sync(0xffffffffu, idx, 0);
    } else {
      // the last reduction barrier of the previous pair orders every read of fetched[] before this write
      if (gw == 0 && lane == 0) fetched[grp] = atomicAdd(next, 1ull);
      emfast::named_bar_sync(1 + grp, 32 * G);
      idx = fetched[grp];
    }
    if (idx >= n_work) break;
    if (sel) idx = sel[idx];
    const uint32_t s1 = C.s1[idx], s2 = C.s2[idx];
    const double *row_a = T.gl + ((size_t)s1 * T.n_pad + lo) * 3, *row_b = T.gl + ((size_t)s2 * T.n_pad + lo) * 3;

    // ---- stage the row tails (async proxy) while the register part is loaded ----
    if (tail_bytes) {
      __syncwarp();  // every lane is done reading the previous pair's tails
      if (lane == 0) {
        emfast::mbar_expect_tx(bar, 2 * tail_bytes);
        emfast::tma_row_load(const_cast<double *>(tail_a), row_a + (size_t)32 * R * 3, tail_bytes, bar);
        emfast::tma_row_load(const_cast<double *>(tail_b), row_b + (size_t)32 * R * 3, tail_bytes, bar);
      }
    }
    Ind g[R];
    uint32_t rmask = 0;  // bit r: register individual r takes part
#pragma unroll
    for (int r = 0; r < R; r++) {
      const uint32_t i = (uint32_t)lane + 32u * r;
      bool ok = i < reg_n;
      g[r].p0 = g[r].p1 = g[r].p2 = g[r].q0 = g[r].q1 = g[r].q2 = 0.0;
      if (ok) {
        const double *pa = row_a + 3 * (size_t)i, *pb = row_b + 3 * (size_t)i;
        g[r].p0 = pa[0]; g[r].p1 = pa[1]; g[r].p2 = pa[2];
        g[r].q0 = pb[0]; g[r].q1 = pb[1]; g[r].q2 = pb[2];
        if (IGN && (gl_missing(g[r].p0, g[r].p1, g[r].p2) || gl_missing(g[r].q0, g[r].q1, g[r].q2))) ok = false;
      }
      rmask |= (ok ? 1u : 0u) << r;
    }
    const bool reg_full = reg_n == 32u * R;  // warp-uniform: no register slot is out of range
    if (tail_bytes) {
      emfast::mbar_wait(bar, phase);
      phase ^= 1;
    }
    // tail mask (only with IGN): bit 2j / 2j+1 = the lane's two individuals of tail iteration j take part
    unsigned long long tmask = ~0ull;
    uint32_t n_used = T.n_ind;
    if (IGN) {
      tmask = 0;
      for (uint32_t j = 0; j < n_iter_tail; j++) {
        const uint32_t i0 = 2u * lane + 64u * j;
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const uint32_t i = i0 + e;
          if (i < tail_n) {
            const double *pa = tail_a + 3 * (size_t)i, *pb = tail_b + 3 * (size_t)i;
            if (!(gl_missing(pa[0], pa[1], pa[2]) || gl_missing(pb[0], pb[1], pb[2]))) tmask |= 1ull << (2 * j + e);
          }
        }
      }
      uint32_t cnt = __popc(rmask) + __popcll(tmask);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (G > 1) {
        if (lane == 0) used_cnt[grp][gw] = cnt;
        emfast::named_bar_sync(1 + grp, 32 * G);
        cnt = 0;
#pragma unroll
        for (int w = 0; w < G; w++) cnt += used_cnt[grp][w];
      }
      n_used = cnt;
    }
    const double inv_x = __ddiv_rn(1.0, (double)n_used);
    const double m1 = T.maf[s1], m2 = T.maf[s2];  // haplo_freq start point, gen_func.cpp:1034-1037
    double f0 = __dmul_rn(__dsub_rn(1.0, m1), __dsub_rn(1.0, m2));
    double f1 = __dmul_rn(__dsub_rn(1.0, m1), m2);
    double f2 = __dmul_rn(m1, __dsub_rn(1.0, m2));
    double f3 = __dmul_rn(m1, m2);
    double A0 = 0, A1 = 0, A2 = 0, A3 = 0;
    uint32_t it = 0;
    bool conv = false;

    for (;;) {
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      // ---- register-resident individuals ----
      if (IGN || !reg_full) {
#pragma unroll
        for (int r = 0; r < R; r++) {
          double i0, i1, i2, i3, s;
          estep(f0, f1, f2, f3, g[r], i0, i1, i2, i3, s);
          double inv = emfast::rcp_fast(s);
          if (!((rmask >> r) & 1u)) inv = 0.0;
          accum(a0, a1, a2, a3, i0, i1, i2, i3, inv);
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; r++) {
          double i0, i1, i2, i3, s;
          estep(f0, f1, f2, f3, g[r], i0, i1, i2, i3, s);
          accum(a0, a1, a2, a3, i0, i1, i2, i3, emfast::rcp_fast(s));
        }
      }
      // ---- individuals streamed from the warp's shared-memory slice ----
      uint32_t j = 0;
      if (!IGN) {
        // iterations in which all 32 lanes own two valid individuals: no masks, U iterations fused for ILP
        for (; j + U <= n_full; j += U) {
          TailVec v[U];
#pragma unroll
          for (int u = 0; u < U; u++) v[u].load(sa + (j + u) * 1536u, sb + (j + u) * 1536u);
#pragma unroll
          for (int u = 0; u < U; u++) v[u].template run<false>(f0, f1, f2, f3, a0, a1, a2, a3, true, true);
        }
      }
      for (; j < n_iter_tail; j++) {
        const uint32_t i0x = 2u * lane + 64u * j;
        if (i0x < tail_n) {
          TailVec v;
          v.load(sa + j * 1536u, sb + j * 1536u);
          bool ok_a = true, ok_b = i0x + 1 < tail_n;  // odd sample size: the pad slot is left out
          if (IGN) {
            ok_a = (tmask >> (2 * j)) & 1ull;
            ok_b = (tmask >> (2 * j + 1)) & 1ull;
          }
          v.template run<true>(f0, f1, f2, f3, a0, a1, a2, a3, ok_a, ok_b);
        }
      }
      emfast::group_sum4<32>(a0, a1, a2, a3, lane);
      if (G > 1) {
        const int par = it & 1;
        if (lane == 0) {
          double *slot = red[grp][par][gw];
          slot[0] = a0; slot[1] = a1; slot[2] = a2; slot[3] = a3;
        }
        emfast::named_bar_sync(1 + grp, 32 * G);
        a0 = a1 = a2 = a3 = 0.0;
#pragma unroll
        for (int w = 0; w < G; w++) {  // same order in every warp of the group: identical bits everywhere
          const double *slot = red[grp][par][w];
          a0 += slot[0]; a1 += slot[1]; a2 += slot[2]; a3 += slot[3];
        }
      }
      // ---- M-step and convergence test (reference gen_func.cpp:1049-1055: eps = max |f - f_last| < 1e-5) ----
      A0 = f0 * a0; A1 = f1 * a1; A2 = f2 * a2; A3 = f3 * a3;
      const double n0 = A0 * inv_x, n1 = A1 * inv_x, n2 = A2 * inv_x, n3 = A3 * inv_x;
      // eps starts at 0 and only a difference that compares greater raises it, as in the reference: NaN frequencies
      // (all-missing site under --ignore_miss_data, an individual with sum == 0) leave eps at 0 and stop the EM at once
      // with nIter = the current pass.  (fmax(NaN, NaN) = NaN would instead run all 100 passes.)
      double eps = fmax(0.0, fabs(n0 - f0));
      eps = fmax(eps, fabs(n1 - f1));
      eps = fmax(eps, fabs(n2 - f2));
      eps = fmax(eps, fabs(n3 - f3));
      f0 = n0; f1 = n1; f2 = n2; f3 = n3;
      conv = eps < NGSLD_EPS;
      if (conv || it == NGSLD_ITER_MAX - 1) break;
      it++;
    }
    if (lane == 0 && gw == 0) {
      // Output M-step in the reference's own arithmetic (gen_func.cpp:1108-1113): true divisions and the
      // sequential renormalisation, so exactly-degenerate pairs land on the same 0/0 -> NaN outcomes.
      const double xd = (double)n_used;
      double gq[4] = {__ddiv_rn(A0, xd), __ddiv_rn(A1, xd), __ddiv_rn(A2, xd), __ddiv_rn(A3, xd)};
#pragma unroll
      for (int k = 0; k < 4; k++)
        gq[k] = __ddiv_rn(gq[k], __dadd_rn(__dadd_rn(__dadd_rn(gq[0], gq[1]), gq[2]), gq[3]));
      derive_and_store(C.rows + idx, gq, conv ? it : (uint32_t)NGSLD_ITER_MAX, n_used);
      my_passes += it + 1;
    }
  }
  if (lane == 0 && my_passes) atomicAdd(&ctr->em_passes, my_passes);
}

struct WarpVariant {
  int r, g;
  const void *fn, *fn_ign, *fn_u1;  // default (2 tail iterations fused), --ignore_miss_data, unfused
};

}  // namespace emwarp
