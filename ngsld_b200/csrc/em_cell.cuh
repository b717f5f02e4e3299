// Class-compressed fast EM (replaces haplo_freq + pair_freq_iter, reference shared/gen_func.cpp:1027-1119).
//
// The reference's E-step treats every individual separately, but its contribution to a pass depends only on the
// individual's two genotype-likelihood triples (p at site 1, q at site 2).  Sequencing data at low depth has few
// DISTINCT triples per site -- a triple is a function of the reads an individual happens to have (no reads: flat;
// one read: two possibilities; called genotypes: three or four values in total) -- so the n_ind terms of
//     ff[k] += tmp_k / sum                                   (gen_func.cpp:1094-1104)
// collapse into one term per distinct (p, q) combination, weighted by the number of individuals that have it:
//     a_k = sum over cells c of  w_c * i_k(p_c, q_c) / s(p_c, q_c)
// This is the same fixed-point iteration as em_warp.cuh with the sum over individuals regrouped (equal terms added
// w times become one product), so results agree with the bit-faithful kernel to ~1e-15 at equal nIter, and the cost
// of a pass falls from n_ind E-steps to n_cells (BASELINE config 3, 500 individuals at 2x depth: ~160 cells).  With
// called genotypes (reference call_geno, gen_func.cpp:886-914) a pair has at most 16 cells whatever the sample size.
//
// Per site, build_palette_kernel lists the distinct triples (the "palette", at most NGSLD_KMAX entries, bitwise
// equality) and codes every individual as one byte.  Per pair, a warp
//   1. counts the joint classes (c1[i], c2[i]) of the individuals in a K x K table of 16-bit counters in shared
//      memory (lanes with equal keys are combined with match.any, so no two lanes touch the same counter), noting every
//      counter that becomes non-zero in a list: the pair's cells;
//   2. loads the cells -- palette triples + weight -- into registers (R per lane; cells beyond 32 R go to a shared-
//      memory tail) and zeroes the counters it used;
//   3. iterates the EM exactly like em_warp.cuh, one weighted E-step per cell.
// (Step 1 was also tried with shared-memory atomics and a sweep of the table instead of match.any and lane-by-lane counter
// updates -- the MATCH instruction alone holds 8 % of the kernel's stall samples -- and measured 9 % SLOWER: sixteen
// atomics per lane and pair compete with the shuffles of the EM reduction for the same data path.  profiles/README.md.)
// Pairs it cannot take (a site with more than NGSLD_KMAX distinct triples, more cells than the warp has room for) are
// appended to a list that the dense warp-per-pair kernel processes right afterwards.
//
// r2_ExpG (pearson.cuh) is order-dependent 80-bit arithmetic and cannot be regrouped; with FUSE it runs inside this
// kernel -- every warp first does the 32 pairs of its batch one pair per lane, then their EMs one after the other -- so
// the integer-only emulation and the FP64-bound EM share every SM without a second kernel to balance against.
#pragma once
#include "common.cuh"
#include "em_fast.cuh"
#include "pearson.cuh"

// compile-time experiment switches (A/B builds: make ALT=_x ALTFLAGS=-D... lib in csrc/Makefile, scripts/gpu_job.sh ab)
// 1: the four new frequencies of a pass go round through shared memory instead of eight shuffles.  Measured slower
// (53.5 vs 58.2 M pairs/s together with a preload of the class words, round 2): stays off.
#ifndef NGSLD_CELL_SMEM_BCAST
#define NGSLD_CELL_SMEM_BCAST 0
#endif
// 1: loop bodies without any shared-memory-tail code for pairs that have no tail.  34 instructions fewer per pass (236
// instead of 270 at six register levels), but measured SLOWER at 500 and 2000 individuals (59.4 vs 61.4 and 22.1 vs 23.0
// M pairs/s, round 2): in that form ptxas keeps the four haplotype frequencies in vector registers instead of uniform
// registers, and an FP64 instruction that reads three vector registers issues at 0.72 of the rate of one that reads two.
#ifndef NGSLD_CELL_NOTAIL_BODIES
#define NGSLD_CELL_NOTAIL_BODIES 0
#endif
// 1: while a pair is iterated, the class row and the palette of the NEXT pair's second site are pulled into L1, so that
// the construction of its cells does not start with a round trip to L2 (+0.3 % at 500, +1.7 % at 2000 individuals).
#ifndef NGSLD_CELL_PREFETCH
#define NGSLD_CELL_PREFETCH 1
#endif
// 1: the per-pair totals wait for the end of the batch in the pair's own output row (global memory, L2) instead of 5 KB of
// shared memory per CTA.  Meant to make room for a fourth CTA per SM when the tail holds 192 cells (2000 individuals, where
// ncu shows three); measured the same there (27.10 vs 27.13 M pairs/s) and -0.3 % at 500 individuals: stays off.
#ifndef NGSLD_CELL_STAGE_GLOBAL
#define NGSLD_CELL_STAGE_GLOBAL 0
#endif
// 1: joint classes with the class words of the next block of individuals requested ahead and the four match.any of a
// block issued together; 0: one match per counter update (+2 % at 500, +10 % at 2000 individuals for 1, round 2).
#ifndef NGSLD_CELL_MATCH4
#define NGSLD_CELL_MATCH4 1
#endif

namespace emcell {

constexpr int WARPS_PER_CTA = 4;
constexpr int CTA_THREADS = 32 * WARPS_PER_CTA;
constexpr int NBINS = NGSLD_KMAX * NGSLD_KMAX;  // largest joint-class table (the kernels size theirs from the data)

using emfast::Ind;
using emfast::estep;

struct CellArgs {
  uint32_t *resid;    // [chunk rows] indices (inside the chunk) of the pairs left to the dense kernel
  uint32_t tcap;      // cells a warp can hold in shared memory beyond its 32 R register cells (multiple of 64)
  int ignore_miss;    // --ignore_miss_data: individuals whose class is flat at either site are left out
  int fuse_pearson;   // compute r2_ExpG in this kernel
  uint32_t kstride;   // row length of the joint-class table: the largest palette of the data set, rounded up to 8
  // Order in which the batches of 32 pairs are worked through.  The chunk's pairs are listed first site by first site;
  // taken in that order, the warps resident at one time share a first site and sweep the whole second-site axis, so every
  // second site's data (x87 terms, class row, palette) comes from DRAM once per first site.  With swz_len = L > 0 the
  // list is read as a matrix of rows of L pairs (about one first site per row) and walked 32 columns at a time, all rows
  // of a column block before the next block: the second sites of a block are then fetched once and serve every first
  // site of the chunk out of L2.  Only the order of the work changes, not where a pair's row is written.
  uint32_t swz_len, swz_rows;
};

// shared memory of one warp: joint-class counters | cell keys | tail cells (7 doubles each, structure of arrays)
__host__ __device__ inline size_t bins_bytes(uint32_t kstride) { return ((size_t)kstride * kstride * 2 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t warp_smem_bytes(int r, uint32_t tcap, uint32_t kstride) {
  const size_t keys = ((size_t)(32 * r + tcap) * 2 + 15) & ~(size_t)15;
  return bins_bytes(kstride) + keys + (size_t)tcap * 7 * 8;
}

__device__ __forceinline__ void wipe_bins(uint16_t *bins, uint32_t kstride, int lane) {
  uint4 *b = reinterpret_cast<uint4 *>(bins);
  for (int k = lane; k < (int)(bins_bytes(kstride) / 16); k += 32) b[k] = make_uint4(0, 0, 0, 0);
  __syncwarp();
}

// Joint classes of a pair.  bins must be all zero on entry; on return bins[c1 * kstride + c2] = individuals with classes
// (c1, c2), keys[0 .. min(n_cells, cap)) = the distinct (c1 << 8 | c2) in order of first appearance, n_used = individuals
// counted.  Returns n_cells (if it exceeds cap the caller has to wipe the whole table).
__device__ __forceinline__ uint32_t joint_classes(const SiteTable &T, uint32_t s1, uint32_t s2, bool ign, uint16_t *bins,
                                                  uint32_t kstride, uint16_t *keys, uint32_t cap, uint32_t &n_used, int lane) {
  const uint8_t *c1 = T.cls + (size_t)s1 * T.n_cpad, *c2 = T.cls + (size_t)s2 * T.n_cpad;
  const uint64_t miss1 = ign ? T.pal_miss[s1] : 0ull, miss2 = ign ? T.pal_miss[s2] : 0ull;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t n_cells = 0, used = 0;
#if NGSLD_CELL_MATCH4
  // The class words of block k + 1 are requested before block k is counted, and the four match.any of a block -- they
  // depend on the keys only, not on the counters -- are issued back to back before the four counter updates, which are
  // the sequential part: the MATCH instruction is slow (it alone held 8 % of the kernel's stall samples when every
  // update waited for its own match).
  uint32_t nw1 = 0, nw2 = 0;
  if (4u * (uint32_t)lane < T.n_ind) {
    nw1 = *reinterpret_cast<const uint32_t *>(c1 + 4u * (uint32_t)lane);
    nw2 = *reinterpret_cast<const uint32_t *>(c2 + 4u * (uint32_t)lane);
  }
  for (uint32_t blk = 0; blk < T.n_ind; blk += 128u) {
    const uint32_t i0 = blk + 4u * (uint32_t)lane;  // this lane's four individuals of the block
    const uint32_t w1 = nw1, w2 = nw2;
    if (i0 + 128u < T.n_ind) {
      nw1 = *reinterpret_cast<const uint32_t *>(c1 + i0 + 128u);
      nw2 = *reinterpret_cast<const uint32_t *>(c2 + i0 + 128u);
    }
    uint32_t key[4], peers[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const uint32_t a = (w1 >> (8 * e)) & 255u, b = (w2 >> (8 * e)) & 255u;
      bool valid = i0 + e < T.n_ind;
      if (valid && ign) valid = !((((miss1 >> a) | (miss2 >> b)) & 1ull) != 0);
      key[e] = valid ? (a << 8 | b) : 0xffffffffu;
    }
#pragma unroll
    for (int e = 0; e < 4; e++) peers[e] = __match_any_sync(0xffffffffu, key[e]);
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const bool valid = key[e] != 0xffffffffu;
      const bool lead = valid && (__ffs(peers[e]) - 1 == lane);
      bool fresh = false;
      if (lead) {
        const uint32_t bin = (key[e] >> 8) * kstride + (key[e] & 255u);
        const uint32_t old = bins[bin];
        bins[bin] = (uint16_t)(old + __popc(peers[e]));
        fresh = old == 0;
      }
      const uint32_t fm = __ballot_sync(0xffffffffu, fresh);
      if (fresh) {
        const uint32_t slot = n_cells + __popc(fm & lt);
        if (slot < cap) keys[slot] = (uint16_t)key[e];
      }
      n_cells += __popc(fm);
      if (ign) used += __popc(__ballot_sync(0xffffffffu, valid));  // (without ignore_miss everybody is counted)
      __syncwarp();  // the counters written here are read by whichever lane leads the key next time
    }
  }
  if (!ign) used = T.n_ind;
#else
  for (uint32_t blk = 0; blk < T.n_ind; blk += 128u) {
    const uint32_t i0 = blk + 4u * (uint32_t)lane;  // this lane's four individuals of the block
    uint32_t w1 = 0, w2 = 0;
    if (i0 < T.n_ind) {
      w1 = *reinterpret_cast<const uint32_t *>(c1 + i0);
      w2 = *reinterpret_cast<const uint32_t *>(c2 + i0);
    }
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const uint32_t a = (w1 >> (8 * e)) & 255u, b = (w2 >> (8 * e)) & 255u;
      bool valid = i0 + e < T.n_ind;
      if (valid && ign) valid = !((((miss1 >> a) | (miss2 >> b)) & 1ull) != 0);
      const uint32_t key = valid ? (a << 8 | b) : 0xffffffffu;
      const uint32_t peers = __match_any_sync(0xffffffffu, key);
      const bool lead = valid && (__ffs(peers) - 1 == lane);
      bool fresh = false;
      if (lead) {
        const uint32_t bin = a * kstride + b;
        const uint32_t old = bins[bin];
        bins[bin] = (uint16_t)(old + __popc(peers));
        fresh = old == 0;
      }
      const uint32_t fm = __ballot_sync(0xffffffffu, fresh);
      if (fresh) {
        const uint32_t slot = n_cells + __popc(fm & lt);
        if (slot < cap) keys[slot] = (uint16_t)key;
      }
      n_cells += __popc(fm);
      used += __popc(__ballot_sync(0xffffffffu, valid));
      __syncwarp();  // the counters written here are read by whichever lane leads the key next time
    }
  }
#endif
  n_used = used;
  return n_cells;
}

struct Cell {
  Ind g;
  double w;
};

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// The second site's class row (n_cpad bytes) and palette (pal_k triples): one 128-byte line per lane.
__device__ __forceinline__ void prefetch_site(const SiteTable &T, uint32_t s, int lane) {
  const uint32_t off = 128u * (uint32_t)lane;
  if (off < T.n_cpad) prefetch_l1(T.cls + (size_t)s * T.n_cpad + off);
  if (off < (uint32_t)NGSLD_KMAX * 24u) prefetch_l1(reinterpret_cast<const char *>(T.pal + (size_t)s * NGSLD_KMAX * 3) + off);
}

__device__ __forceinline__ void cell_default(Cell &c) {  // an empty slot: weight 0, likelihoods that keep s finite
  c.g.p0 = c.g.p1 = c.g.p2 = c.g.q0 = c.g.q1 = c.g.q2 = 1.0;
  c.w = 0.0;
}

__device__ __forceinline__ void cell_load(Cell &c, const SiteTable &T, uint32_t s1, uint32_t s2, uint32_t key, uint16_t *bins,
                                          uint32_t kstride) {
  const uint32_t a = key >> 8, b = key & 255u;
  const double *pa = T.pal + ((size_t)s1 * NGSLD_KMAX + a) * 3;
  const double *pb = T.pal + ((size_t)s2 * NGSLD_KMAX + b) * 3;
  c.g.p0 = __ldg(pa); c.g.p1 = __ldg(pa + 1); c.g.p2 = __ldg(pa + 2);
  c.g.q0 = __ldg(pb); c.g.q1 = __ldg(pb + 1); c.g.q2 = __ldg(pb + 2);
  const uint32_t bin = a * kstride + b;
  c.w = (double)bins[bin];
  bins[bin] = 0;  // every counter in use belongs to exactly one cell: the table is clean again after the loads
}

__device__ __forceinline__ void cell_step(const double f0, const double f1, const double f2, const double f3, const Cell &c,
                                          double &a0, double &a1, double &a2, double &a3) {
  double i0, i1, i2, i3, s;
  estep(f0, f1, f2, f3, c.g, i0, i1, i2, i3, s);
  const double wi = c.w * emfast::rcp_fast(s);
  a0 = __fma_rn(i0, wi, a0);
  a1 = __fma_rn(i1, wi, a1);
  a2 = __fma_rn(i2, wi, a2);
  a3 = __fma_rn(i3, wi, a3);
}

// Sum of a0..a3 over the warp, transposed: lane l receives the total of component q(l) = 2 * bit4(l) + bit3(l) only
// (lanes 0-7: a0, 8-15: a1, 16-23: a2, 24-31: a3).  Same butterfly as emfast::group_sum4 without its final four
// broadcasts; every lane of a quadrant ends up with identical bits (each round adds the two partners' values in both).
__device__ __forceinline__ double warp_sum4_own(double a0, double a1, double a2, double a3, int lane) {
  const bool hi = lane & 16, lo = lane & 8;
  double x0 = hi ? a2 : a0, x1 = hi ? a3 : a1;
  const double y0 = hi ? a0 : a2, y1 = hi ? a1 : a3;
  x0 += __shfl_xor_sync(0xffffffffu, y0, 16);
  x1 += __shfl_xor_sync(0xffffffffu, y1, 16);
  double z = lo ? x1 : x0;
  const double w = lo ? x0 : x1;
  z += __shfl_xor_sync(0xffffffffu, w, 8);
  z += __shfl_xor_sync(0xffffffffu, z, 4);
  z += __shfl_xor_sync(0xffffffffu, z, 2);
  z += __shfl_xor_sync(0xffffffffu, z, 1);
  return z;
}

// The EM of one pair on its cells: L register levels in use (straight-line, so that the L independent dependency chains
// interleave -- a branch per level would serialise them) + the shared-memory tail.  Per pass every lane performs the
// M-step and the convergence test of ITS OWN frequency component only (f_q <- f_q A_q / n_used with A_q the warp total
// of component q, see warp_sum4_own) and the four new frequencies are then broadcast; `fq` / `Aq` are the lane's own
// component.  Convergence (reference gen_func.cpp:1049-1055): eps = max_k |f_k - f_last_k| < 1e-5 with a NaN difference
// never raising eps  <=>  no component has |difference| >= 1e-5.  Returns the index of the converging pass.
template <int L, int R, bool TAIL>
__device__ __forceinline__ uint32_t em_iterate(const Cell (&g)[R], const double *tail, uint32_t tcap, uint32_t n_tail_pad,
                                               double inv_x, int lane, double &f0, double &f1, double &f2, double &f3,
                                               double &Aq, bool &conv, uint32_t fbuf) {
  double fq = (lane & 16) ? ((lane & 8) ? f3 : f2) : ((lane & 8) ? f1 : f0);
  uint32_t it = 0;
  for (;;) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int r = 0; r < L; r++) cell_step(f0, f1, f2, f3, g[r], a0, a1, a2, a3);
    if (TAIL) for (uint32_t t = lane; t < n_tail_pad; t += 64u) {  // two cells per lane and trip
      Cell c, d;
      c.g.p0 = tail[0 * tcap + t]; c.g.p1 = tail[1 * tcap + t]; c.g.p2 = tail[2 * tcap + t];
      c.g.q0 = tail[3 * tcap + t]; c.g.q1 = tail[4 * tcap + t]; c.g.q2 = tail[5 * tcap + t];
      c.w = tail[6 * tcap + t];
      const uint32_t u = t + 32u;
      d.g.p0 = tail[0 * tcap + u]; d.g.p1 = tail[1 * tcap + u]; d.g.p2 = tail[2 * tcap + u];
      d.g.q0 = tail[3 * tcap + u]; d.g.q1 = tail[4 * tcap + u]; d.g.q2 = tail[5 * tcap + u];
      d.w = tail[6 * tcap + u];
      cell_step(f0, f1, f2, f3, c, a0, a1, a2, a3);
      cell_step(f0, f1, f2, f3, d, a0, a1, a2, a3);
    }
    const double z = warp_sum4_own(a0, a1, a2, a3, lane);
    Aq = fq * z;
    const double nq = Aq * inv_x;
    const bool moved = fabs(nq - fq) >= NGSLD_EPS;  // false for a NaN difference
    fq = nq;
#if NGSLD_CELL_SMEM_BCAST
    // the four new frequencies go round through four doubles of the warp's shared memory (one store by the first lane of
    // each quadrant, two 128-bit broadcast loads) instead of eight shuffles
    if ((lane & 7) == 0) asm volatile("st.shared.f64 [%0], %1;" ::"r"(fbuf + (uint32_t)(lane >> 3) * 8u), "d"(nq) : "memory");
    __syncwarp();
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(f0), "=d"(f1) : "r"(fbuf) : "memory");
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(f2), "=d"(f3) : "r"(fbuf + 16u) : "memory");
    __syncwarp();  // every lane has read the four values before the next pass stores again
    conv = !__any_sync(0xffffffffu, moved);
#else
    f0 = __shfl_sync(0xffffffffu, nq, 0);
    f1 = __shfl_sync(0xffffffffu, nq, 8);
    f2 = __shfl_sync(0xffffffffu, nq, 16);
    f3 = __shfl_sync(0xffffffffu, nq, 24);
    conv = !__any_sync(0xffffffffu, moved);
#endif
    if (conv || it == NGSLD_ITER_MAX - 1) break;
    it++;
  }
  return it;
}

// R     cells per lane held in registers (cell slot lane + 32 r); a pair may have up to 32 R + A.tcap cells.
// MINB  CTAs (of four warps) per SM the register allocation aims at: 3 -> 168 registers per thread, 4 -> 128.
template <int R, bool FUSE, int MINB>
__global__ void __launch_bounds__(CTA_THREADS, MINB) em_cell_kernel(SiteTable T, PairChunk C, CellArgs A, DevCounters *ctr) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ __align__(16) double fbuf_all[WARPS_PER_CTA][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t fbuf = (uint32_t)__cvta_generic_to_shared(fbuf_all[warp]);  // 32-bit shared address: one register
  const uint32_t cap = 32u * R + A.tcap;
  unsigned char *mine = dyn_smem + (size_t)warp * warp_smem_bytes(R, A.tcap, A.kstride);
  uint16_t *bins = reinterpret_cast<uint16_t *>(mine);
  uint16_t *keys = reinterpret_cast<uint16_t *>(mine + bins_bytes(A.kstride));
  double *tail = reinterpret_cast<double *>(mine + warp_smem_bytes(R, A.tcap, A.kstride) - (size_t)A.tcap * 56);
  wipe_bins(bins, A.kstride, lane);
  const bool ign = A.ignore_miss != 0;
  // Results of the batch's EMs wait (per pair: the four warp totals A_k of the last pass, n_used, nIter) until the whole
  // batch is iterated; then every lane finishes ONE pair -- the output M-step with its eight true divisions and D, D',
  // r2, chi2 (another six divisions and a square root) -- instead of lane 0 doing that for every pair while 31 lanes
  // wait (~600 instructions per pair in a single lane: measured +8 % for the whole kernel, round 2).  They wait in the
  // pair's own output row (hap[] <- A_k; n_used and nIter are final already) or, NGSLD_CELL_STAGE_GLOBAL = 0, in shared memory.
#if !NGSLD_CELL_STAGE_GLOBAL
  __shared__ __align__(16) double stageA_all[WARPS_PER_CTA][32][4];
  __shared__ uint2 stageB_all[WARPS_PER_CTA][32];
  double(*stageA)[4] = stageA_all[warp];
  uint2 *stageB = stageB_all[warp];
#endif
  // work statistics of this warp (passes, cell passes, cells, pairs, pairs left over): kept in shared memory, not in
  // ten registers that would be live through both phases of every batch
  __shared__ unsigned long long wstat_all[WARPS_PER_CTA][5];
  unsigned long long *wstat = wstat_all[warp];
  if (lane < 5) wstat[lane] = 0;
  __syncwarp();

  const unsigned long long n_batches =
      A.swz_len ? (unsigned long long)A.swz_rows * ((A.swz_len + 31u) / 32u) : (C.n_pairs + 31ull) / 32ull;
  for (;;) {
    unsigned long long b = 0;
    if (lane == 0) b = atomicAdd(&ctr->next_pair, 1ull);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= n_batches) break;
    unsigned long long base = b * 32ull;
    uint32_t width = 32;  // pairs of this batch
    if (A.swz_len) {
      const unsigned long long cb = b / A.swz_rows, r = b - cb * A.swz_rows;  // column block, matrix row
      const uint32_t c0 = (uint32_t)cb * 32u;
      base = r * A.swz_len + c0;
      width = A.swz_len - c0 < 32u ? A.swz_len - c0 : 32u;
    }
    if (base >= C.n_pairs) continue;
    if (C.n_pairs - base < width) width = (uint32_t)(C.n_pairs - base);
    const unsigned long long me = base + lane;
    const bool have = (uint32_t)lane < width;
    uint32_t my_s1 = 0, my_s2 = 0;
    if (have) {
      my_s1 = C.s1[me];
      my_s2 = C.s2[me];
    }
    // ---- r2_ExpG of the batch, one pair per lane ----
    if (FUSE && have) C.rows[me].r2_expg = pearson::pair_r2(T, my_s1, my_s2);
    __syncwarp();
    const int nb = (int)width;

    // ---- the batch's EMs, the whole warp on one pair at a time ----
    uint32_t done = 0;  // bit j: pair j of the batch was iterated here (not left to the dense kernel)
    for (int j = 0; j < nb; j++) {
      const uint32_t s1 = __shfl_sync(0xffffffffu, my_s1, j), s2 = __shfl_sync(0xffffffffu, my_s2, j);
      const unsigned long long idx = base + j;
      const uint32_t k1 = T.pal_k[s1], k2 = T.pal_k[s2];
      uint32_t n_cells = 0, n_used = T.n_ind;
      bool mine_ok = k1 != 0 && k2 != 0;
      if (mine_ok) {
        n_cells = joint_classes(T, s1, s2, ign, bins, A.kstride, keys, cap, n_used, lane);
        if (n_cells > cap) {
          wipe_bins(bins, A.kstride, lane);
          mine_ok = false;
        }
      }
      if (!mine_ok) {  // left to the dense kernel
        if (lane == 0) {
          A.resid[atomicAdd(&ctr->n_resid, 1ull)] = (uint32_t)idx;
          wstat[4]++;
        }
        continue;
      }
      // ---- cells -> registers (slot lane + 32 r) and the shared-memory tail (slots 32 R ...) ----
      Cell g[R];
#pragma unroll
      for (int r = 0; r < R; r++) {
        const uint32_t slot = (uint32_t)lane + 32u * r;
        cell_default(g[r]);
        if (slot < n_cells) cell_load(g[r], T, s1, s2, keys[slot], bins, A.kstride);
      }
      const uint32_t n_tail = n_cells > 32u * R ? n_cells - 32u * R : 0u;
      const uint32_t n_tail_pad = (n_tail + 63u) & ~63u;  // two cells per lane and trip: filled up with empty cells
      const uint32_t n_lev = n_cells >= 32u * R ? (uint32_t)R : (n_cells + 31u) / 32u;  // register levels in use
      for (uint32_t t = lane; t < n_tail_pad; t += 32u) {
        Cell c;
        cell_default(c);
        if (t < n_tail) cell_load(c, T, s1, s2, keys[32u * R + t], bins, A.kstride);
        tail[0 * A.tcap + t] = c.g.p0; tail[1 * A.tcap + t] = c.g.p1; tail[2 * A.tcap + t] = c.g.p2;
        tail[3 * A.tcap + t] = c.g.q0; tail[4 * A.tcap + t] = c.g.q1; tail[5 * A.tcap + t] = c.g.q2;
        tail[6 * A.tcap + t] = c.w;
      }
      __syncwarp();

#if NGSLD_CELL_PREFETCH
      if (j + 1 < nb) prefetch_site(T, __shfl_sync(0xffffffffu, my_s2, j + 1), lane);
#endif
      const double inv_x = __ddiv_rn(1.0, (double)n_used);
      const double m1 = T.maf[s1], m2 = T.maf[s2];  // haplo_freq start point, gen_func.cpp:1034-1037
      double f0 = __dmul_rn(__dsub_rn(1.0, m1), __dsub_rn(1.0, m2));
      double f1 = __dmul_rn(__dsub_rn(1.0, m1), m2);
      double f2 = __dmul_rn(m1, __dsub_rn(1.0, m2));
      double f3 = __dmul_rn(m1, m2);
      double Aq = 0;
      uint32_t it = 0;
      bool conv = false;
      // warp-uniform choice of the loop body: empty register levels are skipped as a whole (only a pair that fills every
      // register level can have cells in the shared-memory tail: n_tail_pad is 0 for the others)
      static_assert(R <= 6, "the switch below lists register levels 0..6");
      constexpr bool NT = NGSLD_CELL_NOTAIL_BODIES != 0;  // bodies for pairs without a tail carry no tail code
      switch (NT && n_tail ? 99u : n_lev) {
        case 0: it = em_iterate<0, R, !NT>(g, tail, A.tcap, n_tail_pad, inv_x, lane, f0, f1, f2, f3, Aq, conv, fbuf); break;
        case 1: it = em_iterate<1, R, !NT>(g, tail, A.tcap, n_tail_pad, inv_x, lane, f0, f1, f2, f3, Aq, conv, fbuf); break;
        case 2: it = em_iterate<(R < 2 ? R : 2), R, !NT>(g, tail, A.tcap, n_tail_pad, inv_x, lane, f0, f1, f2, f3, Aq, conv, fbuf); break;
        case 3: it = em_iterate<(R < 3 ? R : 3), R, !NT>(g, tail, A.tcap, n_tail_pad, inv_x, lane, f0, f1, f2, f3, Aq, conv, fbuf); break;
        case 4: it = em_iterate<(R < 4 ? R : 4), R, !NT>(g, tail, A.tcap, n_tail_pad, inv_x, lane, f0, f1, f2, f3, Aq, conv, fbuf); break;
        case 5: it = em_iterate<(R < 5 ? R : 5), R, !NT>(g, tail, A.tcap, n_tail_pad, inv_x, lane, f0, f1, f2, f3, Aq, conv, fbuf); break;
        case 6: if (NT) { it = em_iterate<(R < 6 ? R : 6), R, false>(g, tail, A.tcap, 0, inv_x, lane, f0, f1, f2, f3, Aq, conv, fbuf); break; }
        default: it = em_iterate<R, R, true>(g, tail, A.tcap, n_tail_pad, inv_x, lane, f0, f1, f2, f3, Aq, conv, fbuf); break;
      }
#if NGSLD_CELL_STAGE_GLOBAL
      if ((lane & 7) == 0) C.rows[idx].hap[lane >> 3] = Aq;  // the first lane of each quadrant holds the total of its component
      if (lane == 0) {
        C.rows[idx].n_used = n_used;
        C.rows[idx].n_iter = conv ? it : (uint32_t)NGSLD_ITER_MAX;
#else
      if ((lane & 7) == 0) stageA[j][lane >> 3] = Aq;
      if (lane == 0) {
        stageB[j] = make_uint2(n_used, conv ? it : (uint32_t)NGSLD_ITER_MAX);
#endif
        wstat[0] += it + 1;
        wstat[1] += (unsigned long long)(it + 1) * n_cells;
        wstat[2] += n_cells;
        wstat[3]++;
      }
      done |= 1u << j;
      __syncwarp();  // every lane is done with the tail before the next pair overwrites it
    }
    // ---- one pair per lane: output M-step in the reference's own arithmetic (gen_func.cpp:1108-1113: true divisions and
    // the sequential renormalisation, so exactly-degenerate pairs land on the same 0/0 -> NaN outcomes), then D, D', r2, chi2
    if ((done >> lane) & 1u) {
#if NGSLD_CELL_STAGE_GLOBAL
      const ngsld_pair_row *row = C.rows + base + lane;  // written by this warp before the __syncwarp that ended the last pair
      const uint2 sb = make_uint2(row->n_used, row->n_iter);
      const double A0 = row->hap[0], A1 = row->hap[1], A2 = row->hap[2], A3 = row->hap[3];
#else
      const uint2 sb = stageB[lane];
      const double A0 = stageA[lane][0], A1 = stageA[lane][1], A2 = stageA[lane][2], A3 = stageA[lane][3];
#endif
      const double xd = (double)sb.x;
      double gq[4] = {__ddiv_rn(A0, xd), __ddiv_rn(A1, xd), __ddiv_rn(A2, xd), __ddiv_rn(A3, xd)};
#pragma unroll
      for (int k = 0; k < 4; k++)
        gq[k] = __ddiv_rn(gq[k], __dadd_rn(__dadd_rn(__dadd_rn(gq[0], gq[1]), gq[2]), gq[3]));
      derive_and_store(C.rows + base + lane, gq, sb.y, sb.x);
    }
    __syncwarp();  // the staging area is free again
  }
  if (lane == 0) {
    if (wstat[0]) atomicAdd(&ctr->em_passes, wstat[0]);
    if (wstat[1]) atomicAdd(&ctr->cell_passes, wstat[1]);
    if (wstat[2]) atomicAdd(&ctr->cells, wstat[2]);
    if (wstat[3]) atomicAdd(&ctr->cell_pairs, wstat[3]);
    if (wstat[4]) atomicAdd(&ctr->resid_pairs, wstat[4]);
  }
}

// ---- per-site palettes ------------------------------------------------------------------------------------------
// One warp per site: the distinct triples of the site's row in order of first appearance (bitwise equality), the class
// of every individual, and which classes are "missing data".  A site with more than NGSLD_KMAX distinct triples gets
// pal_k = 0 and is not coded.
__global__ void __launch_bounds__(CTA_THREADS) build_palette_kernel(const double *gl, uint32_t n_sites, uint32_t n_ind, uint32_t n_pad, uint32_t n_cpad,
                                     uint8_t *cls, double *pal, uint8_t *pal_k, uint64_t *pal_miss, unsigned int *max_k);

// n_samples pseudo-random pairs: out[0] = sum of cells, out[1] = pairs sampled, out[2] = pairs with an uncoded site,
// hist[b] = pairs with 32 b <= cells < 32 (b + 1) (129 buckets).  Decides whether (and with which tail capacity) the cell
// kernel is used for this data set.
__global__ void __launch_bounds__(CTA_THREADS) cell_stats_kernel(SiteTable T, uint32_t n_samples, int ignore_miss, unsigned long long *out,
                                  unsigned int *hist);

struct CellVariant {
  int r, minb;  // cells per lane in registers, CTAs per SM the kernel was compiled for
  const void *fn, *fn_fused;
};
extern const CellVariant cell_variants[];
extern const int cell_variants_count;

}  // namespace emcell
