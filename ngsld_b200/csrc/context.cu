// C-ABI layer of libngsld_b200.so: context, uploads, scan planner and chunked scan driver.
// See include/ngsld_b200.h for the contract and the reference interfaces each entry point replaces.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "aux_kernels.cuh"
#include "em_cell.cuh"
#include "em_kernels.cuh"
#include "em_warp.cuh"
#include "format.cuh"
#include "host/host_prep.h"
#include "host/loader.h"
#include "host/prune.h"

namespace emfast {
#define DECL_LPG(n)                              \
  extern const EmVariant em_variants_lpg##n[];   \
  extern const int em_variants_lpg##n##_count;
DECL_LPG(4) DECL_LPG(8) DECL_LPG(16) DECL_LPG(32) DECL_LPG(64) DECL_LPG(128) DECL_LPG(256)
#undef DECL_LPG
}  // namespace emfast

namespace emwarp {
extern const WarpVariant warp_variants[];
extern const int warp_variants_count;
}  // namespace emwarp

static thread_local std::string g_create_error;

namespace {

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct ChunkBuf {
  uint32_t *d_s1 = nullptr, *d_s2 = nullptr;
  uint32_t *d_resid = nullptr;       // pairs of the chunk the cell kernel left to the dense kernel
  ngsld_edge *d_edges = nullptr;     // rows of the chunk that are graph edges (ngsld_scan_edges), + their number
  unsigned long long *d_n_edges = nullptr;
  ngsld_pair_row *d_rows = nullptr;
  ngsld_pair_row *h_rows = nullptr;  // pinned
  char *d_text = nullptr;            // formatted TSV bytes (slots, then compacted)
  char *d_text_out = nullptr;
  char *h_text = nullptr;            // pinned
  unsigned long long *d_line_off = nullptr;
  unsigned long long *h_text_len = nullptr;  // pinned: {bytes, overflow flag}
  cudaEvent_t ev_ready = nullptr, ev_em0 = nullptr, ev_em1 = nullptr, ev_p0 = nullptr, ev_p1 = nullptr,
              ev_f0 = nullptr, ev_f1 = nullptr, ev_done = nullptr;
  uint64_t n_rows = 0;
  bool pending = false;
};

struct Plan {
  std::vector<uint32_t> cs;                  // compact index -> site
  std::vector<uint32_t> cw_end;              // per compact first site: exclusive partner end (compact)
  std::vector<unsigned long long> row_off;   // nC + 1 prefix of rows per compact first site
  uint32_t n_compact = 0, c_lo = 0, c_hi = 0;
  bool identity = true, sampled = false;
  unsigned long long total = 0;
};

enum ScanMode { MODE_ROWS, MODE_TEXT, MODE_DEVICE };

}  // namespace

struct ngsld_ctx {
  int device = 0;
  int sm_count = 0;
  int smem_optin = 0;
  cudaStream_t own_main = nullptr, s_main = nullptr, s_aux = nullptr, s_copy = nullptr;
  std::string err;
  // sites
  uint64_t n_sites = 0, n_ind = 0, n_pad = 0;
  double *d_gl = nullptr, *d_maf = nullptr, *d_q = nullptr, *d_expg = nullptr;
  uint64_t *d_dx_sig = nullptr, *d_ratio = nullptr;
  uint16_t *d_dx_se = nullptr;
  std::vector<double> h_maf;
  // site palettes + what they say about the data set (em_cell.cuh)
  uint64_t n_cpad = 0;
  uint8_t *d_cls = nullptr, *d_pal_k = nullptr;
  double *d_pal = nullptr;
  uint64_t *d_pal_miss = nullptr;
  unsigned long long *d_cell_stats = nullptr;  // 3 counters + 129 histogram buckets (as 32-bit)
  bool cell_ok = false;        // the class-compressed EM pays for this data set
  bool cell_possible = false;  // palettes exist (NGSLD_EM_PATH=cell can force the kernel)
  double cell_mean = 0, cell_uncoded_frac = 0;
  uint32_t cell_p995 = 0;      // 99.5 % of the sampled pairs have at most this many cells
  uint32_t cell_kstride = NGSLD_KMAX;  // largest site palette, rounded up to 8: row length of the joint-class tables
  // positions / labels
  bool have_pos = false;
  std::vector<double> h_cum;
  std::vector<uint32_t> h_seg;
  double *d_cum = nullptr;
  uint32_t *d_seg = nullptr;
  bool have_labels = false;
  std::vector<std::string> h_labels;
  char *d_label_blob = nullptr;
  uint32_t *d_label_off = nullptr;  // n_sites + 1
  uint32_t max_label_len = 6;       // "(null)"
  size_t label_blob_bytes = 0;
  // plan buffers
  uint32_t *d_cs = nullptr, *d_cw_end = nullptr;
  unsigned long long *d_row_off = nullptr, *d_seeds = nullptr, *d_counts = nullptr;
  uint32_t *d_taus_jump = nullptr;  // jump-ahead tables of the sampling generator (hostprep::taus_jump_tables)
  uint2 *d_tiles = nullptr;
  size_t cap_compact = 0, cap_tiles = 0;
  DevCounters *d_ctr = nullptr;
  // LD pruning (ngsld_scan_edges): when active, every chunk's rows are filtered into an edge list on the device
  bool prune_active = false;
  ngsld_prune_params prune_q;
  unsigned char *d_seen = nullptr;
  ngsld_edge_sink edge_sink = nullptr;
  void *edge_user = nullptr;
  std::vector<ngsld_edge> h_edges;
  // LD-decay bins (ngsld_scan_decay): when active, every chunk is folded into them on the device
  bool decay_active = false;
  double decay_bin_size = 0;
  unsigned long long decay_n_bins = 0;
  ngsld_decay_bin *d_decay_bins = nullptr;
  unsigned long long *d_decay_outside = nullptr;
  // chunks
  uint64_t chunk_rows = 4ull << 20;
  uint64_t alloc_rows = 0;
  uint32_t alloc_slot = 0;  // bytes per TSV row slot the text buffers were sized with
  bool alloc_text = false, alloc_host = false;
  ChunkBuf buf[2];
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  ngsld_scan_stats stats;
  ngsld_ctx() { memset(&stats, 0, sizeof stats); }
};

namespace {

int fail(ngsld_ctx *c, int code, const std::string &msg) {
  if (c) c->err = msg;
  return code;
}

#define CUDA_TRY(ctx, expr)                                                                          \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return fail(ctx, e__ == cudaErrorMemoryAllocation ? NGSLD_E_NOMEM : NGSLD_E_CUDA,              \
                  std::string(#expr) + ": " + cudaGetErrorString(e__));                              \
  } while (0)

template <class T>
void dfree(T *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}

void free_chunks(ngsld_ctx *c) {
  for (auto &b : c->buf) {
    dfree(b.d_s1);
    dfree(b.d_s2);
    dfree(b.d_resid);
    dfree(b.d_edges);
    dfree(b.d_n_edges);
    dfree(b.d_rows);
    dfree(b.d_text);
    dfree(b.d_text_out);
    dfree(b.d_line_off);
    if (b.h_rows) cudaFreeHost(b.h_rows);
    if (b.h_text) cudaFreeHost(b.h_text);
    if (b.h_text_len) cudaFreeHost(b.h_text_len);
    b.h_rows = nullptr;
    b.h_text = nullptr;
    b.h_text_len = nullptr;
  }
  c->alloc_rows = 0;
  c->alloc_slot = 0;
  c->alloc_text = c->alloc_host = false;
}

// need_host: page-locked staging for binary rows; need_text: device text buffers; text_staging: page-locked staging for
// the text as well (the sink variants; ngsld_scan_tsv_into copies straight into the caller's buffer instead)
int ensure_chunks(ngsld_ctx *c, uint64_t rows, bool need_host, bool need_text, uint32_t slot, bool text_staging = true) {
  if (rows < 1) rows = 1;
  // the text buffers are rows * slot bytes: a longer slot (extend_out, longer labels) needs new ones
  if (c->alloc_rows >= rows && (!need_host || c->alloc_host) &&
      (!need_text || (c->alloc_text && c->alloc_slot >= slot && (!text_staging || c->buf[0].h_text))))
    return NGSLD_OK;
  free_chunks(c);
  // successive scans of one job (the slabs of the CLI) differ a little in size: leave headroom, or every slab that is a
  // few rows larger than all before would pay for a reallocation
  if (rows > (1u << 20)) rows = std::min<uint64_t>(rows + rows / 8, std::max<uint64_t>(rows, c->chunk_rows));
  for (auto &b : c->buf) {
    CUDA_TRY(c, cudaMalloc(&b.d_s1, rows * sizeof(uint32_t)));
    CUDA_TRY(c, cudaMalloc(&b.d_s2, rows * sizeof(uint32_t)));
    CUDA_TRY(c, cudaMalloc(&b.d_resid, rows * sizeof(uint32_t)));
    CUDA_TRY(c, cudaMalloc(&b.d_rows, rows * sizeof(ngsld_pair_row)));
    if (need_host && !need_text) CUDA_TRY(c, cudaMallocHost(&b.h_rows, rows * sizeof(ngsld_pair_row)));
    if (need_text) {
      CUDA_TRY(c, cudaMalloc(&b.d_text, rows * (size_t)slot));
      CUDA_TRY(c, cudaMalloc(&b.d_text_out, rows * (size_t)slot));
      CUDA_TRY(c, cudaMalloc(&b.d_line_off, (rows + rows / 1024 + 8) * sizeof(unsigned long long)));
      if (text_staging) CUDA_TRY(c, cudaMallocHost(&b.h_text, rows * (size_t)slot));
      CUDA_TRY(c, cudaMallocHost(&b.h_text_len, 2 * sizeof(unsigned long long)));
    }
  }
  c->alloc_rows = rows;
  c->alloc_host = need_host && !need_text;
  c->alloc_text = need_text;
  c->alloc_slot = need_text ? slot : 0;
  return NGSLD_OK;
}

// Device buffers of the site table for a given shape (kept when the shape is unchanged).  with_expg: also the
// expected-genotype matrix, which only ngsld_set_sites needs (input of the per-site x87 recurrence).
int alloc_site_buffers(ngsld_ctx *c, uint64_t n_sites, uint64_t n_ind, bool with_expg) {
  const uint64_t n_pad = (n_ind + 1) & ~1ull, n_cpad = (n_ind + 15) & ~15ull;
  // pearson::pair_r2 addresses the records of the x87 term tables with 32-bit indices (block * n_sites + site); a site table
  // of that many likelihood triples (> 400 GB) would not fit any GPU anyway
  if (((n_ind + 3) / 4 + 1) * n_sites > 0xffffffffull) return fail(c, NGSLD_E_INVALID, "site table too large (n_sites * n_ind)");
  const size_t row_bytes = n_pad * 24;
  if (n_sites != c->n_sites || n_ind != c->n_ind || !c->d_gl) {
    dfree(c->d_gl);
    dfree(c->d_maf);
    dfree(c->d_q);
    dfree(c->d_dx_sig);
    dfree(c->d_dx_se);
    dfree(c->d_seg);
    dfree(c->d_expg);
    dfree(c->d_ratio);
    dfree(c->d_cls);
    dfree(c->d_pal);
    dfree(c->d_pal_k);
    dfree(c->d_pal_miss);
    c->n_sites = c->n_ind = 0;
    CUDA_TRY(c, cudaMalloc(&c->d_gl, n_sites * row_bytes));
    CUDA_TRY(c, cudaMalloc(&c->d_maf, n_sites * sizeof(double)));
    CUDA_TRY(c, cudaMalloc(&c->d_q, n_sites * sizeof(double)));
    const uint64_t n_blk = (n_ind + 3) / 4;  // x87 terms: blocks of four individuals
    // (one spare block row behind the last: pearson::pair_r2 requests "the next block" without a guard)
    CUDA_TRY(c, cudaMalloc(&c->d_dx_sig, n_sites * (n_blk + 1) * 4 * sizeof(uint64_t)));
    CUDA_TRY(c, cudaMalloc(&c->d_dx_se, n_sites * (n_blk + 1) * 4 * sizeof(uint16_t)));
    CUDA_TRY(c, cudaMemset(c->d_dx_sig + n_sites * n_blk * 4, 0, n_sites * 4 * sizeof(uint64_t)));
    CUDA_TRY(c, cudaMemset(c->d_dx_se + n_sites * n_blk * 4, 0, n_sites * 4 * sizeof(uint16_t)));
    CUDA_TRY(c, cudaMalloc(&c->d_seg, n_sites * sizeof(uint32_t)));
    CUDA_TRY(c, cudaMalloc(&c->d_ratio, n_blk * 4 * sizeof(uint64_t)));
    if (n_ind < 65536) {  // joint-class counters are 16 bits wide
      CUDA_TRY(c, cudaMalloc(&c->d_cls, n_sites * n_cpad));
      CUDA_TRY(c, cudaMalloc(&c->d_pal, n_sites * (size_t)NGSLD_KMAX * 3 * sizeof(double)));
      CUDA_TRY(c, cudaMalloc(&c->d_pal_k, n_sites));
      CUDA_TRY(c, cudaMalloc(&c->d_pal_miss, n_sites * sizeof(uint64_t)));
    }
    if (!c->d_cell_stats) CUDA_TRY(c, cudaMalloc(&c->d_cell_stats, 3 * sizeof(unsigned long long) + 132 * sizeof(unsigned int)));
  }
  if (with_expg && !c->d_expg) CUDA_TRY(c, cudaMalloc(&c->d_expg, n_sites * n_ind * sizeof(double)));
  return NGSLD_OK;
}

SiteTable site_table(const ngsld_ctx *c) {
  SiteTable T;
  T.gl = c->d_gl;
  T.maf = c->d_maf;
  T.dx_sig = c->d_dx_sig;
  T.dx_se = c->d_dx_se;
  T.q = c->d_q;
  T.ratio = c->d_ratio;
  T.cum = c->have_pos ? c->d_cum : nullptr;
  T.seg = c->d_seg;
  T.cls = c->d_cls;
  T.pal = c->d_pal;
  T.pal_k = c->d_pal_k;
  T.pal_miss = c->d_pal_miss;
  T.n_cpad = (uint32_t)c->n_cpad;
  T.n_sites = (uint32_t)c->n_sites;
  T.n_ind = (uint32_t)c->n_ind;
  T.n_pad = (uint32_t)c->n_pad;
  T.n_blk = (uint32_t)((c->n_ind + 3) / 4);
  return T;
}

// (IPL, LPG) for a sample size: the smallest group whose lanes hold <= 8 individuals each.
const emfast::EmVariant *pick_variant(uint64_t n_ind) {
  using namespace emfast;
  struct Tab { const EmVariant *v; int n; } tabs[] = {
      {em_variants_lpg4, em_variants_lpg4_count},     {em_variants_lpg8, em_variants_lpg8_count},
      {em_variants_lpg16, em_variants_lpg16_count},   {em_variants_lpg32, em_variants_lpg32_count},
      {em_variants_lpg64, em_variants_lpg64_count},   {em_variants_lpg128, em_variants_lpg128_count},
      {em_variants_lpg256, em_variants_lpg256_count}};
  const char *force = getenv("NGSLD_EM_VARIANT");  // "ipl,lpg" for experiments
  int fi = 0, fl = 0;
  if (force && sscanf(force, "%d,%d", &fi, &fl) == 2) {
    for (auto &t : tabs)
      for (int k = 0; k < t.n; k++)
        if (t.v[k].ipl == fi && t.v[k].lpg == fl && (uint64_t)fi * fl >= n_ind) return &t.v[k];
  }
  for (auto &t : tabs) {
    const int lpg = t.v[0].lpg;
    const uint64_t ipl = (n_ind + lpg - 1) / lpg;
    if (ipl > 8) continue;
    for (int k = 0; k < t.n; k++)
      if ((uint64_t)t.v[k].ipl >= ipl) return &t.v[k];
  }
  return nullptr;
}

// ---- planning -----------------------------------------------------------------------------------
// Reproduces the control flow of calc_pair_LD's scan (reference ngsLD.cpp:240-282) as index ranges.
struct PlanInput {
  uint64_t n_sites;
  const double *maf;
  bool have_pos;
  const double *cum;    // exact prefix sums of the finite gaps (valid when have_pos)
  const uint32_t *seg;  // chromosome segment ids
};

PlanInput plan_input(const ngsld_ctx *c) {
  PlanInput in;
  in.n_sites = c->n_sites;
  in.maf = c->h_maf.data();
  in.have_pos = c->have_pos;
  in.cum = c->h_cum.data();
  in.seg = c->h_seg.data();
  return in;
}

int make_plan(const PlanInput &in, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params &P, Plan &pl) {
  const uint64_t n = in.n_sites;
  if (s1_hi > n) s1_hi = n;
  if (s1_lo > s1_hi) s1_lo = s1_hi;
  const double *maf = in.maf;
  // sites that may appear in a pair at all: !(maf < min_maf)  (ngsLD.cpp:264,270)
  std::vector<uint32_t> kp(n + 1);
  pl.cs.clear();
  pl.cs.reserve(n);
  pl.identity = true;
  for (uint64_t s = 0; s < n; s++) {
    kp[s] = (uint32_t)pl.cs.size();
    if (maf[s] < P.min_maf)
      pl.identity = false;
    else
      pl.cs.push_back((uint32_t)s);
  }
  kp[n] = (uint32_t)pl.cs.size();
  pl.n_compact = (uint32_t)pl.cs.size();
  pl.c_lo = kp[s1_lo];
  pl.c_hi = kp[s1_hi];
  pl.cw_end.assign(pl.n_compact, 0);
  // window end per first site (two-pointer; both break conditions are monotone in s1)
  const double kb_limit = (double)(P.max_kb_dist * 1000ull);
  uint64_t e = 0;
  for (uint64_t s1 = 0; s1 < n; s1++) {
    if (e < s1 + 1) e = s1 + 1;
    while (e < n) {
      double dist = INFINITY;
      if (in.have_pos && in.seg[s1] == in.seg[e]) dist = in.cum[e] - in.cum[s1];
      if (P.max_kb_dist > 0 && kb_limit < dist) break;            // ngsLD.cpp:252
      if (P.max_snp_dist > 0 && P.max_snp_dist < e - s1) break;   // ngsLD.cpp:258
      e++;
    }
    if (!(maf[s1] < P.min_maf)) pl.cw_end[kp[s1]] = kp[e];
  }
  pl.sampled = P.rnd_sample < 1.0;
  pl.row_off.assign((size_t)pl.n_compact + 1, 0);
  if (!pl.sampled) {
    unsigned long long acc = 0;
    for (uint32_t k = 0; k < pl.n_compact; k++) {
      pl.row_off[k] = acc;
      if (k >= pl.c_lo && k < pl.c_hi && pl.cw_end[k] > k + 1) acc += pl.cw_end[k] - k - 1;
    }
    pl.row_off[pl.n_compact] = acc;
    pl.total = acc;
  }
  return NGSLD_OK;
}

// A candidate pair is kept iff !(get() / 2^32 > rnd_sample) (reference ngsLD.cpp:277, gen_func.cpp:117-119), i.e. iff
// get() <= floor(rnd_sample * 2^32): both scalings by 2^32 are exact in double.  Only used for rnd_sample < 1.
uint32_t keep_max_of(double rnd_sample) {
  const double x = floor(rnd_sample * 4294967296.0);
  return x >= 4294967295.0 ? 4294967295u : (uint32_t)x;
}

// Exact prefix sums of the inter-site gaps: every finite gap must be a non-negative integer and the total below
// 2^53, so that cum[s2]-cum[s1] equals the reference's running double sum (ngsLD.cpp:241) bit for bit.  A +inf gap
// (chromosome change, read_data.cpp:209) starts a new segment.  Returns NULL or the reason for rejection.
const char *build_cum(const double *pos_dist, uint64_t n, double *cum, uint32_t *seg_out) {
  double acc = 0;
  uint32_t seg = 0;
  for (uint64_t s = 0; s < n; s++) {
    const double g = pos_dist[s];
    if (isinf(g) && g > 0) {
      if (s > 0) seg++;
    } else {
      if (!(g >= 0) || g != floor(g)) return "inter-site distances must be non-negative integers or +inf";
      if (s > 0) acc += g;  // pos_dist[0] (distance from the origin) never enters a pair distance
      if (acc > 9007199254740992.0) return "positions exceed 2^53";
    }
    cum[s] = acc;
    seg_out[s] = seg;
  }
  return nullptr;
}

// Host-side kept-pair counts of a sampled scan (reference ngsLD.cpp:277 with the per-site taus streams of
// ngsLD.cpp:165-166).  O(candidate pairs): used only by the device-free planning entry points; scans and
// ngsld_partition count on the device (aux::taus_sample_kernel).
void host_sample_counts(const Plan &pl, const ngsld_scan_params &P, uint64_t n_sites, std::vector<unsigned long long> &counts) {
  std::vector<uint64_t> seeds(n_sites);
  hostprep::site_seeds(P.seed, n_sites, seeds.data());
  counts.assign(pl.c_hi - pl.c_lo, 0);
  for (uint32_t c1 = pl.c_lo; c1 < pl.c_hi; c1++) {
    hostprep::TausStream g(seeds[pl.cs[c1]]);
    unsigned long long kept = 0;
    for (uint32_t c2 = c1 + 1; c2 < pl.cw_end[c1]; c2++)
      if (!(g.uniform() > P.rnd_sample)) kept++;
    counts[c1 - pl.c_lo] = kept;
  }
}

void finish_sampled_plan(Plan &pl, const std::vector<unsigned long long> &counts) {
  unsigned long long acc = 0;
  for (uint32_t k = 0; k < pl.n_compact; k++) {
    pl.row_off[k] = acc;
    if (k >= pl.c_lo && k < pl.c_hi) acc += counts[k - pl.c_lo];
  }
  pl.row_off[pl.n_compact] = acc;
  pl.total = acc;
}

// Equal-row-count first-site boundaries from a whole-range plan.
void bounds_from_plan(const Plan &pl, uint64_t n_sites, int n_parts, uint64_t *bounds) {
  bounds[0] = 0;
  for (int k = 1; k < n_parts; k++) {
    const unsigned long long target = (unsigned long long)((long double)pl.total * k / n_parts);
    // first compact site whose rows start at or after the target
    auto it = std::lower_bound(pl.row_off.begin(), pl.row_off.begin() + pl.n_compact, target);
    const size_t cidx = it - pl.row_off.begin();
    uint64_t s = cidx < pl.n_compact ? pl.cs[cidx] : n_sites;
    if (s < bounds[k - 1]) s = bounds[k - 1];
    bounds[k] = s;
  }
  bounds[n_parts] = n_sites;
}

int ensure_plan_buffers(ngsld_ctx *c, size_t n_compact) {
  if (c->cap_compact >= n_compact + 1) return NGSLD_OK;
  dfree(c->d_cs);
  dfree(c->d_cw_end);
  dfree(c->d_row_off);
  dfree(c->d_counts);
  const size_t cap = n_compact + 1;
  CUDA_TRY(c, cudaMalloc(&c->d_cs, cap * sizeof(uint32_t)));
  CUDA_TRY(c, cudaMalloc(&c->d_cw_end, cap * sizeof(uint32_t)));
  CUDA_TRY(c, cudaMalloc(&c->d_row_off, cap * sizeof(unsigned long long)));
  CUDA_TRY(c, cudaMalloc(&c->d_counts, cap * sizeof(unsigned long long)));
  c->cap_compact = cap;
  return NGSLD_OK;
}

int upload_plan(ngsld_ctx *c, Plan &pl, const ngsld_scan_params &P) {
  int rc = ensure_plan_buffers(c, pl.n_compact);
  if (rc) return rc;
  if (pl.n_compact) {
    CUDA_TRY(c, cudaMemcpyAsync(c->d_cs, pl.cs.data(), pl.n_compact * sizeof(uint32_t), cudaMemcpyHostToDevice, c->s_main));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_cw_end, pl.cw_end.data(), pl.n_compact * sizeof(uint32_t), cudaMemcpyHostToDevice, c->s_main));
  }
  c->stats.h2d_bytes += (uint64_t)pl.n_compact * 8;
  if (pl.sampled) {
    // per-site generator seeds from the master stream (host, serial as in the reference), then the
    // per-first-site kept-pair counts on the device
    std::vector<uint64_t> seeds(c->n_sites);
    hostprep::site_seeds(P.seed, c->n_sites, seeds.data());
    dfree(c->d_seeds);
    CUDA_TRY(c, cudaMalloc(&c->d_seeds, c->n_sites * sizeof(unsigned long long)));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_seeds, seeds.data(), c->n_sites * 8, cudaMemcpyHostToDevice, c->s_main));
    c->stats.h2d_bytes += c->n_sites * 8;
    if (!c->d_taus_jump) {
      std::vector<uint32_t> jump((size_t)hostprep::TAUS_JUMP_LEVELS * 3 * 32);
      hostprep::taus_jump_tables(jump.data());
      CUDA_TRY(c, cudaMalloc(&c->d_taus_jump, jump.size() * sizeof(uint32_t)));
      CUDA_TRY(c, cudaMemcpy(c->d_taus_jump, jump.data(), jump.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    const uint32_t span = pl.c_hi - pl.c_lo;
    std::vector<unsigned long long> counts(span);
    if (span) {
      const unsigned blocks = (unsigned)std::min<uint64_t>(span, (uint64_t)c->sm_count * 32);
      aux::taus_sample_kernel<<<blocks, 256, 0, c->s_main>>>(c->d_seeds, pl.identity ? nullptr : c->d_cs, c->d_cw_end, pl.c_lo,
                                                             pl.c_hi, keep_max_of(P.rnd_sample), 0, c->d_counts, nullptr, 0, 0,
                                                             nullptr, nullptr, c->d_taus_jump);
      c->stats.n_launches++;
      CUDA_TRY(c, cudaGetLastError());
      CUDA_TRY(c, cudaMemcpyAsync(counts.data(), c->d_counts, span * 8ull, cudaMemcpyDeviceToHost, c->s_main));
      CUDA_TRY(c, cudaStreamSynchronize(c->s_main));
      c->stats.d2h_bytes += span * 8ull;
    }
    finish_sampled_plan(pl, counts);
  }
  CUDA_TRY(c, cudaMemcpyAsync(c->d_row_off, pl.row_off.data(), ((size_t)pl.n_compact + 1) * 8, cudaMemcpyHostToDevice, c->s_main));
  c->stats.h2d_bytes += ((uint64_t)pl.n_compact + 1) * 8;
  // the host vectors must outlive the async copies
  CUDA_TRY(c, cudaStreamSynchronize(c->s_main));
  return NGSLD_OK;
}

// first compact site owning global row g
uint32_t row_owner(const Plan &pl, unsigned long long g) {
  auto it = std::upper_bound(pl.row_off.begin(), pl.row_off.begin() + pl.n_compact, g);
  return (uint32_t)(it - pl.row_off.begin()) - 1;
}

struct EmChoice {
  const emfast::EmVariant *v = nullptr;
  const emwarp::WarpVariant *w = nullptr;  // warp-per-pair kernel (preferred where it applies)
  size_t warp_smem = 0;
  int blocks_warp = 0;
  bool tile = false;
  uint32_t TA = 0, TB = 0;
  size_t dyn_smem = 0;
  int blocks_list = 0, blocks_tile = 0;
  // class-compressed kernel (em_cell.cuh); the dense warp kernel `w` then takes the pairs it leaves over
  const emcell::CellVariant *cell = nullptr;
  uint32_t cell_tcap = 0;
  size_t cell_smem = 0;
  int blocks_cell = 0;
  bool cell_fuse = true;
};

// Warp-per-pair kernel: R registers-resident individuals per lane, the rest of the rows in a warp-private
// shared-memory slice.  Used when at least 3 CTAs (12 warps) fit per SM.
int choose_warp(ngsld_ctx *c, EmChoice &ch, int min_occ, bool forced = false) {
  const char *path = getenv("NGSLD_EM_PATH");  // "cell" | "warp" | "list" | "tile" for experiments
  if (forced) path = "warp";
  if (path && strcmp(path, "warp") != 0) return NGSLD_OK;
  const uint64_t min_ind = path ? 1 : 160;  // below: the sub-warp group kernels waste fewer lanes
  if (c->n_ind < min_ind) return NGSLD_OK;
  // smallest group of warps per pair whose per-CTA shared memory lets three CTAs share an SM
  const char *force = getenv("NGSLD_WARP_R"), *force_g = getenv("NGSLD_WARP_G");
  const emwarp::WarpVariant *w = nullptr;
  size_t smem = 0;
  int occ = 0;
  for (int g : {1, 2, 4}) {
    if (force_g && atoi(force_g) != g) continue;
    const uint64_t per_warp = (c->n_ind + g - 1) / g;
    int r = (int)std::min<uint64_t>(6, (per_warp + 31) / 32);
    if (force && atoi(force) >= 1 && atoi(force) <= 8) r = atoi(force);
    const emwarp::WarpVariant *cand = nullptr;
    for (int k = 0; k < emwarp::warp_variants_count; k++)
      if (emwarp::warp_variants[k].r == r && emwarp::warp_variants[k].g == g) cand = &emwarp::warp_variants[k];
    if (!cand) continue;
    size_t sm = (size_t)emwarp::WARPS_PER_CTA * 2 * emwarp::WarpGeom::tail_slots((uint32_t)c->n_pad, g, r) * 24;
    if (const char *pad = getenv("NGSLD_WARP_PAD_SMEM")) sm += (size_t)atoi(pad);  // occupancy experiments
    if (sm + 2048 > (size_t)c->smem_optin) continue;
    int o = 0;
    for (const void *fn : {cand->fn, cand->fn_ign, cand->fn_u1}) {
      CUDA_TRY(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fn, emwarp::CTA_THREADS, sm));
    }
    if (o > occ) {
      w = cand;
      smem = sm;
      occ = o;
    }
    if (occ >= 3) break;
  }
  if (!w || (occ < min_occ && !path)) return NGSLD_OK;
  ch.w = w;
  ch.warp_smem = smem;
  ch.blocks_warp = std::max(1, occ) * c->sm_count;
  return NGSLD_OK;
}

// sel != NULL: only the pairs sel[0 .. d_ctr->n_resid) (the cell kernel's left-overs, counted on the device)
int launch_warp(ngsld_ctx *c, const EmChoice &ch, const SiteTable &T, const PairChunk &C, int ignore_miss,
                const uint32_t *sel = nullptr) {
  SiteTable Tt = T;
  PairChunk Cc = C;
  DevCounters *ctr = c->d_ctr;
  const unsigned long long *n_sel = sel ? &c->d_ctr->n_resid : nullptr;
  void *args[] = {&Tt, &Cc, &ctr, &sel, &n_sel};
  const unsigned long long want = (C.n_pairs + emwarp::WARPS_PER_CTA - 1) / emwarp::WARPS_PER_CTA;
  // (the number of left-overs is only known on the device: CTAs that find nothing to do exit at once)
  const unsigned blocks = (unsigned)std::min<unsigned long long>(want, ch.blocks_warp);
  const char *u1 = getenv("NGSLD_WARP_U1");  // experiments: unfused tail loop
  const void *fn = ignore_miss ? ch.w->fn_ign : (u1 && atoi(u1) ? ch.w->fn_u1 : ch.w->fn);
  CUDA_TRY(c, cudaLaunchKernel(fn, dim3(blocks), dim3(emwarp::CTA_THREADS), args,
                               ch.warp_smem, c->s_main));
  return NGSLD_OK;
}

// The class-compressed kernel on a chunk, then the dense kernel on whatever it left over (same stream).
// row_len: pairs per first site in this chunk (0 = unknown / irregular): shapes the work order, see CellArgs::swz_len
int launch_cell(ngsld_ctx *c, const EmChoice &ch, const SiteTable &T, const PairChunk &C, int ignore_miss, uint32_t *d_resid,
                uint64_t row_len = 0) {
  SiteTable Tt = T;
  PairChunk Cc = C;
  DevCounters *ctr = c->d_ctr;
  emcell::CellArgs A;
  A.resid = d_resid;
  A.tcap = ch.cell_tcap;
  A.ignore_miss = ignore_miss;
  A.fuse_pearson = ch.cell_fuse ? 1 : 0;
  A.kstride = c->cell_kstride;
  A.swz_len = A.swz_rows = 0;
  static const bool no_swizzle = getenv("NGSLD_CELL_LINEAR") && atoi(getenv("NGSLD_CELL_LINEAR"));
  if (!no_swizzle && row_len >= 64 && row_len < C.n_pairs && row_len < (1ull << 31)) {
    A.swz_len = (uint32_t)row_len;
    A.swz_rows = (uint32_t)((C.n_pairs + row_len - 1) / row_len);
  }
  void *args[] = {&Tt, &Cc, &A, &ctr};
  const unsigned long long want = (C.n_pairs + 32ull * emcell::WARPS_PER_CTA - 1) / (32ull * emcell::WARPS_PER_CTA);
  const unsigned blocks = (unsigned)std::min<unsigned long long>(want, ch.blocks_cell);
  CUDA_TRY(c, cudaLaunchKernel(ch.cell_fuse ? ch.cell->fn_fused : ch.cell->fn, dim3(blocks), dim3(emcell::CTA_THREADS), args,
                               ch.cell_smem, c->s_main));
  c->stats.n_launches++;
  return launch_warp(c, ch, T, C, ignore_miss, d_resid);
}

// Class-compressed kernel: used when the palettes say a pair has clearly fewer distinct (p, q) combinations than
// individuals (ngsld_set_sites samples that), or when NGSLD_EM_PATH=cell forces it.
int choose_cell(ngsld_ctx *c, EmChoice &ch) {
  const char *path = getenv("NGSLD_EM_PATH");
  if (path && strcmp(path, "cell") != 0) return NGSLD_OK;
  if (!(path ? c->cell_possible : c->cell_ok)) return NGSLD_OK;
  // cells per lane in registers from the sampled 99.5th percentile; the rest of a pair's cells go to shared memory
  int r = c->cell_p995 <= 64 ? 2 : c->cell_p995 <= 128 ? 4 : 6;
  // CTAs per SM the variant is compiled for: six register cells per lane need 128 registers (4 CTAs, 16 warps per SM:
  // measured +14 % over 3 CTAs with 168 registers); with four or two cells per lane 96 registers do, and the fifth CTA
  // pays (100 individuals, 65 cells per pair: 110 against 102 M pairs/s, round 2)
  int minb = r <= 4 ? 5 : 4;
  if (const char *e = getenv("NGSLD_CELL_R")) r = atoi(e);
  if (const char *e = getenv("NGSLD_CELL_MINB")) minb = atoi(e);
  const emcell::CellVariant *v = nullptr;
  for (int k = 0; k < emcell::cell_variants_count; k++)
    if (emcell::cell_variants[k].r == r && (!v || emcell::cell_variants[k].minb == minb)) v = &emcell::cell_variants[k];
  if (!v) v = &emcell::cell_variants[emcell::cell_variants_count - 1];
  r = v->r;
  uint32_t tcap = c->cell_p995 > 32u * r ? ((c->cell_p995 - 32u * r + 63u) & ~63u) : 0u;
  tcap = std::max<uint32_t>(tcap, 64);  // room for the odd pair beyond the sampled percentile
  if (const char *e = getenv("NGSLD_CELL_TCAP")) tcap = ((uint32_t)atoi(e) + 63u) & ~63u;
  // at least two CTAs per SM: shrink the tail if it does not fit (pairs beyond it go to the dense kernel)
  while (tcap > 0 && 2 * (emcell::WARPS_PER_CTA * emcell::warp_smem_bytes(r, tcap, c->cell_kstride) + 1024) > (size_t)c->smem_optin) tcap -= 64;
  const size_t smem = emcell::WARPS_PER_CTA * emcell::warp_smem_bytes(r, tcap, c->cell_kstride);
  if (smem + 1024 > (size_t)c->smem_optin) return NGSLD_OK;
  int occ = 0;
  for (const void *fn : {v->fn, v->fn_fused}) {
    CUDA_TRY(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, emcell::CTA_THREADS, smem));
  }
  if (occ < 1) return NGSLD_OK;
  ch.cell = v;
  ch.cell_tcap = tcap;
  ch.cell_smem = smem;
  ch.blocks_cell = occ * c->sm_count;
  ch.cell_fuse = true;
  if (const char *e = getenv("NGSLD_CELL_FUSE")) ch.cell_fuse = atoi(e) != 0;
  // the dense kernel for the left-over pairs, whatever the sample size
  return choose_warp(c, ch, 1, true);
}

int choose_em(ngsld_ctx *c, const Plan &pl, EmChoice &ch) {
  int rcw = choose_cell(c, ch);
  if (rcw) return rcw;
  if (ch.cell && ch.w) return NGSLD_OK;
  ch.cell = nullptr;
  rcw = choose_warp(c, ch, 3);
  if (rcw) return rcw;
  ch.v = pick_variant(c->n_ind);
  if (!ch.v) {
    // samples too large for three resident CTAs of the warp kernel and for the group kernels: take the warp kernel
    // at whatever occupancy its shared-memory slots allow (n_ind up to ~5400); beyond that the strict kernel runs
    if (!ch.w) rcw = choose_warp(c, ch, 1);
    return rcw;
  }
  int occ = 0;
  CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ch.v->list_fn, std::max(emfast::CTA_THREADS, ch.v->lpg), 0));
  ch.blocks_list = std::max(1, occ) * c->sm_count;
  // Site-tile staging (TMA) is optional for the group kernels: the rows of a pair are read once per pair and L2
  // serves them at a small fraction of its bandwidth, so the default is the plain pair list, which leaves the shared
  // memory free and lets 3-4 CTAs share an SM.  NGSLD_EM_PATH=tile turns the tiles on (a third of the SM's shared
  // memory per CTA).
  const size_t row_bytes = c->n_pad * 24;
  const size_t budget = ((size_t)c->smem_optin - 4096) / 3;
  uint32_t rows_fit = (uint32_t)std::min<size_t>(budget / row_bytes, 64);
  uint32_t t = rows_fit / 2;
  if (t > 32) t = 32;
  const char *path = getenv("NGSLD_EM_PATH");  // "list" | "tile" for experiments
  const char *tdim = getenv("NGSLD_TILE");
  if (tdim && atoi(tdim) > 0 && (uint32_t)atoi(tdim) <= t) t = atoi(tdim);
  ch.tile = false;
  if (path && !strcmp(path, "tile") && !pl.sampled && t >= 1) ch.tile = true;
  if (ch.tile) {
    ch.TA = ch.TB = t;
    ch.dyn_smem = (size_t)(ch.TA + ch.TB) * row_bytes;
    CUDA_TRY(c, cudaFuncSetAttribute(ch.v->tile_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch.dyn_smem));
    CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ch.v->tile_fn, std::max(emfast::CTA_THREADS, ch.v->lpg), ch.dyn_smem));
    ch.blocks_tile = std::max(1, occ) * c->sm_count;
  }
  return NGSLD_OK;
}

// Tiles covering the pairs of compact first sites [ca, cb): A-blocks of TA first sites from ca, each
// crossed with TB-wide partner blocks up to the block's furthest window end.  block_off[k] = index of
// the first tile of A-block k (block_off.back() = tiles.size()).
void build_tiles(const Plan &pl, uint32_t ca, uint32_t cb, uint32_t TA, uint32_t TB, std::vector<uint2> &tiles,
                 std::vector<size_t> &block_off) {
  tiles.clear();
  block_off.clear();
  for (uint32_t a = ca; a < cb; a += TA) {
    block_off.push_back(tiles.size());
    const uint32_t a_last = std::min(cb, a + TA) - 1;
    const uint32_t bmax = pl.cw_end[a_last];  // cw_end is non-decreasing over first sites
    for (uint32_t b = a + 1; b < bmax; b += TB) tiles.push_back(make_uint2(a, b));
  }
  block_off.push_back(tiles.size());
}

// rows_dst (MODE_ROWS only, may be NULL): the chunk's rows are copied from the device straight to rows_dst + r0 (the
// caller's buffer for the whole scan) instead of the context's page-locked staging chunk.
int launch_chunk(ngsld_ctx *c, const Plan &pl, const ngsld_scan_params &P, const EmChoice &ch, ChunkBuf &b,
                 unsigned long long r0, unsigned long long r1, size_t t0, size_t t1, ScanMode mode,
                 const fmt::FormatArgs *fa, ngsld_pair_row *rows_dst = nullptr) {
  const unsigned long long n = r1 - r0;
  const SiteTable T = site_table(c);
  PairChunk C;
  C.s1 = b.d_s1;
  C.s2 = b.d_s2;
  C.rows = b.d_rows;
  C.n_pairs = n;
  const int threads = 256;
  const unsigned gblocks = (unsigned)std::min<unsigned long long>((n + threads - 1) / threads, (unsigned long long)c->sm_count * 32);
  const uint32_t ca = row_owner(pl, r0), cb = row_owner(pl, r1 - 1) + 1;
  if (pl.sampled) {
    const uint32_t span = cb - ca;
    const unsigned sb = (unsigned)std::min<uint64_t>(span, (uint64_t)c->sm_count * 32);
    aux::taus_sample_kernel<<<sb, 256, 0, c->s_main>>>(c->d_seeds, pl.identity ? nullptr : c->d_cs, c->d_cw_end, ca, cb,
                                                       keep_max_of(P.rnd_sample), 1, nullptr, c->d_row_off, r0, n, b.d_s1, b.d_s2,
                                                       c->d_taus_jump);
  } else {
    aux::expand_window_kernel<<<gblocks, threads, 0, c->s_main>>>(c->d_row_off, pl.identity ? nullptr : c->d_cs,
                                                                   pl.n_compact, r0, n, b.d_s1, b.d_s2);
  }
  aux::fill_rows_kernel<<<gblocks, threads, 0, c->s_main>>>(T, C);
  c->stats.n_launches += 2;
  CUDA_TRY(c, cudaMemsetAsync(c->d_ctr, 0, NGSLD_WORK_COUNTERS * sizeof(unsigned long long), c->s_main));
  CUDA_TRY(c, cudaEventRecord(b.ev_ready, c->s_main));
  const bool use_cell = ch.cell != nullptr && ch.w != nullptr && !P.strict;
  const bool fused = use_cell && ch.cell_fuse;
  // r2_ExpG on the auxiliary stream, launched BEFORE the EM so that its CTAs are resident beside the EM's
  // (not needed when the cell kernel computes it itself)
  CUDA_TRY(c, cudaStreamWaitEvent(c->s_aux, b.ev_ready, 0));
  CUDA_TRY(c, cudaEventRecord(b.ev_p0, c->s_aux));
  if (!fused) {
    const bool beside_em = ch.w != nullptr && !P.strict;
    // four warps per SM: with fewer, the r2_ExpG kernel (which only gets the issue slots the EM leaves) becomes the
    // critical path (measured: 96 threads -> 22 M pairs/s, 64 -> 15.5 M, against 27 M)
    int pth = 128, pctas = 1;
    if (const char *e = getenv("NGSLD_PEARSON_THREADS")) pth = std::max(32, std::min(128, atoi(e) / 32 * 32));
    if (const char *e = getenv("NGSLD_PEARSON_CTAS")) pctas = std::max(1, std::min(16, atoi(e)));
    const unsigned long long want = (n + pth - 1) / pth;
    const unsigned pb = (unsigned)std::min<unsigned long long>(want, (unsigned long long)c->sm_count * (beside_em ? pctas : 16));
    aux::pearson_kernel<<<pb, pth, 0, c->s_aux>>>(T, C, c->d_ctr);
    c->stats.n_launches++;
  }
  CUDA_TRY(c, cudaEventRecord(b.ev_p1, c->s_aux));
  // EM
  if (P.strict || (ch.v == nullptr && ch.w == nullptr))
    snprintf(c->stats.em_kernel, sizeof c->stats.em_kernel, "aux::em_strict_kernel");
  else if (use_cell)
    snprintf(c->stats.em_kernel, sizeof c->stats.em_kernel, "emcell::em_cell_kernel<R=%d,FUSE=%d,MINB=%d>", ch.cell->r, fused ? 1 : 0,
             ch.cell->minb);
  else if (ch.w)
    snprintf(c->stats.em_kernel, sizeof c->stats.em_kernel, "emwarp::em_warp_kernel<R=%d,G=%d>", ch.w->r, ch.w->g);
  else
    snprintf(c->stats.em_kernel, sizeof c->stats.em_kernel, "emfast::em_%s_kernel<IPL=%d,LPG=%d>", ch.tile ? "tile" : "list",
             ch.v->ipl, ch.v->lpg);
  CUDA_TRY(c, cudaEventRecord(b.ev_em0, c->s_main));
  if (P.strict || (ch.v == nullptr && ch.w == nullptr)) {
    const unsigned sb = (unsigned)std::min<unsigned long long>((n + 127) / 128, (unsigned long long)c->sm_count * 64);
    aux::em_strict_kernel<<<sb, 128, 0, c->s_main>>>(T, C, P.ignore_miss_data, c->d_ctr);
  } else if (use_cell) {
    // pairs per first site around here: the window of the chunk's first site (windows shrink or shift by about one pair
    // per site, which the work order tolerates); sampled scans keep the listed order
    uint64_t row_len = 0;
    if (!pl.sampled && pl.cw_end[ca] > ca + 1) row_len = pl.cw_end[ca] - ca - 1;
    int rcc = launch_cell(c, ch, T, C, P.ignore_miss_data, b.d_resid, row_len);
    if (rcc) return rcc;
  } else if (ch.w) {
    int rcw = launch_warp(c, ch, T, C, P.ignore_miss_data);
    if (rcw) return rcw;
  } else if (ch.tile) {
    emfast::TileArgs A;
    A.tiles = c->d_tiles + t0;
    A.n_tiles = t1 - t0;
    A.cs = pl.identity ? nullptr : c->d_cs;
    A.cw_end = c->d_cw_end;
    A.row_off = c->d_row_off;
    A.row_lo = r0;
    A.row_hi = r1;
    A.n_compact = pl.n_compact;
    A.TA = ch.TA;
    A.TB = ch.TB;
    int ign = P.ignore_miss_data;
    ngsld_pair_row *rows = b.d_rows;
    SiteTable Tt = T;
    DevCounters *ctr = c->d_ctr;
    void *args[] = {&Tt, &rows, &A, &ign, &ctr};
    const unsigned blocks = (unsigned)std::min<unsigned long long>(std::max<size_t>(t1 - t0, 1), ch.blocks_tile);
    CUDA_TRY(c, cudaLaunchKernel(ch.v->tile_fn, dim3(blocks), dim3(std::max(emfast::CTA_THREADS, ch.v->lpg)), args, ch.dyn_smem, c->s_main));
  } else {
    int ign = P.ignore_miss_data;
    SiteTable Tt = T;
    PairChunk Cc = C;
    DevCounters *ctr = c->d_ctr;
    void *args[] = {&Tt, &Cc, &ign, &ctr};
    const unsigned long long groups_per_cta = std::max(emfast::CTA_THREADS, ch.v->lpg) / ch.v->lpg;
    const unsigned blocks = (unsigned)std::min<unsigned long long>((n + groups_per_cta - 1) / groups_per_cta, ch.blocks_list);
    CUDA_TRY(c, cudaLaunchKernel(ch.v->list_fn, dim3(blocks), dim3(std::max(emfast::CTA_THREADS, ch.v->lpg)), args, 0, c->s_main));
  }
  c->stats.n_launches++;
  CUDA_TRY(c, cudaEventRecord(b.ev_em1, c->s_main));
  CUDA_TRY(c, cudaStreamWaitEvent(c->s_main, b.ev_p1, 0));
  if (c->decay_active) {
    const unsigned db = (unsigned)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)c->sm_count * 8);
    aux::decay_bins_kernel<<<db, 256, 0, c->s_main>>>(b.d_rows, n, c->decay_bin_size, c->decay_n_bins, c->d_decay_bins,
                                                     c->d_decay_outside);
    c->stats.n_launches++;
  }
  if (c->prune_active) {
    if (!b.d_edges) {
      CUDA_TRY(c, cudaMalloc(&b.d_edges, c->alloc_rows * sizeof(ngsld_edge)));
      CUDA_TRY(c, cudaMalloc(&b.d_n_edges, sizeof(unsigned long long)));
    }
    CUDA_TRY(c, cudaMemsetAsync(b.d_n_edges, 0, sizeof(unsigned long long), c->s_main));
    const unsigned pb = (unsigned)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)c->sm_count * 8);
    double prec = 1;
    for (int k = 0; k < c->prune_q.weight_precision; k++) prec *= 10;
    aux::prune_edges_kernel<<<pb, 256, 0, c->s_main>>>(b.d_rows, n, c->prune_q, prec, b.d_edges, b.d_n_edges, c->d_seen);
    c->stats.n_launches++;
  }
  if (mode == MODE_TEXT) {
    CUDA_TRY(c, cudaEventRecord(b.ev_f0, c->s_main));
    int rc = fmt::launch_format(*fa, T, b.d_rows, n, b.d_text, b.d_line_off, b.d_text_out, c->sm_count, c->s_main);
    if (rc < 0) return fail(c, NGSLD_E_CUDA, "TSV formatter launch failed");
    c->stats.n_launches += rc;
    CUDA_TRY(c, cudaEventRecord(b.ev_f1, c->s_main));
    // total byte count first, then the text itself (size known only after the first copy)
    CUDA_TRY(c, cudaMemcpyAsync(b.h_text_len, b.d_line_off + n, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->s_main));
  } else if (mode == MODE_ROWS) {
    CUDA_TRY(c, cudaEventRecord(b.ev_f0, c->s_main));  // join point
    CUDA_TRY(c, cudaStreamWaitEvent(c->s_copy, b.ev_f0, 0));
    CUDA_TRY(c, cudaMemcpyAsync(rows_dst ? rows_dst + r0 : b.h_rows, b.d_rows, n * sizeof(ngsld_pair_row), cudaMemcpyDeviceToHost,
                                c->s_copy));
    CUDA_TRY(c, cudaEventRecord(b.ev_done, c->s_copy));
    c->stats.d2h_bytes += n * sizeof(ngsld_pair_row);
  }
  if (mode != MODE_ROWS) CUDA_TRY(c, cudaEventRecord(b.ev_done, c->s_main));
  CUDA_TRY(c, cudaGetLastError());
  b.n_rows = n;
  b.pending = true;
  return NGSLD_OK;
}

struct Delivery {
  ScanMode mode;
  int extend_out = 0;
  ngsld_row_sink rows;
  ngsld_text_sink text;
  void *user;
  // MODE_TEXT without a sink: the text goes straight from the device into one caller buffer (ngsld_scan_tsv_into)
  char *text_dst = nullptr;
  uint64_t text_cap = 0;
  uint64_t *text_len = nullptr;
  // MODE_ROWS without a sink: rows go straight from the device into one caller buffer (ngsld_scan_into)
  ngsld_pair_row *rows_dst = nullptr;
};

int deliver_chunk(ngsld_ctx *c, ChunkBuf &b, const Delivery &d) {
  CUDA_TRY(c, cudaEventSynchronize(b.ev_done));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, b.ev_em0, b.ev_em1) == cudaSuccess) c->stats.ms_em += ms;
  if (cudaEventElapsedTime(&ms, b.ev_p0, b.ev_p1) == cudaSuccess) c->stats.ms_pearson += ms;
  b.pending = false;
  c->stats.n_pairs += b.n_rows;
  if (d.mode == MODE_TEXT) {
    if (cudaEventElapsedTime(&ms, b.ev_f0, b.ev_f1) == cudaSuccess) c->stats.ms_format += ms;
    unsigned long long bytes = b.h_text_len[0];
    static const bool force_host = getenv("NGSLD_FORCE_HOST_FORMAT") && atoi(getenv("NGSLD_FORCE_HOST_FORMAT"));  // tests
    if (b.h_text_len[1] == 0 && !force_host) {
      char *dst = b.h_text;
      if (d.text_dst) {  // no staging: device -> the caller's (page-locked) buffer at its running offset
        if (*d.text_len + bytes > d.text_cap) return fail(c, NGSLD_E_INVALID, "text buffer too small for this scan");
        dst = d.text_dst + *d.text_len;
        *d.text_len += bytes;
      }
      CUDA_TRY(c, cudaMemcpyAsync(dst, b.d_text_out, bytes, cudaMemcpyDeviceToHost, c->s_copy));
      CUDA_TRY(c, cudaStreamSynchronize(c->s_copy));
      c->stats.d2h_bytes += bytes + 16;
      if (d.text_dst) return NGSLD_OK;
    } else {
      // some value was outside the device formatter's range: re-format this chunk with the host printf
      std::vector<ngsld_pair_row> host_rows(b.n_rows);  // rare path: pageable is fine
      CUDA_TRY(c, cudaMemcpyAsync(host_rows.data(), b.d_rows, b.n_rows * sizeof(ngsld_pair_row), cudaMemcpyDeviceToHost, c->s_copy));
      CUDA_TRY(c, cudaStreamSynchronize(c->s_copy));
      c->stats.d2h_bytes += b.n_rows * sizeof(ngsld_pair_row);
      std::string txt;
      char line[2048];
      for (uint64_t k = 0; k < b.n_rows; k++) {
        const ngsld_pair_row &r = host_rows[k];
        const char *l1 = c->have_labels ? c->h_labels[r.s1].c_str() : "(null)";
        const char *l2 = c->have_labels ? c->h_labels[r.s2].c_str() : "(null)";
        const int m = fmt::format_row_host(r, l1, l2, c->h_maf[r.s1], c->h_maf[r.s2], d.extend_out, line, sizeof line);
        if (m < 0) return fail(c, NGSLD_E_INVALID, "row too long for the host formatter");
        txt.append(line, m);
      }
      if (d.text_dst) {
        if (*d.text_len + txt.size() > d.text_cap) return fail(c, NGSLD_E_INVALID, "text buffer too small for this scan");
        memcpy(d.text_dst + *d.text_len, txt.data(), txt.size());
        *d.text_len += txt.size();
        return NGSLD_OK;
      }
      if (d.text && d.text(d.user, txt.data(), txt.size(), b.n_rows) != 0) return fail(c, NGSLD_E_SINK, "text sink aborted the scan");
      return NGSLD_OK;
    }
    if (d.text && d.text(d.user, b.h_text, bytes, b.n_rows) != 0) return fail(c, NGSLD_E_SINK, "text sink aborted the scan");
  } else if (d.mode == MODE_ROWS) {
    if (d.rows && d.rows(d.user, b.h_rows, b.n_rows) != 0) return fail(c, NGSLD_E_SINK, "row sink aborted the scan");
  }
  if (c->prune_active) {
    unsigned long long ne = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&ne, b.d_n_edges, sizeof ne, cudaMemcpyDeviceToHost, c->s_copy));
    CUDA_TRY(c, cudaStreamSynchronize(c->s_copy));
    c->h_edges.resize(ne);
    if (ne) {
      CUDA_TRY(c, cudaMemcpyAsync(c->h_edges.data(), b.d_edges, ne * sizeof(ngsld_edge), cudaMemcpyDeviceToHost, c->s_copy));
      CUDA_TRY(c, cudaStreamSynchronize(c->s_copy));
      // the lanes of different warps append concurrently: restore (s1, s2) order inside the chunk
      std::sort(c->h_edges.begin(), c->h_edges.end(),
                [](const ngsld_edge &x, const ngsld_edge &y) { return x.s1 != y.s1 ? x.s1 < y.s1 : x.s2 < y.s2; });
      c->stats.d2h_bytes += ne * sizeof(ngsld_edge) + 8;
      if (c->edge_sink && c->edge_sink(c->edge_user, c->h_edges.data(), ne) != 0) return fail(c, NGSLD_E_SINK, "edge sink aborted the scan");
    }
  }
  return NGSLD_OK;
}

int run_scan(ngsld_ctx *c, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *Pp, const Delivery &d) {
  if (!c) return NGSLD_E_INVALID;
  if (!Pp) return fail(c, NGSLD_E_INVALID, "scan parameters missing");
  if (!c->d_gl) return fail(c, NGSLD_E_INVALID, "ngsld_set_sites must be called before a scan");
  const ngsld_scan_params P = *Pp;
  if (!(P.rnd_sample > 0) || P.rnd_sample > 1) return fail(c, NGSLD_E_INVALID, "proportion of comparisons to sample must be in ]0,1]!");
  if (P.min_maf < 0 || P.min_maf > 1) return fail(c, NGSLD_E_INVALID, "minimum allele frequency must be in [0,1]!");
  CUDA_TRY(c, cudaSetDevice(c->device));
  memset(&c->stats, 0, sizeof c->stats);
  for (auto &b : c->buf) {  // a previous scan that failed midway may have left chunks in flight
    if (b.pending) cudaStreamSynchronize(c->s_main), cudaStreamSynchronize(c->s_aux), cudaStreamSynchronize(c->s_copy);
    b.pending = false;
  }
  const double t_plan0 = now_ms();
  static const bool dbg = getenv("NGSLD_DEBUG_PLAN") && atoi(getenv("NGSLD_DEBUG_PLAN"));
  Plan pl;
  int rc = make_plan(plan_input(c), s1_lo, s1_hi, P, pl);
  if (rc) return rc;
  const double t_plan1 = now_ms();
  rc = upload_plan(c, pl, P);
  if (rc) return rc;
  c->stats.ms_plan = now_ms() - t_plan0;
  const double t_plan2 = now_ms();
  if (pl.total == 0) return NGSLD_OK;
  if (d.rows_dst && pl.total > d.text_cap) return fail(c, NGSLD_E_INVALID, "output buffer too small for this scan");
  EmChoice ch;
  rc = choose_em(c, pl, ch);
  if (rc) return rc;
  fmt::FormatArgs fa;
  memset(&fa, 0, sizeof fa);
  uint32_t slot = 0;
  if (d.mode == MODE_TEXT) {
    fa.labels = c->have_labels ? c->d_label_blob : nullptr;
    fa.label_off = c->d_label_off;
    fa.maf = c->d_maf;
    fa.extend_out = P.extend_out;
    slot = fmt::slot_bytes(c->max_label_len, P.extend_out != 0);
    fa.slot = slot;
  }
  uint64_t chunk = std::min<unsigned long long>(c->chunk_rows, pl.total);
  // text through a sink is staged in page-locked chunk buffers of the context: keep those at 1 M rows
  if (d.mode == MODE_TEXT && !d.text_dst) chunk = std::min<uint64_t>(chunk, 1ull << 20);
  const bool use_tiles = ch.tile && ch.v && !ch.w && !P.strict;
  std::vector<uint2> tiles;
  std::vector<size_t> block_off;
  if (use_tiles) {
    // whole A-blocks per chunk, so no tile is ever staged twice
    build_tiles(pl, pl.c_lo, pl.c_hi, ch.TA, ch.TB, tiles, block_off);
    for (size_t ab = 0; ab + 1 < block_off.size(); ab++) {
      const uint32_t a = pl.c_lo + (uint32_t)ab * ch.TA, a_end = std::min<uint32_t>(pl.c_hi, a + ch.TA);
      chunk = std::max<uint64_t>(chunk, pl.row_off[a_end] - pl.row_off[a]);
    }
    if (tiles.size() > c->cap_tiles) {
      dfree(c->d_tiles);
      c->cap_tiles = tiles.size() + 1024;
      CUDA_TRY(c, cudaMalloc(&c->d_tiles, c->cap_tiles * sizeof(uint2)));
    }
    if (!tiles.empty()) {
      CUDA_TRY(c, cudaMemcpy(c->d_tiles, tiles.data(), tiles.size() * sizeof(uint2), cudaMemcpyHostToDevice));
      c->stats.h2d_bytes += tiles.size() * sizeof(uint2);
    }
  }
  rc = ensure_chunks(c, chunk, d.mode == MODE_ROWS && !d.rows_dst, d.mode == MODE_TEXT, slot, d.text_dst == nullptr);
  if (rc) return rc;
  CUDA_TRY(c, cudaMemsetAsync(c->d_ctr, 0, sizeof(DevCounters), c->s_main));
  c->stats.ms_plan = now_ms() - t_plan0;  // everything on the host before the first launch: plan, kernel choice, buffers
  if (dbg)
    fprintf(stderr, "[plan] rows %llu: make_plan %.2f ms, upload_plan %.2f ms, kernel choice + buffers %.2f ms\n", pl.total,
            t_plan1 - t_plan0, t_plan2 - t_plan1, now_ms() - t_plan2);
  CUDA_TRY(c, cudaEventRecord(c->ev_begin, c->s_main));
  unsigned long long r0 = 0;
  size_t ab0 = 0;
  int k = 0;
  while (r0 < pl.total) {
    ChunkBuf &b = c->buf[k & 1];
    if (b.pending) {
      rc = deliver_chunk(c, b, d);
      if (rc) return rc;
    }
    unsigned long long r1;
    size_t t0 = 0, t1 = 0;
    if (use_tiles) {
      size_t ab1 = ab0;
      r1 = r0;
      while (ab1 + 1 < block_off.size()) {
        const uint32_t a_end = std::min<uint32_t>(pl.c_hi, pl.c_lo + (uint32_t)(ab1 + 1) * ch.TA);
        if (pl.row_off[a_end] - r0 > chunk && ab1 > ab0) break;
        r1 = pl.row_off[a_end];
        ab1++;
        if (r1 - r0 >= chunk) break;
      }
      t0 = block_off[ab0];
      t1 = block_off[ab1];
      ab0 = ab1;
      if (r1 == r0) continue;  // A-blocks without pairs
    } else {
      r1 = std::min<unsigned long long>(pl.total, r0 + chunk);
    }
    rc = launch_chunk(c, pl, P, ch, b, r0, r1, t0, t1, d.mode, &fa, d.rows_dst);
    if (rc) return rc;
    r0 = r1;
    k++;
  }
  for (int j = 0; j < 2; j++) {
    ChunkBuf &b = c->buf[(k + j) & 1];
    if (b.pending) {
      rc = deliver_chunk(c, b, d);
      if (rc) return rc;
    }
  }
  CUDA_TRY(c, cudaEventRecord(c->ev_end, c->s_main));
  CUDA_TRY(c, cudaEventSynchronize(c->ev_end));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, c->ev_begin, c->ev_end) == cudaSuccess) c->stats.ms_device_total = ms;
  DevCounters hc;
  CUDA_TRY(c, cudaMemcpy(&hc, c->d_ctr, sizeof hc, cudaMemcpyDeviceToHost));
  c->stats.sum_em_passes = hc.em_passes;
  c->stats.sum_cells = hc.cells;
  c->stats.sum_cell_passes = hc.cell_passes;
  c->stats.n_cell_pairs = hc.cell_pairs;
  c->stats.n_resid_pairs = hc.resid_pairs;
  return NGSLD_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int ngsld_abi_version(void) { return NGSLD_ABI_VERSION; }

int ngsld_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int ngsld_create(ngsld_ctx **out, int device) {
  if (!out) return NGSLD_E_INVALID;
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) +
                     "); this library has no CPU fallback";
    return NGSLD_E_CUDA;
  }
  if (device < 0 || device >= n_dev) {
    g_create_error = "device index out of range";
    return NGSLD_E_INVALID;
  }
  ngsld_ctx *c = new (std::nothrow) ngsld_ctx();
  if (!c) return NGSLD_E_NOMEM;
  c->device = device;
  auto bail = [&](const char *what, cudaError_t err) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
    ngsld_destroy(c);
    return NGSLD_E_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
  if (prop.major != 10) {
    g_create_error = "this build carries sm_100a code only; device is sm_" + std::to_string(prop.major * 10 + prop.minor);
    ngsld_destroy(c);
    return NGSLD_E_CUDA;
  }
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = (int)prop.sharedMemPerBlockOptin;
  if ((e = cudaStreamCreateWithFlags(&c->own_main, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&c->s_aux, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  c->s_main = c->own_main;
  for (auto &b : c->buf) {
    cudaEvent_t *evs[] = {&b.ev_ready, &b.ev_em0, &b.ev_em1, &b.ev_p0, &b.ev_p1, &b.ev_f0, &b.ev_f1, &b.ev_done};
    for (auto ev : evs)
      if ((e = cudaEventCreate(ev)) != cudaSuccess) return bail("event", e);
  }
  if ((e = cudaEventCreate(&c->ev_begin)) != cudaSuccess) return bail("event", e);
  if ((e = cudaEventCreate(&c->ev_end)) != cudaSuccess) return bail("event", e);
  if ((e = cudaMalloc(&c->d_ctr, sizeof(DevCounters))) != cudaSuccess) return bail("cudaMalloc", e);
  *out = c;
  return NGSLD_OK;
}

void ngsld_destroy(ngsld_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  free_chunks(c);
  dfree(c->d_gl);
  dfree(c->d_expg);
  dfree(c->d_ratio);
  dfree(c->d_maf);
  dfree(c->d_q);
  dfree(c->d_dx_sig);
  dfree(c->d_dx_se);
  dfree(c->d_cls);
  dfree(c->d_pal);
  dfree(c->d_pal_k);
  dfree(c->d_pal_miss);
  dfree(c->d_cell_stats);
  dfree(c->d_cum);
  dfree(c->d_seg);
  dfree(c->d_label_blob);
  dfree(c->d_label_off);
  dfree(c->d_cs);
  dfree(c->d_cw_end);
  dfree(c->d_row_off);
  dfree(c->d_seeds);
  dfree(c->d_counts);
  dfree(c->d_taus_jump);
  dfree(c->d_tiles);
  dfree(c->d_ctr);
  dfree(c->d_decay_bins);
  dfree(c->d_decay_outside);
  dfree(c->d_seen);
  for (auto &b : c->buf) {
    cudaEvent_t evs[] = {b.ev_ready, b.ev_em0, b.ev_em1, b.ev_p0, b.ev_p1, b.ev_f0, b.ev_f1, b.ev_done};
    for (auto ev : evs)
      if (ev) cudaEventDestroy(ev);
  }
  if (c->ev_begin) cudaEventDestroy(c->ev_begin);
  if (c->ev_end) cudaEventDestroy(c->ev_end);
  if (c->own_main) cudaStreamDestroy(c->own_main);
  if (c->s_aux) cudaStreamDestroy(c->s_aux);
  if (c->s_copy) cudaStreamDestroy(c->s_copy);
  delete c;
}

const char *ngsld_last_error(const ngsld_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int ngsld_set_stream(ngsld_ctx *c, void *stream) {
  if (!c) return NGSLD_E_INVALID;
  c->s_main = stream ? (cudaStream_t)stream : c->own_main;
  return NGSLD_OK;
}

int ngsld_set_chunk_rows(ngsld_ctx *c, uint64_t rows) {
  if (!c) return NGSLD_E_INVALID;
  c->chunk_rows = rows ? rows : (4ull << 20);
  return NGSLD_OK;
}

int ngsld_prepare_sites(const double *raw, uint64_t n_sites, uint64_t n_ind, int log_scale, int from_log_cells,
                        int ignore_miss_data, int call_geno, double N_thresh, double call_thresh, int n_threads,
                        double *gl, double *expg, double *maf) {
  if (!raw || !gl || !expg || !maf || n_sites == 0 || n_ind == 0) return NGSLD_E_INVALID;
  if (call_geno && N_thresh > call_thresh) return NGSLD_E_INVALID;  // gen_func.cpp:887-888
  hostprep::PrepOptions o;
  o.log_scale = log_scale != 0;
  o.from_log_cells = from_log_cells != 0;
  o.ignore_miss = ignore_miss_data != 0;
  o.call_geno = call_geno != 0;
  o.n_thresh = N_thresh;
  o.call_thresh = call_thresh;
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  return hostprep::prepare_sites(raw, n_sites, n_ind, o, n_threads, gl, expg, maf) == 0 ? NGSLD_OK : NGSLD_E_DATA;
}

namespace {
// Common head of ngsld_set_sites / ngsld_set_sites_raw: device buffers for the shape, positions and labels dropped.
int begin_sites(ngsld_ctx *c, uint64_t n_sites, uint64_t n_ind) {
  if (n_sites >= (1ull << 32) - 1) return fail(c, NGSLD_E_INVALID, "n_sites must be below 2^32-1");
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaStreamSynchronize(c->s_main));
  dfree(c->d_cum);
  dfree(c->d_label_blob);
  dfree(c->d_label_off);
  c->have_pos = c->have_labels = false;
  const uint64_t n_pad = (n_ind + 1) & ~1ull;    // rows stay 16-byte aligned for the TMA bulk copies
  const uint64_t n_cpad = (n_ind + 15) & ~15ull;  // class rows (one byte per individual) are read as 32-bit words
  int rc_alloc = alloc_site_buffers(c, n_sites, n_ind, true);  // same shape as last time: the device buffers are kept
  if (rc_alloc) return rc_alloc;
  c->n_sites = n_sites;
  c->n_ind = n_ind;
  c->n_pad = n_pad;
  c->n_cpad = n_cpad;
  CUDA_TRY(c, cudaMemsetAsync(c->d_seg, 0, n_sites * sizeof(uint32_t), c->s_main));
  if (n_pad != n_ind) CUDA_TRY(c, cudaMemsetAsync(c->d_gl, 0, n_sites * n_pad * 24, c->s_main));
  return NGSLD_OK;
}

// Common tail: with gl, maf (and expg unless host_expg is used for the cross-check path) on the device, derive the site
// palettes, decide about the class-compressed EM, and compute the per-site x87 terms of r2_ExpG.
int finish_sites(ngsld_ctx *c, const double *host_expg) {
  const uint64_t n_sites = c->n_sites, n_ind = c->n_ind, n_pad = c->n_pad, n_cpad = c->n_cpad;
  const uint64_t n_blk = (n_ind + 3) / 4;
  // site palettes for the class-compressed EM, and a sample of pairs to see whether it pays on this data
  c->cell_ok = c->cell_possible = false;
  c->cell_mean = c->cell_uncoded_frac = 0;
  c->cell_p995 = 0;
  if (c->d_cls && n_sites >= 2) {
    const unsigned pblocks = (unsigned)std::min<uint64_t>((n_sites + 3) / 4, (uint64_t)c->sm_count * 16);
    const size_t stat_bytes = 3 * sizeof(unsigned long long) + 132 * sizeof(unsigned int);
    CUDA_TRY(c, cudaMemsetAsync(c->d_cell_stats, 0, stat_bytes, c->s_main));
    emcell::build_palette_kernel<<<pblocks, emcell::CTA_THREADS, 0, c->s_main>>>(
        c->d_gl, (uint32_t)n_sites, (uint32_t)n_ind, (uint32_t)n_pad, (uint32_t)n_cpad, c->d_cls, c->d_pal, c->d_pal_k, c->d_pal_miss,
        reinterpret_cast<unsigned int *>(c->d_cell_stats + 3) + 130);  // hist[130]: largest palette
    const uint32_t n_samples = 4096;
    SiteTable T = site_table(c);
    emcell::cell_stats_kernel<<<c->sm_count * 2, emcell::CTA_THREADS, 0, c->s_main>>>(
        T, n_samples, 0, c->d_cell_stats, reinterpret_cast<unsigned int *>(c->d_cell_stats + 3));
    CUDA_TRY(c, cudaGetLastError());
    struct {
      unsigned long long sum, n, uncoded;
      unsigned int hist[132];
    } hs;
    static_assert(sizeof(hs) == 3 * sizeof(unsigned long long) + 132 * sizeof(unsigned int), "layout");
    CUDA_TRY(c, cudaMemcpyAsync(&hs, c->d_cell_stats, stat_bytes, cudaMemcpyDeviceToHost, c->s_main));
    CUDA_TRY(c, cudaStreamSynchronize(c->s_main));
    const unsigned long long coded = hs.n - hs.uncoded;
    c->cell_uncoded_frac = hs.n ? (double)hs.uncoded / (double)hs.n : 1.0;
    c->cell_kstride = std::min<uint32_t>(NGSLD_KMAX, std::max<uint32_t>(8, (hs.hist[130] + 7u) & ~7u));
    if (coded) {
      c->cell_mean = (double)hs.sum / (double)coded;
      unsigned long long acc = 0;
      uint32_t b = 0;
      for (; b < 129; b++) {
        acc += hs.hist[b];
        if ((double)acc >= 0.995 * (double)coded) break;
      }
      c->cell_p995 = 32u * (std::min<uint32_t>(b, 128) + 1);
      c->cell_possible = true;
      // pays when a pair has clearly fewer cells than individuals and (almost) every site could be coded (measured at 100
      // individuals, 65 cells per pair: 96.7 M pairs/s against 80.3 M for the sub-warp group kernels); small samples stay
      // with the group kernels
      c->cell_ok = n_ind >= 64 && c->cell_uncoded_frac <= 0.05 && c->cell_mean <= 0.7 * (double)n_ind;
    }
  }
  // per-site x87 terms of the expected-genotype correlation
  const char *host_terms = getenv("NGSLD_HOST_TERMS");
  if (host_expg && host_terms && atoi(host_terms)) {
    // cross-check path: the host FPU's native long double instead of the device's emulation (same bits)
    std::vector<uint64_t> sig(n_sites * n_pad);
    std::vector<uint16_t> se(n_sites * n_pad);
    std::vector<double> q(n_sites);
    const int nt = (int)std::max(1u, std::thread::hardware_concurrency());
    hostprep::pearson_site_terms(host_expg, n_sites, n_ind, n_pad, nt, sig.data(), se.data(), q.data());
    {  // the device table holds blocks of four individuals, site by site inside a block, in mac3's packed exponent form
      std::vector<uint64_t> sig_t(n_sites * n_blk * 4, 0);
      std::vector<uint16_t> se_t(n_sites * n_blk * 4, 0);
      for (uint64_t s = 0; s < n_sites; s++)
        for (uint64_t i = 1; i < n_ind; i++) {
          const uint64_t at = ((i >> 2) * n_sites + s) * 4 + (i & 3);
          sig_t[at] = sig[s * n_pad + i];
          se_t[at] = sig_t[at] ? x87::se14_from_x87(se[s * n_pad + i]) : (uint16_t)0;
        }
      sig.swap(sig_t);
      se.swap(se_t);
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->d_dx_sig, sig.data(), sig.size() * 8, cudaMemcpyHostToDevice, c->s_main));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_dx_se, se.data(), se.size() * 2, cudaMemcpyHostToDevice, c->s_main));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_q, q.data(), q.size() * 8, cudaMemcpyHostToDevice, c->s_main));
    aux::site_terms_kernel<<<8, 128, 0, c->s_main>>>(nullptr, 0, 0, (uint32_t)n_blk, nullptr, nullptr, nullptr, c->d_ratio);
    CUDA_TRY(c, cudaStreamSynchronize(c->s_main));
  } else {
    const unsigned blocks = (unsigned)std::min<uint64_t>((n_sites + 127) / 128, (uint64_t)c->sm_count * 16);
    aux::site_terms_kernel<<<blocks, 128, 0, c->s_main>>>(c->d_expg, (uint32_t)n_sites, (uint32_t)n_ind, (uint32_t)n_blk,
                                                          c->d_dx_sig, c->d_dx_se, c->d_q, c->d_ratio);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(c->s_main));
  }
  c->h_seg.assign(n_sites, 0);
  c->h_cum.assign(n_sites, 0.0);
  return NGSLD_OK;
}
}  // namespace

int ngsld_set_sites(ngsld_ctx *c, const double *gl, const double *expg, const double *maf, uint64_t n_sites,
                    uint64_t n_ind) {
  if (!c) return NGSLD_E_INVALID;
  if (!gl || !expg || !maf || n_sites == 0 || n_ind == 0) return fail(c, NGSLD_E_INVALID, "null or empty site arrays");
  for (uint64_t s = 0; s < n_sites; s++)
    if (maf[s] < 0 || maf[s] > 1) return fail(c, NGSLD_E_DATA, "invalid allele frequencies");  // gen_func.cpp:1030
  int rc = begin_sites(c, n_sites, n_ind);
  if (rc) return rc;
  CUDA_TRY(c, cudaMemcpy2DAsync(c->d_gl, c->n_pad * 24, gl, n_ind * 24, n_ind * 24, n_sites, cudaMemcpyHostToDevice, c->s_main));
  CUDA_TRY(c, cudaMemcpyAsync(c->d_maf, maf, n_sites * sizeof(double), cudaMemcpyHostToDevice, c->s_main));
  CUDA_TRY(c, cudaMemcpyAsync(c->d_expg, expg, n_sites * n_ind * sizeof(double), cudaMemcpyHostToDevice, c->s_main));
  c->h_maf.assign(maf, maf + n_sites);
  return finish_sites(c, expg);
}

int ngsld_set_sites_raw(ngsld_ctx *c, const double *raw, uint64_t n_sites, uint64_t n_ind, int log_scale, int from_log_cells,
                        int ignore_miss_data, int call_geno, double N_thresh, double call_thresh, double *maf_out) {
  if (!c) return NGSLD_E_INVALID;
  if (!raw || n_sites == 0 || n_ind == 0) return fail(c, NGSLD_E_INVALID, "null or empty site arrays");
  if (call_geno && N_thresh > call_thresh)  // gen_func.cpp:887-888
    return fail(c, NGSLD_E_INVALID, "missing data threshold must be smaller than calling genotype threshold!");
  int rc = begin_sites(c, n_sites, n_ind);
  if (rc) return rc;
  // the file's cells go straight into the likelihood table and are prepared there, in place
  CUDA_TRY(c, cudaMemcpy2DAsync(c->d_gl, c->n_pad * 24, raw, n_ind * 24, n_ind * 24, n_sites, cudaMemcpyHostToDevice, c->s_main));
  int *d_flag = nullptr;
  CUDA_TRY(c, cudaMalloc(&d_flag, sizeof(int)));
  CUDA_TRY(c, cudaMemsetAsync(d_flag, 0, sizeof(int), c->s_main));
  const unsigned blocks = (unsigned)std::min<uint64_t>((n_sites + 3) / 4, (uint64_t)c->sm_count * 16);
  aux::prep_sites_kernel<<<blocks, 128, 0, c->s_main>>>(c->d_gl, (uint32_t)n_sites, (uint32_t)n_ind, (uint32_t)c->n_pad,
                                                        (!from_log_cells && !log_scale) ? 1 : 0, ignore_miss_data, call_geno, N_thresh,
                                                        call_thresh, c->d_expg, c->d_maf, d_flag);
  int flag = 0;
  c->h_maf.assign(n_sites, 0.0);
  cudaError_t e1 = cudaGetLastError();
  cudaError_t e2 = cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->s_main);
  cudaError_t e3 = cudaMemcpyAsync(c->h_maf.data(), c->d_maf, n_sites * sizeof(double), cudaMemcpyDeviceToHost, c->s_main);
  cudaError_t e4 = cudaStreamSynchronize(c->s_main);
  cudaFree(d_flag);
  for (cudaError_t e : {e1, e2, e3, e4})
    if (e != cudaSuccess) return fail(c, NGSLD_E_CUDA, cudaGetErrorString(e));
  if (flag) return fail(c, NGSLD_E_DATA, "NaN found! Is the file format correct?");  // read_data.cpp:42-45
  for (uint64_t s = 0; s < n_sites; s++)
    if (c->h_maf[s] < 0 || c->h_maf[s] > 1) return fail(c, NGSLD_E_DATA, "invalid allele frequencies");
  if (maf_out) memcpy(maf_out, c->h_maf.data(), n_sites * sizeof(double));
  return finish_sites(c, nullptr);
}

int ngsld_set_positions(ngsld_ctx *c, const double *pos_dist, const char *const *labels) {
  if (!c) return NGSLD_E_INVALID;
  if (!c->d_gl) return fail(c, NGSLD_E_INVALID, "ngsld_set_sites must come first");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const uint64_t n = c->n_sites;
  c->have_pos = false;
  if (pos_dist) {
    // exact prefix sums: every finite gap must be a non-negative integer and the total below 2^53, so that
    // cum[s2]-cum[s1] equals the reference's running double sum (ngsLD.cpp:241) bit for bit
    const char *why = build_cum(pos_dist, n, c->h_cum.data(), c->h_seg.data());
    if (why) return fail(c, NGSLD_E_DATA, why);
    dfree(c->d_cum);
    CUDA_TRY(c, cudaMalloc(&c->d_cum, n * sizeof(double)));
    CUDA_TRY(c, cudaMemcpy(c->d_cum, c->h_cum.data(), n * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_seg, c->h_seg.data(), n * 4, cudaMemcpyHostToDevice));
    c->have_pos = true;
  }
  dfree(c->d_label_blob);
  dfree(c->d_label_off);
  c->have_labels = false;
  c->max_label_len = 6;
  if (labels) {
    std::vector<uint32_t> off(n + 1);
    std::string blob;
    c->h_labels.assign(n, std::string());
    uint32_t mx = 0;
    for (uint64_t s = 0; s < n; s++) {
      off[s] = (uint32_t)blob.size();
      const char *l = labels[s] ? labels[s] : "(null)";
      const size_t len = strlen(l);
      if (blob.size() + len >= (1ull << 32)) return fail(c, NGSLD_E_INVALID, "labels exceed 4 GiB");
      blob.append(l, len);
      c->h_labels[s].assign(l, len);
      mx = std::max<uint32_t>(mx, (uint32_t)len);
    }
    off[n] = (uint32_t)blob.size();
    CUDA_TRY(c, cudaMalloc(&c->d_label_blob, std::max<size_t>(blob.size(), 1)));
    CUDA_TRY(c, cudaMalloc(&c->d_label_off, (n + 1) * sizeof(uint32_t)));
    CUDA_TRY(c, cudaMemcpy(c->d_label_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_label_off, off.data(), (n + 1) * 4, cudaMemcpyHostToDevice));
    c->have_labels = true;
    c->label_blob_bytes = blob.size();
    c->max_label_len = std::max<uint32_t>(mx, 1);
  }
  return NGSLD_OK;
}

void ngsld_scan_defaults(ngsld_scan_params *p) {  // reference parse_args.cpp:6-29
  if (!p) return;
  memset(p, 0, sizeof *p);
  p->max_kb_dist = 100;
  p->max_snp_dist = 0;
  p->min_maf = 0;
  p->rnd_sample = 1;
  p->seed = 1;
}

int ngsld_scan_count(ngsld_ctx *c, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p, uint64_t *n_rows) {
  if (!c || !p || !n_rows) return NGSLD_E_INVALID;
  if (!c->d_gl) return fail(c, NGSLD_E_INVALID, "ngsld_set_sites must be called before a scan");
  CUDA_TRY(c, cudaSetDevice(c->device));
  Plan pl;
  int rc = make_plan(plan_input(c), s1_lo, s1_hi, *p, pl);
  if (rc) return rc;
  if (pl.sampled) {
    rc = upload_plan(c, pl, *p);
    if (rc) return rc;
  }
  *n_rows = pl.total;
  return NGSLD_OK;
}

int ngsld_partition(ngsld_ctx *c, const ngsld_scan_params *p, int n_parts, uint64_t *bounds) {
  if (!c || !p || !bounds || n_parts < 1) return NGSLD_E_INVALID;
  if (!c->d_gl) return fail(c, NGSLD_E_INVALID, "ngsld_set_sites must be called before a scan");
  CUDA_TRY(c, cudaSetDevice(c->device));
  Plan pl;
  int rc = make_plan(plan_input(c), 0, c->n_sites, *p, pl);
  if (rc) return rc;
  if (pl.sampled) {
    rc = upload_plan(c, pl, *p);
    if (rc) return rc;
  }
  bounds_from_plan(pl, c->n_sites, n_parts, bounds);
  return NGSLD_OK;
}

// ---- device-free planning and input files -------------------------------------------------------
namespace {
int host_plan(const double *maf, const double *pos_dist, uint64_t n_sites, const ngsld_scan_params *p, uint64_t s1_lo,
              uint64_t s1_hi, Plan &pl) {
  if (!maf || !p || n_sites == 0) {
    g_create_error = "null or empty planning input";
    return NGSLD_E_INVALID;
  }
  if (!(p->rnd_sample > 0) || p->rnd_sample > 1) {
    g_create_error = "proportion of comparisons to sample must be in ]0,1]!";
    return NGSLD_E_INVALID;
  }
  std::vector<double> cum(n_sites, 0.0);
  std::vector<uint32_t> seg(n_sites, 0);
  if (pos_dist) {
    const char *why = build_cum(pos_dist, n_sites, cum.data(), seg.data());
    if (why) {
      g_create_error = why;
      return NGSLD_E_DATA;
    }
  }
  PlanInput in;
  in.n_sites = n_sites;
  in.maf = maf;
  in.have_pos = pos_dist != nullptr;
  in.cum = cum.data();
  in.seg = seg.data();
  int rc = make_plan(in, s1_lo, s1_hi, *p, pl);
  if (rc) return rc;
  if (pl.sampled) {
    std::vector<unsigned long long> counts;
    host_sample_counts(pl, *p, n_sites, counts);
    finish_sampled_plan(pl, counts);
  }
  return NGSLD_OK;
}
}  // namespace

int ngsld_plan_count(const double *maf, const double *pos_dist, uint64_t n_sites, const ngsld_scan_params *p,
                     uint64_t s1_lo, uint64_t s1_hi, uint64_t *n_rows) {
  if (!n_rows) return NGSLD_E_INVALID;
  Plan pl;
  int rc = host_plan(maf, pos_dist, n_sites, p, s1_lo, s1_hi, pl);
  if (rc) return rc;
  *n_rows = pl.total;
  return NGSLD_OK;
}

int ngsld_plan_partition(const double *maf, const double *pos_dist, uint64_t n_sites, const ngsld_scan_params *p,
                         int n_parts, uint64_t *bounds) {
  if (!bounds || n_parts < 1) return NGSLD_E_INVALID;
  Plan pl;
  int rc = host_plan(maf, pos_dist, n_sites, p, 0, n_sites, pl);
  if (rc) return rc;
  bounds_from_plan(pl, n_sites, n_parts, bounds);
  return NGSLD_OK;
}

int ngsld_load_geno(const char *path, int is_bin, int probs, int log_scale, uint64_t n_ind, uint64_t n_sites,
                    double *cells, int *log_cells) {
  if (!path || !cells || !log_cells || n_ind == 0 || n_sites == 0) {
    g_create_error = "[read_geno] null or empty arguments";
    return NGSLD_E_INVALID;
  }
  bool lc = false;
  const loader::Failure f = loader::read_geno(path, is_bin != 0, probs != 0, log_scale != 0, n_ind, n_sites, cells, &lc);
  *log_cells = lc ? 1 : 0;
  if (f) {
    g_create_error = std::string("[") + f.func + "] " + f.msg;
    return f.io ? NGSLD_E_IO : NGSLD_E_DATA;
  }
  return NGSLD_OK;
}

int ngsld_load_positions(const char *path, int header, uint64_t n_sites, double *pos_dist, char **label_blob,
                         uint64_t *blob_bytes) {
  if (!path || !pos_dist || !label_blob || n_sites == 0) {
    g_create_error = "[read_dist] null or empty arguments";
    return NGSLD_E_INVALID;
  }
  *label_blob = nullptr;
  std::vector<std::string> labels;
  const loader::Failure f = loader::read_positions(path, header != 0, n_sites, labels, pos_dist);
  if (f) {
    g_create_error = std::string("[") + f.func + "] " + f.msg;
    return f.io ? NGSLD_E_IO : NGSLD_E_DATA;
  }
  size_t total = 0;
  for (auto &l : labels) total += l.size() + 1;
  char *blob = (char *)malloc(total ? total : 1);
  if (!blob) return NGSLD_E_NOMEM;
  size_t o = 0;
  for (auto &l : labels) {
    memcpy(blob + o, l.c_str(), l.size() + 1);
    o += l.size() + 1;
  }
  *label_blob = blob;
  if (blob_bytes) *blob_bytes = total;
  return NGSLD_OK;
}

void ngsld_free(void *p) { free(p); }

int ngsld_scan(ngsld_ctx *c, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p, ngsld_row_sink sink, void *user) {
  Delivery d;
  d.mode = MODE_ROWS;
  d.rows = sink;
  d.text = nullptr;
  d.user = user;
  return run_scan(c, s1_lo, s1_hi, p, d);
}

int ngsld_scan_into(ngsld_ctx *c, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p, ngsld_pair_row *out,
                    uint64_t cap, uint64_t *n_rows) {
  if (!c) return NGSLD_E_INVALID;
  if (!out && cap) return fail(c, NGSLD_E_INVALID, "output buffer missing");
  ngsld_pair_row dummy;
  Delivery d;
  d.mode = MODE_ROWS;
  d.rows = nullptr;
  d.text = nullptr;
  d.user = nullptr;
  d.rows_dst = out ? out : &dummy;  // every chunk is copied from the device to out + its row offset, no staging
  d.text_cap = cap;                 // (capacity in rows)
  int rc = run_scan(c, s1_lo, s1_hi, p, d);
  if (n_rows) *n_rows = rc ? 0 : c->stats.n_pairs;
  return rc;
}

int ngsld_scan_tsv(ngsld_ctx *c, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p, ngsld_text_sink sink,
                   void *user) {
  Delivery d;
  d.mode = MODE_TEXT;
  d.extend_out = p ? p->extend_out : 0;
  d.rows = nullptr;
  d.text = sink;
  d.user = user;
  return run_scan(c, s1_lo, s1_hi, p, d);
}

int ngsld_scan_tsv_into(ngsld_ctx *c, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p, char *buf, uint64_t cap,
                        uint64_t *n_bytes, uint64_t *n_rows) {
  if (!c) return NGSLD_E_INVALID;
  if (!buf && cap) return fail(c, NGSLD_E_INVALID, "text buffer missing");
  uint64_t len = 0;
  Delivery d;
  d.mode = MODE_TEXT;
  d.extend_out = p ? p->extend_out : 0;
  d.rows = nullptr;
  d.text = nullptr;
  d.user = nullptr;
  d.text_dst = buf ? buf : (char *)&len;  // cap 0: any row makes the scan fail with "too small"
  d.text_cap = cap;
  d.text_len = &len;
  const int rc = run_scan(c, s1_lo, s1_hi, p, d);
  if (n_bytes) *n_bytes = len;
  if (n_rows) *n_rows = c->stats.n_pairs;
  return rc;
}

uint64_t ngsld_tsv_row_bound(const ngsld_ctx *c, int extend_out) {
  return fmt::slot_bytes(c ? c->max_label_len : 6, extend_out != 0);
}

uint64_t ngsld_tsv_row_bound_for(uint32_t max_label_len, int extend_out) {
  return fmt::slot_bytes(max_label_len ? max_label_len : 1, extend_out != 0);
}

int ngsld_alloc_host(void **p, size_t bytes) {
  if (!p) return NGSLD_E_INVALID;
  *p = nullptr;
  cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaHostAlloc: ") + cudaGetErrorString(e);
    *p = nullptr;
    return e == cudaErrorMemoryAllocation ? NGSLD_E_NOMEM : NGSLD_E_CUDA;
  }
  return NGSLD_OK;
}

void ngsld_free_host(void *p) {
  if (p) cudaFreeHost(p);
}

// Device-to-device copy of everything ngsld_set_sites + ngsld_set_positions put on src's GPU (NVLink when the two
// devices are peers), instead of a second upload from the host.
int ngsld_share_sites(ngsld_ctx *dst, const ngsld_ctx *src) {
  if (!dst || !src) return NGSLD_E_INVALID;
  ngsld_ctx *c = dst;
  if (dst == src) return NGSLD_OK;
  if (!src->d_gl) return fail(c, NGSLD_E_INVALID, "the source context holds no sites");
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaStreamSynchronize(c->s_main));
  if (c->device != src->device) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, c->device, src->device);
    if (can) {
      cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(c, NGSLD_E_CUDA, cudaGetErrorString(e));
      cudaGetLastError();  // "already enabled" is not an error
    }
  }
  int rc = alloc_site_buffers(c, src->n_sites, src->n_ind, false);
  if (rc) return rc;
  const uint64_t n = src->n_sites, n_pad = src->n_pad, n_cpad = src->n_cpad;
  c->n_sites = n;
  c->n_ind = src->n_ind;
  c->n_pad = n_pad;
  c->n_cpad = n_cpad;
  uint64_t moved = 0;
  auto peer = [&](void *to, const void *from, size_t bytes) -> cudaError_t {
    moved += bytes;
    return cudaMemcpyPeerAsync(to, c->device, from, src->device, bytes, c->s_main);
  };
  CUDA_TRY(c, peer(c->d_gl, src->d_gl, n * n_pad * 24));
  CUDA_TRY(c, peer(c->d_maf, src->d_maf, n * 8));
  CUDA_TRY(c, peer(c->d_q, src->d_q, n * 8));
  const uint64_t n_blk = (src->n_ind + 3) / 4;
  CUDA_TRY(c, peer(c->d_dx_sig, src->d_dx_sig, n * n_blk * 4 * 8));
  CUDA_TRY(c, peer(c->d_dx_se, src->d_dx_se, n * n_blk * 4 * 2));
  CUDA_TRY(c, peer(c->d_seg, src->d_seg, n * 4));
  CUDA_TRY(c, peer(c->d_ratio, src->d_ratio, n_blk * 4 * 8));
  if (c->d_cls && src->d_cls) {
    CUDA_TRY(c, peer(c->d_cls, src->d_cls, n * n_cpad));
    CUDA_TRY(c, peer(c->d_pal, src->d_pal, n * (size_t)NGSLD_KMAX * 24));
    CUDA_TRY(c, peer(c->d_pal_k, src->d_pal_k, n));
    CUDA_TRY(c, peer(c->d_pal_miss, src->d_pal_miss, n * 8));
  }
  dfree(c->d_cum);
  dfree(c->d_label_blob);
  dfree(c->d_label_off);
  if (src->have_pos) {
    CUDA_TRY(c, cudaMalloc(&c->d_cum, n * sizeof(double)));
    CUDA_TRY(c, peer(c->d_cum, src->d_cum, n * 8));
  }
  if (src->have_labels) {
    CUDA_TRY(c, cudaMalloc(&c->d_label_blob, std::max<size_t>(src->label_blob_bytes, 1)));
    CUDA_TRY(c, cudaMalloc(&c->d_label_off, (n + 1) * sizeof(uint32_t)));
    if (src->label_blob_bytes) CUDA_TRY(c, peer(c->d_label_blob, src->d_label_blob, src->label_blob_bytes));
    CUDA_TRY(c, peer(c->d_label_off, src->d_label_off, (n + 1) * 4));
  }
  c->h_maf = src->h_maf;
  c->h_cum = src->h_cum;
  c->h_seg = src->h_seg;
  c->h_labels = src->h_labels;
  c->have_pos = src->have_pos;
  c->have_labels = src->have_labels;
  c->max_label_len = src->max_label_len;
  c->label_blob_bytes = src->label_blob_bytes;
  c->cell_ok = src->cell_ok;
  c->cell_possible = src->cell_possible && c->d_cls;
  c->cell_mean = src->cell_mean;
  c->cell_uncoded_frac = src->cell_uncoded_frac;
  c->cell_p995 = src->cell_p995;
  c->cell_kstride = src->cell_kstride;
  CUDA_TRY(c, cudaStreamSynchronize(c->s_main));
  memset(&c->stats, 0, sizeof c->stats);
  c->stats.h2d_bytes = 0;
  c->stats.d2h_bytes = moved;  // reported as bytes moved between devices
  return NGSLD_OK;
}

int ngsld_scan_device(ngsld_ctx *c, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p) {
  Delivery d;
  d.mode = MODE_DEVICE;
  d.rows = nullptr;
  d.text = nullptr;
  d.user = nullptr;
  return run_scan(c, s1_lo, s1_hi, p, d);
}

int ngsld_scan_decay(ngsld_ctx *c, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p, double bin_size,
                     uint64_t n_bins, ngsld_decay_bin *bins, uint64_t *n_outside) {
  if (!c) return NGSLD_E_INVALID;
  if (!bins || n_bins == 0 || !(bin_size > 0)) return fail(c, NGSLD_E_INVALID, "decay bins: need bin_size > 0 and n_bins > 0");
  CUDA_TRY(c, cudaSetDevice(c->device));
  dfree(c->d_decay_bins);
  dfree(c->d_decay_outside);
  CUDA_TRY(c, cudaMalloc(&c->d_decay_bins, n_bins * sizeof(ngsld_decay_bin)));
  CUDA_TRY(c, cudaMalloc(&c->d_decay_outside, sizeof(unsigned long long)));
  CUDA_TRY(c, cudaMemsetAsync(c->d_decay_bins, 0, n_bins * sizeof(ngsld_decay_bin), c->s_main));
  CUDA_TRY(c, cudaMemsetAsync(c->d_decay_outside, 0, sizeof(unsigned long long), c->s_main));
  c->decay_active = true;
  c->decay_bin_size = bin_size;
  c->decay_n_bins = n_bins;
  Delivery d;
  d.mode = MODE_DEVICE;
  d.rows = nullptr;
  d.text = nullptr;
  d.user = nullptr;
  const int rc = run_scan(c, s1_lo, s1_hi, p, d);
  c->decay_active = false;
  if (rc) return rc;
  unsigned long long outside = 0;
  CUDA_TRY(c, cudaMemcpyAsync(bins, c->d_decay_bins, n_bins * sizeof(ngsld_decay_bin), cudaMemcpyDeviceToHost, c->s_main));
  CUDA_TRY(c, cudaMemcpyAsync(&outside, c->d_decay_outside, sizeof outside, cudaMemcpyDeviceToHost, c->s_main));
  CUDA_TRY(c, cudaStreamSynchronize(c->s_main));
  c->stats.d2h_bytes += n_bins * sizeof(ngsld_decay_bin) + 8;
  if (n_outside) *n_outside = outside;
  return NGSLD_OK;
}

int ngsld_scan_edges(ngsld_ctx *c, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p, const ngsld_prune_params *q,
                     ngsld_edge_sink sink, void *user, uint8_t *seen) {
  if (!c) return NGSLD_E_INVALID;
  if (!q) return fail(c, NGSLD_E_INVALID, "pruning parameters missing");
  if (q->field < 4 || q->field > 7) return fail(c, NGSLD_E_INVALID, "weight field must be 4 (r2_ExpG), 5 (D), 6 (Dp) or 7 (r2)");
  if (q->weight_type != 'a' && q->weight_type != 'e' && q->weight_type != 'n') return fail(c, NGSLD_E_INVALID, "weight type must be 'a', 'e' or 'n'");
  if (q->weight_precision < 0 || q->weight_precision > 8) return fail(c, NGSLD_E_INVALID, "weight precision must be in [0,8]");
  if (!c->d_gl) return fail(c, NGSLD_E_INVALID, "ngsld_set_sites must be called before a scan");
  CUDA_TRY(c, cudaSetDevice(c->device));
  dfree(c->d_seen);
  CUDA_TRY(c, cudaMalloc(&c->d_seen, c->n_sites));
  CUDA_TRY(c, cudaMemsetAsync(c->d_seen, 0, c->n_sites, c->s_main));
  c->prune_active = true;
  c->prune_q = *q;
  c->edge_sink = sink;
  c->edge_user = user;
  Delivery d;
  d.mode = MODE_DEVICE;
  d.rows = nullptr;
  d.text = nullptr;
  d.user = nullptr;
  const int rc = run_scan(c, s1_lo, s1_hi, p, d);
  c->prune_active = false;
  c->edge_sink = nullptr;
  if (rc) return rc;
  if (seen) {
    std::vector<unsigned char> h(c->n_sites);
    CUDA_TRY(c, cudaMemcpyAsync(h.data(), c->d_seen, c->n_sites, cudaMemcpyDeviceToHost, c->s_main));
    CUDA_TRY(c, cudaStreamSynchronize(c->s_main));
    for (uint64_t s = 0; s < c->n_sites; s++)
      if (h[s]) seen[s] = 1;
    c->stats.d2h_bytes += c->n_sites;
  }
  return NGSLD_OK;
}

int ngsld_prune_graph(uint64_t n_sites, const char *const *labels, const uint8_t *seen, const ngsld_edge *edges, uint64_t n_edges,
                      int keep_heavy, uint8_t *kept, uint32_t *excluded, uint64_t *n_excluded) {
  if (!seen || !kept || (!edges && n_edges) || n_sites == 0) return NGSLD_E_INVALID;
  for (uint64_t e = 0; e < n_edges; e++)
    if (edges[e].s1 >= n_sites || edges[e].s2 >= n_sites || edges[e].s1 == edges[e].s2) return NGSLD_E_INVALID;
  std::vector<uint32_t> rank, excl;
  prune::label_ranks(n_sites, labels, rank);
  prune::run(n_sites, seen, rank.data(), edges, n_edges, keep_heavy != 0, kept, excl);
  if (excluded) memcpy(excluded, excl.data(), excl.size() * sizeof(uint32_t));
  if (n_excluded) *n_excluded = excl.size();
  return NGSLD_OK;
}

int ngsld_get_stats(const ngsld_ctx *c, ngsld_scan_stats *out) {
  if (!c || !out) return NGSLD_E_INVALID;
  *out = c->stats;
  return NGSLD_OK;
}

int ngsld_pairs(ngsld_ctx *c, const uint32_t *s1, const uint32_t *s2, uint64_t n_pairs, int ignore_miss_data,
                int strict, ngsld_pair_row *out) {
  if (!c) return NGSLD_E_INVALID;
  if (!c->d_gl) return fail(c, NGSLD_E_INVALID, "ngsld_set_sites must be called first");
  if (n_pairs == 0) return NGSLD_OK;
  if (!s1 || !s2 || !out) return fail(c, NGSLD_E_INVALID, "null pair arrays");
  for (uint64_t k = 0; k < n_pairs; k++)
    if (s1[k] >= c->n_sites || s2[k] >= c->n_sites) return fail(c, NGSLD_E_INVALID, "site index out of range");
  CUDA_TRY(c, cudaSetDevice(c->device));
  memset(&c->stats, 0, sizeof c->stats);
  Plan pl;  // only used for kernel choice (list path)
  pl.sampled = true;
  EmChoice ch;
  int rc = choose_em(c, pl, ch);
  if (rc) return rc;
  const uint64_t chunk = std::min<uint64_t>(c->chunk_rows, n_pairs);
  rc = ensure_chunks(c, chunk, true, false, 0);
  if (rc) return rc;
  CUDA_TRY(c, cudaMemsetAsync(c->d_ctr, 0, sizeof(DevCounters), c->s_main));
  const SiteTable T = site_table(c);
  ngsld_scan_params P;
  ngsld_scan_defaults(&P);
  P.ignore_miss_data = ignore_miss_data;
  P.strict = strict;
  for (uint64_t r0 = 0; r0 < n_pairs; r0 += chunk) {
    const uint64_t n = std::min<uint64_t>(chunk, n_pairs - r0);
    ChunkBuf &b = c->buf[0];
    CUDA_TRY(c, cudaMemcpyAsync(b.d_s1, s1 + r0, n * 4, cudaMemcpyHostToDevice, c->s_main));
    CUDA_TRY(c, cudaMemcpyAsync(b.d_s2, s2 + r0, n * 4, cudaMemcpyHostToDevice, c->s_main));
    PairChunk C;
    C.s1 = b.d_s1;
    C.s2 = b.d_s2;
    C.rows = b.d_rows;
    C.n_pairs = n;
    const unsigned gb = (unsigned)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)c->sm_count * 32);
    aux::fill_rows_kernel<<<gb, 256, 0, c->s_main>>>(T, C);
    const unsigned pb = (unsigned)std::min<unsigned long long>((n + 127) / 128, (unsigned long long)c->sm_count * 64);
    CUDA_TRY(c, cudaMemsetAsync(c->d_ctr, 0, NGSLD_WORK_COUNTERS * sizeof(unsigned long long), c->s_main));
    const bool use_cell = ch.cell && ch.w && !strict;
    if (!(use_cell && ch.cell_fuse)) aux::pearson_kernel<<<pb, 128, 0, c->s_main>>>(T, C, c->d_ctr);
    if (strict || (!ch.v && !ch.w)) {
      aux::em_strict_kernel<<<pb, 128, 0, c->s_main>>>(T, C, ignore_miss_data, c->d_ctr);
    } else if (use_cell) {
      int rcc = launch_cell(c, ch, T, C, ignore_miss_data, b.d_resid);
      if (rcc) return rcc;
    } else if (ch.w) {
      int rcw = launch_warp(c, ch, T, C, ignore_miss_data);
      if (rcw) return rcw;
    } else {
      int ign = ignore_miss_data;
      SiteTable Tt = T;
      PairChunk Cc = C;
      DevCounters *ctr = c->d_ctr;
      void *args[] = {&Tt, &Cc, &ign, &ctr};
      const unsigned long long gpc = std::max(emfast::CTA_THREADS, ch.v->lpg) / ch.v->lpg;
      const unsigned blocks = (unsigned)std::min<unsigned long long>((n + gpc - 1) / gpc, ch.blocks_list);
      CUDA_TRY(c, cudaLaunchKernel(ch.v->list_fn, dim3(blocks), dim3(std::max(emfast::CTA_THREADS, ch.v->lpg)), args, 0, c->s_main));
    }
    c->stats.n_launches += 3;
    CUDA_TRY(c, cudaMemcpyAsync(b.h_rows, b.d_rows, n * sizeof(ngsld_pair_row), cudaMemcpyDeviceToHost, c->s_main));
    CUDA_TRY(c, cudaStreamSynchronize(c->s_main));
    CUDA_TRY(c, cudaGetLastError());
    memcpy(out + r0, b.h_rows, n * sizeof(ngsld_pair_row));
    c->stats.n_pairs += n;
  }
  return NGSLD_OK;
}

int ngsld_site_seeds(uint64_t seed, uint64_t n_sites, uint64_t *out) {
  if (!out && n_sites) return NGSLD_E_INVALID;
  hostprep::site_seeds(seed, n_sites, out);
  return NGSLD_OK;
}

int ngsld_tsv_header(int extend_out, char *buf, size_t cap) {  // reference ngsLD.cpp:77
  const char *base = "site1\tsite2\tdist\tr2_ExpG\tD\tDp\tr2";
  const char *ext = "\tsample_size\tmaf1\tmaf2\thap00\thap01\thap10\thap11\thap_maf1\thap_maf2\tchi2\tloglike\tnIter";
  int n = snprintf(buf, cap, "%s%s\n", base, extend_out ? ext : "");
  return (n < 0 || (size_t)n >= cap) ? NGSLD_E_INVALID : n;
}

int ngsld_probe_fp64(ngsld_ctx *c, double *gflops) {
  if (!c || !gflops) return NGSLD_E_INVALID;
  CUDA_TRY(c, cudaSetDevice(c->device));
  const int blocks = c->sm_count * 8, threads = 256, iters = 20000;
  double *d = nullptr;
  CUDA_TRY(c, cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  aux::fp64_probe_kernel<<<blocks, threads, 0, c->s_main>>>(d, 1000);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0, c->s_main);
    aux::fp64_probe_kernel<<<blocks, threads, 0, c->s_main>>>(d, iters);
    cudaEventRecord(e1, c->s_main);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  CUDA_TRY(c, cudaGetLastError());
  *gflops = (double)blocks * threads * iters * 8.0 * 2.0 / (best * 1e-3) / 1e9;
  return NGSLD_OK;
}

}  // extern "C"
