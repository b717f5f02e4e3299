/* ngsld_b200.h — C ABI of the B200-native pairwise-LD engine (libngsld_b200.so).
 *
 * Drop-in boundary for the pairwise-LD hot path of fgvieira/ngsLD 1.2.1.  The reference has no
 * plugin/FFI layer; the path sits behind the per-first-site task `void calc_pair_LD(void*)`
 * (reference ngsLD.hpp:58, body ngsLD.cpp:229-359) that `main` hands to the pthread pool with
 * `threadpool_add` (reference shared/threadpool.h:139-140, call site ngsLD.cpp:169), and behind the
 * numeric helpers that task calls (reference shared/gen_func.hpp:99-102, ngsLD.hpp:59).  Each entry
 * point below names the reference interface it replaces.  INTEGRATION.md shows the few lines a
 * maintainer of the reference would add to ngsLD.cpp to call this library instead of the pool.
 *
 * Conventions: plain pointers and sizes only; the caller owns every host buffer; device memory is
 * owned by the context; every function returns 0 on success or a negative NGSLD_E_* code, with a
 * message available from ngsld_last_error(); no C++ exception crosses this boundary; one host
 * thread per context at a time; contexts on different GPUs are independent (multi-GPU = one context
 * per device, first-site ranges from ngsld_partition()).  There is NO CPU fallback: without a CUDA
 * device ngsld_create() fails with NGSLD_E_CUDA.
 */
#ifndef NGSLD_B200_H
#define NGSLD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGSLD_ABI_VERSION 2

/* error codes (the reference instead prints "ERROR: [func] msg" and exit(-1), shared/gen_func.cpp:12-18) */
#define NGSLD_OK 0
#define NGSLD_E_INVALID (-1) /* bad argument / call order                                  */
#define NGSLD_E_CUDA (-2)    /* no device, or a CUDA runtime call failed                   */
#define NGSLD_E_NOMEM (-3)   /* host or device allocation failed                           */
#define NGSLD_E_DATA (-4)    /* input data rejected (NaN genotype, maf outside [0,1], bad positions) */
#define NGSLD_E_SINK (-5)    /* the caller's sink returned non-zero                        */
#define NGSLD_E_IO (-6)      /* file could not be read / parsed                            */

typedef struct ngsld_ctx ngsld_ctx;

/* One output row = one site pair.  Field order follows the reference's TSV columns
 * (ngsLD.cpp:314-351): dist r2_ExpG D Dp r2 | sample_size maf1* maf2* hap00 hap01 hap10 hap11
 * hap_maf1 hap_maf2 chi2 loglike* nIter   (* = not stored: maf1/maf2 are maf[s1]/maf[s2] and
 * loglike is the literal 0.0, ngsLD.cpp:347). */
typedef struct {
  double dist;       /* accumulated pos_dist over (s1, s2], ngsLD.cpp:241; +inf across chromosomes */
  double r2_expg;    /* pearson_r(), ngsLD.cpp:290,365-367 — bit-exact (x87 recurrence emulated)   */
  double D, Dp, r2;  /* ngsLD.cpp:300,304,306                                                       */
  double hap[4];     /* haplo_freq() output, shared/gen_func.cpp:1027-1059                          */
  double hap_maf[2]; /* ngsLD.cpp:297-298                                                           */
  float chi2;        /* float arithmetic as in ngsLD.cpp:328-333                                    */
  uint32_t n_iter;   /* haplo_freq() return value: 0-based converging pass, 100 = not converged     */
  uint32_t n_used;   /* individuals used by the EM (sample_size column)                             */
  uint32_t s1, s2;   /* site indices                                                                */
  uint32_t reserved;
} ngsld_pair_row;    /* 112 bytes */

/* The fields of the reference's `params` (ngsLD.hpp:11-44) that shape the pair scan, with the
 * reference's defaults (parse_args.cpp:6-29) filled in by ngsld_scan_defaults(). */
typedef struct {
  uint64_t max_kb_dist;   /* 0 = unlimited; break when max_kb_dist*1000 < dist (ngsLD.cpp:252)   */
  uint64_t max_snp_dist;  /* 0 = unlimited; break when max_snp_dist < s2-s1     (ngsLD.cpp:258)   */
  double min_maf;         /* ngsLD.cpp:264,270                                                    */
  double rnd_sample;      /* keep a candidate pair iff !(u > rnd_sample), ngsLD.cpp:277           */
  uint64_t seed;          /* master gsl_rng_taus seed, ngsLD.cpp:70                               */
  int ignore_miss_data;   /* shared/gen_func.cpp:1089                                             */
  int extend_out;         /* only affects TSV text (ngsld_scan_tsv)                               */
  int strict;             /* 1: EM in the reference's exact operation order (bit-identical hap/D/D'/r2,
                             slower); 0: fast kernel, |delta| <= 1e-9 at equal nIter               */
  int reserved;
} ngsld_scan_params;

/* Timing and work counters of the last scan on this context (device times from CUDA events on the
 * context's own streams). */
typedef struct {
  uint64_t n_pairs;        /* rows produced                                           */
  uint64_t sum_em_passes;  /* total EM passes executed (nIter+1, capped at 100)        */
  uint64_t n_launches;     /* kernels launched by this library during the scan        */
  double ms_em;            /* sum of EM-kernel durations                              */
  double ms_pearson;       /* sum of r2_ExpG-kernel durations                         */
  double ms_format;        /* sum of TSV-formatter durations                          */
  double ms_device_total;  /* first launch -> last result byte in host memory          */
  double ms_plan;          /* host planning (windows, sampling)                       */
  uint64_t h2d_bytes, d2h_bytes;
  char em_kernel[64];      /* EM kernel family the scan used, e.g. "emcell::em_cell_kernel<R=6,FUSE=1>" */
  /* class-compressed EM (0 when another kernel family ran): pairs it computed, sum over them of distinct
   * (p, q) combinations ("cells"), sum of cells x passes, and pairs it left to the dense kernel */
  uint64_t n_cell_pairs, sum_cells, sum_cell_passes, n_resid_pairs;
} ngsld_scan_stats;

/* sinks: called on the scanning host thread, rows in (s1, s2) order — the order the reference
 * produces with --n_threads 1 (FIFO pool, shared/threadpool.c:283-286).  Return non-zero to abort. */
typedef int (*ngsld_row_sink)(void *user, const ngsld_pair_row *rows, uint64_t n_rows);
typedef int (*ngsld_text_sink)(void *user, const char *bytes, uint64_t n_bytes, uint64_t n_rows);

/* ---- context --------------------------------------------------------------------------------- */
int ngsld_abi_version(void);
/* number of usable sm_100 devices (0 if none / no driver); the reference sizes its pool from --n_threads instead. */
int ngsld_device_count(void);
/* replaces threadpool_create() (shared/threadpool.c:51-99, call site ngsLD.cpp:154): binds a GPU. */
int ngsld_create(ngsld_ctx **out, int device);
/* replaces threadpool_destroy() + the free_ptr block (ngsLD.cpp:197-216). */
void ngsld_destroy(ngsld_ctx *ctx);
/* message for the last failure on ctx (ctx == NULL: last ngsld_create failure of this thread). */
const char *ngsld_last_error(const ngsld_ctx *ctx);
/* Optional: run all device work of this context on a caller-provided cudaStream_t (e.g. the
 * framework's current stream) instead of the context's own; NULL restores the default. */
int ngsld_set_stream(ngsld_ctx *ctx, void *cuda_stream);
/* cap on rows per device chunk (0 = default); result buffers scale with it. */
int ngsld_set_chunk_rows(ngsld_ctx *ctx, uint64_t rows);

/* ---- input files (host) --------------------------------------------------------------------------
 * replaces read_geno() (shared/read_data.cpp:13-116): binary = raw little-endian doubles [n_sites][n_ind][3]
 * (plain or gz); text = one site per line, blank/tab separated, non-numeric tokens dropped, a short first line is a
 * header, the LAST n_ind*(probs?3:1) numeric fields are used; genotypes coded -1/0/1/2 when !probs.  cells receives
 * [n_sites][n_ind][3]; *log_cells = 1 when they are already log-space (text input) — pass it on to
 * ngsld_prepare_sites(from_log_cells).  Failure: NGSLD_E_IO / NGSLD_E_DATA, and ngsld_last_error(NULL) holds
 * "[func] message" with the reference's function name and wording. */
int ngsld_load_geno(const char *path, int is_bin, int probs, int log_scale, uint64_t n_ind, uint64_t n_sites,
                    double *cells, int *log_cells);
/* replaces read_dist() + the label fix-up (shared/read_data.cpp:165-218, ngsLD.cpp:119-132): pos_dist[n_sites]
 * (+inf at a chromosome change) and the labels ("chr:pos...", only the first tab replaced) as one malloc'ed blob of
 * n_sites NUL-terminated strings, released with ngsld_free(). */
int ngsld_load_positions(const char *path, int header, uint64_t n_sites, double *pos_dist, char **label_blob,
                         uint64_t *blob_bytes);
void ngsld_free(void *p);

/* ---- per-site preparation (host, bit-identical to the reference's glibc path) ---------------- */
/* replaces the per-cell math of read_geno()'s binary branch (shared/read_data.cpp:28-46), the optional
 * call_geno() pass (ngsLD.cpp:92-98), est_maf() (ngsLD.cpp:103-104, shared/gen_func.cpp:974-1009) and
 * the conv_space/expected-genotype loop (ngsLD.cpp:107-114).  raw = the file's doubles
 * [n_sites][n_ind][3]; outputs gl [n_sites][n_ind][3] (normal space), expg [n_sites][n_ind], maf
 * [n_sites].  from_log_cells = 1: `raw` already holds log-space cells (text input path).  gl may be the same
 * buffer as raw (every cell is read before it is written): one genotype matrix in host memory, as in the reference. */
int ngsld_prepare_sites(const double *raw, uint64_t n_sites, uint64_t n_ind, int log_scale, int from_log_cells,
                        int ignore_miss_data, int call_geno, double N_thresh, double call_thresh, int n_threads,
                        double *gl, double *expg, double *maf);

/* ---- data upload ----------------------------------------------------------------------------- */
/* replaces the shared read-only arrays of `params` (geno_lkl, expected_geno, maf; ngsLD.hpp:36-41):
 * copies them to the device and derives the per-site x87 Pearson terms there (aux::site_terms_kernel).  Device buffers
 * are kept when the shape is unchanged, so calling it once per batch costs only the copies. */
int ngsld_set_sites(ngsld_ctx *ctx, const double *gl, const double *expg, const double *maf, uint64_t n_sites,
                    uint64_t n_ind);
/* Opt-in device-side preparation: takes the FILE's cells (as ngsld_load_geno returns them) and does
 * what ngsld_prepare_sites + ngsld_set_sites do, with the per-cell math (log / normalise / call_geno / est_maf / exp /
 * expected genotype; reference read_data.cpp:28-46, gen_func.cpp:886-1009, ngsLD.cpp:107-114) in a kernel on the
 * uploaded cells.  CUDA's log/exp are not glibc's, so likelihoods and allele frequencies differ from the host path in
 * the last bits and outputs are no longer bit-identical to the reference (still far inside the 1e-9 contract): the
 * host path stays the default.  maf_out (may be NULL) receives the allele frequencies. */
int ngsld_set_sites_raw(ngsld_ctx *ctx, const double *raw, uint64_t n_sites, uint64_t n_ind, int log_scale,
                        int from_log_cells, int ignore_miss_data, int call_geno, double N_thresh, double call_thresh,
                        double *maf_out);
/* replaces params.pos_dist / params.labels (ngsLD.cpp:119-135).  pos_dist NULL = all +inf (no --pos);
 * labels NULL = "(null)" like the reference prints.  labels are only used by ngsld_scan_tsv. */
int ngsld_set_positions(ngsld_ctx *ctx, const double *pos_dist, const char *const *labels);

/* Multi-GPU start-up: give dst everything ngsld_set_sites + ngsld_set_positions put on src's GPU by device-to-device
 * copies (NVLink / NVSwitch when the devices are peers) instead of another upload through the host: the reference's
 * threads share one `params`; here one GPU receives it from the host and passes it on.  Both contexts must belong to
 * the calling process; src must not be scanning meanwhile. */
int ngsld_share_sites(ngsld_ctx *dst, const ngsld_ctx *src);

/* ---- the scan: replaces the threadpool_add(calc_pair_LD) fan-out + threadpool_wait ------------ */
void ngsld_scan_defaults(ngsld_scan_params *p);
/* Planning only: number of rows first sites [s1_lo, s1_hi) will produce. */
int ngsld_scan_count(ngsld_ctx *ctx, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p, uint64_t *n_rows);
/* Equal-row-count first-site ranges for n_parts workers: bounds[0..n_parts], bounds[0]=0, bounds[n_parts]=n_sites. */
int ngsld_partition(ngsld_ctx *ctx, const ngsld_scan_params *p, int n_parts, uint64_t *bounds);
/* The same two planners without a device (host only; with rnd_sample < 1 they walk every candidate pair's draw, so
 * use the context versions for large sampled scans): what a multi-process launcher calls to hand each rank its range. */
int ngsld_plan_count(const double *maf, const double *pos_dist /* NULL = no positions */, uint64_t n_sites,
                     const ngsld_scan_params *p, uint64_t s1_lo, uint64_t s1_hi, uint64_t *n_rows);
int ngsld_plan_partition(const double *maf, const double *pos_dist, uint64_t n_sites, const ngsld_scan_params *p,
                         int n_parts, uint64_t *bounds);
/* Binary rows to a sink. */
int ngsld_scan(ngsld_ctx *ctx, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p, ngsld_row_sink sink,
               void *user);
/* Binary rows into one caller buffer of `cap` rows (fails with NGSLD_E_INVALID if it does not fit). */
int ngsld_scan_into(ngsld_ctx *ctx, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p,
                    ngsld_pair_row *out, uint64_t cap, uint64_t *n_rows);
/* TSV bytes exactly as the reference's fprintf block (ngsLD.cpp:314-351), formatted on the device. */
int ngsld_scan_tsv(ngsld_ctx *ctx, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p, ngsld_text_sink sink,
                   void *user);
/* The same text into ONE caller buffer, no intermediate copy: every device chunk is copied from HBM straight to
 * buf + running offset (true DMA when buf comes from ngsld_alloc_host).  ngsld_tsv_row_bound() * rows is always enough
 * room unless a value needs the host formatter (|x| >= 1e9); a scan that does not fit fails with NGSLD_E_INVALID. */
int ngsld_scan_tsv_into(ngsld_ctx *ctx, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p, char *buf, uint64_t cap,
                        uint64_t *n_bytes, uint64_t *n_rows);
/* upper bound of the bytes of one TSV row the device formatter writes (depends on the longest label). */
uint64_t ngsld_tsv_row_bound(const ngsld_ctx *ctx, int extend_out);
/* the same bound before any context exists, from the length of the longest site label (6 = "(null)" without labels). */
uint64_t ngsld_tsv_row_bound_for(uint32_t max_label_len, int extend_out);
/* page-locked host memory usable from every device (cudaHostAlloc, portable), for result buffers that the GPUs fill
 * by DMA and a writer thread hands to write(2) as they are. */
int ngsld_alloc_host(void **p, size_t bytes);
void ngsld_free_host(void *p);
/* Same work with results left in device memory (no D2H): the HBM-resident timing leg of bench.py. */
int ngsld_scan_device(ngsld_ctx *ctx, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p);
int ngsld_get_stats(const ngsld_ctx *ctx, ngsld_scan_stats *out);

/* ---- fused consumer: LD decay bins ------------------------------------------------
 * The binning step of the reference's downstream scripts/fit_LDdecay.R (lines 133-150) done on the device, so
 * that the pair table never leaves it: pairs with an infinite distance are dropped (line 133: dist < max_kb_dist*1000,
 * default Inf), dist is cut into right-closed bins (k*bin_size, (k+1)*bin_size] (line 145, R's cut()), and each of
 * the four LD statistics is summed per bin over its finite values (lines 138, 149: Inf -> NA, mean(na.rm=TRUE)).
 * mean = sum / n.  Sums are accumulated with floating-point atomics: reproducible to ~1e-12 relative, not bitwise. */
typedef struct {
  uint64_t n[4];  /* finite values of r2_ExpG, D, Dp, r2 that fell into the bin */
  double sum[4];  /* their sums */
} ngsld_decay_bin;
/* bins[n_bins] is overwritten; *n_outside (may be NULL) = rows dropped (infinite distance or beyond the last bin). */
int ngsld_scan_decay(ngsld_ctx *ctx, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p, double bin_size,
                     uint64_t n_bins, ngsld_decay_bin *bins, uint64_t *n_outside);

/* ---- fused consumer: LD pruning ----------------------------------------------------
 * The reference's downstream scripts/prune_graph.pl (and prune_ngsLD.py) read the whole TSV back, keep the rows with
 * dist <= max_dist and weight >= min_weight as edges of a graph (weight = column --field_weight, default 7 = r2; label =
 * int(weight * 10^4)), and greedily delete the "heaviest" node until no edge is left.  Here the filter runs on the
 * device right behind the EM, only the surviving edges (a tiny fraction of the pair table) come back to the host, and
 * ngsld_prune_graph() performs the greedy deletion.  The weight is taken as the scripts see it: the value rounded to the
 * six decimals the TSV prints, parsed back to double, times 10^precision, truncated. */
typedef struct {
  double max_dist;       /* bp; rows with dist > max_dist are no edges (prune_graph.pl: --max_kb_dist * 1000)        */
  double min_weight;     /* rows with weight < min_weight are no edges                                               */
  int field;             /* TSV column of the weight: 4 r2_ExpG, 5 D, 6 Dp, 7 r2 (the scripts' --field_weight [7])     */
  int weight_type;       /* 'a' absolute weight [default], 'e' signed weight, 'n' number of connections             */
  int weight_precision;  /* decimal digits of the integer edge label [4]                                            */
  int reserved;
} ngsld_prune_params;
typedef struct {
  uint32_t s1, s2;
  int32_t label;         /* int(weight * 10^weight_precision), as the scripts compute it                           */
} ngsld_edge;
typedef int (*ngsld_edge_sink)(void *user, const ngsld_edge *edges, uint64_t n_edges);
/* Scan first sites [s1_lo, s1_hi) and deliver only the rows that are edges.  seen (may be NULL) is a host array
 * [n_sites]: seen[s] is set to 1 for every site that occurs in a row at all (the scripts add both nodes of every line). */
int ngsld_scan_edges(ngsld_ctx *ctx, uint64_t s1_lo, uint64_t s1_hi, const ngsld_scan_params *p,
                     const ngsld_prune_params *q, ngsld_edge_sink sink, void *user, uint8_t *seen);
/* prune_graph.pl's prune_graph_idx on an edge list (host only, no device needed).  labels (may be NULL: site order) break
 * ties between equally heavy nodes in case-insensitive string order, as the script does.  kept[n_sites]: 1 = in the
 * pruned set, 0 = excluded, 2 = site occurs in no row; excluded[0 .. *n_excluded) = excluded sites in order of removal
 * (capacity n_sites; may be NULL). */
int ngsld_prune_graph(uint64_t n_sites, const char *const *labels, const uint8_t *seen, const ngsld_edge *edges,
                      uint64_t n_edges, int keep_heavy, uint8_t *kept, uint32_t *excluded, uint64_t *n_excluded);

/* ---- explicit pairs: replaces direct calls of haplo_freq()/pearson_r() (gen_func.hpp:101, ngsLD.hpp:59) */
int ngsld_pairs(ngsld_ctx *ctx, const uint32_t *s1, const uint32_t *s2, uint64_t n_pairs, int ignore_miss_data,
                int strict, ngsld_pair_row *out);

/* ---- pieces exposed for tests / callers that plan themselves --------------------------------- */
/* replaces the per-site generator seeding loop (ngsLD.cpp:165-166). */
int ngsld_site_seeds(uint64_t seed, uint64_t n_sites, uint64_t *out);
/* header line (ngsLD.cpp:77); returns bytes written (excluding NUL) or negative. */
int ngsld_tsv_header(int extend_out, char *buf, size_t cap);
/* FP64 FMA issue-rate probe (GFLOP/s, 2 flop per FMA) — the roofline that actually binds this path. */
int ngsld_probe_fp64(ngsld_ctx *ctx, double *gflops);

#ifdef __cplusplus
}
#endif
#endif /* NGSLD_B200_H */
