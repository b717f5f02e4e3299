// ngsLD-compatible command line on top of libngsld_b200.so.
//
// Same flags, defaults, implications and validation messages as the reference CLI (parse_args.cpp:6-29,35-59,
// 63-132,168-183), same header and TSV bytes on --out / stdout (ngsLD.cpp:77,314-351).  The thread-pool fan-out of
// the reference's main (ngsLD.cpp:153-198) becomes: one ngsld context per GPU (the site table is uploaded once and
// passed on GPU to GPU), the first-site axis split into equal-pair-count slabs (ngsld_partition), one host thread per
// GPU taking slabs in order and letting the device write each slab's text straight into a page-locked buffer
// (ngsld_scan_tsv_into), and writer threads that pwrite() finished slabs at their final offsets -- the row order the
// reference produces with --n_threads 1 -- while the GPUs work on the next ones.  The mutexed fprintf of
// ngsLD.cpp:310-352 has no counterpart: text is produced on the device and copied once, by the kernel's write().
//
// Extra flags (distinct prefix, reference command lines stay valid):
//   --gpu_n INT       GPUs to use (default: all visible)
//   --gpu_strict      bit-faithful EM kernel (hap/D/D'/r2 bit-identical to the reference; slower)
//   --gpu_stats       print pairs, EM passes and device times per GPU to stderr
//   --gpu_out_bin     --out receives the rows as binary records instead of TSV text: back-to-back 112-byte little-endian
//                     ngsld_pair_row structs (include/ngsld_b200.h; numpy: ngsld_b200.ROW_DTYPE), no header, same row order
//   --gpu_out_shards  write one file per slab of first sites, <out>.part-000000, <out>.part-000001, ... (the header is in the
//                     first): `cat <out>.part-*` is the output.  Writes to ONE file are serialised by the kernel (one inode
//                     lock: ~3.4 GB/s into the page cache however many threads write); separate files are written in parallel
//   --gpu_prune FILE  LD pruning fused behind the scan (what scripts/prune_graph.pl does with the TSV): no TSV is written;
//                     FILE receives the labels of the unlinked sites that remain, one per line, in site order.  Edge filter
//                     and options as in the script: --gpu_prune_max_kb_dist KB [inf], --gpu_prune_min_weight W [0],
//                     --gpu_prune_field 4|5|6|7 [7 = r2], --gpu_prune_weight_type a|e|n [a], --gpu_prune_keep_heavy,
//                     --gpu_prune_excl FILE (excluded sites in order of removal)
//   --gpu_prep        per-site preparation (log / normalise / maf / expected genotypes) on the GPU instead of the host:
//                     faster on very large inputs, last-bit differences from the reference (CUDA log/exp are not glibc's)
#include <errno.h>
#include <fcntl.h>
#include <getopt.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ngsld_b200.h"

static const char *kVersion = "1.2.1-b200";

static double wall_s() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// fatal error in the reference's format (shared/gen_func.cpp:12-18): message, perror, exit(-1)
[[noreturn]] static void die(const char *func, const char *msg) {
  fflush(stdout);
  fprintf(stderr, "\n=====\nERROR: [%s] %s\n=====\n\n", func, msg);
  perror("\t");
  fflush(stderr);
  exit(-1);
}
// same, for a library failure whose message already reads "[func] msg"
[[noreturn]] static void die_lib(const char *fallback_func) {
  const char *m = ngsld_last_error(NULL);
  fflush(stdout);
  if (m && m[0] == '[')
    fprintf(stderr, "\n=====\nERROR: %s\n=====\n\n", m);
  else
    fprintf(stderr, "\n=====\nERROR: [%s] %s\n=====\n\n", fallback_func, m ? m : "failure");
  perror("\t");
  fflush(stderr);
  exit(-1);
}

struct Options {
  const char *in_geno = nullptr;
  bool in_probs = false, in_logscale = false;
  uint64_t n_ind = 0, n_sites = 0;
  const char *in_pos = nullptr;
  bool in_pos_header = false;
  uint64_t max_kb_dist = 100, max_snp_dist = 0;
  double min_maf = 0;
  bool ignore_miss_data = false, call_geno = false;
  double N_thresh = 0, call_thresh = 0, rnd_sample = 1;
  uint64_t seed = 0;
  bool extend_out = false;
  const char *out = nullptr;
  int n_threads = 1, verbose = 1;
  int gpu_n = 0;
  bool gpu_strict = false, gpu_stats = false, gpu_prep = false, out_bin = false, out_shards = false;
  const char *prune_out = nullptr, *prune_excl = nullptr;
  double prune_max_kb = INFINITY, prune_min_weight = 0;
  int prune_field = 7, prune_type = 'a';
  bool prune_keep_heavy = false;
};

static void parse(Options &o, int argc, char **argv) {
  o.seed = (uint64_t)(time(NULL) + rand() % 1000);  // parse_args.cpp:23
  static struct option table[] = {{"geno", required_argument, NULL, 'g'},
                                  {"probs", no_argument, NULL, 'p'},
                                  {"log_scale", no_argument, NULL, 'l'},
                                  {"n_ind", required_argument, NULL, 'n'},
                                  {"n_sites", required_argument, NULL, 's'},
                                  {"pos", required_argument, NULL, 'a'},
                                  {"posH", required_argument, NULL, 'A'},
                                  {"max_kb_dist", required_argument, NULL, 'd'},
                                  {"max_snp_dist", required_argument, NULL, 'D'},
                                  {"min_maf", required_argument, NULL, 'f'},
                                  {"ignore_miss_data", no_argument, NULL, 'm'},
                                  {"call_geno", no_argument, NULL, 'c'},
                                  {"N_thresh", required_argument, NULL, 'N'},
                                  {"call_thresh", required_argument, NULL, 'C'},
                                  {"rnd_sample", required_argument, NULL, 'r'},
                                  {"seed", required_argument, NULL, 'S'},
                                  {"extend_out", no_argument, NULL, 'x'},
                                  {"out", required_argument, NULL, 'o'},
                                  {"outH", required_argument, NULL, 'O'},  // in the reference's table without a case: exits
                                  {"n_threads", required_argument, NULL, 't'},
                                  {"verbose", required_argument, NULL, 'V'},
                                  {"gpu_n", required_argument, NULL, 1001},
                                  {"gpu_strict", no_argument, NULL, 1002},
                                  {"gpu_stats", no_argument, NULL, 1003},
                                  {"gpu_prep", no_argument, NULL, 1004},
                                  {"gpu_out_bin", no_argument, NULL, 1005},
                                  {"gpu_out_shards", no_argument, NULL, 1006},
                                  {"gpu_prune", required_argument, NULL, 1010},
                                  {"gpu_prune_max_kb_dist", required_argument, NULL, 1011},
                                  {"gpu_prune_min_weight", required_argument, NULL, 1012},
                                  {"gpu_prune_field", required_argument, NULL, 1013},
                                  {"gpu_prune_weight_type", required_argument, NULL, 1014},
                                  {"gpu_prune_keep_heavy", no_argument, NULL, 1015},
                                  {"gpu_prune_excl", required_argument, NULL, 1016},
                                  {0, 0, 0, 0}};
  int c;
  while ((c = getopt_long_only(argc, argv, "g:pln:s:Z:d:D:f:mcN:C:r:S:xo:t:V:", table, NULL)) != -1) switch (c) {
      case 'g': o.in_geno = optarg; break;
      case 'p': o.in_probs = true; break;
      case 'l': o.in_logscale = o.in_probs = true; break;
      case 'n': o.n_ind = (uint64_t)atoi(optarg); break;
      case 's': o.n_sites = (uint64_t)atoi(optarg); break;
      case 'a': o.in_pos = optarg; o.in_pos_header = false; break;
      case 'A': o.in_pos = optarg; o.in_pos_header = true; break;
      case 'd': o.max_kb_dist = (uint64_t)atoi(optarg); break;
      case 'D': o.max_snp_dist = (uint64_t)atoi(optarg); break;
      case 'f': o.min_maf = atof(optarg); break;
      case 'm': o.ignore_miss_data = true; break;
      case 'c': o.call_geno = true; break;
      case 'N': o.N_thresh = atof(optarg); o.call_geno = true; break;
      case 'C': o.call_thresh = atof(optarg); o.call_geno = true; break;
      case 'r': o.rnd_sample = atof(optarg); break;
      case 'S': o.seed = (uint64_t)atoi(optarg); break;
      case 'x': o.extend_out = true; break;
      case 'o': o.out = optarg; break;
      case 't': o.n_threads = atoi(optarg); break;
      case 'V': o.verbose = atoi(optarg); break;
      case 1001: o.gpu_n = atoi(optarg); break;
      case 1002: o.gpu_strict = true; break;
      case 1003: o.gpu_stats = true; break;
      case 1004: o.gpu_prep = true; break;
      case 1005: o.out_bin = true; break;
      case 1006: o.out_shards = true; break;
      case 1010: o.prune_out = optarg; break;
      case 1011: o.prune_max_kb = atof(optarg); break;
      case 1012: o.prune_min_weight = atof(optarg); break;
      case 1013: o.prune_field = atoi(optarg); break;
      case 1014: o.prune_type = optarg[0]; break;
      case 1015: o.prune_keep_heavy = true; break;
      case 1016: o.prune_excl = optarg; break;
      default: exit(-1);
    }
  if (o.verbose >= 1) {
    fprintf(stderr, "==> Input Arguments:\n");
    fprintf(stderr,
            "\tgeno: %s\n\tprobs: %s\n\tlog_scale: %s\n\tn_ind: %lu\n\tn_sites: %lu\n\tpos: %s (%s header)\n\tmax_kb_dist (kb): "
            "%lu\n\tmax_snp_dist: %lu\n\tmin_maf: %f\n\tignore_miss_data: %s\n\tcall_geno: %s\n\tN_thresh: %f\n\tcall_thresh: "
            "%f\n\trnd_sample: %f\n\tseed: %lu\n\textend_out: %s\n\tout: %s\n\tn_threads: %d\n\tverbose: %d\n\tversion: %s (%s @ "
            "%s)\n\n",
            o.in_geno, o.in_probs ? "true" : "false", o.in_logscale ? "true" : "false", o.n_ind, o.n_sites, o.in_pos,
            o.in_pos_header ? "WITH" : "WITHOUT", o.max_kb_dist, o.max_snp_dist, o.min_maf,
            o.ignore_miss_data ? "true" : "false", o.call_geno ? "true" : "false", o.N_thresh, o.call_thresh, o.rnd_sample,
            o.seed, o.extend_out ? "true" : "false", o.out, o.n_threads, o.verbose, kVersion, __DATE__, __TIME__);
  }
  const char *fn = "parse_cmd_args";
  if (!o.in_geno) die(fn, "genotype input file (--geno) missing!");
  if (o.n_ind == 0) die(fn, "number of individuals (--n_ind) missing!");
  if (o.n_sites == 0) die(fn, "number of sites (--n_sites) missing!");
  if (!o.in_pos && o.max_kb_dist > 0) die(fn, "position file necessary in order to filter by maximum distance!");
  if (o.min_maf < 0 || o.min_maf > 1) die(fn, "minimum allele frequency must be in [0,1]!");
  if (o.call_geno && !o.in_probs) die(fn, "can only call genotypes from likelihoods/probabilities!");
  if (o.rnd_sample <= 0 || o.rnd_sample > 1) die(fn, "proportion of comparisons to sample must be in ]0,1]!");
  if (o.n_threads < 1) die(fn, "number of threads cannot be less than 1!");
}

int main(int argc, char **argv) {
  Options o;
  parse(o, argc, argv);
  const char *fn = "main";

  struct stat st;
  if (stat(o.in_geno, &st) != 0) die(fn, "cannot check GENO file size!");
  const char *dot = strrchr(o.in_geno, '.');
  bool in_bin;
  if (dot && strcmp(dot, ".gz") == 0) {
    if (o.verbose >= 1) fprintf(stderr, "==> GZIP input file (not BINARY)\n");
    in_bin = false;
  } else {
    if (o.verbose >= 1) fprintf(stderr, "==> BINARY input file (always lkl)\n");
    in_bin = true;
    o.in_probs = true;
    if (o.n_sites != (uint64_t)st.st_size / sizeof(double) / o.n_ind / 3) die(fn, "invalid/corrupt genotype input file!");
  }
  // Output: the reference fopen()s --out (or uses stdout) and fprintf()s under a mutex (ngsLD.cpp:73-77, 310-352).
  // Here rows arrive as finished text in page-locked slab buffers, so the file is a plain descriptor: a regular file is
  // written with pwrite() by several writer threads at offsets that are known as soon as all earlier slabs have been
  // formatted; a pipe / terminal / device gets the slabs in order from one writer.
  const bool prune_mode = o.prune_out != nullptr;
  const bool shards = o.out_shards && o.out && !prune_mode;
  int out_fd = STDOUT_FILENO;
  if (o.out && !prune_mode && !shards) out_fd = open(o.out, O_WRONLY | O_CREAT | O_TRUNC, 0666);
  if (out_fd < 0) die(fn, "cannot open output file!");
  struct stat ost;
  const bool seekable = fstat(out_fd, &ost) == 0 && S_ISREG(ost.st_mode);
  auto write_all = [&](int fd, const char *p, size_t n, off_t off) -> bool {  // off < 0: sequential write()
    while (n) {
      const ssize_t w = off >= 0 ? pwrite(fd, p, n, off) : write(fd, p, n);
      if (w < 0) {
        if (errno == EINTR) continue;
        return false;
      }
      p += w;
      n -= (size_t)w;
      if (off >= 0) off += w;
    }
    return true;
  };
  char header[512];
  const int hl_text = ngsld_tsv_header(o.extend_out, header, sizeof header);
  const int hl = o.out_bin ? 0 : hl_text;
  if (!prune_mode && !shards && hl && !write_all(out_fd, header, (size_t)hl, seekable ? 0 : -1)) die(fn, "cannot write output!");

  const double t_start = wall_s();
  // ---- in the background from the first moment: CUDA start-up, one context per GPU (seconds on an 8-GPU box) ----
  int n_dev = 0, n_gpu = 0;
  std::vector<ngsld_ctx *> ctx;
  std::vector<int> create_rc;
  std::mutex mu;                 // guards everything the GPU, writer and allocator threads share
  std::condition_variable cv;
  bool devices_known = false;
  std::thread boot([&]() {
    const int nd = ngsld_device_count();
    const int ng = nd < 1 ? 0 : (o.gpu_n > 0 ? std::min(o.gpu_n, nd) : nd);
    {
      std::lock_guard<std::mutex> lk(mu);
      n_dev = nd;
      n_gpu = ng;
      ctx.assign(ng, nullptr);
      create_rc.assign(ng, 0);
      devices_known = true;
    }
    cv.notify_all();
    std::vector<std::thread> th;
    for (int g = 0; g < ng; g++) th.emplace_back([&, g]() { create_rc[g] = ngsld_create(&ctx[g], g); });
    for (auto &t : th) t.join();
  });
  // positions first (small file): the longest label bounds the length of a TSV row, which sizes the slab buffers, and those
  // are page-locked in the background while the genotypes are read.  A failure is reported where the reference reports it.
  std::vector<double> pos_dist;
  std::vector<const char *> label_ptr;
  char *label_blob = nullptr;
  std::string pos_error;
  uint32_t max_label_len = 6;  // "(null)"
  if (o.in_pos) {
    pos_dist.resize(o.n_sites);
    if (ngsld_load_positions(o.in_pos, o.in_pos_header, o.n_sites, pos_dist.data(), &label_blob, NULL) != NGSLD_OK) {
      pos_error = ngsld_last_error(NULL);
    } else {
      label_ptr.resize(o.n_sites);
      const char *p = label_blob;
      max_label_len = 1;
      for (uint64_t s = 0; s < o.n_sites; s++) {
        label_ptr[s] = p;
        const size_t len = strlen(p);
        max_label_len = std::max<uint32_t>(max_label_len, (uint32_t)len);
        p += len + 1;
      }
    }
  }
  {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&]() { return devices_known; });
  }
  // ---- slab buffers (see below): sized now, page-locked in the background ----
  const uint64_t row_bound = o.out_bin ? sizeof(ngsld_pair_row) : ngsld_tsv_row_bound_for(max_label_len, o.extend_out);
  uint64_t buf_bytes = (n_gpu > 2 ? 256ull : 512ull) << 20;  // page-locking costs ~0.4 s per GB, and again at exit
  if (const char *e = getenv("NGSLD_CLI_BUF_MB"))
    if (atoll(e) > 0) buf_bytes = (uint64_t)atoll(e) << 20;
  uint64_t rows_per_slab = std::max<uint64_t>(1, std::min<uint64_t>(16ull << 20, buf_bytes / row_bound));
  if (const char *e = getenv("NGSLD_CLI_SLAB_ROWS"))  // tests: force many small slabs
    if (atoll(e) > 0) rows_per_slab = (uint64_t)atoll(e);
  // a slab holds its share of the rows (at most rows_per_slab) plus at most the rows of one first site
  const uint64_t cap = std::max<uint64_t>((rows_per_slab + o.n_sites + 1) * row_bound, 4096);
  // one pwrite() stream into the page cache moves ~3.5 GB/s; a GPU produces ~4 GB/s of text
  const int n_writers = shards ? std::max(2, std::min(2 * n_gpu, 16)) : seekable ? 2 : 1;
  const int n_bufs = n_gpu + n_writers + 1;
  struct Buf {
    char *p = nullptr;
    bool pinned = false;
  };
  std::vector<Buf> bufs(n_bufs);
  std::vector<int> free_bufs;
  bool alloc_failed = false, failed = false, write_failed = false;
  const double t_alloc0 = wall_s();
  double t_alloc = 0;
  std::vector<std::thread> allocators;
  if (!prune_mode && n_gpu > 0)
    for (int k = 0; k < n_bufs; k++)
      allocators.emplace_back([&, k]() {
        void *q = nullptr;
        Buf b;
        if (ngsld_alloc_host(&q, cap) == NGSLD_OK) {
          b.p = (char *)q;
          b.pinned = true;
        } else {
          b.p = (char *)malloc(cap);  // pageable: the copies are staged by the driver, everything else works the same
        }
        {
          std::lock_guard<std::mutex> lk(mu);
          bufs[k] = b;
          if (b.p) free_bufs.push_back(k);
          else alloc_failed = failed = true;
          t_alloc = wall_s() - t_alloc0;
        }
        cv.notify_all();
      });

  if (o.verbose >= 1) fprintf(stderr, "> Reading data from file...\n");
  std::vector<double> cells((size_t)o.n_sites * o.n_ind * 3);
  int log_cells = 0;
  if (ngsld_load_geno(o.in_geno, in_bin, o.in_probs, o.in_logscale, o.n_ind, o.n_sites, cells.data(), &log_cells) != NGSLD_OK)
    die_lib("read_geno");
  const double t_read = wall_s();
  if (o.verbose >= 1 && o.call_geno) fprintf(stderr, "> Calling genotypes...\n");
  if (o.verbose >= 1) fprintf(stderr, "==> Calculating MAF for all sites...\n");
  // prepared in place: the normalised likelihoods overwrite the file's cells (one genotype matrix in host memory, like
  // the reference, instead of two)
  std::vector<double> expg, maf(o.n_sites);
  const int host_threads = std::max(o.n_threads, (int)std::thread::hardware_concurrency());
  int rc = NGSLD_OK;
  if (o.call_geno && o.N_thresh > o.call_thresh)  // raised by call_geno() in the reference: after the read
    die("call_geno", "missing data threshold must be smaller than calling genotype threshold!");
  if (!o.gpu_prep) {
    expg.resize((size_t)o.n_sites * o.n_ind);
    rc = ngsld_prepare_sites(cells.data(), o.n_sites, o.n_ind, o.in_logscale, log_cells, o.ignore_miss_data, o.call_geno,
                             o.N_thresh, o.call_thresh, host_threads, cells.data(), expg.data(), maf.data());
  }
  if (rc == NGSLD_E_DATA) die("read_geno", "NaN found! Is the file format correct?");
  if (rc != NGSLD_OK) die(fn, "site preparation failed!");
  const std::vector<double> &gl = cells;
  const double t_prep = wall_s();

  if (o.verbose >= 1) fprintf(stderr, "==> Getting sites coordinates\n");
  if (!pos_error.empty()) {  // (read early, reported here)
    fflush(stdout);
    if (pos_error[0] == '[') fprintf(stderr, "\n=====\nERROR: %s\n=====\n\n", pos_error.c_str());
    else fprintf(stderr, "\n=====\nERROR: [read_dist] %s\n=====\n\n", pos_error.c_str());
    perror("\t");
    exit(-1);
  }

  boot.join();
  if (n_dev < 1) die(fn, "no CUDA device available (this build has no CPU path)!");
  if (o.verbose >= 1) fprintf(stderr, "==> Launching threads...\n");

  ngsld_scan_params P;
  ngsld_scan_defaults(&P);
  P.max_kb_dist = o.max_kb_dist;
  P.max_snp_dist = o.max_snp_dist;
  P.min_maf = o.min_maf;
  P.rnd_sample = o.rnd_sample;
  P.seed = o.seed;
  P.ignore_miss_data = o.ignore_miss_data;
  P.extend_out = o.extend_out;
  P.strict = o.gpu_strict;

  // ---- site table: one upload from the host, then passed on GPU to GPU (NVLink) in a doubling tree ----
  const bool host_upload_all = getenv("NGSLD_CLI_HOST_UPLOAD") && atoi(getenv("NGSLD_CLI_HOST_UPLOAD"));
  auto from_host = [&](int g) -> int {
    int r = create_rc[g];
    if (r) return r;
    if (o.gpu_prep)
      r = ngsld_set_sites_raw(ctx[g], cells.data(), o.n_sites, o.n_ind, o.in_logscale, log_cells, o.ignore_miss_data, o.call_geno,
                              o.N_thresh, o.call_thresh, NULL);
    else
      r = ngsld_set_sites(ctx[g], gl.data(), expg.data(), maf.data(), o.n_sites, o.n_ind);
    if (r) return r;
    return ngsld_set_positions(ctx[g], o.in_pos ? pos_dist.data() : nullptr, o.in_pos ? label_ptr.data() : nullptr);
  };
  auto check_setup = [&](const std::vector<int> &rcs) {
    for (int g = 0; g < n_gpu; g++)
      if (rcs[g]) {
        fprintf(stderr, "GPU %d: %s\n", g, ngsld_last_error(ctx[g]));
        if (rcs[g] == NGSLD_E_DATA && strstr(ngsld_last_error(ctx[g]), "NaN found")) die("read_geno", "NaN found! Is the file format correct?");
        die(fn, rcs[g] == NGSLD_E_DATA ? "invalid allele frequencies" : "failed to initialise the GPU engine!");
      }
  };
  {
    std::vector<int> rcs(n_gpu, 0);
    if (host_upload_all) {
      std::vector<std::thread> th;
      for (int g = 0; g < n_gpu; g++) th.emplace_back([&, g]() { rcs[g] = from_host(g); });
      for (auto &t : th) t.join();
    } else {
      rcs[0] = from_host(0);
      check_setup(rcs);
      for (int have = 1; have < n_gpu; have *= 2) {  // GPUs [0, have) hold the table: each passes it on to one more
        std::vector<std::thread> th;
        for (int g = have; g < std::min(n_gpu, 2 * have); g++)
          th.emplace_back([&, g, have]() {
            rcs[g] = create_rc[g];
            if (!rcs[g]) rcs[g] = ngsld_share_sites(ctx[g], ctx[g - have]);
          });
        for (auto &t : th) t.join();
        check_setup(rcs);
      }
    }
    check_setup(rcs);
  }
  const double t_upload = wall_s();

  if (prune_mode) {
    // ---- LD pruning: every GPU filters the rows of its slabs into edges on the device; the host prunes the graph ----
    ngsld_prune_params Q;
    memset(&Q, 0, sizeof Q);
    Q.max_dist = o.prune_max_kb * 1000.0;
    Q.min_weight = o.prune_min_weight;
    Q.field = o.prune_field;
    Q.weight_type = o.prune_type;
    Q.weight_precision = 4;
    const int n_parts = n_gpu * 8;
    std::vector<uint64_t> pb(n_parts + 1);
    if (ngsld_partition(ctx[0], &P, n_parts, pb.data()) != NGSLD_OK) die(fn, "failed to partition the pair space!");
    std::vector<std::vector<ngsld_edge>> part_edges(n_parts);
    std::vector<std::vector<uint8_t>> seen(n_gpu, std::vector<uint8_t>(o.n_sites, 0));
    std::vector<int> prc(n_gpu, 0);
    std::vector<uint64_t> rows_g(n_gpu, 0);
    std::mutex pm;
    int next_part = 0;
    auto edge_sink = [](void *user, const ngsld_edge *e, uint64_t n) -> int {
      auto *v = (std::vector<ngsld_edge> *)user;
      v->insert(v->end(), e, e + n);
      return 0;
    };
    const double tp0 = wall_s();
    {
      std::vector<std::thread> th;
      for (int g = 0; g < n_gpu; g++)
        th.emplace_back([&, g]() {
          for (;;) {
            int k;
            {
              std::lock_guard<std::mutex> lk(pm);
              if (next_part >= n_parts) return;
              k = next_part++;
            }
            prc[g] = ngsld_scan_edges(ctx[g], pb[k], pb[k + 1], &P, &Q, edge_sink, &part_edges[k], seen[g].data());
            if (prc[g]) return;
            ngsld_scan_stats st;
            ngsld_get_stats(ctx[g], &st);
            rows_g[g] += st.n_pairs;
          }
        });
      for (auto &t : th) t.join();
    }
    for (int g = 0; g < n_gpu; g++)
      if (prc[g]) {
        fprintf(stderr, "GPU %d: %s\n", g, ngsld_last_error(ctx[g]));
        die(fn, "pair scan failed!");
      }
    const double tp1 = wall_s();
    std::vector<ngsld_edge> edges;
    for (auto &v : part_edges) edges.insert(edges.end(), v.begin(), v.end());
    for (int g = 1; g < n_gpu; g++)
      for (uint64_t s = 0; s < o.n_sites; s++) seen[0][s] |= seen[g][s];
    std::vector<uint8_t> kept(o.n_sites);
    std::vector<uint32_t> excl(o.n_sites);
    uint64_t n_excl = 0;
    if (ngsld_prune_graph(o.n_sites, o.in_pos ? label_ptr.data() : nullptr, seen[0].data(), edges.data(), edges.size(),
                          o.prune_keep_heavy, kept.data(), excl.data(), &n_excl) != NGSLD_OK)
      die(fn, "graph pruning failed!");
    const double tp2 = wall_s();
    auto site_name = [&](uint64_t s, char *tmp) -> const char * {
      if (o.in_pos) return label_ptr[s];
      snprintf(tmp, 32, "%lu", (unsigned long)s);
      return tmp;
    };
    FILE *pf = strcmp(o.prune_out, "-") ? fopen(o.prune_out, "w") : stdout;
    if (!pf) die(fn, "cannot open output file!");
    uint64_t n_kept = 0, n_nodes = 0;
    char tmp[32];
    for (uint64_t s = 0; s < o.n_sites; s++) {
      if (kept[s] != 2) n_nodes++;
      if (kept[s] == 1) {
        fprintf(pf, "%s\n", site_name(s, tmp));
        n_kept++;
      }
    }
    if (pf != stdout) fclose(pf);
    if (o.prune_excl) {
      FILE *ef = fopen(o.prune_excl, "w");
      if (!ef) die(fn, "cannot open output file!");
      for (uint64_t k = 0; k < n_excl; k++) fprintf(ef, "%s\n", site_name(excl[k], tmp));
      fclose(ef);
    }
    if (o.gpu_stats) {
      uint64_t rows = 0;
      for (auto r : rows_g) rows += r;
      fprintf(stderr, "[prune] %lu rows scanned on %d GPU(s) in %.2f s, %lu edges (%.4f %% of the rows) between %lu sites came back to the host, pruning %.2f s: %lu sites kept, %lu excluded\n",
              rows, n_gpu, tp1 - tp0, edges.size(), rows ? 100.0 * edges.size() / rows : 0.0, n_nodes, tp2 - tp1, n_kept, n_excl);
    }
    for (auto c : ctx) ngsld_destroy(c);
    ngsld_free(label_blob);
    if (o.verbose >= 1) fprintf(stderr, "Done!\n");
    return 0;
  }

  // ---- work units: slabs of first sites with (about) the same number of rows, taken in order by the GPU threads ----
  uint64_t total_rows = 0;
  if (ngsld_scan_count(ctx[0], 0, o.n_sites, &P, &total_rows) != NGSLD_OK) {
    fprintf(stderr, "%s\n", ngsld_last_error(ctx[0]));
    die(fn, "failed to plan the pair scan!");
  }
  const int n_slabs = (int)std::min<uint64_t>(std::max<uint64_t>((total_rows + rows_per_slab - 1) / rows_per_slab, (uint64_t)n_gpu), 1u << 20);
  std::vector<uint64_t> bounds(n_slabs + 1);
  if (ngsld_partition(ctx[0], &P, n_slabs, bounds.data()) != NGSLD_OK) {
    fprintf(stderr, "%s\n", ngsld_last_error(ctx[0]));
    die(fn, "failed to partition the pair space!");
  }

  if (o.verbose >= 1) fprintf(stderr, "==> Waiting for all threads to finish...\n");
  struct Slab {
    int buf = -1;
    uint64_t bytes = 0, rows = 0;
    off_t offset = -1;          // known once every earlier slab has been formatted
    bool done = false;
    std::string spill;          // only when a slab outgrew its buffer (values the host formatter had to print)
  };
  std::vector<Slab> slab(n_slabs);
  int next_slab = 0, next_off = 0, written = 0;
  std::deque<int> ready;  // slabs that can be written now: formatted and (one file) with their offset known
  off_t cursor = shards ? 0 : hl;
  struct PerGpu {
    uint64_t pairs = 0, passes = 0, launches = 0, slabs = 0, cells = 0, cell_pairs = 0, resid = 0;
    double ms_device = 0, ms_em = 0, ms_pearson = 0, ms_format = 0, ms_plan = 0, s_scan = 0, s_wait = 0;
  };
  std::vector<PerGpu> acc(n_gpu);
  struct PerWriter {
    uint64_t bytes = 0;
    double s_write = 0;
  };
  std::vector<PerWriter> wacc(n_writers);

  std::vector<std::thread> writers;
  for (int w = 0; w < n_writers; w++)
    writers.emplace_back([&, w]() {
      for (;;) {
        int k = -1;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&]() { return failed || written == n_slabs || !ready.empty(); });
          if (failed || written == n_slabs) return;
          k = ready.front();
          ready.pop_front();
        }
        const double t0 = wall_s();
        const char *src = slab[k].spill.empty() ? bufs[slab[k].buf].p : slab[k].spill.data();
        bool ok;
        if (shards) {  // a file of its own: no offset to wait for, no lock shared with the other writers
          char path[4096];
          snprintf(path, sizeof path, "%s.part-%06d", o.out, k);
          const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0666);
          ok = fd >= 0 && (k != 0 || hl == 0 || write_all(fd, header, (size_t)hl, -1)) && write_all(fd, src, slab[k].bytes, -1);
          if (fd >= 0 && close(fd) != 0) ok = false;
        } else {
          ok = write_all(out_fd, src, slab[k].bytes, seekable ? slab[k].offset : -1);
        }
        wacc[w].s_write += wall_s() - t0;
        wacc[w].bytes += slab[k].bytes;
        {
          std::lock_guard<std::mutex> lk(mu);
          if (!ok) failed = write_failed = true;
          free_bufs.push_back(slab[k].buf);
          std::string().swap(slab[k].spill);
          written++;
        }
        cv.notify_all();
      }
    });

  auto append_sink = [](void *user, const char *bytes, uint64_t n_bytes, uint64_t) -> int {
    ((std::string *)user)->append(bytes, n_bytes);
    return 0;
  };
  std::vector<int> rcs(n_gpu, 0);
  const double t_scan0 = wall_s();
  {
    std::vector<std::thread> th;
    for (int g = 0; g < n_gpu; g++)
      th.emplace_back([&, g]() {
        for (;;) {
          int k, bi;
          const double tw0 = wall_s();
          {
            std::unique_lock<std::mutex> lk(mu);
            // slabs are claimed in order by whoever holds a buffer, or the writers could starve for the next slab
            cv.wait(lk, [&]() { return failed || next_slab >= n_slabs || !free_bufs.empty(); });
            if (failed || next_slab >= n_slabs) return;
            k = next_slab++;
            bi = free_bufs.back();
            free_bufs.pop_back();
          }
          const double ts0 = wall_s();
          acc[g].s_wait += ts0 - tw0;
          uint64_t nb = 0, nr = 0;
          std::string spill;
          int rc;
          if (o.out_bin) {
            rc = ngsld_scan_into(ctx[g], bounds[k], bounds[k + 1], &P, (ngsld_pair_row *)bufs[bi].p, cap / sizeof(ngsld_pair_row), &nr);
            nb = nr * sizeof(ngsld_pair_row);
          } else {
            rc = ngsld_scan_tsv_into(ctx[g], bounds[k], bounds[k + 1], &P, bufs[bi].p, cap, &nb, &nr);
          }
          ngsld_scan_stats st;
          ngsld_get_stats(ctx[g], &st);
          if (rc == NGSLD_E_INVALID && strstr(ngsld_last_error(ctx[g]), "too small")) {
            // rows longer than the device formatter's bound (host-formatted values): take this slab through the sink
            rc = ngsld_scan_tsv(ctx[g], bounds[k], bounds[k + 1], &P, append_sink, &spill);
            ngsld_get_stats(ctx[g], &st);
            nb = spill.size();
            nr = st.n_pairs;
          }
          acc[g].s_scan += wall_s() - ts0;
          acc[g].pairs += st.n_pairs; acc[g].passes += st.sum_em_passes; acc[g].launches += st.n_launches; acc[g].slabs++;
          acc[g].ms_device += st.ms_device_total; acc[g].ms_em += st.ms_em; acc[g].ms_pearson += st.ms_pearson;
          acc[g].ms_format += st.ms_format; acc[g].ms_plan += st.ms_plan; acc[g].cells += st.sum_cells; acc[g].cell_pairs += st.n_cell_pairs;
          acc[g].resid += st.n_resid_pairs;
          {
            std::lock_guard<std::mutex> lk(mu);
            if (rc) {
              rcs[g] = rc;
              failed = true;
            } else {
              slab[k].buf = bi;
              slab[k].bytes = nb;
              slab[k].rows = nr;
              slab[k].spill.swap(spill);
              slab[k].done = true;
              if (shards) {
                ready.push_back(k);
                cursor += (off_t)nb;
              }
              while (!shards && next_off < n_slabs && slab[next_off].done) {  // offsets follow from the sizes of all earlier slabs
                slab[next_off].offset = cursor;
                cursor += (off_t)slab[next_off].bytes;
                ready.push_back(next_off);
                next_off++;
              }
            }
          }
          cv.notify_all();
          if (rc) return;
        }
      });
    for (auto &t : th) t.join();
  }
  const double t_scan1 = wall_s();
  cv.notify_all();
  for (auto &t : writers) t.join();
  for (auto &t : allocators) t.join();
  if (alloc_failed) die(fn, "cannot allocate the output buffers!");
  const double t_done = wall_s();
  for (int g = 0; g < n_gpu; g++)
    if (rcs[g]) {
      fprintf(stderr, "GPU %d: %s\n", g, ngsld_last_error(ctx[g]));
      die(fn, "pair scan failed!");
    }
  if (write_failed) die(fn, "cannot write output!");
  if (o.gpu_stats) {
    uint64_t all_pairs = 0;
    for (int g = 0; g < n_gpu; g++) {
      all_pairs += acc[g].pairs;
      fprintf(stderr, "[gpu %d] %lu slabs of %d: %lu pairs, %lu EM passes, %lu launches, scanning %.2f s (planning %.1f ms, device %.1f ms: EM %.1f, r2_ExpG %.1f, format %.1f), waiting for a buffer %.2f s, %.0f pairs/s while scanning",
              g, acc[g].slabs, n_slabs, acc[g].pairs, acc[g].passes, acc[g].launches, acc[g].s_scan, acc[g].ms_plan, acc[g].ms_device, acc[g].ms_em,
              acc[g].ms_pearson, acc[g].ms_format, acc[g].s_wait, acc[g].s_scan > 0 ? acc[g].pairs / acc[g].s_scan : 0.0);
      if (acc[g].cell_pairs)
        fprintf(stderr, "; class-compressed EM: %.1f cells per pair, %lu pairs left to the dense kernel", (double)acc[g].cells / acc[g].cell_pairs, acc[g].resid);
      fprintf(stderr, "\n");
    }
    for (int w = 0; w < n_writers; w++)
      fprintf(stderr, "[writer %d] %.2f GB in %.2f s of %s = %.2f GB/s\n", w, wacc[w].bytes / 1e9, wacc[w].s_write,
              shards ? "write to slab files" : seekable ? "pwrite" : "write", wacc[w].s_write > 0 ? wacc[w].bytes / 1e9 / wacc[w].s_write : 0.0);
    fprintf(stderr, "[time] read %.2f s, prepare (host) %.2f s, positions + site table to %d GPU(s) %.2f s (%s), scan %.2f s, writers done %.2f s after the scan; %lu rows, %.2f GB of text, %.0f rows/s over scan + write, %d slab buffers of %.0f MB (%s, the last one ready %.2f s after the positions were read)\n",
            t_read - t_start, t_prep - t_read, n_gpu, t_upload - t_prep, host_upload_all || n_gpu == 1 ? "from the host" : "one upload, then GPU to GPU",
            t_scan1 - t_scan0, t_done - t_scan1, all_pairs, (double)(cursor - (shards ? 0 : hl)) / 1e9, all_pairs / std::max(t_done - t_scan0, 1e-9), n_bufs, cap / 1e6,
            bufs[0].pinned ? "page-locked" : "pageable", t_alloc);
  }
  if (o.verbose >= 1) fprintf(stderr, "==> Freeing memory...\n");
  for (auto c : ctx) ngsld_destroy(c);
  for (auto &b : bufs) {
    if (b.pinned) ngsld_free_host(b.p);
    else free(b.p);
  }
  ngsld_free(label_blob);
  if (out_fd != STDOUT_FILENO && close(out_fd) != 0) die(fn, "cannot write output!");
  if (o.verbose >= 1) fprintf(stderr, "Done!\n");
  return 0;
}
