// ngsLD-compatible command line on top of libngsld_b200.so.
//
// Same flags, defaults, implications and validation messages as the reference CLI (parse_args.cpp:6-29,35-59,
// 63-132,168-183), same header and TSV bytes on --out / stdout (ngsLD.cpp:77,314-351).  The thread-pool fan-out of
// the reference's main (ngsLD.cpp:153-198) becomes: one ngsld context per GPU, the first-site axis split into
// equal-pair-count slabs (ngsld_partition), one host thread per GPU taking slabs in order and running ngsld_scan_tsv,
// and a writer thread appending the finished slabs in slab order — the row order the reference produces with
// --n_threads 1 — while the GPUs work on the next ones.
//
// Extra flags (distinct prefix, reference command lines stay valid):
//   --gpu_n INT       GPUs to use (default: all visible)
//   --gpu_strict      bit-faithful EM kernel (hap/D/D'/r2 bit-identical to the reference; slower)
//   --gpu_stats       print pairs, EM passes and device times per GPU to stderr
#include <getopt.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ngsld_b200.h"

static const char *kVersion = "1.2.1-b200";

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
  bool gpu_strict = false, gpu_stats = false;
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
  if (o.call_geno && o.N_thresh > o.call_thresh)
    die("call_geno", "missing data threshold must be smaller than calling genotype threshold!");

  FILE *out_fh = stdout;
  if (o.out) out_fh = fopen(o.out, "w");
  if (!out_fh) die(fn, "cannot open output file!");
  char header[512];
  const int hl = ngsld_tsv_header(o.extend_out, header, sizeof header);
  fwrite(header, 1, hl, out_fh);

  if (o.verbose >= 1) fprintf(stderr, "> Reading data from file...\n");
  std::vector<double> raw((size_t)o.n_sites * o.n_ind * 3);
  int log_cells = 0;
  if (ngsld_load_geno(o.in_geno, in_bin, o.in_probs, o.in_logscale, o.n_ind, o.n_sites, raw.data(), &log_cells) != NGSLD_OK)
    die_lib("read_geno");
  if (o.verbose >= 1 && o.call_geno) fprintf(stderr, "> Calling genotypes...\n");
  if (o.verbose >= 1) fprintf(stderr, "==> Calculating MAF for all sites...\n");
  std::vector<double> gl(raw.size()), expg((size_t)o.n_sites * o.n_ind), maf(o.n_sites);
  const int host_threads = std::max(o.n_threads, (int)std::thread::hardware_concurrency());
  int rc = ngsld_prepare_sites(raw.data(), o.n_sites, o.n_ind, o.in_logscale, log_cells, o.ignore_miss_data,
                               o.call_geno, o.N_thresh, o.call_thresh, host_threads, gl.data(), expg.data(), maf.data());
  if (rc == NGSLD_E_DATA) die("read_geno", "NaN found! Is the file format correct?");
  if (rc != NGSLD_OK) die(fn, "site preparation failed!");
  std::vector<double>().swap(raw);

  if (o.verbose >= 1) fprintf(stderr, "==> Getting sites coordinates\n");
  std::vector<double> pos_dist;
  std::vector<const char *> label_ptr;
  char *label_blob = nullptr;
  if (o.in_pos) {
    pos_dist.resize(o.n_sites);
    if (ngsld_load_positions(o.in_pos, o.in_pos_header, o.n_sites, pos_dist.data(), &label_blob, NULL) != NGSLD_OK)
      die_lib("read_dist");
    label_ptr.resize(o.n_sites);
    const char *p = label_blob;
    for (uint64_t s = 0; s < o.n_sites; s++) {
      label_ptr[s] = p;
      p += strlen(p) + 1;
    }
  }

  const int n_dev = ngsld_device_count();
  if (n_dev < 1) die(fn, "no CUDA device available (this build has no CPU path)!");
  int n_gpu = o.gpu_n > 0 ? std::min(o.gpu_n, n_dev) : n_dev;
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

  std::vector<ngsld_ctx *> ctx(n_gpu, nullptr);
  auto setup = [&](int g) -> int {
    int r = ngsld_create(&ctx[g], g);
    if (r) return r;
    r = ngsld_set_sites(ctx[g], gl.data(), expg.data(), maf.data(), o.n_sites, o.n_ind);
    if (r) return r;
    return ngsld_set_positions(ctx[g], o.in_pos ? pos_dist.data() : nullptr, o.in_pos ? label_ptr.data() : nullptr);
  };
  {
    std::vector<std::thread> th;
    std::vector<int> rcs(n_gpu, 0);
    for (int g = 0; g < n_gpu; g++) th.emplace_back([&, g]() { rcs[g] = setup(g); });
    for (auto &t : th) t.join();
    for (int g = 0; g < n_gpu; g++)
      if (rcs[g]) {
        fprintf(stderr, "GPU %d: %s\n", g, ngsld_last_error(ctx[g]));
        die(fn, rcs[g] == NGSLD_E_DATA ? "invalid allele frequencies" : "failed to initialise the GPU engine!");
      }
  }
  // Work units: slabs of first sites with (about) the same number of rows.  GPUs take slabs in order from a shared
  // counter; a writer thread appends finished slabs to the output in slab order, which is first-site order = the
  // reference's --n_threads 1 row order.  The GPUs run at most a pool of buffers ahead of the writer, so the text held
  // in memory stays bounded and nothing is written twice.
  uint64_t total_rows = 0;
  if (ngsld_scan_count(ctx[0], 0, o.n_sites, &P, &total_rows) != NGSLD_OK) {
    fprintf(stderr, "%s\n", ngsld_last_error(ctx[0]));
    die(fn, "failed to plan the pair scan!");
  }
  uint64_t rows_per_slab = 16ull << 20;  // a scan drains its chunk pipeline at the end: keep slabs long
  if (const char *e = getenv("NGSLD_CLI_SLAB_ROWS"))  // tests: force many small slabs
    if (atoll(e) > 0) rows_per_slab = (uint64_t)atoll(e);
  const int n_slabs = (int)std::min<uint64_t>(std::max<uint64_t>((total_rows + rows_per_slab - 1) / rows_per_slab, (uint64_t)n_gpu), 1u << 16);
  std::vector<uint64_t> bounds(n_slabs + 1);
  if (ngsld_partition(ctx[0], &P, n_slabs, bounds.data()) != NGSLD_OK) {
    fprintf(stderr, "%s\n", ngsld_last_error(ctx[0]));
    die(fn, "failed to partition the pair space!");
  }

  if (o.verbose >= 1) fprintf(stderr, "==> Waiting for all threads to finish...\n");
  struct Slab {
    std::string text;
    bool done = false;
  };
  std::vector<Slab> slab(n_slabs);
  std::mutex mu;
  std::condition_variable cv;
  int next_slab = 0, written = 0;
  bool failed = false, write_failed = false;
  // Text buffers are recycled through a pool (no page faults on fresh memory for every slab); its size is the
  // look-ahead window: a GPU thread that finds the pool empty waits for the writer.
  std::vector<std::string> pool(n_gpu + 2);
  struct PerGpu {
    uint64_t pairs = 0, passes = 0, launches = 0, slabs = 0;
    double ms_device = 0, ms_em = 0, ms_pearson = 0, ms_format = 0;
  };
  std::vector<PerGpu> acc(n_gpu);

  std::thread writer([&]() {
    for (;;) {
      std::string text;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&]() { return failed || written == n_slabs || slab[written].done; });
        if (failed || written == n_slabs) return;
        text.swap(slab[written].text);
      }
      const bool ok = fwrite(text.data(), 1, text.size(), out_fh) == text.size();
      text.clear();  // keeps the capacity
      {
        std::lock_guard<std::mutex> lk(mu);
        if (!ok) failed = write_failed = true;
        written++;
        pool.emplace_back(std::move(text));
      }
      cv.notify_all();
    }
  });
  auto append_sink = [](void *user, const char *bytes, uint64_t n_bytes, uint64_t) -> int {
    ((std::string *)user)->append(bytes, n_bytes);
    return 0;
  };
  std::vector<int> rcs(n_gpu, 0);
  {
    std::vector<std::thread> th;
    for (int g = 0; g < n_gpu; g++)
      th.emplace_back([&, g]() {
        for (;;) {
          int k;
          std::string text;
          {
            std::unique_lock<std::mutex> lk(mu);
            // slabs must be claimed in order by whoever holds a buffer, or the writer could starve for the next slab
            cv.wait(lk, [&]() { return failed || next_slab >= n_slabs || !pool.empty(); });
            if (failed || next_slab >= n_slabs) return;
            k = next_slab++;
            text.swap(pool.back());
            pool.pop_back();
          }
          uint64_t rows = 0;
          if (ngsld_scan_count(ctx[g], bounds[k], bounds[k + 1], &P, &rows) == NGSLD_OK)
            text.reserve(rows * (o.extend_out ? 176 : 96) + 4096);  // typical row length; append() grows it if needed
          const int rc = ngsld_scan_tsv(ctx[g], bounds[k], bounds[k + 1], &P, append_sink, &text);
          ngsld_scan_stats st;
          ngsld_get_stats(ctx[g], &st);
          acc[g].pairs += st.n_pairs; acc[g].passes += st.sum_em_passes; acc[g].launches += st.n_launches; acc[g].slabs++;
          acc[g].ms_device += st.ms_device_total; acc[g].ms_em += st.ms_em; acc[g].ms_pearson += st.ms_pearson;
          acc[g].ms_format += st.ms_format;
          {
            std::lock_guard<std::mutex> lk(mu);
            if (rc) {
              rcs[g] = rc;
              failed = true;
            } else {
              slab[k].text.swap(text);
              slab[k].done = true;
            }
          }
          cv.notify_all();
          if (rc) return;
        }
      });
    for (auto &t : th) t.join();
  }
  cv.notify_all();
  writer.join();
  for (int g = 0; g < n_gpu; g++)
    if (rcs[g]) {
      fprintf(stderr, "GPU %d: %s\n", g, ngsld_last_error(ctx[g]));
      die(fn, "pair scan failed!");
    }
  if (write_failed) die(fn, "cannot write output!");
  if (o.gpu_stats)
    for (int g = 0; g < n_gpu; g++)
      fprintf(stderr, "[gpu %d] %lu slabs of %d: %lu pairs, %lu EM passes, %lu launches, scan %.1f ms incl. waiting for the writer (EM %.1f, r2_ExpG %.1f, format %.1f), %.0f pairs/s\n",
              g, acc[g].slabs, n_slabs, acc[g].pairs, acc[g].passes, acc[g].launches, acc[g].ms_device, acc[g].ms_em, acc[g].ms_pearson,
              acc[g].ms_format, acc[g].ms_device > 0 ? acc[g].pairs / (acc[g].ms_device * 1e-3) : 0.0);
  if (o.verbose >= 1) fprintf(stderr, "==> Freeing memory...\n");
  for (auto c : ctx) ngsld_destroy(c);
  ngsld_free(label_blob);
  if (out_fh != stdout) fclose(out_fh);
  else fflush(stdout);
  if (o.verbose >= 1) fprintf(stderr, "Done!\n");
  return 0;
}
