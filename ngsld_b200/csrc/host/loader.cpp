#include "loader.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <thread>

namespace loader {

static const size_t kLineCap = 500000;  // reference BUFF_LEN, shared/gen_func.hpp:17
static const double kBig = 1e15;        // reference INF, shared/gen_func.hpp:15

static Failure fail(const char *func, const char *msg, bool io = false) {
  Failure f;
  f.func = func;
  f.msg = msg;
  f.io = io;
  return f;
}

static gzFile open_any(const char *path, const char *mode) {
  gzFile fh = strcmp(path, "-") == 0 ? gzdopen(fileno(stdin), mode) : gzopen(path, mode);
  if (fh) gzbuffer(fh, 1 << 20);
  return fh;
}

// a regular file that does not start with the gzip magic: can be read without zlib
static bool plain_regular_file(const char *path) {
  if (strcmp(path, "-") == 0) return false;
  struct stat st;
  if (stat(path, &st) != 0 || !S_ISREG(st.st_mode)) return false;
  FILE *fh = fopen(path, "rb");
  if (!fh) return false;
  unsigned char m[2] = {0, 0};
  const size_t n = fread(m, 1, 2, fh);
  fclose(fh);
  return !(n == 2 && m[0] == 0x1f && m[1] == 0x8b);
}

// drop ONE trailing '\n' or '\r' like the reference's chomp() (shared/gen_func.cpp:192-199)
static void chomp1(char *s) {
  const size_t n = strlen(s);
  if (n && (s[n - 1] == '\n' || s[n - 1] == '\r')) s[n - 1] = '\0';
}

// numeric fields of a line split on blanks/tabs; tokens strtod does not fully consume are dropped
// (reference split(), shared/gen_func.cpp:388-412)
static void numeric_fields(char *line, std::vector<double> &out) {
  out.clear();
  char *p = line;
  while (*p) {
    while (*p == ' ' || *p == '\t') p++;
    if (!*p) break;
    char *q = p;
    while (*q && *q != ' ' && *q != '\t') q++;
    const char keep = *q;
    *q = '\0';
    char *end = nullptr;
    const double v = strtod(p, &end);
    if (end != p && *end == '\0') out.push_back(v);
    *q = keep;
    p = q;
  }
}

namespace {
struct GzCloser {
  gzFile fh;
  ~GzCloser() {
    if (fh) gzclose(fh);
  }
};
}  // namespace

Failure read_geno(const char *path, bool is_bin, bool probs, bool log_scale, uint64_t n_ind, uint64_t n_sites,
                  double *cells, bool *log_cells) {
  const char *fn = "read_geno";
  gzFile fh = open_any(path, is_bin ? "rb" : "r");
  if (!fh) return fail(fn, "cannot open GENO file!", true);
  GzCloser closer{fh};
  const size_t n_cells = (size_t)n_sites * n_ind * 3;
  if (is_bin) {
    *log_cells = false;
    if (plain_regular_file(path)) {
      // An uncompressed regular file: zlib would only pass the bytes through (at ~1.8 GB/s); read it directly, in parallel
      // slices.  Same outcomes as the zlib path: a short file is "premature EOF", a longer one "not at EOF".
      const int fd = open(path, O_RDONLY);
      if (fd < 0) return fail(fn, "cannot open GENO file!", true);
      struct stat st;
      const size_t bytes = n_cells * sizeof(double);
      if (fstat(fd, &st) != 0) {
        close(fd);
        return fail(fn, "cannot read binary GENO file. Check GENO file and number of sites!", true);
      }
      if ((size_t)st.st_size < bytes) {
        close(fd);
        return fail(fn, "GENO file at premature EOF. Check GENO file and number of sites!");
      }
      const int n_thr = bytes > (64u << 20) ? 8 : 1;
      std::vector<int> bad(n_thr, 0);
      std::vector<std::thread> pool;
      for (int t = 0; t < n_thr; t++)
        pool.emplace_back([&, t]() {
          size_t lo = bytes / n_thr * t, hi = t + 1 == n_thr ? bytes : bytes / n_thr * (t + 1);
          char *dst = reinterpret_cast<char *>(cells);
          while (lo < hi) {
            const ssize_t got = pread(fd, dst + lo, std::min<size_t>(hi - lo, (size_t)1 << 30), (off_t)lo);
            if (got <= 0) {
              if (got < 0 && errno == EINTR) continue;
              bad[t] = 1;
              return;
            }
            lo += (size_t)got;
          }
        });
      for (auto &t : pool) t.join();
      close(fd);
      for (int b : bad)
        if (b) return fail(fn, "cannot read binary GENO file. Check GENO file and number of sites!", true);
      if ((size_t)st.st_size > bytes) return fail(fn, "GENO file not at EOF. Check GENO file and number of sites!");
      return Failure();
    }
    char *dst = reinterpret_cast<char *>(cells);
    size_t left = n_cells * sizeof(double);
    while (left) {
      const unsigned want = (unsigned)(left < (1u << 30) ? left : (1u << 30));
      const int got = gzread(fh, dst, want);
      if (got <= 0) {
        if (gzeof(fh)) return fail(fn, "GENO file at premature EOF. Check GENO file and number of sites!");
        return fail(fn, "cannot read binary GENO file. Check GENO file and number of sites!", true);
      }
      dst += got;
      left -= got;
    }
  } else {
    *log_cells = true;
    const uint64_t per_ind = probs ? 3 : 1, need = n_ind * per_ind;
    std::vector<char> buf(kLineCap);
    std::vector<double> f;
    uint64_t s = 0;
    while (s < n_sites) {
      if (gzgets(fh, buf.data(), (int)kLineCap) == NULL) {
        if (gzeof(fh)) return fail(fn, "GENO file at premature EOF. Check GENO file and number of sites!");
        return fail(fn, "cannot read GZip GENO file. Check GENO file and number of sites!", true);
      }
      chomp1(buf.data());
      if (buf[0] == '\0') {  // the reference consumes a site slot for an empty line (read_data.cpp:58-59)
        double *c = cells + s * n_ind * 3;
        for (uint64_t k = 0; k < n_ind * 3; k++) c[k] = -kBig;
        s++;
        continue;
      }
      numeric_fields(buf.data(), f);
      if (f.empty() || (s == 0 && f.size() < need)) {
        fprintf(stderr, "> Header found! Skipping line...\n");
        if (s != 0) fprintf(stderr, "\n=======\nWARNING: [%s] %s\n=======\n\n", fn, " header found but not on first line. Is this an error?");
        continue;
      }
      if (f.size() < need) return fail(fn, "wrong GENO file format. Less fields than expected!");
      const double *v = f.data() + (f.size() - need);
      double *c = cells + s * n_ind * 3;
      for (uint64_t i = 0; i < n_ind; i++) {
        if (probs) {
          for (int g = 0; g < 3; g++) c[3 * i + g] = log_scale ? v[3 * i + g] : log(v[3 * i + g]);
        } else {
          const int g = (int)v[i];
          if (g >= 0) {
            if (g > 2) return fail(fn, "wrong GENO file format. Genotypes must be coded as {-1,0,1,2} !");
            c[3 * i] = c[3 * i + 1] = c[3 * i + 2] = -kBig;
            c[3 * i + g] = log(1);
          } else {
            c[3 * i] = c[3 * i + 1] = c[3 * i + 2] = log((double)1 / 3);
          }
        }
      }
      s++;
    }
  }
  char one;
  gzread(fh, &one, 1);
  if (!gzeof(fh)) return fail(fn, "GENO file not at EOF. Check GENO file and number of sites!");
  return Failure();
}

Failure read_positions(const char *path, bool header, uint64_t n_sites, std::vector<std::string> &labels,
                       double *pos_dist) {
  const char *fn = "read_dist";
  gzFile fh = open_any(path, "r");
  if (!fh) return fail("read_file", "cannot open file!", true);
  std::vector<char> buf(kLineCap);
  labels.clear();
  uint64_t skip = header ? 1 : 0;
  for (;;) {
    buf[0] = '\0';
    gzgets(fh, buf.data(), (int)kLineCap);
    if (gzeof(fh) && buf[0] == '\0') break;
    const bool last = gzeof(fh);
    chomp1(buf.data());
    if (buf[0] != '\0' && buf[0] != '#') {
      if (skip)
        skip--;
      else
        labels.emplace_back(buf.data());
    }
    if (last) break;
  }
  gzclose(fh);
  if (labels.size() != n_sites) return fail(fn, "wrong number of lines in POS file!");
  for (uint64_t s = 0; s < n_sites; s++) pos_dist[s] = INFINITY;
  std::string prev_chr;
  bool have_chr = false;
  uint64_t prev_pos = 0;
  for (uint64_t s = 0; s < n_sites; s++) {
    std::string &ln = labels[s];
    const size_t t1 = ln.find('\t');
    if (t1 == std::string::npos) return fail(fn, "wrong POS file format!");
    const size_t t2 = ln.find('\t', t1 + 1);
    const std::string chr = ln.substr(0, t1);
    const std::string ptxt = ln.substr(t1 + 1, t2 == std::string::npos ? std::string::npos : t2 - t1 - 1);
    const double pos = strtod(ptxt.c_str(), NULL);
    // The reference treats such a line as a header and loops forever (read_data.cpp:188-196); reject instead.
    if (pos == 0) return fail(fn, "position 0 or non-numeric position in POS file (use --posH for a header line)!");
    if (!have_chr) {
      prev_chr = chr;
      have_chr = true;
    }
    if (prev_chr == chr) {
      pos_dist[s] = pos - (double)prev_pos;
      if (pos_dist[s] < 1) return fail(fn, "invalid distance between adjacent sites!");
    } else {
      pos_dist[s] = INFINITY;
      prev_chr = chr;
    }
    prev_pos = strtoul(ptxt.c_str(), NULL, 0);
    ln[t1] = ':';  // only the first tab becomes ':' (ngsLD.cpp:127-132)
  }
  return Failure();
}

}  // namespace loader
