// Input readers of the ngsLD-compatible CLI: genotype file (binary doubles / gz text) and position file.
// Behaviour follows the reference's readers (shared/read_data.cpp:13-116,165-218; shared/gen_func.cpp:238-282)
// but produces flat arrays for the C ABI instead of pointer-of-pointer tables.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace loader {

// A reader failure: the reference would call error(func, msg) (shared/gen_func.cpp:12-18) and exit; the library
// reports it to the caller instead (the CLI then prints it in the reference's format).
struct Failure {
  const char *func = nullptr;  // reference function name the message belongs to
  const char *msg = nullptr;
  bool io = false;             // true: file could not be opened/read; false: content rejected
  explicit operator bool() const { return msg != nullptr; }
};

// cells: [n_sites][n_ind][3]; *log_cells = true when the cells are log-space values (text input), false when they are
// the file's raw doubles.  is_bin: raw little-endian doubles [n_sites][n_ind][3]; otherwise gz/plain text, one site per line, the last
// n_ind*(probs?3:1) numeric fields used, leading header line skipped.
Failure read_geno(const char *path, bool is_bin, bool probs, bool log_scale, uint64_t n_ind, uint64_t n_sites,
                  double *cells, bool *log_cells);

// labels ("chr:pos", only the first tab replaced) and inter-site distances (+inf at a chromosome change)
Failure read_positions(const char *path, bool header, uint64_t n_sites, std::vector<std::string> &labels,
                       double *pos_dist);

}  // namespace loader
