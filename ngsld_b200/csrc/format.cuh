// Device-side TSV formatter: reproduces the reference's fprintf block (ngsLD.cpp:314-351) byte for
// byte, i.e. glibc's "%s\t%s\t%.0f\t%f\t%f\t%f\t%f" [+ "\t%lu" 10x"\t%f" "\t%lu"] "\n".
#pragma once
#include "common.cuh"

namespace fmt {

struct FormatArgs {
  const char *labels;         // concatenated site labels (NULL -> every label prints as "(null)")
  const uint32_t *label_off;  // [n_sites + 1] offsets into `labels`
  const double *maf;          // [n_sites] (maf1 / maf2 columns)
  int extend_out;
  uint32_t slot;              // bytes reserved per row in the scratch buffer
};

// worst-case bytes of one formatted row (values the device formats itself are bounded, see format.cu)
uint32_t slot_bytes(uint32_t max_label_len, bool extend_out);

// Formats rows[0..n) into `packed` (rows back to back, in order).  line_off[0..n] receives the
// exclusive prefix of row lengths (line_off[n] = total bytes); line_off[n+1] is set non-zero if some
// value was outside the device formatter's range (caller must fall back to host formatting).
// Returns the number of kernels launched, or -1.
int launch_format(const FormatArgs &fa, const SiteTable &T, const ngsld_pair_row *rows, unsigned long long n,
                  char *slots, unsigned long long *line_off, char *packed, int sm_count, cudaStream_t stream);

// Host reference formatter for one row (snprintf; used for the out-of-range fallback).
int format_row_host(const ngsld_pair_row &r, const char *l1, const char *l2, double maf1, double maf2, int extend_out,
                    char *buf, size_t cap);

}  // namespace fmt
