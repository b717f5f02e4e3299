// TSV formatter kernels (see format.cuh).
//
// "%f" is reproduced exactly: a double is m * 2^e with an integer m < 2^53, so |v| * 10^6 =
// (m * 10^6) * 2^e is formed in 128-bit integer arithmetic and rounded half-to-even on the exact
// value -- what glibc's printf does in the default rounding mode.  The device path covers
// |v| < 1e9 (every probability / ratio this program prints in practice); a larger finite value sets
// the overflow flag and the host re-formats that chunk with snprintf.  NaN prints as "-nan": every
// NaN the x86 reference can produce on this path is the negative default NaN (0/0, inf-inf,
// sqrt(<0)) or a propagated copy of it, whereas the GPU's canonical NaN is positive.
#include <stdio.h>
#include <string.h>

#include "fixed6.cuh"
#include "format.cuh"

namespace fmt {

namespace {

constexpr int F_MAX = 18;  // sign + 9 integer digits + '.' + 6 decimals (+1 spare)

__device__ __forceinline__ int put_u64(char *out, unsigned long long v) {
  char tmp[20];
  int n = 0;
  do {
    tmp[n++] = (char)('0' + (int)(v % 10));
    v /= 10;
  } while (v);
  for (int k = 0; k < n; k++) out[k] = tmp[n - 1 - k];
  return n;
}

// "%f"; returns length, sets *ovf when |v| >= 1e9 (finite)
__device__ int put_f6(char *out, double v, unsigned int *ovf) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const bool neg = (bits >> 63) != 0;
  const int be = (int)((bits >> 52) & 0x7ff);
  const unsigned long long frac = bits & 0xfffffffffffffull;
  int n = 0;
  if (be == 0x7ff) {
    if (frac) {
      out[0] = '-'; out[1] = 'n'; out[2] = 'a'; out[3] = 'n';
      return 4;
    }
    if (neg) out[n++] = '-';
    out[n++] = 'i'; out[n++] = 'n'; out[n++] = 'f';
    return n;
  }
  if (neg) out[n++] = '-';
  unsigned long long N = 0;  // round_half_even(|v| * 1e6)
  if (!fixed6(v, N)) {
    *ovf = 1;
    out[n++] = '?';
    return n;
  }
  const unsigned long long ip = N / 1000000ull;
  unsigned int fp = (unsigned int)(N % 1000000ull);
  n += put_u64(out + n, ip);
  out[n++] = '.';
  for (int k = 5; k >= 0; k--) {
    out[n + k] = (char)('0' + (int)(fp % 10));
    fp /= 10;
  }
  return n + 6;
}

// "%.0f" for the accumulated distance: a non-negative integer below 2^53, or +inf
__device__ __forceinline__ int put_dist(char *out, double v) {
  if (isinf(v)) {
    out[0] = 'i'; out[1] = 'n'; out[2] = 'f';
    return 3;
  }
  return put_u64(out, (unsigned long long)v);
}

__device__ __forceinline__ int put_label(char *out, const FormatArgs &fa, uint32_t site) {
  if (!fa.labels) {
    const char nul[6] = {'(', 'n', 'u', 'l', 'l', ')'};
    for (int k = 0; k < 6; k++) out[k] = nul[k];
    return 6;
  }
  const uint32_t a = fa.label_off[site], b = fa.label_off[site + 1];
  for (uint32_t k = a; k < b; k++) out[k - a] = fa.labels[k];
  return (int)(b - a);
}

__global__ void __launch_bounds__(128) format_rows_kernel(FormatArgs fa, const ngsld_pair_row *rows,
                                                          unsigned long long n, char *slots,
                                                          unsigned long long *line_len, unsigned long long *ovf_flag) {
  for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < n;
       p += (unsigned long long)gridDim.x * blockDim.x) {
    const ngsld_pair_row r = rows[p];
    char *out = slots + p * fa.slot;
    unsigned int ovf = 0;
    int k = 0;
    k += put_label(out + k, fa, r.s1);
    out[k++] = '\t';
    k += put_label(out + k, fa, r.s2);
    out[k++] = '\t';
    k += put_dist(out + k, r.dist);
    const double std4[4] = {r.r2_expg, r.D, r.Dp, r.r2};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      out[k++] = '\t';
      k += put_f6(out + k, std4[j], &ovf);
    }
    if (fa.extend_out) {
      out[k++] = '\t';
      k += put_u64(out + k, r.n_used);
      const double ext[10] = {fa.maf[r.s1], fa.maf[r.s2], r.hap[0], r.hap[1], r.hap[2], r.hap[3],
                              r.hap_maf[0], r.hap_maf[1], (double)r.chi2, 0.0};
#pragma unroll
      for (int j = 0; j < 10; j++) {
        out[k++] = '\t';
        k += put_f6(out + k, ext[j], &ovf);
      }
      out[k++] = '\t';
      k += put_u64(out + k, r.n_iter);
    }
    out[k++] = '\n';
    line_len[p] = (unsigned long long)k;
    if (ovf) atomicExch(ovf_flag, 1ull);
  }
}

// ---- exclusive scan of row lengths: block sums, scan of block sums, final pass -------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;  // rows per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long *total) {
  __shared__ unsigned long long warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
    for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = w;
  }
  __syncthreads();
  const unsigned long long base = warp ? warp_sums[warp - 1] : 0;
  if (total) *total = warp_sums[SCAN_THREADS / 32 - 1];
  __syncthreads();
  return base + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const unsigned long long *len,
                                                                      unsigned long long n, unsigned long long *sums) {
  const unsigned long long base = (unsigned long long)blockIdx.x * SCAN_TILE + (unsigned long long)threadIdx.x * SCAN_ITEMS;
  unsigned long long s = 0;
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (base + k < n) s += len[base + k];
  unsigned long long total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(unsigned long long *sums, unsigned int n_tiles,
                                                                 unsigned long long *grand_total) {
  // single block: sequential over strips of SCAN_THREADS tile sums
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (unsigned int base = 0; base < n_tiles; base += SCAN_THREADS) {
    const unsigned int i = base + threadIdx.x;
    const unsigned long long v = i < n_tiles ? sums[i] : 0;
    unsigned long long total;
    const unsigned long long ex = block_exclusive_scan(v, &total);
    const unsigned long long c = carry;
    if (i < n_tiles) sums[i] = c + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_final_kernel(unsigned long long *len_to_off, unsigned long long n,
                                                                  const unsigned long long *sums) {
  const unsigned long long base = (unsigned long long)blockIdx.x * SCAN_TILE + (unsigned long long)threadIdx.x * SCAN_ITEMS;
  unsigned long long v[SCAN_ITEMS], s = 0;
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = base + k < n ? len_to_off[base + k] : 0;
    s += v[k];
  }
  unsigned long long run = sums[blockIdx.x] + block_exclusive_scan(s, nullptr);
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) len_to_off[base + k] = run;
    run += v[k];
  }
}

// one warp per row: slot -> packed position
__global__ void __launch_bounds__(256) pack_rows_kernel(const char *slots, uint32_t slot, const unsigned long long *off,
                                                        unsigned long long n, char *packed) {
  const int lane = threadIdx.x & 31;
  const unsigned long long warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long p = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n; p += warps) {
    const unsigned long long a = off[p], b = off[p + 1];
    const char *src = slots + p * slot;
    for (unsigned long long k = lane; k < b - a; k += 32) packed[a + k] = src[k];
  }
}

}  // namespace

uint32_t slot_bytes(uint32_t max_label_len, bool extend_out) {
  uint32_t n = 2 * max_label_len + 2 + 17 + 4 * (1 + F_MAX) + 1;
  if (extend_out) n += 2 * (1 + 11) + 10 * (1 + F_MAX);
  return (n + 15) & ~15u;
}

int launch_format(const FormatArgs &fa, const SiteTable &T, const ngsld_pair_row *rows, unsigned long long n,
                  char *slots, unsigned long long *line_off, char *packed, int sm_count, cudaStream_t stream) {
  (void)T;
  if (n == 0) return 0;
  // line_off[0..n): lengths then offsets; [n]: total; [n+1]: overflow flag; [n+2 ...): tile sums
  const unsigned int n_tiles = (unsigned int)((n + SCAN_TILE - 1) / SCAN_TILE);
  unsigned long long *sums = line_off + n + 2;
  cudaMemsetAsync(line_off + n, 0, 2 * sizeof(unsigned long long), stream);
  const unsigned fb = (unsigned)((n + 127) / 128 < (unsigned long long)sm_count * 32 ? (n + 127) / 128
                                                                                     : (unsigned long long)sm_count * 32);
  format_rows_kernel<<<fb, 128, 0, stream>>>(fa, rows, n, slots, line_off, line_off + n + 1);
  scan_tile_sums_kernel<<<n_tiles, SCAN_THREADS, 0, stream>>>(line_off, n, sums);
  scan_sums_kernel<<<1, SCAN_THREADS, 0, stream>>>(sums, n_tiles, line_off + n);
  scan_final_kernel<<<n_tiles, SCAN_THREADS, 0, stream>>>(line_off, n, sums);
  const unsigned long long want = (n * 32 + 255) / 256;
  const unsigned pb = (unsigned)(want < (unsigned long long)sm_count * 16 ? want : (unsigned long long)sm_count * 16);
  pack_rows_kernel<<<pb, 256, 0, stream>>>(slots, fa.slot, line_off, n, packed);
  return cudaGetLastError() == cudaSuccess ? 5 : -1;
}

// Every NaN the x86 reference produces on this path has its sign bit set and prints as "-nan"; the GPU's canonical
// NaN is positive, so the sign is forced before glibc sees it.
static inline double neg_nan(double v) { return v != v ? -__builtin_nan("") : v; }

int format_row_host(const ngsld_pair_row &r0, const char *l1, const char *l2, double maf1, double maf2, int extend_out,
                    char *buf, size_t cap) {
  ngsld_pair_row r = r0;
  r.r2_expg = neg_nan(r.r2_expg); r.D = neg_nan(r.D); r.Dp = neg_nan(r.Dp); r.r2 = neg_nan(r.r2);
  for (int k = 0; k < 4; k++) r.hap[k] = neg_nan(r.hap[k]);
  r.hap_maf[0] = neg_nan(r.hap_maf[0]); r.hap_maf[1] = neg_nan(r.hap_maf[1]);
  const double chi2 = neg_nan((double)r.chi2);
  maf1 = neg_nan(maf1); maf2 = neg_nan(maf2);
  int n = snprintf(buf, cap, "%s\t%s\t%.0f\t%f\t%f\t%f\t%f", l1, l2, r.dist, r.r2_expg, r.D, r.Dp, r.r2);
  if (n < 0 || (size_t)n >= cap) return -1;
  if (extend_out) {
    int m = snprintf(buf + n, cap - n, "\t%lu\t%f\t%f\t%f\t%f\t%f\t%f\t%f\t%f\t%f\t%f\t%lu", (unsigned long)r.n_used, maf1,
                     maf2, r.hap[0], r.hap[1], r.hap[2], r.hap[3], r.hap_maf[0], r.hap_maf[1], chi2, 0.0,
                     (unsigned long)r.n_iter);
    if (m < 0 || (size_t)(n + m) >= cap) return -1;
    n += m;
  }
  if ((size_t)n + 1 >= cap) return -1;
  buf[n++] = '\n';
  return n;
}

}  // namespace fmt
