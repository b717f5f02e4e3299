// FP64 pipe micro-benchmark 2: does DFMA throughput depend on how many distinct register operands it reads?
#include <cstdio>
#include <cuda_runtime.h>
// MODE 0: a = fma(a, m, c) (2 shared operands)   MODE 1: a[j] = fma(b[j], c[j], a[j]) (3 distinct register operands)
// MODE 2: a[j] = fma(b[j], m, a[j])  MODE 3: estep-like mix: DMUL + DFMA with distinct operands
template <int ILP, int MODE>
__global__ void k(double *out, const double *in, int iters) {
  double a[ILP], b[ILP], c[ILP];
#pragma unroll
  for (int j = 0; j < ILP; j++) { a[j] = in[threadIdx.x + j]; b[j] = in[threadIdx.x + 64 + j]; c[j] = in[threadIdx.x + 128 + j]; }
  const double m = in[300], cc = in[301];
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < ILP; j++) {
      if (MODE == 0) a[j] = __fma_rn(a[j], m, cc);
      if (MODE == 1) a[j] = __fma_rn(b[j], c[j], a[j]);
      if (MODE == 2) a[j] = __fma_rn(b[j], m, a[j]);
      if (MODE == 3) { a[j] = __fma_rn(b[j], c[(j + 1) % ILP], a[j]); }
    }
    if (MODE == 3) {
#pragma unroll
      for (int j = 0; j < ILP; j++) b[j] = __dmul_rn(a[j], c[j]);
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) s += a[j] + b[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int MODE>
void run(int warps_per_sm, int sms, double *d, double *in) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP, MODE><<<sms, warps_per_sm * 32>>>(d, in, 100);
  cudaEventRecord(e0);
  k<ILP, MODE><<<sms, warps_per_sm * 32>>>(d, in, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cycles = ms * 1e-3 * clk * 1e3;
  double n = (MODE == 3 ? 2.0 : 1.0);
  printf("MODE %d ILP %d warps/SM %2d: %.1f FP64 instr-lanes/clk/SM\n", MODE, ILP, warps_per_sm,
         (double)warps_per_sm * 32 * ILP * iters * n / cycles);
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *d, *in; cudaMalloc(&d, sms * 1024 * sizeof(double)); cudaMalloc(&in, 4096 * sizeof(double));
  cudaMemset(in, 0, 4096 * sizeof(double));
  for (int w : {4, 8, 12, 16}) {
    run<4, 0>(w, sms, d, in); run<4, 1>(w, sms, d, in); run<4, 2>(w, sms, d, in); run<4, 3>(w, sms, d, in);
    run<8, 0>(w, sms, d, in); run<8, 1>(w, sms, d, in); run<8, 2>(w, sms, d, in); run<8, 3>(w, sms, d, in);
  }
  return 0;
}
