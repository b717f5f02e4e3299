// Does the 0.73x rate of 3-register DFMAs come from register-bank conflicts (allocation-dependent) or from a
// read-port limit?  Same instruction count, different operand index patterns -> different register assignments.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int SB, int SC, int PAD>
__global__ void k(double *out, const double *in, int iters) {
  double a[ILP], b[ILP], c[ILP], pad[PAD + 1];
#pragma unroll
  for (int j = 0; j < ILP; j++) { a[j] = in[threadIdx.x + j]; b[j] = in[threadIdx.x + 64 + j]; c[j] = in[threadIdx.x + 128 + j]; }
#pragma unroll
  for (int j = 0; j <= PAD; j++) pad[j] = in[threadIdx.x + 200 + j];
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < ILP; j++) a[j] = __fma_rn(b[(j + SB) % ILP], c[(j + SC) % ILP], a[j]);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) s += a[j] + b[j] + c[j];
#pragma unroll
  for (int j = 0; j <= PAD; j++) s += pad[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int SB, int SC, int PAD>
void run(int sms, double *d, double *in) {
  const int iters = 20000, w = 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP, SB, SC, PAD><<<sms, w * 32>>>(d, in, 100);
  cudaEventRecord(e0);
  k<ILP, SB, SC, PAD><<<sms, w * 32>>>(d, in, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("ILP %d SB %d SC %d PAD %d: %.1f DFMA lanes/clk/SM\n", ILP, SB, SC, PAD, (double)w * 32 * ILP * iters / (ms * 1e-3 * clk * 1e3));
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *d, *in; cudaMalloc(&d, sms * 1024 * sizeof(double)); cudaMalloc(&in, 4096 * sizeof(double));
  cudaMemset(in, 0, 4096 * sizeof(double));
  run<8, 0, 0, 0>(sms, d, in); run<8, 1, 0, 0>(sms, d, in); run<8, 0, 1, 0>(sms, d, in); run<8, 1, 2, 0>(sms, d, in);
  run<8, 0, 0, 1>(sms, d, in); run<8, 1, 0, 1>(sms, d, in); run<8, 0, 1, 1>(sms, d, in); run<8, 3, 5, 1>(sms, d, in);
  run<8, 0, 0, 2>(sms, d, in); run<8, 1, 3, 2>(sms, d, in); run<8, 2, 1, 3>(sms, d, in); run<8, 3, 2, 3>(sms, d, in);
  run<6, 0, 0, 0>(sms, d, in); run<6, 1, 2, 1>(sms, d, in); run<7, 1, 2, 0>(sms, d, in); run<5, 1, 2, 0>(sms, d, in);
  return 0;
}
