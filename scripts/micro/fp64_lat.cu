// FP64 pipe micro-benchmark: DFMA throughput per SM as a function of independent chains per warp (ILP)
// and resident warps per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double m, double c) {
  double a[ILP];
#pragma unroll
  for (int j = 0; j < ILP; j++) a[j] = threadIdx.x * 1e-9 + j;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < ILP; j++) a[j] = __fma_rn(a[j], m, c);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int warps_per_sm, int sms, double *d) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP><<<sms, warps_per_sm * 32>>>(d, 100, 1.0000001, 1e-7);
  cudaEventRecord(e0);
  k<ILP><<<sms, warps_per_sm * 32>>>(d, iters, 1.0000001, 1e-7);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cycles = ms * 1e-3 * clk * 1e3;
  double fma_per_clk_sm = (double)warps_per_sm * 32 * ILP * iters / cycles;
  printf("ILP %d warps/SM %2d: %.1f DFMA/clk/SM, cycles per dependent step per warp %.1f\n", ILP, warps_per_sm, fma_per_clk_sm,
         cycles / iters);
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *d; cudaMalloc(&d, sms * 1024 * sizeof(double));
  for (int w : {4, 8, 12, 16, 32}) {
    run<1>(w, sms, d); run<2>(w, sms, d); run<4>(w, sms, d); run<8>(w, sms, d);
  }
  return 0;
}
