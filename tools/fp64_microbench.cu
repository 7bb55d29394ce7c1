// fp64 pipe microbenchmark (B200): dependent-DFMA latency and throughput versus ILP and warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench tools/fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, int iters, double a, double b) {
  double x[ILP];
  for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-3 + k;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int k = 0; k < ILP; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(t1 - t0) * 0.0;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int ILP> void run(int warps_per_sm, int sms) {
  double* d; cudaMalloc(&d, sizeof(double) * 4096 * 64);
  const int iters = 4096;
  const int threads = 32 * warps_per_sm;
  chain<ILP><<<sms, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  chain<ILP><<<sms, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  const double fma_total = (double)sms * threads * iters * ILP;
  printf("ILP %d warps/SM %2d: %.2f cycles per dependent DFMA step (per warp), %.2f warp-DFMA/clk/SM, %.2f TFLOP/s\n", ILP, warps_per_sm,
         cyc / iters, (double)warps_per_sm * ILP * iters / cyc, 2.0 * fma_total / (ms * 1e-3) / 1e12);
  cudaFree(d);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  for (int w : {1, 4, 8, 16, 32}) { run<1>(w, sms); run<2>(w, sms); run<4>(w, sms); run<8>(w, sms); }
  return 0;
}
