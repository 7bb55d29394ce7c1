// tools/dmma_microbench.cu -- throughput / latency of the fp64 tensor-core instruction (mma.sync.m8n8k4.f64, SASS DMMA) on this GPU,
// next to the vector DFMA numbers of tools/fp64_microbench.cu.  One m8n8k4 is 256 FMA = 8 warp-wide DFMA instructions of work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma tools/dmma_microbench.cu && /tmp/dmma
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, int iters, double a0, double b0) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(int warps_per_sm, int sms, double* d_out, double clock_ghz) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP><<<sms, 32 * warps_per_sm>>>(d_out, 100, 1e-3, 1e-3);
  cudaEventRecord(e0);
  k<ILP><<<sms, 32 * warps_per_sm>>>(d_out, iters, 1e-3, 1e-3);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double cycles = ms * 1e-3 * clock_ghz * 1e9;
  const double per_step = cycles / iters;                       // cycles per ILP-wide step per warp
  const double dmma_per_clk_sm = (double)ILP * warps_per_sm * iters / cycles;
  printf("ILP %d warps/SM %2d: %.2f cycles per dependent DMMA step, %.3f DMMA/clk/SM = %.2f warp-DFMA-equivalents/clk/SM, %.2f TFLOP/s\n", ILP, warps_per_sm,
         per_step, dmma_per_clk_sm, 8 * dmma_per_clk_sm, 512.0 * ILP * warps_per_sm * sms * iters / (ms * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
  double* d; cudaMalloc(&d, sizeof(double) * p.multiProcessorCount * 1024);
  for (int w : {1, 4, 8, 16}) { run<1>(w, p.multiProcessorCount, d, ghz); run<2>(w, p.multiProcessorCount, d, ghz); run<4>(w, p.multiProcessorCount, d, ghz); run<8>(w, p.multiProcessorCount, d, ghz); }
  return 0;
}
