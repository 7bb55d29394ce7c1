// tools/tma_store_microbench.cu -- how fast can an SM ISSUE bulk stores (cp.async.bulk.global.shared::cta, SASS UBLKCP)?
// k_static_rs hands every finished 912-B row to the TMA engine with its own bulk store (32 per tile in caller order); the phase timing of
// round 2 (profiles/r2j_phase_timing.txt) puts 13 % of a tile on that issue loop.  This measures cycles per bulk store for W warps per SM
// issuing stores of BYTES bytes to scattered rows, against plain st.global.v2.f64 stores of the same bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma tools/tma_store_microbench.cu && /tmp/tma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void bulk_store(double* g, const double* s, unsigned bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(sa), "r"(bytes) : "memory");
}
template <int MODE>      // 0: one bulk store per lane (32 per iteration); 1: one bulk store per warp (lane 0, 32 rows contiguous); 2: cooperative 16-B st.global
__global__ void k(double* out, int iters, int row_doubles, long long n_rows, long long* cycles) {
  extern __shared__ __align__(128) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* wbase = smem + (size_t)warp * 32 * row_doubles;
  for (int i = lane; i < 32 * row_doubles; i += 32) wbase[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp, nw = (long long)gridDim.x * (blockDim.x >> 5);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const long long tile = (gw + (long long)it * nw) % (n_rows / 32);
    if (MODE == 0) {
      const long long row = (tile * 32 + (lane * 7919LL) % 32 + ((tile * 2654435761LL) % (n_rows / 32)) * 0) ;
      const long long r2 = (row * 2654435761LL) % n_rows;                 // scattered destination rows
      bulk_store(out + r2 * row_doubles, wbase + lane * row_doubles, row_doubles * 8);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    } else if (MODE == 1) {
      if (lane == 0) {
        bulk_store(out + tile * 32 * row_doubles, wbase, 32 * row_doubles * 8);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
    } else {
      for (int rr = 0; rr < 32; ++rr) {
        const long long r2 = ((tile * 32 + rr) * 2654435761LL) % n_rows;
        const double2* src = reinterpret_cast<const double2*>(wbase + rr * row_doubles);
        double2* dst = reinterpret_cast<double2*>(out + r2 * row_doubles);
        for (int c = lane; c < row_doubles / 2; c += 32) dst[c] = src[c];
      }
    }
  }
  const long long t1 = clock64();
  if (lane == 0 && gw == 0) *cycles = t1 - t0;
}

template <int MODE> void run(const char* name, int warps, int sms, double* d_out, long long n_rows, long long* d_cyc) {
  const int row_doubles = 114, iters = 200;
  const size_t smem = (size_t)warps * 32 * row_doubles * 8;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<MODE><<<sms, 32 * warps, smem>>>(d_out, 10, row_doubles, n_rows, d_cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<sms, 32 * warps, smem>>>(d_out, iters, row_doubles, n_rows, d_cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cyc; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  const double bytes = (double)sms * warps * iters * 32 * row_doubles * 8;
  printf("%-34s warps/SM %d: %8.0f cycles per 32-row tile per warp, %6.1f cycles per tile per SM, %7.1f GB/s chip-wide\n", name, warps, (double)cyc / iters,
         (double)cyc / iters / warps, bytes / (ms * 1e-3) / 1e9);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const long long n_rows = 500000;
  double* d; cudaMalloc(&d, n_rows * 114 * 8);
  long long* c; cudaMalloc(&c, 8);
  printf("%s, %d SMs; 912-B rows, 32 rows per warp iteration\n", p.name, p.multiProcessorCount);
  for (int w : {1, 2, 4, 7}) run<0>("bulk store per row (32 per tile)", w, p.multiProcessorCount, d, n_rows, c);
  for (int w : {1, 2, 4, 7}) run<1>("one bulk store per tile (29 KB)", w, p.multiProcessorCount, d, n_rows, c);
  for (int w : {1, 2, 4, 7}) run<2>("cooperative st.global.v2 (16 B)", w, p.multiProcessorCount, d, n_rows, c);
  return 0;
}
