// tools/scatter4_probe.cu -- does cp.async.bulk.tensor.2d ... tile::scatter4 store four 912-B fp64 rows from shared memory to four
// arbitrary rows of a [n][114] matrix, and with which box shape of the tensor map?  (No public doc in this image: probe, then use.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o scatter4_probe tools/scatter4_probe.cu && timeout 60 ./scatter4_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

constexpr int kCols = 114, kRows = 64;

__global__ void k_probe(const __grid_constant__ CUtensorMap tmap, int r0, int r1, int r2, int r3, int pad_doubles) {
  extern __shared__ __align__(128) double smem[];
  double* src = smem + pad_doubles;
  for (int i = threadIdx.x; i < 4 * kCols; i += blockDim.x) src[i] = 1000.0 * (i / kCols + 1) + (i % kCols);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(src);
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile::scatter4.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                 ::"l"(&tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(saddr) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  double* d = nullptr;
  cudaMalloc(&d, sizeof(double) * kRows * kCols);
  for (int box_rows : {1, 4}) {
    for (int pad : {0, 8}) {            // source 128-B aligned / 64-B aligned
      cudaMemset(d, 0, sizeof(double) * kRows * kCols);
      CUtensorMap tmap;
      const cuuint64_t dims[2] = {kCols, kRows}, strides[1] = {kCols * sizeof(double)};
      const cuuint32_t box[2] = {kCols, (cuuint32_t)box_rows}, estr[2] = {1, 1};
      const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("box_rows=%d: encode failed (%d)\n", box_rows, (int)r); continue; }
      k_probe<<<1, 128, (4 * kCols + 16) * sizeof(double)>>>(tmap, 5, 40, 17, 63, pad);
      const cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("box_rows=%d pad=%d: kernel failed: %s\n", box_rows, pad, cudaGetErrorString(e)); return 2; }
      std::vector<double> h(kRows * kCols);
      cudaMemcpy(h.data(), d, sizeof(double) * h.size(), cudaMemcpyDeviceToHost);
      int ok = 1, written = 0;
      const int rows[4] = {5, 40, 17, 63};
      for (int k = 0; k < 4; ++k) for (int c = 0; c < kCols; ++c) ok &= h[rows[k] * kCols + c] == 1000.0 * (k + 1) + c;
      for (double v : h) written += v != 0.0;
      printf("box_rows=%d pad=%d: rows %s, %d doubles written (expect %d)\n", box_rows, pad, ok ? "OK" : "WRONG", written, 4 * kCols - 0);
    }
  }
  // out-of-range row index: is the row skipped?
  return 0;
}
