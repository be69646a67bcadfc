// Where do the warps of 1-warp thread blocks land?  Prints, for blocks of 1 / 2 / 4 warps at several residencies, the
// histogram of %warpid % 4 (the SM sub-partition of a warp slot) and the time of an FP64-throughput-bound loop — if
// 1-warp blocks all sat on one sub-partition the loop would take 4x as long as with 4-warp blocks.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a smsp_probe.cu -o smsp_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int* hist, int iters) {
  unsigned wid, smid;
  asm volatile("mov.u32 %0, %warpid;" : "=r"(wid));
  asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
  if ((threadIdx.x & 31) == 0) atomicAdd(&hist[(smid * 4 + (wid & 3)) % (256 * 4)], 1);
  double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
  double* o; int* h;
  cudaMalloc(&o, 8 * 148 * 64 * 128); cudaMalloc(&h, 4096 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int wpb : {1, 2, 4})
    for (int wps : {4, 8, 12, 16}) {
      const int blocks = 148 * wps / wpb;
      cudaMemset(h, 0, 4096 * 4);
      k<<<blocks, 32 * wpb>>>(o, h, 1000);
      cudaDeviceSynchronize();
      cudaMemset(h, 0, 4096 * 4);
      cudaEventRecord(e0);
      k<<<blocks, 32 * wpb>>>(o, h, 200000);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      int hh[1024]; cudaMemcpy(hh, h, 4096, cudaMemcpyDeviceToHost);
      int mx = 0, mn = 1 << 30, tot[4] = {0, 0, 0, 0};
      for (int s = 0; s < 148; ++s) for (int q = 0; q < 4; ++q) { int v = hh[s * 4 + q]; tot[q] += v; if (v > mx) mx = v; if (v < mn) mn = v; }
      printf("warps/block %d, warps/SM %2d: %.2f ms; per-SMSP warps min %d max %d; totals %d %d %d %d\n", wpb, wps, ms, mn, mx, tot[0], tot[1], tot[2], tot[3]);
    }
  return 0;
}
