// Dependent-chain latencies of the instructions the ViGO kernels are made of (B200, sm_100a).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a --fmad=false lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double* out, long long* cyc, double a, double b, int ia) {
  __shared__ double sm[1024];
  __shared__ int smi[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) { sm[i] = 1.0 + i * 1e-9; smi[i] = (i * 7 + 1) & 1023; }
  __syncthreads();
  double x = a; long long t0, t1; int idx = ia;
  int t = 0;
#define RUN(name, body)                                   \
  t0 = clock64();                                         \
  _Pragma("unroll 8") for (int i = 0; i < N; ++i) { body; } \
  t1 = clock64();                                         \
  if (threadIdx.x == 0) cyc[t] = t1 - t0;                 \
  ++t;
  RUN(dadd, x = x + b)
  RUN(dmul, x = x * b)
  RUN(dfma, x = __fma_rn(x, b, a))
  RUN(ddiv, x = a / x + b)
  RUN(dsqrt, x = sqrt(x) + b)
  RUN(lds, idx = smi[idx])
  RUN(ldsd, x = sm[((int)x) & 1023] + 1.0)
  RUN(shfl64, x = __shfl_xor_sync(0xffffffffu, x, 1))
  RUN(shfl32, idx = __shfl_xor_sync(0xffffffffu, idx, 1) + 1)
  RUN(dsetp, if (x > b) idx += 1; x = (double)idx)
  RUN(idiv, idx = idx / (ia | 3) + 1000000)
  RUN(bar, __syncthreads())
  out[threadIdx.x] = x + idx;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 8192); cudaMalloc(&c, 256);
  const char* names[] = {"dadd","dmul","dfma","ddiv+dadd","dsqrt+dadd","lds(int chase)","lds.f64+d2i+dadd","shfl64","shfl32+iadd","dsetp+i2d","idiv","bar.sync(128thr)"};
  for (int threads : {32, 128}) {
    k<<<1, threads>>>(o, c, 1.000001, 1.0000001, 5);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, c, 12 * 8, cudaMemcpyDeviceToHost);
    printf("threads=%d\n", threads);
    for (int i = 0; i < 12; ++i) printf("  %-22s %.1f cycles\n", names[i], (double)h[i] / N);
  }
  return 0;
}
