// TMEM as lane-private scratch: allocation by several co-resident blocks, tcgen05.st / tcgen05.ld round trip with
// run-time column addresses, and the latency of a dependent ld / st+wait chain.  Development probe for the
// TMEM-resident L-BFGS history of tp_lbfgs_fast.cuh.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tmem_rt.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_st1(uint32_t taddr, double v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__double2loint(v)), "r"(__double2hiint(v)) : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_ld8(uint32_t taddr, double (&v)[8]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}

__global__ void __launch_bounds__(128, 4) k_tmem(int ncols, int rounds, int* errors, long long* cyc) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(&slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
  int bad = 0;
  for (int r = 0; r < rounds; ++r) {
    // run-time column order: write doubles d = 0 .. ncols/2-1 in a rotated order
    for (int d0 = 0; d0 < ncols / 2; ++d0) {
      const int d = (d0 + r + blockIdx.x) % (ncols / 2);
      tm_st1(base + 2 * d, (double)(blockIdx.x * 1000003 + tid * 1009 + d * 7 + r) + 0.25);
    }
    tm_wait_st();
    for (int c = 0; c < ncols; c += 16) {
      double v[8];
      tm_ld8(base + c, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int d = c / 2 + k;
        if (v[k] != (double)(blockIdx.x * 1000003 + tid * 1009 + d * 7 + r) + 0.25) ++bad;
      }
    }
    __syncthreads();   // other blocks on the SM keep running: exercises co-residency
  }
  // latency probes (one block only reports)
  long long t0 = clock64();
  double acc = 0.0;
  uint32_t a = base;
  for (int it = 0; it < 256; ++it) {
    double v[8];
    tm_ld8(a, v);
    acc += v[0];
    a = base + ((__double2loint(acc) & 1) ? 16 : 0) * 0 + (it & 3) * 16;   // address depends on the loaded value
  }
  long long t1 = clock64();
  for (int it = 0; it < 256; ++it) {
    tm_st1(base + 2 * (it & 31), acc + it);
    tm_wait_st();
  }
  long long t2 = clock64();
  // throughput: 8 x (x16 loads) = all 128 columns, summed
  double s2 = 0.0;
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int c = 0; c < 128; c += 16) {
      double v[8];
      tm_ld8(base + c, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) s2 += v[k];
    }
  }
  long long t3 = clock64();
  if (blockIdx.x == 0 && tid == 0) { cyc[0] = (t1 - t0) / 256; cyc[1] = (t2 - t1) / 256; cyc[2] = (t3 - t2) / 64; }
  if (acc + s2 == 1.2345) errors[1] = 1;
  if (bad) atomicAdd(errors, bad);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(ncols) : "memory");
}

int main() {
  int* err; long long* cyc;
  cudaMallocManaged(&err, 8); cudaMallocManaged(&cyc, 32);
  err[0] = err[1] = 0;
  for (int grid : {1, 148 * 4, 148 * 8}) {
    cyc[0] = cyc[1] = cyc[2] = 0;
    k_tmem<<<grid, 128>>>(128, 8, err, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("grid %d: %s errors %d | cycles: ld x16+wait (dependent) %lld, st x2+wait %lld, 8 x ld x16 (128 cols) %lld\n", grid,
           cudaGetErrorString(e), err[0], cyc[0], cyc[1], cyc[2]);
  }
  return err[0] != 0;
}
