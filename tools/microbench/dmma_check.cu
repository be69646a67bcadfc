#include <cstdio>
__global__ void k(double* out) {
  int l = threadIdx.x;
  double a = (l / 4) * 10 + (l % 4);      // A[row=l/4][k=l%4]
  double b = (l % 4) + 100 * (l / 4);     // B[k=l%4][col=l/4]
  double c0 = 0, c1 = 0;
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  out[2 * l] = c0; out[2 * l + 1] = c1;
}
int main() { double* o; cudaMalloc(&o, 512); k<<<1, 32>>>(o); double h[64]; cudaMemcpy(h, o, 512, cudaMemcpyDeviceToHost);
  // expected C[r][c] = sum_k (10r+k)*(k+100c)
  int bad = 0;
  for (int l = 0; l < 32; ++l) for (int i = 0; i < 2; ++i) { int r = l / 4, c = 2 * (l % 4) + i; double e = 0; for (int kk = 0; kk < 4; ++kk) e += (10.0 * r + kk) * (kk + 100.0 * c); if (e != h[2 * l + i]) ++bad; }
  printf("dmma m8n8k4 layout check: %s\n", bad ? "MISMATCH" : "ok"); return bad; }
