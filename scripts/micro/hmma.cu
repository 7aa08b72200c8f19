// Micro-benchmark: mma.sync m16n8k16 bf16 throughput per SM on sm_100a (legacy tensor path).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, long long* cyc) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  long long* cyc; cudaMallocManaged(&cyc, 148 * 8);
  for (int warps : {1, 4, 8, 16}) {
    k<<<148, warps * 32>>>(out, 2000, cyc); cudaDeviceSynchronize();
    double c = cyc[0];
    printf("warps/SM %2d: %.1f cycles per HMMA per warp; %.2f cycles per HMMA per SM -> %.0f MAC/cycle/SM\n", warps, c / (2000 * 8), c / (2000.0 * 8 * warps), 4096.0 * 2000 * 8 * warps / c);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
