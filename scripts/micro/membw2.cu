// Micro-benchmark 2: COLD per-SM streaming of 147 KB tiles scattered over a large buffer
// (the fused FFN kernel's residual hand-over pattern), vs. loads in flight and grid size.
#include <cstdio>
#include <cuda_runtime.h>
template <int DEPTH>
__global__ void __launch_bounds__(128, 1) rd(const float4* __restrict__ x, int tiles_per_cta, int stride_tiles, long long* out, float* sink) {
  const int r = threadIdx.x;
  float acc = 0.f;
  long long tot = 0;
  for (int t = 0; t < tiles_per_cta; ++t) {
    const float4* xr = x + (size_t)((blockIdx.x * 131 + t * stride_tiles) % 12000) * 72 * 128 + r;
    long long t0 = clock64();
    for (int h = 0; h < 72 / DEPTH; ++h) {
      float4 b[DEPTH];
#pragma unroll
      for (int k = 0; k < DEPTH; ++k) b[k] = xr[(size_t)(h * DEPTH + k) * 128];
#pragma unroll
      for (int k = 0; k < DEPTH; ++k) acc += b[k].x + b[k].w;
    }
    tot += clock64() - t0;
    // idle gap so tiles are not back-to-back (lets the memory system go quiet, like the real kernel)
    long long tw = clock64(); while (clock64() - tw < 20000) {}
  }
  if (r == 0) out[blockIdx.x] = tot;
  if (acc == 12345.f) *sink = acc;
}
__global__ void flush(float4* p, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; for (; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_float4(1, 2, 3, 4); }
int main() {
  const size_t ntiles = 12000;  // 1.77 GB
  float4* x; cudaMalloc(&x, ntiles * 72 * 128 * 16);
  long long* out; cudaMallocManaged(&out, 256 * 8);
  float* sink; cudaMalloc(&sink, 4);
  auto run = [&](int depth, int grid) {
    flush<<<592, 256>>>(x, ntiles * 72 * 128); cudaDeviceSynchronize();
    const int tpc = 6;
    if (depth == 4) rd<4><<<grid, 128>>>(x, tpc, 997, out, sink);
    else if (depth == 12) rd<12><<<grid, 128>>>(x, tpc, 997, out, sink);
    else rd<36><<<grid, 128>>>(x, tpc, 997, out, sink);
    cudaDeviceSynchronize();
    double m = 0; for (int i = 0; i < grid; ++i) m += out[i]; m /= grid;
    printf("cold read depth %2d (%.0f KB in flight) grid %3d: %.0f cycles/tile -> %.1f B/cycle/SM\n", depth, depth * 2.0, grid, m / tpc, 147456.0 * tpc / m);
  };
  for (int grid : {1, 16, 148}) for (int depth : {4, 12, 36}) run(depth, grid);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
