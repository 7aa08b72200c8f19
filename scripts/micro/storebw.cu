// Micro-benchmark 3: per-SM global store rate: st.global.v4 from registers vs cp.async.bulk (TMA) from smem.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) st_regs(float4* x, int tiles, int stride, long long* out) {
  const int r = threadIdx.x; long long tot = 0;
  for (int t = 0; t < tiles; ++t) {
    float4* xr = x + (size_t)((blockIdx.x * 131 + t * stride) % 12000) * 72 * 128 + r;
    long long t0 = clock64();
#pragma unroll
    for (int k = 0; k < 72; ++k) xr[(size_t)k * 128] = make_float4(k, t, r, 1.f);
    tot += clock64() - t0;
    long long tw = clock64(); while (clock64() - tw < 20000) {}
  }
  if (r == 0) out[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(128, 1) st_bulk(float4* x, int tiles, int stride, long long* out, int chunk_bytes) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int r = threadIdx.x; long long tot = 0;
  for (int i = r; i < 147456 / 16; i += 128) reinterpret_cast<float4*>(sm)[i] = make_float4(i, 1, 2, 3);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  for (int t = 0; t < tiles; ++t) {
    uint8_t* dst = reinterpret_cast<uint8_t*>(x + (size_t)((blockIdx.x * 131 + t * stride) % 12000) * 72 * 128);
    long long t0 = clock64();
    if (r == 0) {
      for (int off = 0; off < 147456; off += chunk_bytes)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(s32(sm + off)), "r"(chunk_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    tot += clock64() - t0;
    long long tw = clock64(); while (clock64() - tw < 20000) {}
  }
  if (r == 0) out[blockIdx.x] = tot;
}
int main() {
  float4* x; cudaMalloc(&x, (size_t)12000 * 72 * 128 * 16);
  long long* out; cudaMallocManaged(&out, 256 * 8);
  cudaFuncSetAttribute(st_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (int grid : {1, 148}) {
    st_regs<<<grid, 128>>>(x, 6, 997, out); cudaDeviceSynchronize();
    double m = 0; for (int i = 0; i < grid; ++i) m += out[i]; m /= grid;
    printf("st.global.v4     grid %3d: %.0f cycles/147KB -> %.1f B/cycle/SM\n", grid, m / 6, 147456.0 * 6 / m);
    for (int cb : {4096, 16384, 147456}) {
      st_bulk<<<grid, 128, 147456>>>(x, 6, 997, out, cb); cudaDeviceSynchronize();
      m = 0; for (int i = 0; i < grid; ++i) m += out[i]; m /= grid;
      printf("bulk s2g %6d B  grid %3d: %.0f cycles/147KB -> %.1f B/cycle/SM\n", cb, grid, m / 6, 147456.0 * 6 / m);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
