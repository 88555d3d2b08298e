// HBM bandwidth by read:write mix (development probe; result kept in profiles/r02_hbm_mix.txt).
// Each thread streams 16-byte vectors: R input streams are read and summed, W output streams are written.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probe_hbm_mix scripts/probe_hbm_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int R, int W>
__global__ void mix(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {  // n = vectors per stream
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint4 a = make_uint4(1, 2, 3, 4);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      uint4 v = __ldcs(in + r * n + i);
      a.x += v.x; a.y ^= v.y; a.z += v.z; a.w ^= v.w;
    }
    if (W == 0) { if (a.x == 0x12345678u && a.y == 77u) out[i] = a; }
#pragma unroll
    for (int w = 0; w < W; ++w) __stcs(out + w * n + i, a);
  }
}
template <int R, int W>
void run(uint4* in, uint4* out, size_t n) {
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  float best = 1e9f;
  for (int it = 0; it < 8; ++it) {
    cudaEventRecord(s);
    mix<R, W><<<148 * 16, 512>>>(in, out, n);
    cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e);
    if (it >= 2 && ms < best) best = ms;
  }
  printf("read %d : write %d   %8.1f GB/s total  (%.1f read, %.1f write)\n", R, W, (R + W) * n * 16.0 / best / 1e6,
         R * n * 16.0 / best / 1e6, W * n * 16.0 / best / 1e6);
}
int main() {
  const size_t n = (size_t)512 << 20 >> 4;  // 512 MiB per stream
  uint4 *in, *out;
  cudaMalloc(&in, 4 * n * 16); cudaMalloc(&out, 4 * n * 16);
  cudaMemset(in, 1, 4 * n * 16);
  run<1, 0>(in, out, 4 * n); run<0, 1>(in, out, 4 * n); run<1, 1>(in, out, 4 * n);
  run<1, 3>(in, out, n); run<1, 2>(in, out, n); run<2, 1>(in, out, n); run<3, 1>(in, out, n); run<1, 4>(in, out, n);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
