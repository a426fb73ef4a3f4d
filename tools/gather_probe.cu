// Probe: cost of random 4-byte gathers/scatters over an array much larger than L2 under different
// cudaLimitMaxL2FetchGranularity settings.  nvcc -arch=sm_100a -O3 tools/gather_probe.cu -o /tmp/gather_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__global__ void init(uint32_t *idx, uint64_t n, uint64_t range) {
  uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x;
  if (i < n) { uint64_t x = i * 0x9E3779B97F4A7C15ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; idx[i] = (uint32_t) (x % range); }
}
__global__ void gather(const uint32_t *__restrict__ idx, const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, uint64_t n) {
  uint64_t i = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n) { uint4 j = *(const uint4 *) (idx + i); uint4 o; o.x = src[j.x]; o.y = src[j.y]; o.z = src[j.z]; o.w = src[j.w]; *(uint4 *) (dst + i) = o; }
}
__global__ void scatter(const uint32_t *__restrict__ idx, uint32_t *__restrict__ dst, uint64_t n) {
  uint64_t i = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n) { uint4 j = *(const uint4 *) (idx + i); dst[j.x] = i; dst[j.y] = i + 1; dst[j.z] = i + 2; dst[j.w] = i + 3; }
}
__global__ void gather_bytes(const uint32_t *__restrict__ idx, const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint64_t n) {
  uint64_t i = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n) { uint4 j = *(const uint4 *) (idx + i); uchar4 o; o.x = src[j.x]; o.y = src[j.y]; o.z = src[j.z]; o.w = src[j.w]; *(uchar4 *) (dst + i) = o; }
}
int main() {
  const uint64_t n = 100000000;
  uint32_t *idx, *src, *dst; uint8_t *bsrc, *bdst;
  cudaMalloc(&idx, n * 4); cudaMalloc(&src, n * 4); cudaMalloc(&dst, n * 4); cudaMalloc(&bsrc, n); cudaMalloc(&bdst, n);
  cudaMemset(src, 1, n * 4); cudaMemset(bsrc, 1, n);
  init<<<(n + 255) / 256, 256>>>(idx, n, n);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  size_t lims[4] = {0, 32, 64, 128};
  for (int t = 0; t < 4; t++) {
    if (lims[t]) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, lims[t]); if (e) printf("setlimit %zu: %s\n", lims[t], cudaGetErrorString(e)); }
    size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    float ms[3];
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(a); gather<<<(n / 4 + 255) / 256, 256>>>(idx, src, dst, n); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms[0], a, b);
      cudaEventRecord(a); scatter<<<(n / 4 + 255) / 256, 256>>>(idx, dst, n); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms[1], a, b);
      cudaEventRecord(a); gather_bytes<<<(n / 4 + 255) / 256, 256>>>(idx, bsrc, bdst, n); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms[2], a, b);
    }
    printf("limit=%zu  gather4 %.3f ms  scatter4 %.3f ms  gather1(100MB src) %.3f ms   (100M elements each)\n", got, ms[0], ms[1], ms[2]);
  }
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
