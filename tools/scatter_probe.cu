// scatter_probe.cu -- what does HBM deliver for the ACCESS PATTERN of a onesweep radix pass, without any of its compute?
// One coalesced read of (key, value) pairs, one write in which every tile of TILE pairs is cut into 256 runs that go to
// 256 different output streams (run length TILE/256 elements = 128 bytes at TILE = 8192).  Compared with a plain copy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/scatter_probe tools/scatter_probe.cu && build/scatter_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32;
typedef uint64_t u64;

template <int ITEMS>
__global__ void __launch_bounds__(512, 2) scatter_kernel(const u32 *__restrict__ kin, const u32 *__restrict__ vin,
                                                         u32 *__restrict__ kout, u32 *__restrict__ vout, u32 ntiles, u32 streams)
{
  constexpr int TILE = 512 * ITEMS;
  const u32 tile = blockIdx.x;
  const u32 run = TILE / streams;  // elements per run
  const u64 stream_len = (u64) ntiles * run;
#pragma unroll
  for (int j = 0; j < ITEMS; j++)
  {
    const u32 p = j * 512 + threadIdx.x;
    const u32 d = p / run, o = p % run;
    const u64 src = (u64) tile * TILE + p;
    const u64 dst = (u64) d * stream_len + (u64) tile * run + o;
    kout[dst] = kin[src];
    vout[dst] = vin[src];
  }
}

__global__ void copy_kernel(const uint4 *__restrict__ a, uint4 *__restrict__ b, u64 n16)
{
  for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (u64) gridDim.x * blockDim.x)
  {
    b[i] = a[i];
  }
}

int main()
{
  const u64 n = 100663296;  // 12288 tiles of 8192
  u32 *kin, *vin, *kout, *vout;
  cudaMalloc(&kin, n * 4), cudaMalloc(&vin, n * 4), cudaMalloc(&kout, n * 4), cudaMalloc(&vout, n * 4);
  cudaMemset(kin, 1, n * 4), cudaMemset(vin, 2, n * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  auto time = [&](const char *name, auto launch) {
    float best = 1e9;
    for (int r = 0; r < 5; r++)
    {
      cudaEventRecord(e0);
      launch();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    printf("%-44s %.3f ms  %.0f GB/s (read + write of %.0f MB)  err=%s\n", name, best, 16.0 * n / best / 1e6, 16.0 * n / 1e6,
           cudaGetErrorString(cudaGetLastError()));
  };
  time("plain copy (uint4, 148*8 CTAs)", [&] {
    copy_kernel<<<148 * 8, 512>>>((const uint4 *) kin, (uint4 *) kout, n / 4);
    copy_kernel<<<148 * 8, 512>>>((const uint4 *) vin, (uint4 *) vout, n / 4);
  });
  time("tile 8192 -> 256 streams (128 B runs)", [&] { scatter_kernel<16><<<n / 8192, 512>>>(kin, vin, kout, vout, n / 8192, 256); });
  time("tile 8192 -> 128 streams (256 B runs)", [&] { scatter_kernel<16><<<n / 8192, 512>>>(kin, vin, kout, vout, n / 8192, 128); });
  time("tile 8192 -> 64 streams (512 B runs)", [&] { scatter_kernel<16><<<n / 8192, 512>>>(kin, vin, kout, vout, n / 8192, 64); });
  time("tile 8192 -> 16 streams (2 KB runs)", [&] { scatter_kernel<16><<<n / 8192, 512>>>(kin, vin, kout, vout, n / 8192, 16); });
  time("tile 8192 -> 1 stream (plain tile copy)", [&] { scatter_kernel<16><<<n / 8192, 512>>>(kin, vin, kout, vout, n / 8192, 1); });
  time("tile 8192 -> 512 streams (64 B runs)", [&] { scatter_kernel<16><<<n / 8192, 512>>>(kin, vin, kout, vout, n / 8192, 512); });
  time("tile 8192 -> 2048 streams (16 B runs)", [&] { scatter_kernel<16><<<n / 8192, 512>>>(kin, vin, kout, vout, n / 8192, 2048); });
  time("tile 12288 -> 256 streams (192 B runs)", [&] { scatter_kernel<24><<<n / 12288, 512>>>(kin, vin, kout, vout, n / 12288, 256); });
  return 0;
}
