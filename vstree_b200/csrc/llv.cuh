// llv.cuh -- ordered compaction of the lcp values >= 255 into (index, value) pairs (.llv), shared by the two
// finalize paths.  reference: src/Mkvtree/bese.c:557-566 (outlcpsubtab) + mkvprocess.c:214-237 (makelargelcp).
#pragma once

#include "scan.cuh"

constexpr int FN_THREADS = 256;
constexpr int FN_ITEMS = 8;
constexpr int FN_TILE = FN_THREADS * FN_ITEMS;
constexpr int LLV_SCAN_THREADS = 1024;

// scalars: [0] maxbranchdepth [1] longest [2] nllv
// single CTA: per-tile counts of large values -> exclusive offsets; total -> scalars[2]
static __global__ void __launch_bounds__(LLV_SCAN_THREADS)
large_scan_kernel(u32 *tile_large, u32 ntiles, unsigned long long *scalars)
{
  constexpr int ITEMS = 8;  // tiles per thread and step
  __shared__ u64 scratch[LLV_SCAN_THREADS / 32 + 1];
  u64 carry = 0;
  for (u32 b = 0; b < ntiles; b += LLV_SCAN_THREADS * ITEMS)
  {
    const u32 i0 = b + threadIdx.x * ITEMS;
    u32 v[ITEMS];
    u64 sum = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; k++)
    {
      v[k] = i0 + k < ntiles ? tile_large[i0 + k] : 0u;
      sum += v[k];
    }
    u64 total;
    u64 run = carry + block_exclusive_sum<u64, LLV_SCAN_THREADS>(sum, scratch, &total);
#pragma unroll
    for (int k = 0; k < ITEMS; k++)
    {
      if (i0 + k < ntiles)
      {
        tile_large[i0 + k] = (u32) run;
      }
      run += v[k];
    }
    carry += total;
  }
  if (threadIdx.x == 0)
  {
    scalars[2] = carry;
  }
}

// writes the (index, value) pairs in increasing index order.  The exact value of entry r comes either from the
// plcp array (v[sa[r]] - sa[r]) or from lcpfull[r] (direct path; only entries with lcp8 == 255 are defined).
static __global__ void __launch_bounds__(FN_THREADS)
llv_emit_kernel(u64 N, const u32 *__restrict__ sa, const u32 *__restrict__ v, const u32 *__restrict__ lcpfull,
                const u8 *__restrict__ lcp8, const u32 *__restrict__ tile_offset, u64 nllv, u64 *__restrict__ llv,
                u64 index_base /* global index of local entry 0 */)
{
  __shared__ __align__(8) u8 sflag[FN_TILE];
  __shared__ u32 scratch[FN_THREADS / 32 + 1];
  const u64 tile_base = (u64) blockIdx.x * FN_TILE;
  const u64 first = tile_offset[blockIdx.x];
  const u64 next = blockIdx.x + 1 < gridDim.x ? tile_offset[blockIdx.x + 1] : nllv;
  if (first == next)
  {
    return;  // uniform for the whole CTA
  }
#pragma unroll
  for (int j = 0; j < FN_ITEMS; j++)
  {
    const int l = j * FN_THREADS + threadIdx.x;
    const u64 r = tile_base + l;
    sflag[l] = (r < N && lcp8[r] == 255) ? 1 : 0;
  }
  __syncthreads();
  const int b0 = threadIdx.x * FN_ITEMS;
  const u64 f8 = *reinterpret_cast<const u64 *>(sflag + b0);
  const u32 cnt = (u32) __popcll(f8);
  u32 o = block_exclusive_sum<u32, FN_THREADS>(cnt, scratch, nullptr);
#pragma unroll
  for (int i = 0; i < FN_ITEMS; i++)
  {
    if ((f8 >> (8 * i)) & 1)
    {
      const u64 r = tile_base + b0 + i;
      u64 value;
      if (v != nullptr)
      {
        const u32 s = sa[r];
        value = (u64) (v[s] - s);
      } else
      {
        value = lcpfull[r];
      }
      llv[2 * (first + o)] = index_base + r;
      llv[2 * (first + o) + 1] = value;
      o++;
    }
  }
}
