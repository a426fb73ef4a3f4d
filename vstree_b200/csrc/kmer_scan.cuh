// kmer_scan.cuh -- rolling padded h-mer codes and "distance to the next special" for a tile of text positions.
//
// Semantics restated from the reference's bucketsortsuffixes (src/Mkvtree/ppsort.c:134-153,192-217): scanning
// right to left, code = code/sigma + a*sigma^(h-1); at a special symbol (>= 254; the sentinel behind the text
// counts as one) the code is reset to sigma^h-1, i.e. the h-mer is padded with the largest letter.  dist is the
// number of letters before the first special, capped at h: dist == 0 special-initial suffix, 0 < dist < h suffix
// with a special inside its h-prefix, dist == h full h-mer.
#pragma once

#include "common.cuh"

constexpr int KS_THREADS = 256;
constexpr int KS_ITEMS = 16;
constexpr int KS_TILE = KS_THREADS * KS_ITEMS;  // text positions per CTA
constexpr int KS_HALO = 32;                     // >= h0 - 1 look-ahead symbols (h0 <= 31)

struct KmerParams
{
  u32 sigma;
  u32 sigma_shift;  // log2(sigma) when sigma is a power of two, else 0
  u32 h;            // symbols per code
  u32 top_power;    // sigma^(h-1)
  u32 all_pad;      // sigma^h - 1
};

// Cooperative load of the tile's symbols [tile_base, tile_base + KS_TILE + KS_HALO) into shared memory;
// positions >= n read as 255 (sentinel and beyond).  `tis` is padded so that 16-byte loads up to
// round_up(n,16)+16 are legal.
__device__ __forceinline__ void ks_load_tile(const u8 *__restrict__ tis, u64 n, u64 tile_base, u8 *stext)
{
  constexpr int VEC = (KS_TILE + KS_HALO) / 16;
  for (int v = threadIdx.x; v < VEC; v += KS_THREADS)
  {
    const u64 p = tile_base + (u64) v * 16;
    uint4 w;
    if (p + 16 <= n)
    {
      w = *reinterpret_cast<const uint4 *>(tis + p);
    } else
    {
      u8 tmp[16];
#pragma unroll
      for (int b = 0; b < 16; b++)
      {
        tmp[b] = (p + b < n) ? tis[p + b] : (u8) 255;
      }
      w = *reinterpret_cast<uint4 *>(tmp);
    }
    *reinterpret_cast<uint4 *>(stext + v * 16) = w;
  }
}

// Rolling scan of one thread's KS_ITEMS consecutive positions (local offsets lo .. lo+KS_ITEMS-1 in stext).
// code[i], dist[i] describe the suffix at local offset lo + i.
__device__ __forceinline__ void ks_scan_thread(const u8 *stext, int lo, const KmerParams &kp, u32 *code, u32 *dist)
{
  u32 c = kp.all_pad, d = 0;
  // the thread's 16 symbols and the 32 behind them as three 16-byte shared-memory loads (conflict free; byte loads
  // at a 16-byte stride were 4-way bank conflicts: 1.1e8 conflicts per 100 Mbp in the round-1 profile)
  u32 w[12];
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(stext + lo);
    const uint4 x0 = src[0], x1 = src[1], x2 = src[2];
    w[0] = x0.x, w[1] = x0.y, w[2] = x0.z, w[3] = x0.w;
    w[4] = x1.x, w[5] = x1.y, w[6] = x1.z, w[7] = x1.w;
    w[8] = x2.x, w[9] = x2.y, w[10] = x2.z, w[11] = x2.w;
  }
  // warm-up over the h-1 symbols behind the thread's last position
#pragma unroll
  for (int j = KS_HALO - 2; j >= 0; j--)
  {
    if (j < (int) kp.h - 1)
    {
      const int q = KS_ITEMS + j;
      const u32 a = (w[q >> 2] >> ((q & 3) * 8)) & 0xffu;
      if (a >= ESA_WILDCARD)
      {
        c = kp.all_pad;
        d = 0;
      } else
      {
        c = (kp.sigma_shift ? (c >> kp.sigma_shift) : (c / kp.sigma)) + a * kp.top_power;
        d = min(d + 1, kp.h);
      }
    }
  }
#pragma unroll
  for (int i = KS_ITEMS - 1; i >= 0; i--)
  {
    const u32 a = (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
    if (a >= ESA_WILDCARD)
    {
      c = kp.all_pad;
      d = 0;
    } else
    {
      c = (kp.sigma_shift ? (c >> kp.sigma_shift) : (c / kp.sigma)) + a * kp.top_power;
      d = min(d + 1, kp.h);
    }
    code[i] = c;
    dist[i] = d;
  }
}

static inline KmerParams make_kmer_params(u32 sigma, u32 h)
{
  KmerParams kp;
  kp.sigma = sigma;
  kp.sigma_shift = 0;
  if (sigma >= 2 && (sigma & (sigma - 1)) == 0)
  {
    u32 s = 0;
    while ((1u << s) < sigma)
    {
      s++;
    }
    kp.sigma_shift = s;
  }
  kp.h = h;
  u64 p = 1;
  for (u32 i = 1; i < h; i++)
  {
    p *= sigma;
  }
  kp.top_power = (u32) p;
  kp.all_pad = (u32) (p * sigma - 1);
  return kp;
}
