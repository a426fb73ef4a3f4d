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
  // warm-up over the h-1 symbols behind the thread's last position
  for (int q = lo + KS_ITEMS - 1 + (int) kp.h - 1; q >= lo + KS_ITEMS; q--)
  {
    const u32 a = stext[q];
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
#pragma unroll
  for (int i = KS_ITEMS - 1; i >= 0; i--)
  {
    const u32 a = stext[lo + i];
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
