// textcmp.cuh -- random-access views of the transformed sequence for suffix comparison.
//
// Comparison semantics are the reference's (src/Mkvtree/bese.c:26-55, GETLCP :91-95): letters match by code,
// a special symbol (254/255, and everything behind position n) matches nothing, so an lcp stops there.
//   TextBytes   any alphabet: the tis bytes themselves, 8 symbols per 64-bit window.
//   TextPacked2 alphabets of at most 4 letters (DNA): 2 bits per symbol, 32 symbols per 64-bit window, plus a
//               1-bit special mask; 0.375 B/symbol keeps a 300 Mbp text inside the 126 MB L2, which is what turns
//               the random suffix-start accesses of the lcp and bwt kernels into L2 hits.
#pragma once

#include "common.cuh"

struct TextBytes
{
  const u8 *tis;  // padded with >= 64 bytes of 255 behind position n

  __device__ __forceinline__ u64 window(u64 p) const
  {
    const u64 *w = reinterpret_cast<const u64 *>(tis + (p & ~(u64) 7));
    const u32 sh = (u32) (p & 7) * 8;
    const u64 lo = w[0];
    if (sh == 0)
    {
      return lo;
    }
    return (lo >> sh) | (w[1] << (64 - sh));
  }
  struct Probe
  {
    u64 w;
    u32 before;
  };
  __device__ __forceinline__ Probe probe(u64 p) const
  {
    Probe r;
    r.w = window(p);
    r.before = before(p);
    return r;
  }
  static __device__ __forceinline__ u32 probe_stop(const Probe &a, const Probe &b)
  {
    const u64 diff = a.w ^ b.w;
    const u64 t = ~(a.w | 0x0101010101010101ull);
    const u64 z = (t - 0x0101010101010101ull) & ~t & 0x8080808080808080ull;
    const u32 stop_mis = diff ? (u32) (__ffsll((long long) diff) - 1) >> 3 : 8u;
    const u32 stop_sp = z ? (u32) (__ffsll((long long) z) - 1) >> 3 : 8u;
    return min(stop_mis, stop_sp);
  }
  static constexpr u32 WINDOW = 8;  // symbols per comparison step
  // equal leading letters inside one window pair: 0..8 (8 = no mismatch and no special in the window)
  __device__ __forceinline__ u32 step(u64 pa, u64 pb) const
  {
    const u64 x = window(pa), y = window(pb);
    const u64 diff = x ^ y;
    // special bytes of x: zero bytes of ~(x | 0x01..); an equal special on both sides does not match either
    const u64 t = ~(x | 0x0101010101010101ull);
    const u64 z = (t - 0x0101010101010101ull) & ~t & 0x8080808080808080ull;
    const u32 stop_mis = diff ? (u32) (__ffsll((long long) diff) - 1) >> 3 : 8u;
    const u32 stop_sp = z ? (u32) (__ffsll((long long) z) - 1) >> 3 : 8u;
    return min(stop_mis, stop_sp);
  }
  // number of equal leading letters of the suffixes at a and b (a != b), given that the first `start` match
  __device__ __forceinline__ u32 lcp(u64 a, u64 b, u32 start) const
  {
    u32 l = start;
    while (true)
    {
      const u32 stop = step(a + l, b + l);
      l += stop;
      if (stop < 8)
      {
        return l;
      }
    }
  }
  // symbol in front of position p as stored in tis (255 for p == 0: never equal to a letter)
  __device__ __forceinline__ u32 before(u64 p) const { return p > 0 ? tis[p - 1] : 255u; }
};

struct TextPacked2
{
  const u64 *pk;   // 32 symbols per word, symbol i of the word in bits [2i, 2i+2); specials stored as 0
  const u32 *spm;  // 1 bit per symbol: special, or at/behind position n
  const u8 *tis;   // the bytes, only consulted to tell 254 from 255

  __device__ __forceinline__ u64 window(u64 p) const
  {
    const u64 w = p >> 5;
    const u32 sh = (u32) (p & 31) * 2;
    const u64 lo = pk[w];
    if (sh == 0)
    {
      return lo;
    }
    return (lo >> sh) | (pk[w + 1] << (64 - sh));
  }
  __device__ __forceinline__ u32 mask_window(u64 p) const
  {
    const u64 w = p >> 5;
    return __funnelshift_r(spm[w], spm[w + 1], (u32) (p & 31));
  }
  // One suffix start, loaded once and used for both neighbour comparisons and for the bwt symbol: the first
  // window, its special mask and the symbol in front of the suffix (the lcp/bwt sweep is bound by the number of
  // L2 sector requests, not by bytes: 10 requests per suffix-array entry before this, about 3 with it).
  struct Probe
  {
    u64 w;
    u32 m;
    u32 before;
  };
  __device__ __forceinline__ Probe probe(u64 p) const
  {
    Probe r;
    const u64 wi = p >> 5;
    const u32 o = (u32) (p & 31);
    const u64 lo = pk[wi];
    const u32 mlo = spm[wi];
    if (o == 0)
    {
      r.w = lo;
      r.m = mlo;
      r.before = before(p);
    } else
    {
      r.w = (lo >> (2 * o)) | (pk[wi + 1] << (64 - 2 * o));
      r.m = __funnelshift_r(mlo, spm[wi + 1], o);
      r.before = ((mlo >> (o - 1)) & 1u) ? (u32) tis[p - 1] : (u32) (lo >> (2 * (o - 1))) & 3u;
    }
    return r;
  }
  static __device__ __forceinline__ u32 probe_stop(const Probe &a, const Probe &b)
  {
    const u64 diff = a.w ^ b.w;
    const u64 mm = (diff | (diff >> 1)) & 0x5555555555555555ull;
    const u32 sp = a.m | b.m;
    const u32 stop_mis = mm ? (u32) (__ffsll((long long) mm) - 1) >> 1 : 32u;
    const u32 stop_sp = sp ? (u32) (__ffs((int) sp) - 1) : 32u;
    return min(stop_mis, stop_sp);
  }
  static constexpr u32 WINDOW = 32;  // symbols per comparison step
  // equal leading letters inside one window pair: 0..32 (32 = no mismatch and no special in the window)
  __device__ __forceinline__ u32 step(u64 pa, u64 pb) const
  {
    const u64 diff = window(pa) ^ window(pb);
    const u64 mm = (diff | (diff >> 1)) & 0x5555555555555555ull;
    const u32 sp = mask_window(pa) | mask_window(pb);
    const u32 stop_mis = mm ? (u32) (__ffsll((long long) mm) - 1) >> 1 : 32u;
    const u32 stop_sp = sp ? (u32) (__ffs((int) sp) - 1) : 32u;
    return min(stop_mis, stop_sp);
  }
  __device__ __forceinline__ u32 lcp(u64 a, u64 b, u32 start) const
  {
    u32 l = start;
    while (true)
    {
      const u32 stop = step(a + l, b + l);
      l += stop;
      if (stop < 32)
      {
        return l;
      }
    }
  }
  __device__ __forceinline__ u32 before(u64 p) const
  {
    if (p == 0)
    {
      return 255u;
    }
    const u64 q = p - 1;
    if ((spm[q >> 5] >> (q & 31)) & 1u)
    {
      return tis[q];
    }
    return (u32) (pk[q >> 5] >> ((q & 31) * 2)) & 3u;
  }
};
