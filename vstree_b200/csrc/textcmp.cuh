// textcmp.cuh -- random-access views of the transformed sequence for suffix comparison.
//
// Comparison semantics are the reference's (src/Mkvtree/bese.c:26-55, GETLCP :91-95): letters match by code,
// a special symbol (254/255, and everything behind position n) matches nothing, so an lcp stops there.
//   TextBytes   any alphabet: the tis bytes themselves, 8 symbols per 64-bit window.
//   TextPacked2 alphabets of at most 4 letters (DNA): 2 bits per symbol in 16-byte records of 60 symbols (one load
//               per suffix start, 28 symbols per comparison step), plus a 1-bit special mask that is consulted only
//               when a record's summary bits say so; 0.625 B/symbol keeps a 200 Mbp text inside the 126 MB L2.
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
  static __device__ __forceinline__ Probe probe_from_lane_below(u32 mask, const Probe &x)
  {
    Probe r;
    r.w = __shfl_up_sync(mask, x.w, 1);
    r.before = __shfl_up_sync(mask, x.before, 1);
    return r;
  }
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

// Record layout of the 2-bit text.  One 16-byte record per 32 symbols, but every record carries 60 symbols: record i
// holds the text symbols [32i - 1, 32i + 59) (symbol j of the record in bits [2j, 2j+2) of the 128-bit word, specials
// stored as 0) and, in its top 8 bits, a summary: bit g is set when a special (or a position outside [0, n)) lies among
// the record symbols [8g, 8g + 8).  So ONE 16-byte load yields, for any position p, the symbol in front of p and a
// window of 28 symbols behind it, and tells whether the special mask has to be consulted at all.  The random accesses
// of the lcp/bwt sweep and of the text keys are bound by L1 wavefronts (one per distinct line and load instruction,
// about 136 G/s on a B200), not by bytes: the previous layout (64-bit words + separate mask words) needed four loads
// per suffix start, this one needs one.  Same footprint: 16 + 4 bytes per 32 symbols.
struct TextPacked2
{
  const u64 *pk;   // records, two words each (lo, hi)
  const u32 *spm;  // 1 bit per symbol: special, or at/behind position n (consulted only when a summary bit says so)
  const u8 *tis;   // the bytes, only consulted to tell 254 from 255

  static constexpr u32 WINDOW = 28;  // symbols per comparison step
  static constexpr u64 WINDOW_BITS_MASK = ((u64) 1 << (2 * WINDOW)) - 1;
  static constexpr u32 WINDOW_MASK = (1u << WINDOW) - 1;

  struct Rec
  {
    u64 lo, hi;
    u32 o;  // index of position p inside the record: 1..32
  };
  __device__ __forceinline__ Rec load(u64 p) const
  {
    const ulonglong2 r = __ldg(reinterpret_cast<const ulonglong2 *>(pk) + (p >> 5));
    Rec x;
    x.lo = r.x;
    x.hi = r.y;
    x.o = (u32) (p & 31) + 1;
    return x;
  }
  static __device__ __forceinline__ u64 rec_window(const Rec &x)
  {
    const u32 sh = 2 * x.o;  // 2..64
    const u64 w = sh < 64 ? ((x.lo >> sh) | (x.hi << (64 - sh))) : x.hi;
    return w & WINDOW_BITS_MASK;
  }
  // does a special lie among the record symbols [first, last]?
  static __device__ __forceinline__ bool rec_has_special(const Rec &x, u32 first, u32 last)
  {
    const u32 summary = (u32) (x.hi >> 56);
    const u32 g0 = first >> 3, g1 = last >> 3;
    return ((summary >> g0) & ((2u << (g1 - g0)) - 1)) != 0;
  }
  __device__ __forceinline__ u32 mask_slow(u64 p) const
  {
    const u64 w = p >> 5;
    return __funnelshift_r(spm[w], spm[w + 1], (u32) (p & 31)) & WINDOW_MASK;
  }
  // window of WINDOW symbols at p (symbol k in bits [2k, 2k+2), upper bits zero) and its special mask
  struct Win
  {
    u64 w;
    u32 m;
  };
  __device__ __forceinline__ Win win(u64 p) const
  {
    const Rec x = load(p);
    Win r;
    r.w = rec_window(x);
    r.m = rec_has_special(x, x.o, x.o + WINDOW - 1) ? mask_slow(p) : 0u;
    return r;
  }
  __device__ __forceinline__ u64 window(u64 p) const { return rec_window(load(p)); }
  __device__ __forceinline__ u32 mask_window(u64 p) const
  {
    const Rec x = load(p);
    return rec_has_special(x, x.o, x.o + WINDOW - 1) ? mask_slow(p) : 0u;
  }
  __device__ __forceinline__ u32 before_of(const Rec &x, u64 p) const
  {
    if (p == 0)
    {
      return 255u;
    }
    const u32 j = x.o - 1;
    if (rec_has_special(x, j, j))
    {
      const u64 q = p - 1;
      if ((spm[q >> 5] >> (q & 31)) & 1u)
      {
        return tis[q];
      }
    }
    return (u32) ((j < 32 ? x.lo >> (2 * j) : x.hi >> (2 * (j - 32))) & 3u);
  }
  // One suffix start, loaded once and used for both neighbour comparisons and for the bwt symbol: the first
  // window, its special mask and the symbol in front of the suffix.
  struct Probe
  {
    u64 w;
    u32 m;
    u32 before;
  };
  static __device__ __forceinline__ Probe probe_from_lane_below(u32 mask, const Probe &x)
  {
    Probe r;
    r.w = __shfl_up_sync(mask, x.w, 1);
    r.m = __shfl_up_sync(mask, x.m, 1);
    r.before = __shfl_up_sync(mask, x.before, 1);
    return r;
  }
  __device__ __forceinline__ Probe probe(u64 p) const
  {
    const Rec x = load(p);
    Probe r;
    r.w = rec_window(x);
    r.m = rec_has_special(x, x.o, x.o + WINDOW - 1) ? mask_slow(p) : 0u;
    r.before = before_of(x, p);
    return r;
  }
  static __device__ __forceinline__ u32 stop_of(u64 wa, u64 wb, u32 sp)
  {
    const u64 diff = wa ^ wb;
    const u64 mm = (diff | (diff >> 1)) & 0x5555555555555555ull;
    const u32 stop_mis = mm ? (u32) (__ffsll((long long) mm) - 1) >> 1 : WINDOW;
    const u32 stop_sp = sp ? (u32) (__ffs((int) sp) - 1) : WINDOW;
    return min(stop_mis, stop_sp);
  }
  static __device__ __forceinline__ u32 probe_stop(const Probe &a, const Probe &b)
  {
    return stop_of(a.w, b.w, a.m | b.m);
  }
  // equal leading letters inside one window pair: 0..WINDOW (WINDOW = no mismatch and no special in the window)
  __device__ __forceinline__ u32 step(u64 pa, u64 pb) const
  {
    const Win a = win(pa), b = win(pb);
    return stop_of(a.w, b.w, a.m | b.m);
  }
  __device__ __forceinline__ u32 lcp(u64 a, u64 b, u32 start) const
  {
    u32 l = start;
    while (true)
    {
      const u32 stop = step(a + l, b + l);
      l += stop;
      if (stop < WINDOW)
      {
        return l;
      }
    }
  }
  __device__ __forceinline__ u32 before(u64 p) const { return before_of(load(p), p); }
};

// 5-bit records for alphabets of 5 to 30 letters (protein: 20).  One 16-byte record per 16 positions that carries 24
// symbols: record i holds the text symbols [16i - 1, 16i + 23), symbol j in bits [5j, 5j + 5) of the 128-bit word; codes
// 0 .. sigma-1 are the letters, 30 the wildcard (254), 31 the separator (255) and everything outside [0, n).  ONE 16-byte
// load yields, for any position, the symbol in front of it and a window of 8 symbols behind it with the specials inline --
// the byte view needs two unaligned 8-byte loads for the window and a third load for the symbol in front, and these random
// accesses are bound by L1 wavefronts per load instruction, not by bytes (same reasoning as TextPacked2).  1 B/symbol.
struct TextPacked5
{
  const u64 *pk;  // records, two words each (lo, hi)

  static constexpr u32 WINDOW = 8;  // symbols per comparison step
  static constexpr u64 WINDOW_BITS_MASK = ((u64) 1 << (5 * WINDOW)) - 1;
  static constexpr u64 GROUP_BIT1 = 0x0842108421ull << 1;  // bit 1 of each of the 8 five-bit groups
  static constexpr u32 CODE_WILDCARD = 30, CODE_SEPARATOR = 31;

  struct Rec
  {
    u64 lo, hi;
    u32 o;  // index of position p inside the record: 1..16
  };
  __device__ __forceinline__ Rec load(u64 p) const
  {
    const ulonglong2 r = __ldg(reinterpret_cast<const ulonglong2 *>(pk) + (p >> 4));
    Rec x;
    x.lo = r.x;
    x.hi = r.y;
    x.o = (u32) (p & 15) + 1;
    return x;
  }
  static __device__ __forceinline__ u64 rec_window(const Rec &x)
  {
    const u32 sh = 5 * x.o;  // 5..80
    const u64 w = sh < 64 ? ((x.lo >> sh) | (x.hi << (64 - sh))) : (x.hi >> (sh - 64));
    return w & WINDOW_BITS_MASK;
  }
  static __device__ __forceinline__ u32 rec_symbol(const Rec &x, u32 j)  // record symbol j (0..23)
  {
    const u32 sh = 5 * j;
    u64 w;
    if (sh < 64)
    {
      w = x.lo >> sh;
      if (sh > 59)
      {
        w |= x.hi << (64 - sh);  // symbol 12 straddles the two words
      }
    } else
    {
      w = x.hi >> (sh - 64);
    }
    return (u32) (w & 31u);
  }
  __device__ __forceinline__ u64 window(u64 p) const { return rec_window(load(p)); }
  // symbol k of a window as a tis byte value (letters as they are, 254, 255)
  static __device__ __forceinline__ u32 window_symbol(u64 w, u32 k)
  {
    const u32 c = (u32) (w >> (5 * k)) & 31u;
    return c < CODE_WILDCARD ? c : (c == CODE_WILDCARD ? (u32) ESA_WILDCARD : (u32) ESA_SEPARATOR);
  }
  struct Probe
  {
    u64 w;
    u32 before;
  };
  static __device__ __forceinline__ Probe probe_from_lane_below(u32 mask, const Probe &x)
  {
    Probe r;
    r.w = __shfl_up_sync(mask, x.w, 1);
    r.before = __shfl_up_sync(mask, x.before, 1);
    return r;
  }
  static __device__ __forceinline__ u32 before_of(const Rec &x, u64 p)
  {
    if (p == 0)
    {
      return 255u;
    }
    const u32 c = rec_symbol(x, x.o - 1);
    return c < CODE_WILDCARD ? c : (c == CODE_WILDCARD ? (u32) ESA_WILDCARD : (u32) ESA_SEPARATOR);
  }
  __device__ __forceinline__ Probe probe(u64 p) const
  {
    const Rec x = load(p);
    Probe r;
    r.w = rec_window(x);
    r.before = before_of(x, p);
    return r;
  }
  // equal leading letters of two windows: 0..WINDOW; a special (codes 30, 31: the four high bits of the group are set)
  // on the first side ends the match even when the other side holds the same code
  static __device__ __forceinline__ u32 stop_of(u64 wa, u64 wb)
  {
    const u64 diff = wa ^ wb;
    const u64 sp = wa & (wa >> 1) & (wa >> 2) & (wa >> 3) & GROUP_BIT1;
    const u32 stop_mis = diff ? (u32) (__ffsll((long long) diff) - 1) / 5u : WINDOW;
    const u32 stop_sp = sp ? (u32) (__ffsll((long long) sp) - 1) / 5u : WINDOW;
    return min(stop_mis, stop_sp);
  }
  static __device__ __forceinline__ u32 probe_stop(const Probe &a, const Probe &b) { return stop_of(a.w, b.w); }
  __device__ __forceinline__ u32 step(u64 pa, u64 pb) const { return stop_of(window(pa), window(pb)); }
  __device__ __forceinline__ u32 lcp(u64 a, u64 b, u32 start) const
  {
    u32 l = start;
    while (true)
    {
      const u32 stop = step(a + l, b + l);
      l += stop;
      if (stop < WINDOW)
      {
        return l;
      }
    }
  }
  __device__ __forceinline__ u32 before(u64 p) const { return before_of(load(p), p); }
};
