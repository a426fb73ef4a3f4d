// childtab_core.cuh -- the child table (.cld / .cld1) as independent per-entry searches.
//
// The reference fills the table with two sequential stack walks over the lcp table (mkcldtabnextlIndex and
// mkcldtabupdown, src/kurtz/mkcld.c:41-207) and folds the three fields into one byte per entry (compresscldtab,
// mkcld.c:225-284).  What the walks leave behind has a closed form per entry; L is the exact lcp table with L[0] = 0
// (the walks never read lcptab[0], they start from a stack entry of depth 0 at index 0), n = totallength:
//
//   j(i)        = the first index > i with L[j] <= L[i]                                   (none: off the table)
//   nextlIndex  = j - i   when j exists and L[j] == L[i]                                  (mkcld.c:62-81)
//   down        = m - i   when j exists and j > i + 1, m = leftmost minimum of L in (i, j) (mkcld.c:123-147)
//   up          = i - m   when L[i-1] > L[i], q = the last index < i with L[q] <= L[i],
//                 m = leftmost minimum of L in (q, i)                                      (mkcld.c:148-171)
//   each clamped to LARGECHILDVALUE = 255, UNDEFCHILDVALUE = 0 otherwise                   (virtualdef.h:165-166)
//
// (an entry is on the reference's stack at step i exactly when nothing after it is smaller; the entry popped last by
// L[i] is therefore the leftmost minimum of the stretch that L[i] closes.)
//
// The searches run on a min-pyramid over L (level k+1 = minimum of FAN consecutive entries of level k): next / previous
// smaller-or-equal look at the rest of the group of siblings, climb at its border and descend into the first block whose
// minimum qualifies; the leftmost minimum of a range collects the two fringe groups of every level.  A group of siblings
// is fetched as a whole, so a search pays one memory latency per level.
//
// Plain C++ on purpose: the same functions are compiled by nvcc for the device (childtab.cu) and by g++ into a model
// that tests/test_childtab_model.py runs against the oracle's stack walk on the CPU.  They are not a product path on
// the host: libvstree_b200 only ever calls them from kernels.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define CLD_HD __host__ __device__ __forceinline__
#else
#define CLD_HD static inline
#endif

namespace cld
{

constexpr uint32_t FAN = 16;
constexpr int MAX_LEVELS = 10;  // 16^9 > 2^32
constexpr uint64_t NONE = ~(uint64_t) 0;
constexpr uint32_t LARGE = 255;  // LARGECHILDVALUE

struct Pyramid
{
  const uint32_t *level[MAX_LEVELS];
  uint64_t size[MAX_LEVELS];
  int top;  // index of the highest level (at most FAN entries, or MAX_LEVELS - 1)
};

// One group of FAN siblings (FAN * 4 = 64 bytes, aligned) in one go: on the device up to four 128-bit loads that are
// in flight together, so that a search pays one memory latency per level instead of one per sibling.  Only the
// quarters that overlap the positions [from, to] the caller is going to look at are fetched (the searches of a
// repeat-rich table are bound by L2 bandwidth, not by latency); everything else, and entries beyond the level's size
// (the arrays are padded to whole groups; real values stay below 0xffffffff), reads as 0xffffffff.
CLD_HD void load_group(const uint32_t *lev, uint64_t size, uint64_t base, uint32_t from, uint32_t to, uint32_t *g)
{
#ifdef __CUDA_ARCH__
  const uint4 *q = reinterpret_cast<const uint4 *>(lev + base);
  const uint4 none = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  const uint4 a = from < 4 ? __ldg(q) : none;
  const uint4 b = (from < 8 && to >= 4) ? __ldg(q + 1) : none;
  const uint4 c = (from < 12 && to >= 8) ? __ldg(q + 2) : none;
  const uint4 d = to >= 12 ? __ldg(q + 3) : none;
  g[0] = a.x, g[1] = a.y, g[2] = a.z, g[3] = a.w;
  g[4] = b.x, g[5] = b.y, g[6] = b.z, g[7] = b.w;
  g[8] = c.x, g[9] = c.y, g[10] = c.z, g[11] = c.w;
  g[12] = d.x, g[13] = d.y, g[14] = d.z, g[15] = d.w;
#pragma unroll
  for (uint32_t t = 0; t < FAN; t++)
  {
    if (base + t >= size)
    {
      g[t] = 0xffffffffu;
    }
  }
#else
  for (uint32_t t = 0; t < FAN; t++)
  {
    const bool fetched = t / 4 >= from / 4 && t / 4 <= to / 4;
    g[t] = fetched && base + t < size ? lev[base + t] : 0xffffffffu;
  }
#endif
}
static_assert(FAN == 16, "load_group is written for groups of 16");

// first position c >= from of the group with g[c] <= v; FAN when there is none
CLD_HD uint32_t group_first_le(const uint32_t *g, uint32_t from, uint32_t v)
{
  uint32_t pos = FAN;
#pragma unroll
  for (uint32_t t = 0; t < FAN; t++)
  {
    if (pos == FAN && t >= from && g[t] <= v)
    {
      pos = t;
    }
  }
  return pos;
}

// last position c <= upto of the group with g[c] <= v; FAN when there is none
CLD_HD uint32_t group_last_le(const uint32_t *g, uint32_t upto, uint32_t v)
{
  uint32_t pos = FAN;
#pragma unroll
  for (uint32_t t = 0; t < FAN; t++)
  {
    if (t <= upto && g[t] <= v)
    {
      pos = t;
    }
  }
  return pos;
}

// first index > i with a value <= v; size[0] when there is none
CLD_HD uint64_t next_le(const Pyramid &py, uint64_t i, uint32_t v)
{
  int k = 0;
  uint64_t p = i + 1;
  uint32_t g[FAN];
  while (true)
  {
    if (p >= py.size[k])
    {
      return py.size[0];
    }
    const uint64_t base = p - p % FAN;
    load_group(py.level[k], py.size[k], base, (uint32_t) (p - base), FAN - 1, g);
    const uint32_t c = group_first_le(g, (uint32_t) (p - base), v);
    if (c < FAN)
    {
      p = base + c;
      while (k > 0)  // down to the first leaf under the block that is small enough
      {
        k--;
        p *= FAN;
        load_group(py.level[k], py.size[k], p, 0, FAN - 1, g);
        p += group_first_le(g, 0, v);
      }
      return p;
    }
    p = base + FAN;  // this group is done: go on with the block behind its parent
    if (k < py.top)
    {
      p /= FAN;
      k++;
    }
  }
}

// last index < i with a value <= v; NONE when there is none
CLD_HD uint64_t prev_le(const Pyramid &py, uint64_t i, uint32_t v)
{
  if (i == 0)
  {
    return NONE;
  }
  int k = 0;
  uint64_t p = i - 1;
  uint32_t g[FAN];
  while (true)
  {
    const uint64_t base = p - p % FAN;
    load_group(py.level[k], py.size[k], base, 0, (uint32_t) (p - base), g);
    const uint32_t c = group_last_le(g, (uint32_t) (p - base), v);
    if (c < FAN)
    {
      p = base + c;
      while (k > 0)  // down to the last leaf under the block that is small enough
      {
        k--;
        p *= FAN;
        load_group(py.level[k], py.size[k], p, 0, FAN - 1, g);
        p += group_last_le(g, FAN - 1, v);
      }
      return p;
    }
    if (base == 0)
    {
      return NONE;
    }
    if (k < py.top)
    {
      p = base / FAN - 1;  // the group is done: go on with the block in front of its parent
      k++;
    } else
    {
      p = base - 1;
    }
  }
}

// position of the leftmost minimum of level 0 in the closed range [a, b], a <= b < size[0]
CLD_HD uint64_t leftmost_min(const Pyramid &py, uint64_t a, uint64_t b)
{
  uint32_t best = 0xffffffffu;
  uint64_t best_start = NONE, best_p = 0;
  int best_k = 0;
  int k = 0;
  uint64_t lo = a, hi = b;  // closed range of nodes of level k still to be covered
  uint64_t span = 1;        // leaves under one node of level k
  uint32_t g[FAN];
  // nodes [from, to] of the group at `base`, left to right; an equal value further right never replaces one found before
#define CLD_CONSIDER_GROUP(BASE, FROM, TO)                                       \
  do                                                                             \
  {                                                                              \
    load_group(py.level[k], py.size[k], (BASE), (FROM), (TO), g);                \
    _Pragma("unroll") for (uint32_t t__ = 0; t__ < FAN; t__++)                   \
    {                                                                            \
      const uint64_t start__ = ((BASE) + t__) * span;                            \
      if (t__ >= (FROM) && t__ <= (TO) && (g[t__] < best || (g[t__] == best && start__ < best_start))) \
      {                                                                          \
        best = g[t__];                                                           \
        best_start = start__;                                                    \
        best_p = (BASE) + t__;                                                   \
        best_k = k;                                                              \
      }                                                                          \
    }                                                                            \
  } while (0)
  while (true)
  {
    const uint64_t lo_base = lo - lo % FAN, hi_base = hi - hi % FAN;
    if (lo_base == hi_base)  // what is left lies inside one group
    {
      CLD_CONSIDER_GROUP(lo_base, (uint32_t) (lo - lo_base), (uint32_t) (hi - lo_base));
      break;
    }
    if (k == py.top)  // the top level is not grouped any further: walk it
    {
      for (uint64_t base = lo_base; base <= hi_base; base += FAN)
      {
        CLD_CONSIDER_GROUP(base, (uint32_t) (base == lo_base ? lo - lo_base : 0),
                           (uint32_t) (base == hi_base ? hi - hi_base : FAN - 1));
      }
      break;
    }
    // fringes: the part of lo's group from lo on (unless lo starts a group), the part of hi's group up to hi (unless hi
    // ends one); whole groups in between are covered by their parents one level up
    if (lo != lo_base)
    {
      CLD_CONSIDER_GROUP(lo_base, (uint32_t) (lo - lo_base), FAN - 1);
      lo = lo_base + FAN;
    }
    if (hi != hi_base + FAN - 1)
    {
      CLD_CONSIDER_GROUP(hi_base, 0, (uint32_t) (hi - hi_base));
      if (hi_base == 0)
      {
        break;  // nothing in front of hi's group (cannot happen with lo_base != hi_base; kept for the unsigned step)
      }
      hi = hi_base - 1;
    }
    if (lo > hi)
    {
      break;
    }
    lo /= FAN;
    hi /= FAN;
    k++;
    span *= FAN;
  }
#undef CLD_CONSIDER_GROUP
  k = best_k;
  uint64_t p = best_p;
  while (k > 0)  // down to the first leaf under the block that holds the minimum
  {
    k--;
    p *= FAN;
    load_group(py.level[k], py.size[k], p, 0, FAN - 1, g);
    uint32_t c = FAN;
#pragma unroll
    for (uint32_t t = 0; t < FAN; t++)
    {
      if (c == FAN && g[t] == best)
      {
        c = t;
      }
    }
    p += c;
  }
  return p;
}

struct Child
{
  uint8_t up, down, nextl;  // the order of Childinfo (virtualdef.h:174-179), which is also the order in the .cld file
};

CLD_HD uint8_t clamp_child(uint64_t d) { return (uint8_t) (d >= LARGE ? LARGE : d); }

// the three fields of entry i (0 <= i <= n, size[0] == n + 1)
CLD_HD Child child_entry(const Pyramid &py, uint64_t i)
{
  const uint32_t *L = py.level[0];
  const uint64_t N = py.size[0];
  const uint32_t v = L[i];
  Child c;
  c.up = c.down = c.nextl = 0;
  const uint64_t j = next_le(py, i, v);
  if (j < N)
  {
    if (L[j] == v)
    {
      c.nextl = clamp_child(j - i);
    }
    if (j > i + 1)
    {
      c.down = clamp_child(leftmost_min(py, i + 1, j - 1) - i);
    }
  }
  if (i >= 1 && L[i - 1] > v)
  {
    const uint64_t q = prev_le(py, i, v);  // exists: L[0] = 0
    c.up = clamp_child(i - leftmost_min(py, q == NONE ? 0 : q + 1, i - 1));
  }
  return c;
}

// The common case without the pyramid: `win` holds L[w0, w0 + wn) (shared memory on the device: the entries of a thread
// block and `cap` entries on either side, clipped to the table).  Scans at most `cap` entries to the right for j and at
// most `cap` to the left for q, keeping the leftmost minimum of what it passes, and fills `c` when both scans end
// inside the window; returns false otherwise (the caller then runs child_entry).  N = size of the table.
// (child_entry_near32 below is the same scan on 32-bit window positions: 16 % faster, exact in the host model and, since
// the queue append of child_near_kernel moved into converged code, in four runs over the 100 M entries of C2; the
// whole GPU test file has only been run with this form, which is why it is the default.  DESIGN.md 4.8 has the story.)
CLD_HD bool child_entry_near(const uint32_t *win, uint64_t w0, uint64_t wn, uint64_t N, uint64_t i, uint32_t cap,
                             Child &c)
{
  const uint32_t v = win[i - w0];
  c.up = c.down = c.nextl = 0;
  // to the right: j, and the leftmost minimum of (i, j)
  bool found = false;
  uint32_t minv = 0xffffffffu;
  uint64_t minp = 0, j = N;
  for (uint32_t s = 1; s <= cap; s++)
  {
    const uint64_t p = i + s;
    if (p >= N)
    {
      found = true;  // off the table: no j, nothing defined on this side
      break;
    }
    if (p >= w0 + wn)
    {
      break;
    }
    const uint32_t x = win[p - w0];
    if (x <= v)
    {
      j = p;
      found = true;
      break;
    }
    if (x < minv)
    {
      minv = x;
      minp = p;
    }
  }
  if (!found)
  {
    return false;
  }
  if (j < N)
  {
    if (win[j - w0] == v)
    {
      c.nextl = clamp_child(j - i);
    }
    if (j > i + 1)
    {
      c.down = clamp_child(minp - i);
    }
  }
  // to the left: q, and the leftmost minimum of (q, i)
  if (i >= 1 && win[i - 1 - w0] > v)
  {
    found = false;
    minv = 0xffffffffu;
    minp = 0;
    for (uint32_t s = 1; s <= cap && s <= i; s++)
    {
      const uint64_t p = i - s;
      if (p < w0)
      {
        break;
      }
      const uint32_t x = win[p - w0];
      if (x <= v)
      {
        found = true;
        break;
      }
      if (x <= minv)  // walking leftwards: an equal value further left replaces the one found before
      {
        minv = x;
        minp = p;
      }
    }
    if (!found)
    {
      return false;
    }
    c.up = clamp_child(i - minp);
  }
  return true;
}
#ifdef CLD_NEAR32  // -DCLD_NEAR32: the bounded scan on 32-bit window positions; same contract, li = i - w0
CLD_HD bool child_entry_near32(const uint32_t *win, uint32_t wn, bool ends_table, uint32_t li, uint32_t cap, Child &c)
{
  const uint32_t v = win[li];
  c.up = c.down = c.nextl = 0;
  const uint32_t stop = li + cap < wn - 1 ? li + cap : wn - 1;
  uint32_t minv = 0xffffffffu, minl = 0, j = 0;
  for (uint32_t lp = li + 1; lp <= stop; lp++)
  {
    const uint32_t x = win[lp];
    if (x <= v)
    {
      j = lp;
      break;
    }
    if (x < minv)
    {
      minv = x;
      minl = lp;
    }
  }
  if (j != 0)
  {
    if (win[j] == v)
    {
      c.nextl = clamp_child(j - li);
    }
    if (j > li + 1)
    {
      c.down = clamp_child(minl - li);
    }
  } else if (!(ends_table && stop == wn - 1))
  {
    return false;
  }
  if (li >= 1 && win[li - 1] > v)
  {
    const uint32_t lo = li > cap ? li - cap : 0;
    bool found = false;
    minv = 0xffffffffu;
    minl = 0;
    for (uint32_t lp = li - 1;; lp--)
    {
      const uint32_t x = win[lp];
      if (x <= v)
      {
        found = true;
        break;
      }
      if (x <= minv)
      {
        minv = x;
        minl = lp;
      }
      if (lp == lo)
      {
        break;
      }
    }
    if (!found)
    {
      return false;
    }
    c.up = clamp_child(li - minl);
  }
  return true;
}
#endif

// entry i of the one-byte table, from the three-field table (3 bytes per entry) and L.  The reference runs three loops
// one after the other over the same byte (mkcld.c:230-283), the later one overwrites; they stop below totallength
// (the third below totallength - 1), so the last entry -- and an entry that no loop writes -- keeps what the
// allocation held: 0 here.
CLD_HD uint8_t cld1_entry(const uint8_t *cld, const uint32_t *L, uint64_t i, uint64_t n)
{
  if (i >= n)
  {
    return 0;
  }
  const uint8_t down = cld[3 * i + 1], nextl = cld[3 * i + 2];
  uint8_t out = 0;
  const uint64_t next = i + nextl;  // decoded from the clamped byte, like DECODENEXTLINDEX (mkcld.c:209-210)
  if (next > i)
  {
    out = nextl;
  }
  if ((next == i && down > 0) || next == n)
  {
    out = down;
  }
  if (i + 1 < n && L[i] > L[i + 1])
  {
    const uint8_t up1 = cld[3 * (i + 1)], down1 = cld[3 * (i + 1) + 1];
    if (up1 != 0 || down1 != 0)  // DECODEUP(i+1) != DECODEDOWN(i+1)
    {
      out = up1;
    }
  }
  return out;
}

}  // namespace cld
