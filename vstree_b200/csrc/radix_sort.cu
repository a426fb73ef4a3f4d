// radix_sort.cu -- least-significant-digit radix sort of (key, 32-bit value) pairs, one memory sweep per digit.
//
// This is the dominant kernel family of the suffix sorter (north star: "LSD radix passes ... coalesced,
// vectorised HBM loads, shared-memory histogramming and warp-shuffle scans").  Structure per sort:
//   1. radix_hist_kernel   one read of the keys -> histograms of EVERY digit (shared-memory bins, plain atomics:
//                          ATOMS.POPC.INC counts the lanes of a warp that hit the same bin in one operation)
//   2. radix_scan_kernel   exclusive scan of each digit histogram -> global bin bases
//   3. radix_pass_kernel   per digit ONE read + ONE write of the pairs: each 512-thread CTA takes a tile
//                          (dynamic ticket), ranks its keys with warp match/ballot, learns the number of equal
//                          digits in all earlier tiles by decoupled look-back over a status array (single 64-bit
//                          word = 2 flag bits + 62 value bits), reorders the tile in shared memory and writes
//                          each digit run coalesced.  Stable.  The pass is bound by instruction issue, not by HBM
//                          (88 warp instructions per 32 pairs): see DESIGN.md section 5 for what was measured.
// Algorithmic bytes: hist = n*sizeof(key); pass = 2*n*(sizeof(key)+sizeof(value)).
#include "scan.cuh"
#include "textcmp.cuh"
#include <algorithm>

namespace
{

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
#ifndef ESA_SORT_THREADS
#define ESA_SORT_THREADS 512
#endif
#ifndef ESA_SORT_ITEMS32
#define ESA_SORT_ITEMS32 16
#endif
#ifndef ESA_SORT_ITEMS64
#define ESA_SORT_ITEMS64 12
#endif
#ifndef ESA_SORT_MINBLOCKS
#define ESA_SORT_MINBLOCKS 2
#endif
#ifndef ESA_SORT_EARLY
#define ESA_SORT_EARLY 2     // publish the tile's digit counts before the ranking phase (2: plain popc-inc atomics)
#endif
#ifndef ESA_HIST_PLAIN
#define ESA_HIST_PLAIN 1     // histogram with plain shared-memory atomics (no warp-uniformity test)
#endif
#ifndef ESA_SORT_CHAIN
#define ESA_SORT_CHAIN 1     // serial chain through the warp counters without shuffles
#endif
#ifndef ESA_SORT_MATCH
#define ESA_SORT_MATCH 0  // 0: eight ballots per key, 1: shared-memory atomicOr lane masks, 2: MATCH.ANY
#endif
#ifndef ESA_SORT_OVERLAP
#define ESA_SORT_OVERLAP 1  // look-back (warps 0-7) runs while the other warps already reorder their keys
#endif
#ifndef ESA_SORT_FULLTILE
#define ESA_SORT_FULLTILE 1  // full tiles take a path without bounds checks
#endif
#ifndef ESA_SORT_XMASK
#define ESA_SORT_XMASK 1  // ballot loop: peers &= ballot ^ (bit ? 0 : ~0), one LOP3 per bit
#endif
#ifndef ESA_SORT_PERSISTENT
#define ESA_SORT_PERSISTENT 0  // 1: one CTA per slot of the device loops over tile tickets; measured 5 % SLOWER (the
#endif                         //    hardware starts the next CTA while the last warps of the old one drain)
#ifndef ESA_SORT_KV64
#define ESA_SORT_KV64 1  // 32-bit key + value travel through shared memory as ONE 64-bit word: fewer shared-memory
#endif                   // wavefronts (ncu: the LSU data pipe is the busiest unit, 72 %, 60 % of its wavefronts are
                         // bank-conflict replays of the scattered tile reorder); measured neutral in time
#ifndef ESA_SORT_LBK
#define ESA_SORT_LBK 4  // status words of this many predecessor tiles are fetched per round trip of the look-back walk
#endif
#ifndef ESA_SORT_SPIN_NS
#define ESA_SORT_SPIN_NS 0  // nanosleep while a predecessor tile has not published yet (0: busy spin)
#endif
constexpr int SORT_THREADS = ESA_SORT_THREADS;
constexpr int SORT_WARPS = SORT_THREADS / 32;
template <int ITEMS> constexpr int pass_tile_items() { return SORT_THREADS * ITEMS; }
constexpr int HIST_THREADS = 512;
constexpr int HIST_ITEMS = 8;
// look-back status word: 2 flag bits + 62 value bits (counts reach 2^32 - 2, so 32-bit words with 30 value bits
// would force inputs beyond 2^30 pairs to be cut into separately histogrammed parts)
constexpr u64 STATUS_AGG = (u64) 1 << 62;
constexpr u64 STATUS_PREFIX = (u64) 2 << 62;
constexpr u64 STATUS_VALUE_MASK = ((u64) 1 << 62) - 1;

template <typename KeyT> struct SortTraits;
template <> struct SortTraits<u32>
{
  static constexpr int ITEMS = ESA_SORT_ITEMS32;
};
template <> struct SortTraits<u64>
{
  static constexpr int ITEMS = ESA_SORT_ITEMS64;
};

template <typename KeyT> __device__ __forceinline__ u32 digit_of(KeyT k, u32 shift, u32 mask)
{
  return (u32) (k >> shift) & mask;
}

// Lanes of the warp whose digit equals mine, one ballot per digit bit.  Per bit: test with an immediate mask (no shift),
// select of the xor mask, ballot of the negated predicate, one three-input LOP3 -- four instructions where the plain
// C++ forms compile to five (shift, and, negate, vote, lop3).  nd = ~digit.
template <int B> __device__ __forceinline__ u32 match_digit_bits(u32 nd, u32 peers)
{
  if constexpr (B < RADIX_BITS)
  {
    u32 xm, bal;
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                 "and.b32 t, %2, %3;\n\t"
                 "setp.ne.u32 p, t, 0;\n\t"
                 "selp.b32 %0, 0xffffffff, 0, p;\n\t"
                 "vote.sync.ballot.b32 %1, !p, 0xffffffff;\n\t}"
                 : "=r"(xm), "=r"(bal)
                 : "r"(nd), "n"(1u << B));
    return match_digit_bits<B + 1>(nd, peers & (bal ^ xm));
  } else
  {
    return peers;
  }
}

// ---------------------------------------------------------------------------------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(HIST_THREADS)
radix_hist_kernel(const KeyT *__restrict__ keys, u32 count, u32 begin_bit, u32 end_bit, u32 npass,
                  u32 *__restrict__ hist /* [npass][256] of this part */)
{
  __shared__ u32 sh[sizeof(KeyT) * RADIX];
  for (u32 i = threadIdx.x; i < npass * RADIX; i += HIST_THREADS)
  {
    sh[i] = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  constexpr u32 STEP = HIST_THREADS * HIST_ITEMS;
  for (u64 base = (u64) blockIdx.x * STEP; base < count; base += (u64) gridDim.x * STEP)
  {
    KeyT k[HIST_ITEMS];
    bool valid[HIST_ITEMS];
#pragma unroll
    for (int i = 0; i < HIST_ITEMS; i++)
    {
      u64 idx = base + (u64) i * HIST_THREADS + threadIdx.x;
      valid[i] = idx < count;
      k[i] = valid[i] ? keys[idx] : (KeyT) 0;
    }
#pragma unroll
    for (int i = 0; i < HIST_ITEMS; i++)
    {
      const u32 vmask = __ballot_sync(FULL_MASK, valid[i]);
      if (valid[i])
      {
        const int first = __ffs(vmask) - 1;
        for (u32 p = 0; p < npass; p++)
        {
          const u32 shift = begin_bit + p * RADIX_BITS;
          const u32 bits = min((u32) RADIX_BITS, end_bit - shift);
          const u32 d = digit_of(k[i], shift, (1u << bits) - 1);
#if ESA_HIST_PLAIN
          atomicAdd(&sh[p * RADIX + d], 1u);  // ATOMS.POPC.INC counts the lanes that hit the same bin in one go
          continue;
#endif
          // a digit that is the same in the whole warp (constant high bits, runs of equal keys) would
          // serialise 32 ways on one bin: count it once
          if (__all_sync(vmask, d == __shfl_sync(vmask, d, first)))
          {
            if (lane == first)
            {
              atomicAdd(&sh[p * RADIX + d], (u32) __popc(vmask));
            }
          } else
          {
            atomicAdd(&sh[p * RADIX + d], 1u);
          }
        }
      }
    }
  }
  __syncthreads();
  for (u32 i = threadIdx.x; i < npass * RADIX; i += HIST_THREADS)
  {
    const u32 v = sh[i];
    if (v != 0)
    {
      atomicAdd(&hist[i], v);
    }
  }
}

// one block per pass, one thread per digit
__global__ void __launch_bounds__(RADIX)
radix_scan_kernel(const u32 *__restrict__ hist, u32 *__restrict__ base, u32 npass, u32 parts,
                  const u32 *__restrict__ given /* if set (npass == 1): first output index of every digit */)
{
  __shared__ u32 scratch[RADIX / 32 + 1];
  const u32 p = blockIdx.x, d = threadIdx.x;
  u32 total = 0;
  for (u32 part = 0; part < parts; part++)
  {
    total += hist[((u64) part * npass + p) * RADIX + d];
  }
  u32 run = block_exclusive_sum<u32, RADIX>(total, scratch, nullptr);
  if (given != nullptr)
  {
    run = given[d];
  }
  for (u32 part = 0; part < parts; part++)
  {
    const u64 o = ((u64) part * npass + p) * RADIX + d;
    base[o] = run;
    run += hist[o];
  }
}

// ---------------------------------------------------------------------------------------------------------
// REMOTE: the pass is the partition step of the sharded build fused with its all-to-all: digit d belongs to rank
// remote->owner[d] and its run is stored straight into that rank's receive buffers (peer memory over NVLink);
// base[d] then is the first index of THIS rank's run of digit d inside the owner's buffers.
#ifdef ESA_SORT_TRACE
// phase timeline of the pass kernel (tuning builds only): per tile, SM clock at the phase borders as seen by thread 0
// (warp 0: walks the look-back) and by the last thread (warp 15: reorders at once)
constexpr int TRACE_SLOTS = 12;
__device__ u32 g_trace[2 * TRACE_SLOTS * 16384];
#define TRACE(slot)                                                                                         \
  do                                                                                                        \
  {                                                                                                         \
    if ((tid == 0 || tid == SORT_THREADS - 1) && tile < 16384)                                              \
    {                                                                                                       \
      g_trace[(tile * 2 + (tid != 0)) * TRACE_SLOTS + (slot)] = (u32) clock64();                            \
    }                                                                                                       \
  } while (0)
#else
#define TRACE(slot)
#endif

template <typename KeyT, bool HAS_VALS, int ITEMS, bool REMOTE = false>
__global__ void __launch_bounds__(SORT_THREADS, ESA_SORT_MINBLOCKS)
radix_pass_kernel(const KeyT *__restrict__ kin, KeyT *__restrict__ kout, const u32 *__restrict__ vin,
                  u32 *__restrict__ vout, u32 count, u32 shift, u32 mask, const u32 *__restrict__ base,
                  volatile u64 *status, u32 *ticket, const RadixRemoteDst *__restrict__ remote = nullptr)
{
  constexpr int TILE = SORT_THREADS * ITEMS;
  constexpr bool KV64 = ESA_SORT_KV64 && HAS_VALS && sizeof(KeyT) == 4;  // (key, value) as one 64-bit shared-memory word
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KeyT *skeys = (KeyT *) smem_raw;
  u32 *svals = (u32 *) (skeys + TILE);
  u32 *wc = svals + (HAS_VALS ? TILE : 0);     // [SORT_WARPS][RADIX] per-warp digit counters, then offsets
  u32 *digit_start = wc + SORT_WARPS * RADIX;  // [RADIX] first slot of the digit in the reordered tile
  u32 *gbase = digit_start + RADIX;            // [RADIX] global index of slot 0 of the digit, minus digit_start
  u32 *misc = gbase + RADIX;                   // [0] tile ticket, [1..] scan scratch
  u32 *thist = misc + 16;                      // [RADIX] digit counts of the tile, published BEFORE the ranking
  // REMOTE: the "digit" of a key is the rank that owns its bin (a look-up of the real digit), so that one tile's
  // output consists of nranks long runs instead of 256 short ones -- long stores are what NVLink wants
  __shared__ u8 sowner[RADIX];
  if (REMOTE && threadIdx.x < RADIX)
  {
    sowner[threadIdx.x] = remote->owner[threadIdx.x];
  }
#define DIGIT(k) (REMOTE ? (u32) sowner[digit_of((k), shift, mask)] : digit_of((k), shift, mask))

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u32 *mywc = wc + warp * RADIX;
#if ESA_SORT_PERSISTENT
  const u32 ntiles = (u32) (((u64) count + TILE - 1) / TILE);
  while (true)
  {
#endif
  if (tid == 0)
  {
    misc[0] = atomicAdd(ticket, 1u);
  }
#pragma unroll
  for (int d = lane; d < RADIX; d += 32)
  {
    mywc[d] = 0;
#if ESA_SORT_MATCH == 1
    thist[RADIX + warp * RADIX + d] = 0;
#endif
  }
  if (tid < RADIX)
  {
    thist[tid] = 0;
  }
  __syncthreads();
  const u32 tile = misc[0];
#if ESA_SORT_PERSISTENT
  if (tile >= ntiles)
  {
    break;  // the same for every thread of the CTA
  }
#endif
  const u32 tile_base = tile * (u32) TILE;  // < count < 2^32
  const u32 tile_count = min((u32) TILE, count - tile_base);
  TRACE(0);

  // ---- load, warp-striped: warp w owns [w*32*ITEMS, (w+1)*32*ITEMS) of the tile
  KeyT keys[ITEMS];
  u32 ranks[ITEMS];
  const u32 wbase = warp * (32 * ITEMS);
  const bool full_tile = ESA_SORT_FULLTILE && tile_count == (u32) TILE;
  if (full_tile)
  {
    const KeyT *kp = kin + tile_base + wbase + lane;
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
    {
      keys[i] = kp[i * 32];
    }
  } else
  {
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
    {
      const u32 idx = wbase + i * 32 + lane;
      keys[i] = idx < tile_count ? kin[tile_base + idx] : (KeyT) ~(KeyT) 0;  // padding sorts last
    }
  }
#if ESA_SORT_EARLY
  // ---- early counts: the tile's digit histogram is published before the (long) ranking phase, so that the
  // look-back of the following tiles, which runs after THEIR ranking, rarely waits.  Measured on B200: 2 % faster
  // than publishing after the ranking; a warp-cooperative look-back (32 predecessor tiles per round trip) was
  // 2.5x SLOWER than the one-thread-per-digit walk below -- chains are short in steady state, and the wide
  // windows only add L2 requests.
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    const u32 d = DIGIT(keys[i]);
#if ESA_SORT_EARLY == 2
    atomicAdd(&thist[d], 1u);  // ATOMS.POPC.INC: the hardware counts the lanes that hit the same word
#else
    if (__all_sync(FULL_MASK, d == __shfl_sync(FULL_MASK, d, 0)))
    {
      if (lane == 0)
      {
        atomicAdd(&thist[d], 32u);  // constant digit (high bits of small keys, runs of equal keys): count once
      }
    } else
    {
      atomicAdd(&thist[d], 1u);
    }
#endif
  }
  __syncthreads();
  u32 valid_count = 0;
  if (tid < RADIX)
  {
    valid_count = thist[tid];
    if ((u32) tid == (REMOTE ? (u32) sowner[mask] : mask))
    {
      valid_count -= (u32) TILE - tile_count;  // padding keys carry the largest digit
    }
    status[(u64) tile * RADIX + tid] = (tile == 0 ? STATUS_PREFIX : STATUS_AGG) | (u64) valid_count;
  }
#endif
  TRACE(1);  // keys loaded, early histogram published
  // ---- rank inside the warp: stable in (item, lane) order = memory order
  const u32 lt_mask = (1u << lane) - 1;
  const u32 gt_mask = 0xfffffffeu << lane;  // lanes above mine
  // lanes holding the same digit, found with one ballot per digit bit.  MATCH.ANY does the same in one
  // instruction but its cost grows with the number of distinct values in the warp (ncu, round 1: the instruction
  // behind it held 35 % of all stall samples and splitting it from its use changed nothing).
#if ESA_SORT_MATCH == 1
  // Variant: peers through a per-warp, per-digit lane mask in shared memory (one atomicOr per key, read back after
  // a warp barrier) instead of eight ballots: far fewer instructions, three shared-memory round trips per step.
  u32 *mymm = thist + RADIX + warp * RADIX;
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    const u32 d = DIGIT(keys[i]);
    atomicOr(&mymm[d], 1u << lane);
    __syncwarp();
    const u32 peers = mymm[d];
    __syncwarp();  // every lane has read the mask before its leader clears it
    const int leader = __ffs(peers) - 1;
    u32 old = 0;
    if (lane == leader)
    {
      old = mywc[d];
      mywc[d] = old + __popc(peers);
      mymm[d] = 0;
    }
    old = __shfl_sync(FULL_MASK, old, leader);
    ranks[i] = old + __popc(peers & lt_mask);
    __syncwarp();
  }
#else
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    const u32 d = DIGIT(keys[i]);
    u32 peers = FULL_MASK;
#if ESA_SORT_MATCH == 2
    peers = __match_any_sync(FULL_MASK, d);  // one MATCH.ANY instead of eight ballots (tuning variant)
#elif ESA_SORT_XMASK
    peers = match_digit_bits<0>(~d, peers);
#else
#pragma unroll
    for (int b = 0; b < RADIX_BITS; b++)
    {
      const bool bit = (d >> b) & 1;
      const u32 bal = __ballot_sync(FULL_MASK, bit);
      peers &= bit ? bal : ~bal;
    }
#endif
    ranks[i] = peers;
  }
  // then the short serial chain through the warp's shared counters
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    const u32 d = DIGIT(keys[i]);
    const u32 peers = ranks[i];
#if ESA_SORT_CHAIN == 1
    // every lane of a peer group reads the group's counter (one broadcast word), the highest lane of the group
    // writes it back: no shuffle, no find-first-set
    u32 *cp = mywc + d;
    const u32 r = *cp + __popc(peers & lt_mask);
    __syncwarp();  // all reads before the write
    if ((peers & gt_mask) == 0)
    {
      *cp = r + 1;
    }
    __syncwarp();
    ranks[i] = r;
#else
    const int leader = __ffs(peers) - 1;
    u32 old = 0;
    if (lane == leader)
    {
      old = mywc[d];
      mywc[d] = old + __popc(peers);
    }
    old = __shfl_sync(FULL_MASK, old, leader);
    ranks[i] = old + __popc(peers & lt_mask);
    __syncwarp();
#endif
  }
#endif
  TRACE(2);  // ranked
  __syncthreads();
  TRACE(3);  // all warps ranked
  // ---- per digit: tile count, then (after the scan over digits) the first slot of every warp's run
  u32 bcount = 0;
  if (tid < RADIX)
  {
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++)
    {
      bcount += wc[w * RADIX + tid];
    }
  }
  // exclusive scan over the 256 digit counts with ONE barrier: warp scans, eight warp totals through shared memory
  u32 dstart = 0;
  {
    const u32 incl = warp_inclusive_sum(bcount);
    if (tid < RADIX && lane == 31)
    {
      misc[1 + warp] = incl;
    }
    __syncthreads();
    if (tid < RADIX)
    {
      u32 before = 0;
#pragma unroll
      for (int w = 0; w < RADIX / 32; w++)
      {
        const u32 t = misc[1 + w];
        before += w < warp ? t : 0u;
      }
      dstart = before + incl - bcount;
    }
  }
  if (tid < RADIX)
  {
    u32 run = dstart;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++)
    {
      const u32 c = wc[w * RADIX + tid];
      wc[w * RADIX + tid] = run;  // slot of the warp's first key with this digit in the reordered tile
      run += c;
    }
  }
#if ESA_SORT_OVERLAP
  __syncthreads();  // warp offsets visible: warps 8.. reorder right away, warps 0-7 after their look-back
#endif
  TRACE(4);  // scanned
  // ---- decoupled look-back, one thread per digit
  if (tid < RADIX)
  {
    u32 cnt = bcount;
    if ((u32) tid == (REMOTE ? (u32) sowner[mask] : mask))
    {
      cnt -= (u32) TILE - tile_count;  // padding keys carry the largest digit
    }
    u32 excl = 0;
#if !ESA_SORT_EARLY
    if (tile == 0)
    {
      status[tid] = STATUS_PREFIX | (u64) cnt;
    } else
    {
      status[(u64) tile * RADIX + tid] = STATUS_AGG | (u64) cnt;
    }
#endif
    if (tile != 0)
    {
#if ESA_SORT_LBK > 1
      // The walk is a serial chain across tiles: with one status word per round trip it settles at
      // (latency until a published prefix is visible) x (round trip) / (gap between the starts of consecutive tiles),
      // measured 6300 clocks = a quarter of a tile's life (per-phase clock trace).  So every round trip
      // fetches the words of the next ESA_SORT_LBK predecessors at once and consumes them in order.
      u32 t = tile - 1;
      bool found = false;
      while (!found)
      {
        u64 sv[ESA_SORT_LBK];
#pragma unroll
        for (int j = 0; j < ESA_SORT_LBK; j++)
        {
          const u32 tj = t >= (u32) j ? t - (u32) j : 0u;  // clamped: tile 0 always holds a prefix, the walk ends there
          sv[j] = status[(u64) tj * RADIX + tid];
        }
        u32 used = 0;
#pragma unroll
        for (int j = 0; j < ESA_SORT_LBK; j++)
        {
          const u32 flag = (u32) (sv[j] >> 62);
          const bool take = !found && used == (u32) j && flag != 0;
          if (take)
          {
            excl += (u32) (sv[j] & STATUS_VALUE_MASK);
            used++;
            found = flag == 2;
          }
        }
        t -= found ? 0u : used;
      }
#else
      u32 t = tile - 1;
      while (true)
      {
        const u64 s = status[(u64) t * RADIX + tid];
        if ((s >> 62) == 0)
        {
#if ESA_SORT_SPIN_NS > 0
          __nanosleep(ESA_SORT_SPIN_NS);
#endif
          continue;  // predecessor has not published yet
        }
        excl += (u32) (s & STATUS_VALUE_MASK);
        if ((s >> 62) == 2)
        {
          break;
        }
        t--;
      }
#endif
      status[(u64) tile * RADIX + tid] = STATUS_PREFIX | (u64) (excl + cnt);
    }
    digit_start[tid] = dstart;
    gbase[tid] = base[tid] + excl - dstart;  // wraps modulo 2^32 by design
  }
#if !ESA_SORT_OVERLAP
  __syncthreads();
#endif
  TRACE(5);  // walked (warps 0-7)
  // ---- reorder the tile in shared memory
  if (full_tile)
  {
    const u32 *vp = vin + tile_base + wbase + lane;
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
    {
      const u32 d = DIGIT(keys[i]);
      const u32 pos = mywc[d] + ranks[i];
      if (KV64)
      {
        reinterpret_cast<u64 *>(smem_raw)[pos] = (u64) keys[i] | ((u64) vp[i * 32] << 32);
      } else
      {
        skeys[pos] = keys[i];
        if (HAS_VALS)
        {
          svals[pos] = vp[i * 32];
        }
      }
    }
  } else
  {
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
    {
      const u32 d = DIGIT(keys[i]);
      const u32 pos = mywc[d] + ranks[i];
      const u32 idx = wbase + i * 32 + lane;
      if (KV64)
      {
        const u32 v = idx < tile_count ? vin[tile_base + idx] : 0u;
        reinterpret_cast<u64 *>(smem_raw)[pos] = (u64) keys[i] | ((u64) v << 32);
      } else
      {
        skeys[pos] = keys[i];
        if (HAS_VALS)
        {
          svals[pos] = idx < tile_count ? vin[tile_base + idx] : 0u;
        }
      }
    }
  }
  TRACE(6);  // reordered
  __syncthreads();
  TRACE(7);  // join
  // ---- write every digit run coalesced
#pragma unroll
  for (int j = 0; j < ITEMS; j++)
  {
    const u32 p = j * SORT_THREADS + tid;
    if (full_tile || p < tile_count)
    {
      KeyT k;
      u32 v = 0;
      if (KV64)
      {
        const u64 kv = reinterpret_cast<const u64 *>(smem_raw)[p];
        k = (KeyT) (u32) kv;
        v = (u32) (kv >> 32);
      } else
      {
        k = skeys[p];
        if (HAS_VALS)
        {
          v = svals[p];
        }
      }
      const u32 d = DIGIT(k);
      const u32 g = gbase[d] + p;
      if (REMOTE)
      {
        reinterpret_cast<KeyT *>(remote->keys[d])[g] = (KeyT) ((k - remote->key_sub[d]) << remote->key_shl);  // d = owner
        if (HAS_VALS)
        {
          remote->vals[d][g] = v;
        }
      } else
      {
        kout[g] = k;
        if (HAS_VALS)
        {
          vout[g] = v;
        }
      }
    }
  }
  TRACE(8);  // written
#if ESA_SORT_PERSISTENT
  __syncthreads();  // the tile's shared memory is dead: the next ticket may overwrite it
  }
#endif
  if (REMOTE)
  {
    __threadfence_system();  // the stores crossed NVLink: visible to the owner before this kernel counts as done
  }
#undef DIGIT
}



// ---------------------------------------------------------------------------------------------------------
// First pass of the initial DNA sort with the keys taken FROM THE PACKED TEXT.
// classify_dna_emit used to write a (key, position) pair per suffix (8 B), the histogram kernel re-read the keys (4 B)
// and pass 1 re-read the pairs (8 B).  Key and position of a "full" suffix (no special among its first h symbols) are a
// function of the text position alone, so the histogram kernel and pass 1 compute them from the 2-bit records
// (0.5 B per position, staged in shared memory: the 32 positions of a warp row share one record) and the pairs exist
// for the first time in the OUTPUT of pass 1.  Positions that are not full (special-initial, special inside the
// prefix, behind the text) are simply not ranked; the few suffixes with a special inside their prefix ("side",
// already ordered by the caller) follow as extra tiles read from arrays, so that they stay behind the full suffixes
// with the same padded code, as the stable sort needs (suffix_sort.cu, step 1).
struct DnaKeySource
{
  TextPacked2 text;
  u64 p_begin;        // first text position (a multiple of 32): 0 on one GPU, the rank's slice in a sharded build
  u64 N;              // positions p_begin .. N-1 are classified (N - 1 may be the sentinel)
  u32 h;              // symbols per key
  u32 key_shl;        // 1: key = code << 1 (flag bit 0); 0: raw 16-symbol code (sharded builds with wide keys)
  u32 text_tiles;     // tiles of tile_items positions
  const u32 *side_key, *side_pos;
  u32 nside;
};

// key (padded code << key_shl, flag never set for a full suffix) of the position that is symbol o_rec of record words
// x0..x3; full: no special among the h symbols
__device__ __forceinline__ u32 dna_full_key(const TextPacked2 &text, u32 x0, u32 x1, u32 x2, u32 x3, u32 o_rec, u32 h, u64 p,
                                            bool &full, u32 key_shl = 1)
{
  const u32 bitoff = 2 * o_rec, wi = bitoff >> 5, bo = bitoff & 31;
  const u32 lo32 = wi == 0 ? x0 : (wi == 1 ? x1 : x2), hi32 = wi == 0 ? x1 : (wi == 1 ? x2 : x3);
  const u32 w32 = __funnelshift_r(lo32, hi32, bo);
  const u32 gmask = ((2u << (((o_rec + h - 1) >> 3) - (o_rec >> 3))) - 1) << (o_rec >> 3);
  full = true;
  if (((x3 >> 24) & gmask) != 0)  // a special among the record symbols around [o, o + h): ask the mask
  {
    const u32 msk = text.mask_slow(p);
    const u32 d = msk ? (u32) (__ffs((int) msk) - 1) : 32u;
    full = d >= h;
  }
  u32 r = __brev(w32);
  r = ((r & 0xAAAAAAAAu) >> 1) | ((r & 0x55555555u) << 1);
  return (r >> (32 - 2 * h)) << key_shl;
}

__global__ void __launch_bounds__(HIST_THREADS)
radix_hist_dna_kernel(DnaKeySource src, u32 begin_bit, u32 end_bit, u32 npass, u32 *__restrict__ hist /* [npass][256] */)
{
  __shared__ u32 sh[4 * RADIX];
  for (u32 i = threadIdx.x; i < npass * RADIX; i += HIST_THREADS)
  {
    sh[i] = 0;
  }
  __syncthreads();
  const u32 lane = threadIdx.x & 31;
  const u64 rows = (src.N + 31) / 32, first_row = src.p_begin >> 5;  // one warp row = the 32 positions of one record
  constexpr int ROWS = 8;              // records in flight per warp (one record per load: latency, not bytes, is the limit)
  const u64 warp_global = ((u64) blockIdx.x * HIST_THREADS + threadIdx.x) >> 5, nwarps = ((u64) gridDim.x * HIST_THREADS) >> 5;
  for (u64 row0 = first_row + warp_global * ROWS; row0 < rows; row0 += nwarps * ROWS)
  {
    ulonglong2 rr[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; j++)
    {
      rr[j] = row0 + j < rows ? __ldg(reinterpret_cast<const ulonglong2 *>(src.text.pk) + row0 + j) : make_ulonglong2(0, 0);
    }
#pragma unroll
    for (int j = 0; j < ROWS; j++)
    {
      const u64 p = (row0 + j) * 32 + lane;
      if (p < src.N)
      {
        bool full;
        const u32 key = dna_full_key(src.text, (u32) rr[j].x, (u32) (rr[j].x >> 32), (u32) rr[j].y, (u32) (rr[j].y >> 32), lane + 1,
                                     src.h, p, full, src.key_shl);
        if (full)
        {
          for (u32 q = 0; q < npass; q++)
          {
            const u32 shift = begin_bit + q * RADIX_BITS;
            const u32 bits = min((u32) RADIX_BITS, end_bit - shift);
            atomicAdd(&sh[q * RADIX + ((key >> shift) & ((1u << bits) - 1))], 1u);
          }
        }
      }
    }
  }
  __syncthreads();
  for (u32 i = threadIdx.x; i < npass * RADIX; i += HIST_THREADS)
  {
    const u32 v = sh[i];
    if (v != 0)
    {
      atomicAdd(&hist[i], v);
    }
  }
}

// REMOTE: the pass is the partition step of a sharded build fused with its all-to-all (see radix_pass_kernel): the digit
// of a key is the rank that owns its bin, the runs are stored into the owners' buffers.
template <int ITEMS, bool REMOTE = false>
__global__ void __launch_bounds__(SORT_THREADS, ESA_SORT_MINBLOCKS)
radix_pass_dna_kernel(DnaKeySource src, u32 *__restrict__ kout, u32 *__restrict__ vout, u32 shift, u32 mask,
                      const u32 *__restrict__ base, volatile u64 *status, u32 *ticket,
                      const RadixRemoteDst *__restrict__ remote = nullptr)
{
  constexpr int TILE = SORT_THREADS * ITEMS;
  __shared__ u8 sowner[RADIX];
  if (REMOTE && threadIdx.x < RADIX)
  {
    sowner[threadIdx.x] = remote->owner[threadIdx.x];
  }
#define DIG(k) (REMOTE ? (u32) sowner[((k) >> shift) & mask] : (((k) >> shift) & mask))
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64 *skv = (u64 *) smem_raw;                  // [TILE] reordered (key, position) pairs
  u32 *wc = (u32 *) (skv + TILE);               // [SORT_WARPS][RADIX]
  u32 *digit_start = wc + SORT_WARPS * RADIX;   // [RADIX]
  u32 *gbase = digit_start + RADIX;             // [RADIX]
  u32 *misc = gbase + RADIX;                    // [0] ticket, [1..8] scan scratch, [9] valid pairs of the tile
  u32 *thist = misc + 16;                       // [RADIX]
  ulonglong2 *srec = (ulonglong2 *) (thist + RADIX);  // [TILE / 32] records of a text tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0)
  {
    misc[0] = atomicAdd(ticket, 1u);
  }
  u32 *mywc = wc + warp * RADIX;
#pragma unroll
  for (int d = lane; d < RADIX; d += 32)
  {
    mywc[d] = 0;
  }
  if (tid < RADIX)
  {
    thist[tid] = 0;
  }
  __syncthreads();
  const u32 tile = misc[0];
  const bool text_tile = tile < src.text_tiles;
  const u64 tile_base = src.p_begin + (u64) tile * TILE;  // text position of item 0 (text tiles)
  if (text_tile && tid < TILE / 32 && ((tile_base >> 5) + tid) * 32 < src.N + 64)  // (records end 4 behind the text)
  {
    srec[tid] = __ldg(reinterpret_cast<const ulonglong2 *>(src.text.pk) + (tile_base >> 5) + tid);
  }
  __syncthreads();
  u32 keys[ITEMS], ranks[ITEMS];
  u32 valid = 0;  // bit i: item i takes part
  const u32 wbase = warp * (32 * ITEMS);
  const u32 side_base = text_tile ? 0u : (tile - src.text_tiles) * (u32) TILE;
  if (text_tile)
  {
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
    {
      const u64 p = tile_base + wbase + i * 32 + lane;
      keys[i] = 0;
      if (p < src.N)
      {
        const ulonglong2 rr = srec[warp * ITEMS + i];  // one record per warp row: a broadcast read
        bool full;
        keys[i] = dna_full_key(src.text, (u32) rr.x, (u32) (rr.x >> 32), (u32) rr.y, (u32) (rr.y >> 32), (u32) lane + 1, src.h, p,
                               full, src.key_shl);
        valid |= full ? (1u << i) : 0u;
      }
    }
  } else
  {
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
    {
      const u32 idx = side_base + wbase + i * 32 + lane;
      keys[i] = idx < src.nside ? src.side_key[idx] : 0u;
      valid |= idx < src.nside ? (1u << i) : 0u;
    }
  }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    if ((valid >> i) & 1u)
    {
      atomicAdd(&thist[DIG(keys[i])], 1u);
    }
  }
  __syncthreads();
  if (tid < RADIX)
  {
    status[(u64) tile * RADIX + tid] = (tile == 0 ? STATUS_PREFIX : STATUS_AGG) | (u64) thist[tid];
  }
  const u32 lt_mask = (1u << lane) - 1;
  const u32 gt_mask = 0xfffffffeu << lane;
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    const bool v = (valid >> i) & 1u;
    const u32 vb = __ballot_sync(FULL_MASK, v);
    ranks[i] = match_digit_bits<0>(~DIG(keys[i]), vb);  // valid lanes with my digit
  }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    const bool v = (valid >> i) & 1u;
    const u32 peers = ranks[i];
    u32 *cp = mywc + DIG(keys[i]);
    const u32 r = *cp + __popc(peers & lt_mask);
    __syncwarp();
    if (v && (peers & gt_mask) == 0)
    {
      *cp = r + 1;
    }
    __syncwarp();
    ranks[i] = r;
  }
  __syncthreads();
  u32 bcount = 0;
  if (tid < RADIX)
  {
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++)
    {
      bcount += wc[w * RADIX + tid];
    }
  }
  u32 dstart = 0;
  {
    const u32 incl = warp_inclusive_sum(bcount);
    if (tid < RADIX && lane == 31)
    {
      misc[1 + warp] = incl;
    }
    __syncthreads();
    if (tid < RADIX)
    {
      u32 before = 0;
#pragma unroll
      for (int w = 0; w < RADIX / 32; w++)
      {
        const u32 t = misc[1 + w];
        before += w < warp ? t : 0u;
      }
      dstart = before + incl - bcount;
    }
  }
  if (tid < RADIX)
  {
    u32 run = dstart;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++)
    {
      const u32 c = wc[w * RADIX + tid];
      wc[w * RADIX + tid] = run;
      run += c;
    }
    if (tid == RADIX - 1)
    {
      misc[9] = dstart + bcount;  // pairs of the tile
    }
  }
  __syncthreads();
  if (tid < RADIX)
  {
    u32 excl = 0;
    if (tile != 0)
    {
      u32 t = tile - 1;
      bool found = false;
      while (!found)
      {
        u64 sv[ESA_SORT_LBK];
#pragma unroll
        for (int j = 0; j < ESA_SORT_LBK; j++)
        {
          const u32 tj = t >= (u32) j ? t - (u32) j : 0u;
          sv[j] = status[(u64) tj * RADIX + tid];
        }
        u32 used = 0;
#pragma unroll
        for (int j = 0; j < ESA_SORT_LBK; j++)
        {
          const u32 flag = (u32) (sv[j] >> 62);
          const bool take = !found && used == (u32) j && flag != 0;
          if (take)
          {
            excl += (u32) (sv[j] & STATUS_VALUE_MASK);
            used++;
            found = flag == 2;
          }
        }
        t -= found ? 0u : used;
      }
      status[(u64) tile * RADIX + tid] = STATUS_PREFIX | (u64) (excl + bcount);
    }
    digit_start[tid] = dstart;
    gbase[tid] = base[tid] + excl - dstart;
  }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
  {
    if ((valid >> i) & 1u)
    {
      const u32 idx = wbase + i * 32 + lane;
      const u32 pos = mywc[DIG(keys[i])] + ranks[i];
      const u32 val = text_tile ? (u32) (tile_base + idx) : src.side_pos[side_base + idx];
      skv[pos] = (u64) keys[i] | ((u64) val << 32);
    }
  }
  __syncthreads();
  const u32 tile_pairs = misc[9];
#pragma unroll
  for (int j = 0; j < ITEMS; j++)
  {
    const u32 p = j * SORT_THREADS + tid;
    if (p < tile_pairs)
    {
      const u64 kv = skv[p];
      const u32 k = (u32) kv;
      const u32 d = DIG(k);
      const u32 g = gbase[d] + p;
      if (REMOTE)
      {
        remote->keys[d][g] = (k - remote->key_sub[d]) << remote->key_shl;  // d = owner; rebased to its key range
        remote->vals[d][g] = (u32) (kv >> 32);
      } else
      {
        kout[g] = k;
        vout[g] = (u32) (kv >> 32);
      }
    }
  }
  if (REMOTE)
  {
    __threadfence_system();  // the stores crossed NVLink: visible to the owner before this kernel counts as done
  }
#undef DIG
}

constexpr size_t pass_dna_smem_bytes()
{
  constexpr int TILE = pass_tile_items<SortTraits<u32>::ITEMS>();
  return (size_t) TILE * 8 + (size_t) SORT_WARPS * RADIX * 4 + 2 * RADIX * 4 + 16 * 4 + RADIX * 4 + (size_t) (TILE / 32) * 16;
}

template <typename KeyT> constexpr size_t pass_smem_bytes(bool has_vals)
{
  constexpr int TILE = pass_tile_items<SortTraits<KeyT>::ITEMS>();
  return (size_t) TILE * sizeof(KeyT) + (has_vals ? (size_t) TILE * 4 : 0) + (size_t) SORT_WARPS * RADIX * 4
         + 2 * RADIX * 4 + 16 * 4 + RADIX * 4 + (ESA_SORT_MATCH == 1 ? (size_t) SORT_WARPS * RADIX * 4 : 0);
}

// the whole input is one part: indices and counts are 32-bit (n < 2^32 is guaranteed by the API) and the look-back
// words carry 62 value bits.  (Round 1 cut inputs beyond 2^30 pairs into parts with 30-bit status words; the parts of
// passes after the first were histogrammed from the wrong slices -- found on the 3 Gbp configuration.)
template <typename KeyT> u64 part_items_max() { return (u64) 1 << 32; }

struct TempLayout
{
  u32 parts, npass_max;
  size_t hist_off, base_off, status_off, ticket_off, total;
};

template <typename KeyT> TempLayout temp_layout(u64 n)
{
  constexpr u64 TILE = (u64) pass_tile_items<SortTraits<KeyT>::ITEMS>();
  TempLayout L;
  const u64 pim = part_items_max<KeyT>();
  L.parts = (u32) (n == 0 ? 1 : div_up_u64(n, pim));
  L.npass_max = (u32) sizeof(KeyT);
  const u64 tiles_per_part = div_up_u64(n < pim ? (n == 0 ? 1 : n) : pim, TILE);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t) 255;
    return o;
  };
  L.hist_off = take((size_t) L.parts * L.npass_max * RADIX * 4);
  L.ticket_off = take((size_t) L.parts * L.npass_max * 4);
  L.base_off = take((size_t) L.parts * L.npass_max * RADIX * 4);
  L.status_off = take((size_t) tiles_per_part * RADIX * 8);
  L.total = off;
  return L;
}

template <typename KeyT>
int radix_sort_impl(esa_ctx *ctx, KeyT *ka, KeyT *kb, u32 *va, u32 *vb, u64 n, u32 begin_bit, u32 end_bit, void *temp,
                    size_t temp_bytes)
{
  constexpr int ITEMS = SortTraits<KeyT>::ITEMS;
  constexpr u64 TILE = (u64) pass_tile_items<ITEMS>();
  if (n <= 1 || end_bit <= begin_bit)
  {
    return 0;
  }
  const bool has_vals = va != nullptr;
  const u32 npass = (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS;
  const TempLayout L = temp_layout<KeyT>(n);
  if (temp_bytes < L.total)
  {
    esa_throw(ESA_ERR_ARGUMENT, "radix sort scratch too small: %zu < %zu", temp_bytes, L.total);
  }
  char *t = (char *) temp;
  u32 *hist = (u32 *) (t + L.hist_off), *tickets = (u32 *) (t + L.ticket_off), *base = (u32 *) (t + L.base_off),
      *status_words = (u32 *) (t + L.status_off);
  u64 *status = (u64 *) status_words;
  const u64 pim = part_items_max<KeyT>();
  cudaStream_t s = ctx->stream;

  // hist and tickets are adjacent: one memset clears both
  ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(hist, 0, L.base_off - L.hist_off, s)));
  for (u32 part = 0; part < L.parts; part++)
  {
    const u64 off = (u64) part * pim;
    const u32 cnt = (u32) (n - off < pim ? n - off : pim);
    const u32 grid = (u32) min((u64) ctx->num_sms * 4, div_up_u64(cnt, (u64) HIST_THREADS * HIST_ITEMS));
    ESA_LAUNCH(ctx, ESA_K_RADIX_HIST, (double) cnt * sizeof(KeyT),
               (radix_hist_kernel<KeyT><<<grid, HIST_THREADS, 0, s>>>(ka + off, cnt, begin_bit, end_bit, npass,
                                                                        hist + (u64) part * npass * RADIX)));
  }
  ESA_LAUNCH(ctx, ESA_K_MISC, 0, (radix_scan_kernel<<<npass, RADIX, 0, s>>>(hist, base, npass, L.parts, nullptr)));

  const size_t smem = pass_smem_bytes<KeyT>(has_vals);
  KeyT *ksrc = ka, *kdst = kb;
  u32 *vsrc = va, *vdst = vb;
  for (u32 p = 0; p < npass; p++)
  {
    const u32 shift = begin_bit + p * RADIX_BITS;
    const u32 bits = min((u32) RADIX_BITS, end_bit - shift);
    const u32 mask = (1u << bits) - 1;
    for (u32 part = 0; part < L.parts; part++)
    {
      const u64 off = (u64) part * pim;
      const u32 cnt = (u32) (n - off < pim ? n - off : pim);
      const u32 tiles = (u32) div_up_u64(cnt, TILE);
      const u32 grid = ESA_SORT_PERSISTENT ? min(tiles, (u32) ctx->num_sms * ESA_SORT_MINBLOCKS) : tiles;
      ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(status, 0, (size_t) tiles * RADIX * 8, s)));
      const u32 *pbase = base + ((u64) part * npass + p) * RADIX;
      u32 *ticket = tickets + (u64) part * npass + p;
      const double bytes = 2.0 * cnt * (sizeof(KeyT) + (has_vals ? 4 : 0));
      if (has_vals)
      {
        ESA_LAUNCH(ctx, ESA_K_RADIX_PASS, bytes,
                   (radix_pass_kernel<KeyT, true, ITEMS><<<grid, SORT_THREADS, smem, s>>>(
                       ksrc + off, kdst, vsrc + off, vdst, cnt, shift, mask, pbase, status, ticket)));
      } else
      {
        ESA_LAUNCH(ctx, ESA_K_RADIX_PASS, bytes,
                   (radix_pass_kernel<KeyT, false, ITEMS><<<grid, SORT_THREADS, smem, s>>>(
                       ksrc + off, kdst, nullptr, nullptr, cnt, shift, mask, pbase, status, ticket)));
      }
    }
#ifdef ESA_SORT_TRACE
    if (p == 0 && n > 50000000 && getenv("ESA_SORT_TRACE_DUMP") != nullptr)
    {
      CUDA_CHECK(cudaStreamSynchronize(s));
      static std::vector<u32> h(2 * TRACE_SLOTS * 16384);
      CUDA_CHECK(cudaMemcpyFromSymbol(h.data(), g_trace, h.size() * sizeof(u32)));
      const u32 ntr = (u32) std::min<u64>(16384, div_up_u64(n, TILE)) - 1;
      for (int who = 0; who < 2; who++)
      {
        double acc[TRACE_SLOTS] = {0};
        for (u32 t = 300; t < ntr; t++)
        {
          const u32 *r = &h[(t * 2 + who) * TRACE_SLOTS];
          for (int k = 1; k <= 8; k++)
          {
            acc[k] += (double) (u32) (r[k] - r[k - 1]);
          }
        }
        fprintf(stderr, "TRACE %s: load+hist %.0f rank %.0f wait1 %.0f scan %.0f walk %.0f reorder %.0f wait2 %.0f write %.0f clocks (avg over %u tiles)\n",
                who ? "warp15" : "warp0 ", acc[1] / (ntr - 300), acc[2] / (ntr - 300), acc[3] / (ntr - 300), acc[4] / (ntr - 300),
                acc[5] / (ntr - 300), acc[6] / (ntr - 300), acc[7] / (ntr - 300), acc[8] / (ntr - 300), ntr - 300);
      }
    }
#endif
    KeyT *kt = ksrc;
    ksrc = kdst;
    kdst = kt;
    u32 *vt = vsrc;
    vsrc = vdst;
    vdst = vt;
  }
  return (int) (npass & 1);
}

template <typename KeyT> void set_smem_attr()
{
  if (sizeof(KeyT) == 4)
  {
    CUDA_CHECK(cudaFuncSetAttribute(radix_pass_kernel<u32, true, SortTraits<u32>::ITEMS, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pass_smem_bytes<u32>(true)));
  }
  CUDA_CHECK(cudaFuncSetAttribute(radix_pass_kernel<KeyT, true, SortTraits<KeyT>::ITEMS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pass_smem_bytes<KeyT>(true)));
  CUDA_CHECK(cudaFuncSetAttribute(radix_pass_kernel<KeyT, false, SortTraits<KeyT>::ITEMS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pass_smem_bytes<KeyT>(false)));
}

}  // namespace

void radix_sort_init()
{
  CUDA_CHECK(cudaFuncSetAttribute(radix_pass_dna_kernel<SortTraits<u32>::ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int) pass_dna_smem_bytes()));
  CUDA_CHECK(cudaFuncSetAttribute(radix_pass_dna_kernel<SortTraits<u32>::ITEMS, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pass_dna_smem_bytes()));

  set_smem_attr<u32>();
  set_smem_attr<u64>();
}

// Stable partition of (key, value) pairs by the digit in key bits [begin_bit, end_bit) (at most 8 bits) whose output
// goes to OTHER ranks' buffers: see the REMOTE note at radix_pass_kernel.  d_start[256] (device) holds, per digit, the
// first index of this rank's run inside the owner's buffers (entry r of d_start for owner rank r; the kernel
// partitions by OWNER, the look-up of the digit, not by the digit itself).
void radix_partition_remote_u32(esa_ctx *ctx, const u32 *keys, const u32 *vals, u64 n, u32 begin_bit, u32 end_bit,
                                const RadixRemoteDst *d_remote, const u32 *d_start, void *temp, size_t temp_bytes)
{
  constexpr int ITEMS = SortTraits<u32>::ITEMS;
  constexpr u64 TILE = (u64) pass_tile_items<ITEMS>();
  if (n == 0)
  {
    return;
  }
  if (end_bit <= begin_bit || end_bit - begin_bit > RADIX_BITS)
  {
    esa_throw(ESA_ERR_ARGUMENT, "remote partition needs a digit of 1..8 bits");
  }
  const TempLayout L = temp_layout<u32>(n);
  if (temp_bytes < L.total)
  {
    esa_throw(ESA_ERR_ARGUMENT, "radix sort scratch too small: %zu < %zu", temp_bytes, L.total);
  }
  char *t = (char *) temp;
  u32 *hist = (u32 *) (t + L.hist_off), *tickets = (u32 *) (t + L.ticket_off), *base = (u32 *) (t + L.base_off),
      *status_words = (u32 *) (t + L.status_off);
  u64 *status = (u64 *) status_words;
  const u64 pim = part_items_max<u32>();
  cudaStream_t s = ctx->stream;
  // one part (n < 2^32): the first destination index of every owner's run is given, no histogram pass is needed
  ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(hist, 0, L.base_off - L.hist_off, s)));
  ESA_LAUNCH(ctx, ESA_K_MEMOP, 0,
             CUDA_CHECK(cudaMemcpyAsync(base, d_start, RADIX * sizeof(u32), cudaMemcpyDeviceToDevice, s)));
  (void) hist;
  const u32 mask = (1u << (end_bit - begin_bit)) - 1;
  const size_t smem = pass_smem_bytes<u32>(true);
  for (u32 part = 0; part < L.parts; part++)
  {
    const u64 off = (u64) part * pim;
    const u32 cnt = (u32) (n - off < pim ? n - off : pim);
    const u32 tiles = (u32) div_up_u64(cnt, TILE);
    const u32 grid = ESA_SORT_PERSISTENT ? min(tiles, (u32) ctx->num_sms * ESA_SORT_MINBLOCKS) : tiles;
    ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(status, 0, (size_t) tiles * RADIX * 8, s)));
    // algorithmic bytes: one local read, one store into peer memory
    ESA_LAUNCH(ctx, ESA_K_RADIX_PASS, 2.0 * cnt * 8,
               (radix_pass_kernel<u32, true, ITEMS, true><<<grid, SORT_THREADS, smem, s>>>(
                   keys + off, nullptr, vals + off, nullptr, cnt, begin_bit, mask, base + (u64) part * RADIX, status,
                   tickets + part, d_remote)));
  }
}


// The initial sort of a DNA text: M = nfull + nside pairs that exist only implicitly (see radix_pass_dna_kernel).
// Pass 1 reads the text and writes (keys_b, vals_b); the remaining passes ping-pong as in radix_sort_impl.  Returns 0
// when the result is in (keys_a, vals_a), 1 when it is in (keys_b, vals_b).
int radix_sort_dna_initial(esa_ctx *ctx, const u64 *d_pk, const u32 *d_spm, const u8 *d_tis, u64 N, u32 h, u64 nfull,
                           const u32 *side_key, const u32 *side_pos, u64 nside, u32 *keys_a, u32 *keys_b, u32 *vals_a,
                           u32 *vals_b, u32 key_bits, void *temp, size_t temp_bytes)
{
  constexpr int ITEMS = SortTraits<u32>::ITEMS;
  constexpr u64 TILE = (u64) pass_tile_items<ITEMS>();
  const u64 M = nfull + nside;
  const u32 npass = (key_bits + RADIX_BITS - 1) / RADIX_BITS;
  DnaKeySource src;
  src.text = TextPacked2{d_pk, d_spm, d_tis};
  src.p_begin = 0;
  src.N = N;
  src.h = h;
  src.key_shl = 1;
  src.text_tiles = (u32) div_up_u64(N, TILE);
  src.side_key = side_key;
  src.side_pos = side_pos;
  src.nside = (u32) nside;
  const u32 tiles1 = src.text_tiles + (u32) div_up_u64(nside, TILE);
  const TempLayout L = temp_layout<u32>(N + nside + TILE);  // the status rows of pass 1 cover text tiles + side tiles
  if (temp_bytes < L.total || npass > 4)
  {
    esa_throw(ESA_ERR_ARGUMENT, "radix sort scratch too small: %zu < %zu", temp_bytes, L.total);
  }
  char *t = (char *) temp;
  u32 *hist = (u32 *) (t + L.hist_off), *tickets = (u32 *) (t + L.ticket_off), *base = (u32 *) (t + L.base_off);
  u64 *status = (u64 *) (t + L.status_off);
  cudaStream_t s = ctx->stream;
  ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(hist, 0, L.base_off - L.hist_off, s)));
  {
    const u32 grid = (u32) min((u64) ctx->num_sms * 4, div_up_u64(N, (u64) HIST_THREADS * HIST_ITEMS));
    ESA_LAUNCH(ctx, ESA_K_RADIX_HIST, 0.5 * (double) N,
               (radix_hist_dna_kernel<<<grid, HIST_THREADS, 0, s>>>(src, 0, key_bits, npass, hist)));
    if (nside > 0)
    {
      const u32 g2 = (u32) min((u64) ctx->num_sms * 4, div_up_u64(nside, (u64) HIST_THREADS * HIST_ITEMS));
      ESA_LAUNCH(ctx, ESA_K_RADIX_HIST, 4.0 * nside,
                 (radix_hist_kernel<u32><<<g2, HIST_THREADS, 0, s>>>(side_key, (u32) nside, 0, key_bits, npass, hist)));
    }
  }
  ESA_LAUNCH(ctx, ESA_K_MISC, 0, (radix_scan_kernel<<<npass, RADIX, 0, s>>>(hist, base, npass, 1, nullptr)));
  // pass 1: text -> b
  {
    const u32 bits = min((u32) RADIX_BITS, key_bits);
    ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(status, 0, (size_t) tiles1 * RADIX * 8, s)));
    ESA_LAUNCH(ctx, ESA_K_RADIX_PASS_TEXT, 0.5 * (double) N + 8.0 * (double) M + 8.0 * (double) nside,
               (radix_pass_dna_kernel<ITEMS><<<tiles1, SORT_THREADS, pass_dna_smem_bytes(), s>>>(
                   src, keys_b, vals_b, 0u, (1u << bits) - 1, base, status, tickets)));
  }
  u32 *ksrc = keys_b, *kdst = keys_a, *vsrc = vals_b, *vdst = vals_a;
  const size_t smem = pass_smem_bytes<u32>(true);
  const u32 tiles = (u32) div_up_u64(M, TILE);
  for (u32 p = 1; p < npass; p++)
  {
    const u32 shift = p * RADIX_BITS;
    const u32 bits = min((u32) RADIX_BITS, key_bits - shift);
    ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(status, 0, (size_t) tiles * RADIX * 8, s)));
    ESA_LAUNCH(ctx, ESA_K_RADIX_PASS, 16.0 * (double) M,
               (radix_pass_kernel<u32, true, ITEMS><<<tiles, SORT_THREADS, smem, s>>>(
                   ksrc, kdst, vsrc, vdst, (u32) M, shift, (1u << bits) - 1, base + (u64) p * RADIX, status, tickets + p)));
    u32 *kt = ksrc;
    ksrc = kdst;
    kdst = kt;
    u32 *vt = vsrc;
    vsrc = vdst;
    vdst = vt;
  }
  return ksrc == keys_a ? 0 : 1;
}


// Sharded DNA builds: the splitter histogram (256 bins = key bits [bin_shift, bin_shift + 8)) of the full suffixes of the
// text slice [p_begin, p_end), and the partition pass fused with the all-to-all (radix_partition_remote_u32), both with
// the keys taken from the packed text instead of an emitted list.
static DnaKeySource dna_slice_source(const u64 *d_pk, const u32 *d_spm, const u8 *d_tis, u64 p_begin, u64 p_end, u32 h,
                                     u32 key_shl)
{
  constexpr u64 TILE = (u64) pass_tile_items<SortTraits<u32>::ITEMS>();
  DnaKeySource src;
  src.text = TextPacked2{d_pk, d_spm, d_tis};
  src.p_begin = p_begin;
  src.N = p_end;
  src.h = h;
  src.key_shl = key_shl;
  src.text_tiles = (u32) div_up_u64(p_end - p_begin, TILE);
  src.side_key = src.side_pos = nullptr;
  src.nside = 0;
  return src;
}

void radix_bin_hist_dna(esa_ctx *ctx, const u64 *d_pk, const u32 *d_spm, const u8 *d_tis, u64 p_begin, u64 p_end, u32 h,
                        u32 key_shl, u32 bin_shift, u32 *d_hist256)
{
  if (p_end <= p_begin)
  {
    return;
  }
  const DnaKeySource src = dna_slice_source(d_pk, d_spm, d_tis, p_begin, p_end, h, key_shl);
  const u32 grid = (u32) min((u64) ctx->num_sms * 4, div_up_u64(p_end - p_begin, (u64) HIST_THREADS * HIST_ITEMS));
  ESA_LAUNCH(ctx, ESA_K_EXCHANGE, 0.5 * (double) (p_end - p_begin),
             (radix_hist_dna_kernel<<<grid, HIST_THREADS, 0, ctx->stream>>>(src, bin_shift, bin_shift + 8, 1, d_hist256)));
}

void radix_partition_remote_dna(esa_ctx *ctx, const u64 *d_pk, const u32 *d_spm, const u8 *d_tis, u64 p_begin, u64 p_end,
                                u32 h, u32 key_shl, u64 nfull, u32 bin_shift, const RadixRemoteDst *d_remote,
                                const u32 *d_start, void *temp, size_t temp_bytes)
{
  constexpr int ITEMS = SortTraits<u32>::ITEMS;
  constexpr u64 TILE = (u64) pass_tile_items<ITEMS>();
  if (p_end <= p_begin)
  {
    return;
  }
  const DnaKeySource src = dna_slice_source(d_pk, d_spm, d_tis, p_begin, p_end, h, key_shl);
  const TempLayout L = temp_layout<u32>(p_end - p_begin + TILE);
  if (temp_bytes < L.total)
  {
    esa_throw(ESA_ERR_ARGUMENT, "radix sort scratch too small: %zu < %zu", temp_bytes, L.total);
  }
  char *t = (char *) temp;
  u32 *hist = (u32 *) (t + L.hist_off), *tickets = (u32 *) (t + L.ticket_off), *base = (u32 *) (t + L.base_off);
  u64 *status = (u64 *) (t + L.status_off);
  cudaStream_t s = ctx->stream;
  ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(hist, 0, L.base_off - L.hist_off, s)));
  ESA_LAUNCH(ctx, ESA_K_MEMOP, 0,
             CUDA_CHECK(cudaMemcpyAsync(base, d_start, RADIX * sizeof(u32), cudaMemcpyDeviceToDevice, s)));
  ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(status, 0, (size_t) src.text_tiles * RADIX * 8, s)));
  // algorithmic bytes: the text slice, one store of the pairs into peer memory
  ESA_LAUNCH(ctx, ESA_K_RADIX_PASS_TEXT, 0.5 * (double) (p_end - p_begin) + 8.0 * (double) nfull,
             (radix_pass_dna_kernel<ITEMS, true><<<src.text_tiles, SORT_THREADS, pass_dna_smem_bytes(), s>>>(
                 src, nullptr, nullptr, bin_shift, 0xffu, base, status, tickets, d_remote)));
}

size_t radix_sort_temp_bytes(u64 n, u32 key_bytes)
{
  return key_bytes == 8 ? temp_layout<u64>(n).total : temp_layout<u32>(n).total;
}

int radix_sort_pairs_u32(esa_ctx *ctx, u32 *keys_a, u32 *keys_b, u32 *vals_a, u32 *vals_b, u64 n, u32 begin_bit,
                         u32 end_bit, void *temp, size_t temp_bytes)
{
  return radix_sort_impl<u32>(ctx, keys_a, keys_b, vals_a, vals_b, n, begin_bit, end_bit, temp, temp_bytes);
}

int radix_sort_pairs_u64(esa_ctx *ctx, u64 *keys_a, u64 *keys_b, u32 *vals_a, u32 *vals_b, u64 n, u32 begin_bit,
                         u32 end_bit, void *temp, size_t temp_bytes)
{
  return radix_sort_impl<u64>(ctx, keys_a, keys_b, vals_a, vals_b, n, begin_bit, end_bit, temp, temp_bytes);
}
