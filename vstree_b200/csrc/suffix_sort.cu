// suffix_sort.cu -- suffix array + inverse suffix array by radix-sort based prefix doubling, and the .bck table.
//
// Replaces bucketsortsuffixes + bentleysedgewick (reference: src/Mkvtree/ppsort.c:83-313, bese.c:643-1105) with
// a data-parallel algorithm that yields the same total order (bese.c:26-55: letters by code; a special at
// position p, and the sentinel, compare as p+256):
//   1. classify every position by the distance to its next special and build the padded h0-mer code
//      (h0 = largest h with sigma^h <= 2^31): full h0-mers -> 32-bit key code<<1; suffixes with a special at
//      offset 0 < j < h0 are already uniquely ordered: key code<<1|1, appended BEHIND the full ones in
//      (h0-j ascending, position ascending) order so that the stable LSD sort places them right
//      (same padded code: full first, then larger j first, then smaller position first);
//      special-initial suffixes and the sentinel go straight to the tail of the table in position order.
//   2. one 32-bit LSD radix sort of (key, position).
//   3. .bck from the sorted keys (k <= h0 always holds): left = lower bound of the k-mer, mid = next left minus
//      the number of suffixes of the bucket with a special inside the k-prefix (mkvprocess.c:251-327).
//   4. regroup: equal keys form a group; rank[pos] = index of the group's first element; singletons are final.
//   5. refinement rounds on the ACTIVE set only: sort by (rank[s], key2) as one 64-bit key, regroup.
//      Round 1 (ESA_TEXT_ROUNDS rounds, default 1) takes key2 straight from the text: the next L symbols behind the h
//      the group agrees on, L as large as the free key bits allow (DNA: 16 symbols for n < 2^27), h += L.  Inputs
//      without long repeats end here, and then no rank array was ever needed: the N-wide scatter rank[pos] = slot of
//      the first regroup is DEFERRED (the ranks are kept in suffix-array order, a sequential write) and only carried
//      out when a doubling round follows.  Later rounds: key2 = rank[s+h], h *= 2.
#include <time.h>

#include "kmer_scan.cuh"
#include "scan.cuh"
#include "textcmp.cuh"

void radix_sort_init();

namespace
{

constexpr u32 NONE32 = 0xffffffffu;

// =========================================================================================================
// 1. classification
// scalars (device u64[16]): [0] nfull [1] nside [2] nspec [3] num_sorted M (= nfull + nside) [4] active count
// [5] carry (unused) [6] longest ...

__global__ void __launch_bounds__(KS_THREADS)
classify_count_kernel(const u8 *__restrict__ tis, u64 n, u64 p_begin, u64 p_end, KmerParams kp,
                      u32 *__restrict__ tile_full, u32 *__restrict__ tile_side, u32 *__restrict__ tile_spec)
{
  // positions [p_begin, p_end) are classified: the whole text [0, n+1) on one GPU, a slice of it when sharded
  __shared__ __align__(16) u8 stext[KS_TILE + KS_HALO];
  __shared__ u32 red[3][KS_THREADS / 32];
  const u64 tile_base = p_begin + (u64) blockIdx.x * KS_TILE;
  const u64 N = p_end;
  ks_load_tile(tis, n, tile_base, stext);
  __syncthreads();
  u32 code[KS_ITEMS], dist[KS_ITEMS];
  const int lo = threadIdx.x * KS_ITEMS;
  ks_scan_thread(stext, lo, kp, code, dist);
  u32 nf = 0, ns = 0, np = 0;
#pragma unroll
  for (int i = 0; i < KS_ITEMS; i++)
  {
    const u64 p = tile_base + lo + i;
    if (p < N)
    {
      if (dist[i] == 0)
      {
        np++;
      } else if (dist[i] < kp.h)
      {
        ns++;
      } else
      {
        nf++;
      }
    }
  }
  nf = warp_reduce_sum(nf);
  ns = warp_reduce_sum(ns);
  np = warp_reduce_sum(np);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
  {
    red[0][warp] = nf;
    red[1][warp] = ns;
    red[2][warp] = np;
  }
  __syncthreads();
  if (threadIdx.x < 3)
  {
    u32 s = 0;
    for (int w = 0; w < KS_THREADS / 32; w++)
    {
      s += red[threadIdx.x][w];
    }
    (threadIdx.x == 0 ? tile_full : threadIdx.x == 1 ? tile_side : tile_spec)[blockIdx.x] = s;
  }
}

// In-place exclusive sums of up to three per-tile count arrays, one CTA per array (blockIdx.x), eight tiles per thread
// and step; totals go to scalars[first_scalar + a].  (One CTA looping over the arrays with one tile per thread took
// 72 us for the 24 K tiles of 100 Mbp: three such single-CTA scans were 2 % of a build.)
constexpr int SCAN1_THREADS = 1024;
constexpr int SCAN1_ITEMS = 8;
__global__ void __launch_bounds__(SCAN1_THREADS)
tile_exclusive_sum_kernel(u32 *a0, u32 *a1, u32 *a2, u32 ntiles, u64 *scalars, int first_scalar)
{
  __shared__ u64 scratch[SCAN1_THREADS / 32 + 1];
  u32 *arr = blockIdx.x == 0 ? a0 : blockIdx.x == 1 ? a1 : a2;
  if (arr == nullptr)
  {
    return;
  }
  u64 carry = 0;
  for (u32 b = 0; b < ntiles; b += SCAN1_THREADS * SCAN1_ITEMS)
  {
    const u32 i0 = b + threadIdx.x * SCAN1_ITEMS;
    u32 v[SCAN1_ITEMS];
    u64 sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN1_ITEMS; k++)
    {
      v[k] = i0 + k < ntiles ? arr[i0 + k] : 0u;
      sum += v[k];
    }
    u64 total;
    u64 run = carry + block_exclusive_sum<u64, SCAN1_THREADS>(sum, scratch, &total);
#pragma unroll
    for (int k = 0; k < SCAN1_ITEMS; k++)
    {
      if (i0 + k < ntiles)
      {
        arr[i0 + k] = (u32) run;  // offsets stay below 2^32 because N does
      }
      run += v[k];
    }
    carry += total;
  }
  if (threadIdx.x == 0)
  {
    scalars[first_scalar + blockIdx.x] = carry;
  }
}

__global__ void __launch_bounds__(KS_THREADS)
classify_emit_kernel(const u8 *__restrict__ tis, u64 n, u64 p_begin, u64 p_end, KmerParams kp,
                     const u32 *__restrict__ tile_full,
                     const u32 *__restrict__ tile_side, const u32 *__restrict__ tile_spec,
                     u64 M /* full + side suffixes of the whole table */, u32 *__restrict__ main_key, u32 *__restrict__ main_pos,
                     u32 *__restrict__ side_key, u32 *__restrict__ side_pos, u32 *__restrict__ side_dj,
                     u32 *__restrict__ sa, u32 *__restrict__ rank, u32 *__restrict__ spec_list, u32 key_shl = 1)
{
  // key_shl = 1: keys are code << 1 | special-inside flag.  key_shl = 0 (sharded builds with 16-symbol DNA keys): the raw
  // padded code; the owner of the suffix rebases it to its own key range and adds the flag (suffix_sort_sharded).
  __shared__ __align__(16) u8 stext[KS_TILE + KS_HALO];
  // staging for coalesced output; index e is stored at e + e/32 so that the 16-strided per-thread writes hit
  // 32 different banks
  __shared__ u32 skey[KS_TILE + KS_TILE / 32];
  __shared__ u32 spos[KS_TILE + KS_TILE / 32];
  __shared__ u32 scratch[KS_THREADS / 32 + 1];
  const u64 tile_base = p_begin + (u64) blockIdx.x * KS_TILE;
  const u64 N = p_end;
  ks_load_tile(tis, n, tile_base, stext);
  __syncthreads();
  u32 code[KS_ITEMS], dist[KS_ITEMS];
  const int lo = threadIdx.x * KS_ITEMS;
  ks_scan_thread(stext, lo, kp, code, dist);
  u32 nf = 0, ns = 0, np = 0;
#pragma unroll
  for (int i = 0; i < KS_ITEMS; i++)
  {
    const u64 p = tile_base + lo + i;
    if (p < N)
    {
      if (dist[i] == 0)
      {
        np++;
      } else if (dist[i] < kp.h)
      {
        ns++;
      } else
      {
        nf++;
      }
    }
  }
  // counts per thread are <= 16, per tile <= 4096: pack full and side counts into one word for one scan
  u32 tot_fs, tot_p;
  const u32 ex_fs = block_exclusive_sum<u32, KS_THREADS>(nf | (ns << 16), scratch, &tot_fs);
  const u32 ex_p = block_exclusive_sum<u32, KS_THREADS>(np, scratch, &tot_p);
  u32 rf = ex_fs & 0xffffu, rs = ex_fs >> 16, rp = ex_p;
  const u32 tfull = tot_fs & 0xffffu;
  const u64 off_full = tile_full[blockIdx.x], off_side = tile_side[blockIdx.x], off_spec = tile_spec[blockIdx.x];
#pragma unroll
  for (int i = 0; i < KS_ITEMS; i++)
  {
    const u64 p = tile_base + lo + i;
    if (p < N)
    {
      if (dist[i] == 0)
      {
        if (spec_list != nullptr)
        {
          spec_list[off_spec + rp] = (u32) p;  // sharded build: the tail is assembled after the exchange
        } else
        {
          const u64 slot = M + off_spec + rp;  // tail of the table, position order
          sa[slot] = (u32) p;
          rank[p] = (u32) slot;
        }
        rp++;
      } else if (dist[i] < kp.h)
      {
        const u64 o = off_side + rs;
        side_key[o] = (code[i] << key_shl) | key_shl;
        side_pos[o] = (u32) p;
        side_dj[o] = kp.h - dist[i];
        rs++;
      } else
      {
        skey[rf + (rf >> 5)] = code[i] << key_shl;
        spos[rf + (rf >> 5)] = (u32) p;
        rf++;
      }
    }
  }
  __syncthreads();
  for (u32 r = threadIdx.x; r < tfull; r += KS_THREADS)
  {
    main_key[off_full + r] = skey[r + (r >> 5)];
    main_pos[off_full + r] = spos[r + (r >> 5)];
  }
}

// ---------------------------------------------------------------------------------------------------------
// Classification for the 4-letter alphabet from the PACKED text (sigma == 4): every position is independent -- one
// 16-byte record load (shared by the 32 consecutive positions of a warp) gives the h0 symbols and the distance to
// the next special, where the generic kernels above roll a base-sigma code over 16 + h0 symbols per thread and run
// twice (count, scan, emit).  The class totals are known beforehand from the special mask alone, the per-tile
// offsets of the rare classes come from a chained scan, so the text is swept once.
constexpr int CD_THREADS = 256;
constexpr int CD_ITEMS = 16;
constexpr int CD_TILE = CD_THREADS * CD_ITEMS;
constexpr int CD_WARPS = CD_THREADS / 32;

// One CTA per tile of CD_TILE positions (128 mask words), positions [p_begin, p_end), p_begin a multiple of CD_TILE.
// special-initial = mask bit; side = no mask bit but one among the next h-1 positions.  Per tile the two counts
// (packed: nside | nspec << 16), per group of 2^group_shift tiles their sums (packed: nside | nspec << 32, added
// atomically), and the totals in scalars[0] (nside) and scalars[1] (nspec).
__global__ void __launch_bounds__(CD_TILE / 32)
classify_dna_count_kernel(const u32 *__restrict__ spm, u64 p_begin, u64 p_end, u32 h, u32 group_shift,
                          u32 *__restrict__ tile_cnt, unsigned long long *__restrict__ group_cnt,
                          unsigned long long *__restrict__ scalars)
{
  __shared__ u32 red[2][CD_TILE / 32 / 32];
  const u64 w = (p_begin >> 5) + (u64) blockIdx.x * (CD_TILE / 32) + threadIdx.x;
  const u64 p0 = w * 32;
  u32 ns = 0, np = 0;
  if (p0 < p_end)
  {
    const u64 W = (u64) spm[w] | ((u64) spm[w + 1] << 32);
    const u32 valid = p_end - p0 >= 32 ? 0xffffffffu : ((1u << (u32) (p_end - p0)) - 1);
    u64 S = 0;
    for (u32 j = 1; j < h; j++)
    {
      S |= W >> j;
    }
    np = __popc((u32) W & valid);
    ns = __popc((u32) (S & ~W) & valid);
  }
  ns = warp_reduce_sum(ns);
  np = warp_reduce_sum(np);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
  {
    red[0][warp] = ns;
    red[1][warp] = np;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    u32 a = 0, b = 0;
    for (int k = 0; k < CD_TILE / 32 / 32; k++)
    {
      a += red[0][k];
      b += red[1][k];
    }
    tile_cnt[blockIdx.x] = a | (b << 16);
    if ((a | b) != 0)
    {
      atomicAdd(&group_cnt[blockIdx.x >> group_shift], (unsigned long long) a | ((unsigned long long) b << 32));
      atomicAdd(&scalars[0], (unsigned long long) a);
      atomicAdd(&scalars[1], (unsigned long long) b);
    }
  }
}

__global__ void __launch_bounds__(CD_THREADS, 4)
classify_dna_emit_kernel(TextPacked2 text, u64 p_begin, u64 p_end, u32 h, u64 M, u32 *__restrict__ main_key,
                         u32 *__restrict__ main_pos, u32 *__restrict__ side_key, u32 *__restrict__ side_pos,
                         u32 *__restrict__ side_dj, u32 *__restrict__ sa, u32 *__restrict__ rank,
                         u32 *__restrict__ spec_list, u32 group_shift, const u32 *__restrict__ tile_cnt,
                         const unsigned long long *__restrict__ group_cnt, u32 key_shl = 1)
{
  // key_shl = 1: keys are code << 1 | special-inside flag; 0: raw padded codes of h = 16 symbols (sharded builds with wide
  // keys, see classify_emit_kernel)
  __shared__ u32 cnt[3][CD_ITEMS * CD_WARPS];  // per (item, warp): members of the class; then exclusive offsets
  __shared__ u64 base[3];
  __shared__ u64 part[CD_WARPS];
  __shared__ ulonglong2 srec[CD_TILE / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 tile = blockIdx.x;
  const u64 tile_base = p_begin + (u64) tile * CD_TILE;
  const u32 lt_mask = (1u << lane) - 1;
  if (main_key == nullptr && tile_cnt[tile] == 0)
  {
    return;  // only the rare classes are wanted (the sort takes the full suffixes from the text) and the tile has none
  }
  // the tile's 128 records, one coalesced 16-byte load per thread (the 32 positions of a warp share ONE record: loaded
  // per position, a warp-wide load fetched 16 bytes and the sweep ran at 1 ms per 100 Mbp on memory latency alone)
  if (tid < CD_TILE / 32 && ((tile_base >> 5) + tid) * 32 < p_end + 64)  // (records end 4 behind the last position)
  {
    srec[tid] = __ldg(reinterpret_cast<const ulonglong2 *>(text.pk) + (tile_base >> 5) + tid);
  }
  // side and special suffixes in front of this tile: whole groups of tiles, then the tiles of my group (no chained
  // scan: with some hundred tiles in flight a look-back walks back over all of them, 20 us per tile)
  {
    const u32 g = tile >> group_shift;
    u64 acc = 0;  // nside | nspec << 32 (each below 2^32)
    for (u32 i = tid; i < g; i += CD_THREADS)
    {
      acc += group_cnt[i];
    }
    for (u32 t = (g << group_shift) + tid; t < tile; t += CD_THREADS)
    {
      const u32 c = tile_cnt[t];
      acc += (u64) (c & 0xffffu) | ((u64) (c >> 16) << 32);
    }
    acc = warp_reduce_sum(acc);
    if (lane == 0)
    {
      part[warp] = acc;
    }
  }
  __syncthreads();  // records staged
  const u32 o_rec = (u32) lane + 1, bitoff = 2 * o_rec, wi = bitoff >> 5, bo = bitoff & 31;
  const u32 gmask = ((2u << (((o_rec + h - 1) >> 3) - (o_rec >> 3))) - 1) << (o_rec >> 3);
  u32 key[CD_ITEMS];
  u64 dpk[2] = {0, 0};  // 8 bits per item: min(distance to the next special, h), h <= 16
#pragma unroll
  for (int i = 0; i < CD_ITEMS; i++)
  {
    const u64 p = tile_base + (u64) i * CD_THREADS + tid;
    u32 d = 0, code = 0;
    const bool valid = p < p_end;
    if (valid)
    {
      // position p is symbol o = lane + 1 of its record (tiles start at multiples of 32): the h <= 15 symbols behind
      // it are one 32-bit funnel shift out of the record's four words, at a per-thread constant offset
      const ulonglong2 rr = srec[(i * CD_THREADS + tid) >> 5];
      const u32 x0 = (u32) rr.x, x1 = (u32) (rr.x >> 32), x2 = (u32) rr.y, x3 = (u32) (rr.y >> 32);
      const u32 lo32 = wi == 0 ? x0 : (wi == 1 ? x1 : x2), hi32 = wi == 0 ? x1 : (wi == 1 ? x2 : x3);
      const u32 w32 = __funnelshift_r(lo32, hi32, bo);
      d = 32;
      if (((x3 >> 24) & gmask) != 0)  // a special among the record symbols [o, o + h): ask the mask
      {
        const u32 msk = text.mask_slow(p);
        d = msk ? (u32) (__ffs((int) msk) - 1) : 32u;
      }
      d = min(d, h);
      u32 r = __brev(w32);
      r = ((r & 0xAAAAAAAAu) >> 1) | ((r & 0x55555555u) << 1);
      code = r >> (32 - 2 * h);
      if (d < h)
      {
        code |= (u32) (((u64) 1 << (2 * (h - d))) - 1);  // padded with the largest letter behind the special
      }
    }
    key[i] = (code << key_shl) | (d < h ? key_shl : 0u);
    dpk[i >> 3] |= (u64) d << (8 * (i & 7));
    const u32 bf = __ballot_sync(FULL_MASK, valid && d == h);
    const u32 bs = __ballot_sync(FULL_MASK, valid && d > 0 && d < h);
    const u32 bp = __ballot_sync(FULL_MASK, valid && d == 0);
    if (lane == 0)
    {
      cnt[0][i * CD_WARPS + warp] = __popc(bf);
      cnt[1][i * CD_WARPS + warp] = __popc(bs);
      cnt[2][i * CD_WARPS + warp] = __popc(bp);
    }
  }
  __syncthreads();
  if (warp == 0)
  {
    // exclusive scan over the CD_ITEMS * CD_WARPS = 128 counters of every class, in position order (item, warp)
    constexpr int PER = CD_ITEMS * CD_WARPS / 32;
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
      u32 v[PER], sum = 0;
#pragma unroll
      for (int k = 0; k < PER; k++)
      {
        v[k] = cnt[c][lane * PER + k];
        sum += v[k];
      }
      const u32 incl = warp_inclusive_sum(sum);
      u32 run = incl - sum;
#pragma unroll
      for (int k = 0; k < PER; k++)
      {
        cnt[c][lane * PER + k] = run;
        run += v[k];
      }
    }
    if (lane == 0)
    {
      u64 acc = 0;
      for (int k = 0; k < CD_WARPS; k++)
      {
        acc += part[k];
      }
      const u64 ex_side = acc & 0xffffffffull, ex_spec = acc >> 32;
      base[0] = (tile_base - p_begin) - ex_side - ex_spec;  // every position in front of the tile has one class
      base[1] = ex_side;
      base[2] = ex_spec;
    }
  }
  __syncthreads();
  const u64 bfull = base[0], bside = base[1], bspec = base[2];
#pragma unroll
  for (int i = 0; i < CD_ITEMS; i++)
  {
    const u64 p = tile_base + (u64) i * CD_THREADS + tid;
    const bool valid = p < p_end;
    const u32 d = (u32) (dpk[i >> 3] >> (8 * (i & 7))) & 255u;
    const bool is_full = valid && d == h, is_side = valid && d > 0 && d < h, is_spec = valid && d == 0;
    const u32 bf = __ballot_sync(FULL_MASK, is_full);
    const u32 bs = __ballot_sync(FULL_MASK, is_side);
    const u32 bp = __ballot_sync(FULL_MASK, is_spec);
    const int e = i * CD_WARPS + warp;
    if (is_full)
    {
      if (main_key != nullptr)
      {
        const u64 o = bfull + cnt[0][e] + __popc(bf & lt_mask);
        main_key[o] = key[i];
        main_pos[o] = (u32) p;
      }
    } else if (is_side)
    {
      const u64 o = bside + cnt[1][e] + __popc(bs & lt_mask);
      side_key[o] = key[i];
      side_pos[o] = (u32) p;
      side_dj[o] = h - d;
    } else if (is_spec)
    {
      const u64 o = bspec + cnt[2][e] + __popc(bp & lt_mask);
      if (spec_list != nullptr)
      {
        spec_list[o] = (u32) p;  // sharded build: the tail is assembled after the exchange
      } else
      {
        const u64 slot = M + o;  // tail of the table, position order
        sa[slot] = (u32) p;
        rank[p] = (u32) slot;
      }
    }
  }
}

// per-bucket count of suffixes with a special inside the k-prefix (side suffixes with j < k)
__global__ void side_bucket_count_kernel(const u32 *__restrict__ side_key, const u32 *__restrict__ side_dj, u64 nside,
                                         u32 h0, u32 k, u32 kdiv, u32 *__restrict__ cnt, u32 key_shl = 1)
{
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nside)
  {
    const u32 j = h0 - side_dj[t];
    if (j < k)
    {
      atomicAdd(&cnt[(side_key[t] >> key_shl) / kdiv], 1u);
    }
  }
}

// =========================================================================================================
// 3. bck
__global__ void bck_left_kernel(const u32 *__restrict__ sorted_key, u64 m, u64 numofcodes, u32 kdiv,
                                u64 *__restrict__ bck, u64 key_base = 0)
{
  const u64 c = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= numofcodes)
  {
    return;
  }
  // keys are (code - key_base) << 1 | flag (key_base != 0: a rank's range of 16-symbol codes in a sharded build)
  if (c * kdiv <= key_base)
  {
    bck[2 * c] = 0;  // the k-mer starts at or in front of this range
    return;
  }
  const u64 target = (c * kdiv - key_base) << 1;  // first key of the k-mer c; beyond every 32-bit key when outside
  u64 lo = 0, hi = m;
  while (lo < hi)
  {
    const u64 mid = (lo + hi) >> 1;
    if ((u64) sorted_key[mid] < target)
    {
      lo = mid + 1;
    } else
    {
      hi = mid;
    }
  }
  bck[2 * c] = lo;
}

__global__ void bck_mid_kernel(u64 numofcodes, u64 N, u64 nspecial_tail, const u32 *__restrict__ cnt,
                               u64 *__restrict__ bck)
{
  const u64 c = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= numofcodes)
  {
    return;
  }
  u64 right, special = cnt[c];
  if (c + 1 == numofcodes)
  {
    right = N;                 // leftborder[numofcodes] = totallength + 1 (ppsort.c:229-230)
    special += nspecial_tail;  // special-initial suffixes and the sentinel live in the last bucket
  } else
  {
    right = bck[2 * (c + 1)];
  }
  bck[2 * c + 1] = right - special;
}

// =========================================================================================================
// 4. regroup
constexpr int RG_THREADS = 256;
constexpr int RG_ITEMS = 8;
constexpr int RG_TILE = RG_THREADS * RG_ITEMS;

// Loads the tile's keys (plus one halo key on each side) and slots into shared memory and derives per item
// flag bit0 = group head, bit1 = active (member of a group of size > 1).
template <typename KeyT, bool FIRST>
__device__ __forceinline__ void rg_flags(const KeyT *__restrict__ keys, const u32 *__restrict__ slots, u64 m,
                                         u64 tile_base, KeyT *skey /* [RG_TILE+2] */, u32 *sslot, u8 *sflag,
                                         u32 slot_base, KeyT single_mask)
{
#pragma unroll
  for (int j = 0; j < RG_ITEMS; j++)
  {
    const int l = j * RG_THREADS + threadIdx.x;
    const u64 t = tile_base + l;
    if (t < m)
    {
      skey[l + 1] = keys[t];
      sslot[l] = FIRST ? (u32) t + slot_base : slots[t];  // slots are global suffix-array indices
    }
  }
  if (threadIdx.x == 0)
  {
    skey[0] = tile_base > 0 ? keys[tile_base - 1] : (KeyT) 0;
  }
  if (threadIdx.x == 1)
  {
    const u64 t = tile_base + RG_TILE;
    skey[RG_TILE + 1] = t < m ? keys[t] : (KeyT) 0;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RG_ITEMS; j++)
  {
    const int l = j * RG_THREADS + threadIdx.x;
    const u64 t = tile_base + l;
    if (t < m)
    {
      const KeyT k = skey[l + 1];
      bool head = (t == 0) || (k != skey[l]);
      bool next_head = (t + 1 == m) || (skey[l + 2] != k);
      // a key with flag bits set (initial sort: bit 0; text rounds: the low flag field) belongs to a suffix with a
      // special inside the symbols the key covers: such suffixes are unique, and equal keys among them already stand
      // in text order (every sort is stable), which is their final order (bese.c:26-55: specials compare by position)
      head = head || ((k & single_mask) != 0);
      next_head = next_head || ((t + 1 < m) && ((skey[l + 2] & single_mask) != 0));
      sflag[l] = (u8) ((head ? 1 : 0) | ((head && next_head) ? 0 : 2));
    } else
    {
      sflag[l] = 0;
    }
  }
  __syncthreads();
}

template <typename KeyT, bool FIRST>
__global__ void __launch_bounds__(RG_THREADS)
regroup_summary_kernel(const KeyT *__restrict__ keys, const u32 *__restrict__ slots, u64 m,
                       u32 *__restrict__ tile_maxrank, u32 *__restrict__ tile_active, u32 slot_base, KeyT single_mask)
{
  __shared__ KeyT skey[RG_TILE + 2];
  __shared__ u32 sslot[RG_TILE];
  __shared__ __align__(8) u8 sflag[RG_TILE];
  __shared__ u32 red[2][RG_THREADS / 32];
  const u64 tile_base = (u64) blockIdx.x * RG_TILE;
  rg_flags<KeyT, FIRST>(keys, slots, m, tile_base, skey, sslot, sflag, slot_base, single_mask);
  u32 mx = 0, cnt = 0;
#pragma unroll
  for (int j = 0; j < RG_ITEMS; j++)
  {
    const int l = j * RG_THREADS + threadIdx.x;
    const u8 f = sflag[l];
    if (f & 1)
    {
      mx = max(mx, sslot[l]);
    }
    cnt += (f >> 1) & 1;
  }
  mx = warp_reduce_max(mx);
  cnt = warp_reduce_sum(cnt);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
  {
    red[0][warp] = mx;
    red[1][warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    u32 a = 0, b = 0;
    for (int w = 0; w < RG_THREADS / 32; w++)
    {
      a = max(a, red[0][w]);
      b += red[1][w];
    }
    tile_maxrank[blockIdx.x] = a;
    tile_active[blockIdx.x] = b;
  }
}

// single CTA, eight tiles per thread and step: tile_maxrank -> exclusive running max (the rank carried into the tile),
// tile_active -> exclusive sum (first output slot of the tile); scalars[4] = total active
__global__ void __launch_bounds__(SCAN1_THREADS)
regroup_scan_kernel(u32 *tile_maxrank, u32 *tile_active, u32 ntiles, u64 *scalars)
{
  __shared__ u64 scratch[SCAN1_THREADS / 32 + 1];
  __shared__ u32 mscratch[SCAN1_THREADS / 32];
  __shared__ u32 sincl[SCAN1_THREADS];
  u64 carry_sum = 0;
  u32 carry_max = 0;
  for (u32 b = 0; b < ntiles; b += SCAN1_THREADS * SCAN1_ITEMS)
  {
    const u32 i0 = b + threadIdx.x * SCAN1_ITEMS;
    u32 mv[SCAN1_ITEMS], cv[SCAN1_ITEMS];
    u64 sum = 0;
    u32 mx = 0;
#pragma unroll
    for (int k = 0; k < SCAN1_ITEMS; k++)
    {
      mv[k] = i0 + k < ntiles ? tile_maxrank[i0 + k] : 0u;
      cv[k] = i0 + k < ntiles ? tile_active[i0 + k] : 0u;
      sum += cv[k];
      mx = max(mx, mv[k]);
    }
    u64 total;
    u64 run = carry_sum + block_exclusive_sum<u64, SCAN1_THREADS>(sum, scratch, &total);
    const u32 incl = block_inclusive_max<u32, SCAN1_THREADS>(mx, mscratch);
    sincl[threadIdx.x] = incl;
    __syncthreads();
    u32 runmax = max(carry_max, threadIdx.x > 0 ? sincl[threadIdx.x - 1] : 0u);
    const u32 chunk_max = sincl[SCAN1_THREADS - 1];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SCAN1_ITEMS; k++)
    {
      if (i0 + k < ntiles)
      {
        tile_maxrank[i0 + k] = runmax;
        tile_active[i0 + k] = (u32) run;
      }
      runmax = max(runmax, mv[k]);
      run += cv[k];
    }
    carry_sum += total;
    carry_max = max(carry_max, chunk_max);
  }
  if (threadIdx.x == 0)
  {
    scalars[4] = carry_sum;
  }
}

template <typename KeyT, bool FIRST>
__global__ void __launch_bounds__(RG_THREADS)
regroup_apply_kernel(const KeyT *__restrict__ keys, const u32 *__restrict__ pos, const u32 *__restrict__ slots, u64 m,
                     const u32 *__restrict__ tile_maxrank, const u32 *__restrict__ tile_active, u32 *__restrict__ sa,
                     u32 *__restrict__ rank, u64 *__restrict__ next_key, u32 *__restrict__ next_pos,
                     u32 *__restrict__ next_slots, u32 rank_shift /* of the keys read */,
                     u32 next_shift /* of the keys written for the next round */, u8 *__restrict__ hround, u32 round,
                     u32 *__restrict__ rank_out /* if set: rank_out[t] = rank instead of the scatter rank[pos] */,
                     u32 slot_base /* first global slot of this shard; sa and hround are shard-local arrays */,
                     KeyT single_mask,
                     u32 *__restrict__ rank_slot /* if set: rank_slot[slot - slot_base] = rank and nothing else (the
                                                    rank list in suffix-array order of a deferred scatter) */)
{
  __shared__ KeyT skey[RG_TILE + 2];
  __shared__ u32 sslot[RG_TILE];
  __shared__ __align__(8) u8 sflag[RG_TILE];
  __shared__ u32 srank[RG_TILE];
  __shared__ u32 sdest[RG_TILE];
  __shared__ u32 scratch[RG_THREADS / 32 + 1];
  __shared__ u32 sincl[RG_THREADS];
  const u64 tile_base = (u64) blockIdx.x * RG_TILE;
  rg_flags<KeyT, FIRST>(keys, slots, m, tile_base, skey, sslot, sflag, slot_base, single_mask);
  // blocked phase: thread owns local items [tid*8, tid*8+8)
  const int b0 = threadIdx.x * RG_ITEMS;
  const u64 f8 = *reinterpret_cast<const u64 *>(sflag + b0);
  u32 lastrank = 0, nact = 0;
#pragma unroll
  for (int i = 0; i < RG_ITEMS; i++)
  {
    const u32 f = (u32) (f8 >> (8 * i)) & 0xffu;
    if (f & 1)
    {
      lastrank = sslot[b0 + i];  // slots increase, so the last head carries the maximum
    }
    nact += (f >> 1) & 1;
  }
  const u32 ex_act = block_exclusive_sum<u32, RG_THREADS>(nact, scratch, nullptr);
  const u32 incl = block_inclusive_max<u32, RG_THREADS>(lastrank, scratch);
  sincl[threadIdx.x] = incl;
  __syncthreads();
  u32 cur = max(tile_maxrank[blockIdx.x], threadIdx.x > 0 ? sincl[threadIdx.x - 1] : 0u);
  u32 dst = tile_active[blockIdx.x] + ex_act;
#pragma unroll
  for (int i = 0; i < RG_ITEMS; i++)
  {
    const u32 f = (u32) (f8 >> (8 * i)) & 0xffu;
    if (f & 1)
    {
      cur = sslot[b0 + i];
    }
    srank[b0 + i] = cur;
    if (f & 2)
    {
      sdest[b0 + i] = dst++;
    } else
    {
      sdest[b0 + i] = NONE32;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RG_ITEMS; j++)
  {
    const int l = j * RG_THREADS + threadIdx.x;
    const u64 t = tile_base + l;
    if (t < m)
    {
      const u32 p = pos[t], slot = sslot[l], r = srank[l], d = sdest[l];
      if (!FIRST)
      {
        sa[slot - slot_base] = p;  // FIRST: the sorted position buffer already is sa[0..m)
      }
      if (rank_slot != nullptr)
      {
        rank_slot[slot - slot_base] = r;
      } else if (rank_out != nullptr)
      {
        rank_out[t] = r;
      } else
      {
        rank[p] = r;
      }
      if (d != NONE32)
      {
        next_key[d] = (u64) r << next_shift;
        next_pos[d] = p;
        next_slots[d] = slot;
      }
      if (!FIRST && hround != nullptr && t > 0 && (sflag[l] & 1))
      {
        // a head whose predecessor belongs to the same old group was split off in this round: the two
        // suffixes agree on the depth of the round (lcp lower bound for the direct lcp kernel)
        if ((u64) (skey[l + 1] >> rank_shift) == (u64) (skey[l] >> rank_shift))
        {
          hround[slot - slot_base] = (u8) round;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Register-blocked regroup kernels (default).  Same contract and tile size as regroup_summary_kernel /
// regroup_apply_kernel above, but every thread keeps its eight consecutive entries in registers (16-byte loads and
// stores), takes the keys of its neighbours from a two-word shared-memory exchange and derives flags, ranks and
// destinations without staging the tile in shared memory.  The staged kernels execute 55 (summary) and 97 (apply)
// instructions per entry and are issue-bound (ncu: 80 % issue utilisation, 9e7 bank conflicts per 100 M entries).
template <typename KeyT> struct RgVec;
template <> struct RgVec<u32>
{
  static __device__ __forceinline__ void load8(const u32 *p, u32 *k)
  {
    const uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1];
    k[0] = a.x, k[1] = a.y, k[2] = a.z, k[3] = a.w, k[4] = b.x, k[5] = b.y, k[6] = b.z, k[7] = b.w;
  }
};
template <> struct RgVec<u64>
{
  static __device__ __forceinline__ void load8(const u64 *p, u64 *k)
  {
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
      const ulonglong2 a = reinterpret_cast<const ulonglong2 *>(p)[j];
      k[2 * j] = a.x;
      k[2 * j + 1] = a.y;
    }
  }
};
__device__ __forceinline__ void rg_store8(u32 *p, const u32 *v)
{
  reinterpret_cast<uint4 *>(p)[0] = make_uint4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<uint4 *>(p)[1] = make_uint4(v[4], v[5], v[6], v[7]);
}

// Loads the thread's eight entries (keys, and slots unless FIRST) and derives heads / active as bit masks.
// sx: shared exchange [2 * RG_THREADS] of KeyT.  Entries at or behind m get head = active = 0.
template <typename KeyT, bool FIRST>
__device__ __forceinline__ void rg2_flags(const KeyT *__restrict__ keys, const u32 *__restrict__ slots, u64 m, u64 t0,
                                          u32 slot_base, KeyT single_mask, KeyT *sx, KeyT *k, u32 *slot, KeyT &prev,
                                          u32 &heads, u32 &act)
{
  const int tid = threadIdx.x;
  const bool full = t0 + RG_ITEMS <= m;
  if (full)
  {
    RgVec<KeyT>::load8(keys + t0, k);
    if (!FIRST)
    {
      RgVec<u32>::load8(slots + t0, slot);
    }
  } else
  {
#pragma unroll
    for (int i = 0; i < RG_ITEMS; i++)
    {
      k[i] = t0 + i < m ? keys[t0 + i] : (KeyT) 0;
      if (!FIRST)
      {
        slot[i] = t0 + i < m ? slots[t0 + i] : 0u;
      }
    }
  }
  if (FIRST)
  {
#pragma unroll
    for (int i = 0; i < RG_ITEMS; i++)
    {
      slot[i] = (u32) (t0 + i) + slot_base;  // slots are global suffix-array indices
    }
  }
  sx[tid] = k[0];
  sx[RG_THREADS + tid] = k[RG_ITEMS - 1];
  __syncthreads();
  KeyT next;
  if (tid > 0)
  {
    prev = sx[RG_THREADS + tid - 1];
  } else
  {
    prev = t0 > 0 ? keys[t0 - 1] : (KeyT) 0;
  }
  if (tid < RG_THREADS - 1)
  {
    next = sx[tid + 1];
  } else
  {
    next = t0 + RG_ITEMS < m ? keys[t0 + RG_ITEMS] : (KeyT) 0;
  }
  heads = 0;
  act = 0;
  u32 headbits = 0;  // bit i: entry i is a head; bit 8: the entry behind my last one is a head (or the end)
#pragma unroll
  for (int i = 0; i <= RG_ITEMS; i++)
  {
    const u64 t = t0 + i;
    const KeyT cur = i < RG_ITEMS ? k[i] : next;
    const KeyT before = i == 0 ? prev : k[i - 1];
    // a key with flag bits set belongs to a suffix that is unique and already in its final place (see rg_flags)
    const bool h = t >= m || t == 0 || cur != before || (cur & single_mask) != 0;
    headbits |= h ? (1u << i) : 0u;
  }
#pragma unroll
  for (int i = 0; i < RG_ITEMS; i++)
  {
    if (t0 + i < m)
    {
      const bool h = (headbits >> i) & 1u, nh = (headbits >> (i + 1)) & 1u;
      heads |= h ? (1u << i) : 0u;
      act |= (h && nh) ? 0u : (1u << i);
    }
  }
}

template <typename KeyT, bool FIRST>
__global__ void __launch_bounds__(RG_THREADS)
regroup2_summary_kernel(const KeyT *__restrict__ keys, const u32 *__restrict__ slots, u64 m,
                        u32 *__restrict__ tile_maxrank, u32 *__restrict__ tile_active, u32 slot_base, KeyT single_mask)
{
  __shared__ KeyT sx[2 * RG_THREADS];
  __shared__ u32 red[2][RG_THREADS / 32];
  const u64 t0 = (u64) blockIdx.x * RG_TILE + (u64) threadIdx.x * RG_ITEMS;
  KeyT k[RG_ITEMS], prev;
  u32 slot[RG_ITEMS], heads, act;
  rg2_flags<KeyT, FIRST>(keys, slots, m, t0, slot_base, single_mask, sx, k, slot, prev, heads, act);
  u32 mx = heads ? slot[31 - __clz((int) heads)] : 0u;  // slots increase: the last head carries the maximum
  u32 cnt = __popc(act);
  mx = warp_reduce_max(mx);
  cnt = warp_reduce_sum(cnt);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
  {
    red[0][warp] = mx;
    red[1][warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    u32 a = 0, b = 0;
    for (int w = 0; w < RG_THREADS / 32; w++)
    {
      a = max(a, red[0][w]);
      b += red[1][w];
    }
    tile_maxrank[blockIdx.x] = a;
    tile_active[blockIdx.x] = b;
  }
}

template <typename KeyT, bool FIRST>
__global__ void __launch_bounds__(RG_THREADS)
regroup2_apply_kernel(const KeyT *__restrict__ keys, const u32 *__restrict__ pos, const u32 *__restrict__ slots, u64 m,
                      const u32 *__restrict__ tile_maxrank, const u32 *__restrict__ tile_active, u32 *__restrict__ sa,
                      u32 *__restrict__ rank, u64 *__restrict__ next_key, u32 *__restrict__ next_pos,
                      u32 *__restrict__ next_slots, u32 rank_shift, u32 next_shift, u8 *__restrict__ hround, u32 round,
                      u32 *__restrict__ rank_out, u32 slot_base, KeyT single_mask, u32 *__restrict__ rank_slot)
{
  __shared__ KeyT sx[2 * RG_THREADS];
  __shared__ u32 scratch[RG_THREADS / 32 + 1];
  __shared__ u32 sincl[RG_THREADS];
  const u64 t0 = (u64) blockIdx.x * RG_TILE + (u64) threadIdx.x * RG_ITEMS;
  KeyT k[RG_ITEMS], prev;
  u32 slot[RG_ITEMS], heads, act;
  rg2_flags<KeyT, FIRST>(keys, slots, m, t0, slot_base, single_mask, sx, k, slot, prev, heads, act);
  const u32 lastrank = heads ? slot[31 - __clz((int) heads)] : 0u;
  const u32 nact = __popc(act);
  const u32 ex_act = block_exclusive_sum<u32, RG_THREADS>(nact, scratch, nullptr);
  const u32 incl = block_inclusive_max<u32, RG_THREADS>(lastrank, scratch);
  sincl[threadIdx.x] = incl;
  __syncthreads();
  u32 cur = max(tile_maxrank[blockIdx.x], threadIdx.x > 0 ? sincl[threadIdx.x - 1] : 0u);
  u32 dst = tile_active[blockIdx.x] + ex_act;
  const bool full = t0 + RG_ITEMS <= m;
  u32 p[RG_ITEMS], r[RG_ITEMS];
  if (full)
  {
    RgVec<u32>::load8(pos + t0, p);
  } else
  {
#pragma unroll
    for (int i = 0; i < RG_ITEMS; i++)
    {
      p[i] = t0 + i < m ? pos[t0 + i] : 0u;
    }
  }
#pragma unroll
  for (int i = 0; i < RG_ITEMS; i++)
  {
    if ((heads >> i) & 1u)
    {
      cur = slot[i];
    }
    r[i] = cur;
  }
  // the ranks: in suffix-array order (vector store), by slot (deferred scatter list) or scattered by position
  if (rank_slot == nullptr && rank_out != nullptr && full)
  {
    rg_store8(rank_out + t0, r);
  }
#pragma unroll
  for (int i = 0; i < RG_ITEMS; i++)
  {
    const u64 t = t0 + i;
    if (t < m)
    {
      if (!FIRST)
      {
        sa[slot[i] - slot_base] = p[i];  // FIRST: the sorted position buffer already is sa[0..m)
      }
      if (rank_slot != nullptr)
      {
        rank_slot[slot[i] - slot_base] = r[i];
      } else if (rank_out != nullptr)
      {
        if (!full)
        {
          rank_out[t] = r[i];
        }
      } else
      {
        rank[p[i]] = r[i];
      }
      if ((act >> i) & 1u)
      {
        next_key[dst] = (u64) r[i] << next_shift;
        next_pos[dst] = p[i];
        next_slots[dst] = slot[i];
        dst++;
      }
      if (!FIRST && hround != nullptr && t > 0 && ((heads >> i) & 1u))
      {
        // a head whose predecessor belongs to the same old group was split off in this round: the two
        // suffixes agree on the depth of the round (lcp lower bound for the direct lcp kernel)
        const KeyT before = i == 0 ? prev : k[i - 1];
        if ((u64) (k[i] >> rank_shift) == (u64) (before >> rank_shift))
        {
          hround[slot[i] - slot_base] = (u8) round;
        }
      }
    }
  }
}

// 5a. text round: low key bits = the next L symbols behind the h the group agrees on, as an order-preserving code.
// Window of L letters: code (base sigma, first symbol in the top digit) << F.  A special (or the end of the text) at
// offset d < L: the letters in front of it padded with the largest letter, flag field = L - d > 0 in the low F bits.
// This orders exactly like the symbols do (bese.c:26-55: a special is larger than every letter): the padded code is
// >= the code of every window with the same d letters in front, equal codes are split by the flag (full window
// first, then the later special first), and equal flagged keys differ only in the POSITION of their special, i.e.
// stand in text order already -- regroup makes every flagged key a singleton.
struct TextKeyParams
{
  const u8 *tis;
  TextPacked2 packed;  // packed.pk == nullptr: byte path
  TextPacked5 packed5; // packed5.pk != nullptr: 5-bit records (L <= 8)
  u32 len;             // L
  u32 flag_bits;       // F = bits_for(L)
  u32 sigma;
};
__global__ void __launch_bounds__(256)
text_key_kernel(u64 *__restrict__ key, const u32 *__restrict__ pos, u64 m, u64 h, TextKeyParams tk)
{
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m)
  {
    return;
  }
  const u64 q = (u64) pos[t] + h;  // <= n: an active suffix has no special among its first h symbols
  const u32 L = tk.len;
  u64 code = 0;
  u32 flag = 0;
  if (tk.packed.pk != nullptr)
  {
    const TextPacked2::Win tw = tk.packed.win(q);  // 28 symbols >= L
    const u64 w = tw.w;
    const u32 msk = tw.m;
    const u32 d = msk ? (u32) (__ffs((int) msk) - 1) : 32u;
    u64 r = __brevll(w);
    r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);
    code = r >> (64 - 2 * L);
    if (d < L)
    {
      code |= ((u64) 1 << (2 * (L - d))) - 1;
      flag = L - d;
    }
  } else if (tk.packed5.pk != nullptr)
  {
    const u64 w = tk.packed5.window(q);  // 8 symbols >= L
    for (u32 i = 0; i < L; i++)
    {
      const u32 a = TextPacked5::window_symbol(w, i);
      if (a >= ESA_WILDCARD)
      {
        for (u32 j = i; j < L; j++)
        {
          code = code * tk.sigma + (tk.sigma - 1);
        }
        flag = L - i;
        break;
      }
      code = code * tk.sigma + a;
    }
  } else
  {
    for (u32 i = 0; i < L; i++)
    {
      const u32 a = tk.tis[q + i];  // tis is padded with 255 behind position n
      if (a >= ESA_WILDCARD)
      {
        for (u32 j = i; j < L; j++)
        {
          code = code * tk.sigma + (tk.sigma - 1);
        }
        flag = L - i;
        break;
      }
      code = code * tk.sigma + a;
    }
  }
  key[t] |= (code << tk.flag_bits) | (u64) flag;
}

// number of symbols a text key can hold in `free_bits` bits, and the width of the key
struct TextKeyShape
{
  u32 len, code_bits, flag_bits;
};
TextKeyShape text_key_shape(u32 sigma, u32 free_bits, bool packed)
{
  TextKeyShape best{0, 0, 0};
  const u32 max_len = 28;
  double pw = 1.0;
  for (u32 L = 1; L <= max_len; L++)
  {
    pw *= (double) sigma;
    u32 cb = 0;
    if (packed || sigma == 4)
    {
      cb = 2 * L;
    } else
    {
      if (pw > 72057594037927936.0)  // 2^56
      {
        break;
      }
      cb = bits_for((u64) pw - 1);
    }
    const u32 fb = bits_for(L);
    if (cb + fb > free_bits || cb + fb > 56)
    {
      break;
    }
    best = TextKeyShape{L, cb, fb};
  }
  return best;
}

u32 text_rounds_setting()
{
  static const int v = [] {
    const char *e = getenv("ESA_TEXT_ROUNDS");
    return e != nullptr ? atoi(e) : 1;
  }();
  return (u32) (v < 0 ? 0 : (v > 8 ? 8 : v));
}

// 5b. doubling key: low bits = rank of the suffix h symbols further
__global__ void gather_rank_kernel(u64 *__restrict__ key, const u32 *__restrict__ pos, u64 m, u64 h,
                                   const u32 *__restrict__ rank)
{
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < m)
  {
    key[t] |= (u64) rank[(u64) pos[t] + h];
  }
}

// second half of the two-level rank scatter: (pos, rank) pairs arrive partitioned by the top bits of pos, so
// the writes of one stretch of the input fall into one L2-sized window of the rank array
__global__ void scatter_rank_kernel(const u32 *__restrict__ pos, const u32 *__restrict__ val, u64 m,
                                    u32 *__restrict__ rank)
{
  const u64 t = ((u64) blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (t + 4 <= m)
  {
    const uint4 p = *reinterpret_cast<const uint4 *>(pos + t), v = *reinterpret_cast<const uint4 *>(val + t);
    rank[p.x] = v.x;
    rank[p.y] = v.y;
    rank[p.z] = v.z;
    rank[p.w] = v.w;
  } else
  {
    for (u64 q = t; q < m; q++)
    {
      rank[pos[q]] = val[q];
    }
  }
}

// Writing rank[pos[t]] = val[t] for m pairs in suffix-array order is a random 4-byte scatter over 4N bytes:
// every write costs a 32-byte sector fetch plus a 32-byte write-back (measured 3.7 ms per 100 M on B200).
// Two levels instead: one radix pass partitions the pairs by the top bits of pos (sequential traffic), then
// the scatter runs window by window inside L2.  Worth it only for large m.
constexpr u64 PARTITIONED_SCATTER_MIN = (u64) 1 << 22;
constexpr u32 SCATTER_WINDOW_BITS = 23;  // 2^23 positions = 32 MB of the rank array per window

bool scatter_ranks_partitioned(esa_ctx *ctx, u32 *pos, u32 *val, u64 m, u32 pos_bits, u32 *rank, u32 *tmp_pos,
                               u32 *tmp_val, void *sort_temp, size_t temp_bytes)
{
  if (m < PARTITIONED_SCATTER_MIN || pos_bits <= SCATTER_WINDOW_BITS)
  {
    return false;
  }
  const u32 pbits = min(8u, pos_bits - SCATTER_WINDOW_BITS);
  const int w = radix_sort_pairs_u32(ctx, pos, tmp_pos, val, tmp_val, m, pos_bits - pbits, pos_bits, sort_temp,
                                     temp_bytes);
  const u32 *p = w ? tmp_pos : pos, *v = w ? tmp_val : val;
  ESA_LAUNCH(ctx, ESA_K_REGROUP, 12.0 * m,
             (scatter_rank_kernel<<<div_up_u32(div_up_u64(m, 4), 256), 256, 0, ctx->stream>>>(p, v, m, rank)));
  return true;
}

// scratch of the two-level scatter for one doubling round, carved from the arena when there is room (else the
// ranks are scattered directly by regroup_apply); released with ar.rewind(mark) after the scatter has been queued
struct ScatterScratch
{
  u32 *rank_out = nullptr, *tp = nullptr, *tv = nullptr;
};
ScatterScratch round_scatter_scratch(Arena &ar, u64 m, u32 pos_bits)
{
  ScatterScratch sc;
  if (m >= PARTITIONED_SCATTER_MIN && pos_bits > SCATTER_WINDOW_BITS
      && ar.high - ar.top > 12 * (m + 64) + ((size_t) 8 << 20))
  {
    sc.rank_out = ar.alloc<u32>(m + 64);
    sc.tp = ar.alloc<u32>(m + 64);
    sc.tv = ar.alloc<u32>(m + 64);
  }
  return sc;
}

__global__ void iota_kernel(u32 *__restrict__ v, u64 count)
{
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count)
  {
    v[t] = (u32) t;
  }
}

__global__ void scatter_iota_kernel(const u32 *__restrict__ sa, u64 count, u32 *__restrict__ isa)
{
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count)
  {
    isa[sa[t]] = (u32) t;
  }
}

// Shape of the refinement rounds of one build: which rounds take their secondary key from the text, how wide the
// keys of a round are, and the state of the deferred rank scatter.
struct RefinePlan
{
  u32 rank_shift = 0;   // bits of a rank: secondary-key width of the doubling rounds
  u32 text_rounds = 0;  // rounds 1..text_rounds are text rounds
  TextKeyShape tks{0, 0, 0};
  TextKeyParams tk;
  u32 *rank_sa = nullptr;  // ranks in suffix-array order (one entry per sorted suffix of this device)
  bool pending = false;    // the scatter rank[sa[t]] = rank_sa[t] has not been carried out yet

  bool is_text(u32 round) const { return round >= 1 && round <= text_rounds; }
  // a round's key is rank << shift | secondary key
  u32 shift_of(u32 round) const { return is_text(round) ? tks.code_bits + tks.flag_bits : rank_shift; }
  u64 mask_of(u32 round) const { return is_text(round) ? ((u64) 1 << tks.flag_bits) - 1 : 0; }
};

RefinePlan make_refine_plan(esa_ctx *ctx, u32 rank_shift, u32 h0)
{
  RefinePlan plan;
  plan.rank_shift = rank_shift;
  const bool packed = ctx->d_pk != nullptr;
  plan.tks = text_key_shape(ctx->sigma, 64 - rank_shift, packed);
  plan.text_rounds = plan.tks.len > 0 ? text_rounds_setting() : 0;
  plan.tk.tis = ctx->d_tis;
  plan.tk.packed = TextPacked2{packed ? ctx->d_pk : nullptr, ctx->d_spm, ctx->d_tis};
  plan.tk.packed5 = TextPacked5{(!packed && plan.tks.len <= TextPacked5::WINDOW) ? ctx->d_pk5 : nullptr};
  plan.tk.len = plan.tks.len;
  plan.tk.flag_bits = plan.tks.flag_bits;
  plan.tk.sigma = ctx->sigma;
  ctx->sched = DepthSched{h0, plan.text_rounds > 0 ? plan.tks.len : 0, plan.text_rounds};
  return plan;
}

// The scatter of the first regroup is deferred when a text round follows it (or no round at all) and the arena has
// room for the rank list; M = sorted suffixes of this device, m = active ones.
void plan_deferred_scatter(esa_ctx *ctx, RefinePlan &plan, u64 M, u64 m)
{
  Arena &ar = ctx->arena;
  plan.pending = false;
  if (plan.text_rounds == 0 && m > 0)
  {
    return;
  }
  const size_t need = 4 * (M + 64) + 20 * (m + 64) + ((size_t) 64 << 20);
  if (ar.high - ar.top < need)
  {
    return;
  }
  plan.rank_sa = ar.alloc<u32>(M + 64);
  plan.pending = true;
}

// carries the deferred scatter out (before the first round that reads rank[]); sa = this device's suffix array
void flush_deferred_scatter(esa_ctx *ctx, RefinePlan &plan, const u32 *sa, u64 M, u32 *rank, void *sort_temp,
                            size_t temp_bytes)
{
  if (!plan.pending)
  {
    return;
  }
  plan.pending = false;
  if (M == 0)
  {
    return;
  }
  Arena &ar = ctx->arena;
  const size_t mark = ar.mark();
  bool done = false;
  if (M >= PARTITIONED_SCATTER_MIN && plan.rank_shift > SCATTER_WINDOW_BITS
      && ar.high - ar.top > 8 * (M + 64) + ((size_t) 8 << 20))
  {
    u32 *tp = ar.alloc<u32>(M + 64), *tv = ar.alloc<u32>(M + 64);
    done = scatter_ranks_partitioned(ctx, const_cast<u32 *>(sa), plan.rank_sa, M, plan.rank_shift, rank, tp, tv,
                                     sort_temp, temp_bytes);
  }
  if (!done)
  {
    ESA_LAUNCH(ctx, ESA_K_REGROUP, 12.0 * M,
               (scatter_rank_kernel<<<div_up_u32(div_up_u64(M, 4), 256), 256, 0, ctx->stream>>>(sa, plan.rank_sa, M,
                                                                                                   rank)));
  }
  ar.rewind(mark);
}

u64 read_scalar(esa_ctx *ctx, const u64 *d_scalars, int idx)
{
  read_back_u64(ctx, d_scalars + idx, 1);
  return ctx->h_scalars[0];
}

// scratch of the regroup step: per-tile carries
struct RegroupScratch
{
  u32 *tile_maxrank = nullptr, *tile_active = nullptr;
};
RegroupScratch alloc_regroup_scratch(Arena &ar, u64 max_items)
{
  const size_t tiles = (size_t) div_up_u64(max_items, RG_TILE) + 1;
  RegroupScratch rs;
  rs.tile_maxrank = ar.alloc<u32>(tiles);
  rs.tile_active = ar.alloc<u32>(tiles);
  return rs;
}

// One regroup step over m sorted (key, position[, slot]) entries: summary kernel, single-CTA scan over the tiles,
// apply kernel; returns the number of active suffixes it left.
// (A single-pass version -- the apply kernel with chained scans over the tiles instead of the summary kernel and the
// scan -- was measured and dropped: with some hundred tiles of 2048 entries in flight every tile walks back over all
// of them before it meets a prefix, and even with the whole CTA reading 256 predecessors per round trip the sweep
// saved over the keys was lost again: 1.40 vs 1.41 ms per 100 Mbp, 33.9 vs 29.6 ms on the repeat-rich 250 Mbp.)
// The register-blocked kernels serve the FIRST regroup (its rank list is one vector store per thread: 1.22 -> 1.06 ms
// per 100 Mbp); the rounds keep the staged kernels, whose thread-striped stores by slot coalesce where the blocked
// ones write with a stride of eight entries (repeat-rich 250 Mbp: 29.6 ms staged, 36.6 ms blocked).
// ESA_REGROUP_STAGED=1 / ESA_REGROUP_BLOCKED=1 force one kind everywhere.
template <bool FIRST> bool staged_regroup_for()
{
  static const bool staged = getenv("ESA_REGROUP_STAGED") != nullptr, blocked = getenv("ESA_REGROUP_BLOCKED") != nullptr;
  return staged ? true : (blocked ? false : !FIRST);
}

// first half: per-tile carries and the number of active suffixes the step will leave
template <typename KeyT, bool FIRST>
u64 run_regroup_summary(esa_ctx *ctx, const RegroupScratch &rs, u64 *d_scalars, const KeyT *keys, const u32 *slots,
                        u64 m, u32 slot_base, KeyT single_mask)
{
  cudaStream_t s = ctx->stream;
  const u32 rtiles = (u32) div_up_u64(m, RG_TILE);
  if (staged_regroup_for<FIRST>())
  {
    ESA_LAUNCH(ctx, ESA_K_REGROUP, (FIRST ? 4.0 : 12.0) * m,
               (regroup_summary_kernel<KeyT, FIRST><<<rtiles, RG_THREADS, 0, s>>>(keys, slots, m, rs.tile_maxrank,
                                                                                   rs.tile_active, slot_base,
                                                                                   single_mask)));
  } else
  {
    ESA_LAUNCH(ctx, ESA_K_REGROUP, (FIRST ? 4.0 : 12.0) * m,
               (regroup2_summary_kernel<KeyT, FIRST><<<rtiles, RG_THREADS, 0, s>>>(keys, slots, m, rs.tile_maxrank,
                                                                                    rs.tile_active, slot_base,
                                                                                    single_mask)));
  }
  ESA_LAUNCH(ctx, ESA_K_MISC, 0,
             (regroup_scan_kernel<<<1, SCAN1_THREADS, 0, s>>>(rs.tile_maxrank, rs.tile_active, rtiles, d_scalars)));
  return read_scalar(ctx, d_scalars, 4);
}

// second half: ranks, suffix array entries, lcp lower bounds and the next active list (m_next entries)
template <typename KeyT, bool FIRST>
void run_regroup_apply(esa_ctx *ctx, const RegroupScratch &rs, u64 m_next, const KeyT *keys, const u32 *pos,
                       const u32 *slots, u64 m, u32 *sa, u32 *rank, u64 *next_key, u32 *next_pos, u32 *next_slots,
                       u32 cur_shift, u32 next_shift, u8 *hround, u32 round, u32 *rank_out, u32 slot_base,
                       KeyT single_mask, u32 *rank_slot)
{
  cudaStream_t s = ctx->stream;
  const u32 rtiles = (u32) div_up_u64(m, RG_TILE);
  const double bytes = (FIRST ? 12.0 : 24.0) * m + 16.0 * m_next;
  if (staged_regroup_for<FIRST>())
  {
    ESA_LAUNCH(ctx, ESA_K_REGROUP, bytes,
               (regroup_apply_kernel<KeyT, FIRST><<<rtiles, RG_THREADS, 0, s>>>(
                   keys, pos, slots, m, rs.tile_maxrank, rs.tile_active, sa, rank, next_key, next_pos, next_slots,
                   cur_shift, next_shift, hround, round, rank_out, slot_base, single_mask, rank_slot)));
  } else
  {
    ESA_LAUNCH(ctx, ESA_K_REGROUP, bytes,
               (regroup2_apply_kernel<KeyT, FIRST><<<rtiles, RG_THREADS, 0, s>>>(
                   keys, pos, slots, m, rs.tile_maxrank, rs.tile_active, sa, rank, next_key, next_pos, next_slots,
                   cur_shift, next_shift, hround, round, rank_out, slot_base, single_mask, rank_slot)));
  }
}

template <typename KeyT, bool FIRST>
u64 run_regroup(esa_ctx *ctx, const RegroupScratch &rs, u64 *d_scalars, const KeyT *keys, const u32 *pos,
                const u32 *slots, u64 m, u32 *sa, u32 *rank, u64 *next_key, u32 *next_pos, u32 *next_slots,
                u32 cur_shift, u32 next_shift, u8 *hround, u32 round, u32 *rank_out, u32 slot_base, KeyT single_mask,
                u32 *rank_slot)
{
  const u64 m_next = run_regroup_summary<KeyT, FIRST>(ctx, rs, d_scalars, keys, slots, m, slot_base, single_mask);
  run_regroup_apply<KeyT, FIRST>(ctx, rs, m_next, keys, pos, slots, m, sa, rank, next_key, next_pos, next_slots,
                                 cur_shift, next_shift, hround, round, rank_out, slot_base, single_mask, rank_slot);
  return m_next;
}

}  // namespace

// choose h0: the largest h <= 31 with sigma^h <= 2^31
static u32 choose_h0(u32 sigma)
{
  u32 h = 0;
  u64 p = 1;
  while (h < 31 && p * sigma <= ((u64) 1 << 31))
  {
    p *= sigma;
    h++;
  }
  return h == 0 ? 1 : h;
}

void suffix_sort(esa_ctx *ctx, u32 prefixlength, bool want_bck, SortOutcome *outcome, size_t scratch_mark)
{
  const u64 n = ctx->n, N = n + 1;
  const u32 sigma = ctx->sigma;
  cudaStream_t s = ctx->stream;
  Arena &ar = ctx->arena;
  const u32 h0 = choose_h0(sigma);
  const KmerParams kp = make_kmer_params(sigma, h0);
  const u32 key_bits = bits_for(((u64) kp.all_pad << 1) | 1);
  u32 *sa = ctx->d_sa, *rank = ctx->d_isa;
  (void) scratch_mark;

  u64 *d_scalars = ar.alloc<u64>(16);
  ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(d_scalars, 0, 16 * sizeof(u64), s)));

  // ---- 1. classification
  static const bool no_dna_classify = getenv("ESA_NO_DNA_CLASSIFY") != nullptr;
  const bool dna_classify = sigma == 4 && ctx->d_pk != nullptr && !no_dna_classify;
  const u32 ktiles = (u32) div_up_u64(N, KS_TILE);
  u32 *tile_full = nullptr, *tile_side = nullptr, *tile_spec = nullptr;
  const u32 dtiles = (u32) div_up_u64(N, CD_TILE);
  u32 dna_group_shift = 0, *dna_tile_cnt = nullptr;
  unsigned long long *dna_group_cnt = nullptr;
  if (dna_classify)
  {
    // class counts per tile, per group of tiles and in total from the special mask alone (12.5 MB per 100 Mbp)
    dna_group_shift = 0;
    while (((u64) 1 << (2 * dna_group_shift)) < dtiles)
    {
      dna_group_shift++;  // groups of about sqrt(tiles) tiles
    }
    dna_tile_cnt = ar.alloc<u32>(dtiles + 1);
    dna_group_cnt = reinterpret_cast<unsigned long long *>(ar.alloc<u64>((dtiles >> dna_group_shift) + 2));
    ESA_LAUNCH(ctx, ESA_K_MEMOP, 0,
               CUDA_CHECK(cudaMemsetAsync(dna_group_cnt, 0, ((dtiles >> dna_group_shift) + 2) * sizeof(u64), s)));
    ESA_LAUNCH(ctx, ESA_K_CLASSIFY, (double) N / 8,
               (classify_dna_count_kernel<<<dtiles, CD_TILE / 32, 0, s>>>(
                   ctx->d_spm, 0, N, h0, dna_group_shift, dna_tile_cnt, dna_group_cnt,
                   reinterpret_cast<unsigned long long *>(d_scalars + 1))));
    read_back_u64(ctx, d_scalars, 4);
    ctx->h_scalars[0] = N - ctx->h_scalars[1] - ctx->h_scalars[2];
  } else
  {
    tile_full = ar.alloc<u32>(ktiles), tile_side = ar.alloc<u32>(ktiles), tile_spec = ar.alloc<u32>(ktiles);
    ESA_LAUNCH(ctx, ESA_K_CLASSIFY, (double) N,
               (classify_count_kernel<<<ktiles, KS_THREADS, 0, s>>>(ctx->d_tis, n, 0, N, kp, tile_full, tile_side,
                                                                     tile_spec)));
    ESA_LAUNCH(ctx, ESA_K_MISC, 0,
               (tile_exclusive_sum_kernel<<<3, SCAN1_THREADS, 0, s>>>(tile_full, tile_side, tile_spec, ktiles,
                                                                       d_scalars, 0)));
    read_back_u64(ctx, d_scalars, 4);
  }
  const u64 nfull = ctx->h_scalars[0], nside = ctx->h_scalars[1], nspec = ctx->h_scalars[2];
  const u64 M = nfull + nside;  // suffixes that take part in the sort
  if (M + nspec != N)
  {
    esa_throw(ESA_ERR_CUDA, "classification lost suffixes: %llu + %llu + %llu != %llu", (unsigned long long) nfull,
              (unsigned long long) nside, (unsigned long long) nspec, (unsigned long long) N);
  }

  // region X: ping-pong buffers of the initial 32-bit sort (16 bytes per sorted suffix)
  const u64 Mal = M + 64;
  // the position buffer the four-pass sort ends in IS the suffix array (no copy afterwards)
  u32 *key_a = ar.alloc<u32>(Mal), *key_b = ar.alloc<u32>(Mal), *pos_a = sa, *pos_b = ar.alloc<u32>(Mal);
  u32 *side_key = ar.alloc<u32>(nside + 1), *side_pos = ar.alloc<u32>(nside + 1), *side_dj = ar.alloc<u32>(nside + 1),
      *side_tmp = ar.alloc<u32>(nside + 1);
  // (the text-keyed first pass of the DNA sort walks tiles of text positions: N of them, plus the side tiles)
  const size_t temp_bytes = max(radix_sort_temp_bytes(N + nside + 16384, 4), radix_sort_temp_bytes(M, 8));
  void *sort_temp = ar.alloc<char>(temp_bytes);

  // DNA: the full suffixes never exist as an emitted list -- the histogram and the first pass of the sort compute key
  // and position from the packed text (radix_sort.cu, radix_pass_dna_kernel); only the rare classes are emitted
  static const bool no_text_pass = getenv("ESA_NO_TEXT_PASS") != nullptr;
  const bool text_pass = dna_classify && !no_text_pass && M > 1;
  u32 *side_key_s = nullptr, *side_pos_s = nullptr;  // text pass: the side suffixes in their final relative order
  if (dna_classify)
  {
    const TextPacked2 text{ctx->d_pk, ctx->d_spm, ctx->d_tis};
    ESA_LAUNCH(ctx, ESA_K_CLASSIFY, text_pass ? 0.125 * N + 12.0 * nside + 8.0 * nspec : 0.5 * N + 8.0 * M + 8.0 * nspec,
               (classify_dna_emit_kernel<<<dtiles, CD_THREADS, 0, s>>>(text, 0, N, h0, M, text_pass ? nullptr : key_a, pos_a,
                                                                        side_key, side_pos, side_dj, sa, rank, nullptr,
                                                                        dna_group_shift, dna_tile_cnt, dna_group_cnt)));
  } else
  {
    ESA_LAUNCH(ctx, ESA_K_CLASSIFY, (double) N + 8.0 * M + 8.0 * nspec,
               (classify_emit_kernel<<<ktiles, KS_THREADS, 0, s>>>(ctx->d_tis, n, 0, N, kp, tile_full, tile_side,
                                                                    tile_spec, M, key_a, pos_a, side_key, side_pos,
                                                                    side_dj, sa, rank, nullptr)));
  }
  // side suffixes: stable sort by h0-j, written straight behind the full ones (text pass: into their own arrays)
  u32 *side_key_dst = key_a + nfull, *side_pos_dst = pos_a + nfull;
  if (text_pass)
  {
    side_key_s = ar.alloc<u32>(nside + 1);
    side_pos_s = ar.alloc<u32>(nside + 1);
    side_key_dst = side_key_s;
    side_pos_dst = side_pos_s;
  }
  if (nside == 1)
  {
    CUDA_CHECK(cudaMemcpyAsync(side_key_dst, side_key, 4, cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(side_pos_dst, side_pos, 4, cudaMemcpyDeviceToDevice, s));
  } else if (nside > 1)
  {
    int w = radix_sort_pairs_u32(ctx, side_dj, side_tmp, side_pos, side_pos_dst, nside, 0, 8, sort_temp, temp_bytes);
    int w2 = radix_sort_pairs_u32(ctx, side_dj, side_tmp, side_key, side_key_dst, nside, 0, 8, sort_temp, temp_bytes);
    if (w != 1 || w2 != 1)
    {
      esa_throw(ESA_ERR_CUDA, "internal: side sort expected exactly one pass");
    }
  }

  // ---- bucket table, part 1: special counts per bucket
  u32 kdiv = 1;
  u32 *bck_cnt = nullptr;
  if (want_bck)
  {
    for (u32 i = prefixlength; i < h0; i++)
    {
      kdiv *= sigma;
    }
    bck_cnt = ar.alloc<u32>(ctx->numofcodes);
    ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(bck_cnt, 0, ctx->numofcodes * sizeof(u32), s)));
    if (nside > 0)
    {
      ESA_LAUNCH(ctx, ESA_K_BCK, 8.0 * nside,
                 (side_bucket_count_kernel<<<div_up_u32(nside, 256), 256, 0, s>>>(side_key, side_dj, nside, h0,
                                                                                   prefixlength, kdiv, bck_cnt)));
    }
  }

  // ---- 2. the initial sort
  u32 *skey = key_a, *spos = pos_a;
  if (text_pass)
  {
    if (radix_sort_dna_initial(ctx, ctx->d_pk, ctx->d_spm, ctx->d_tis, N, h0, nfull, side_key_s, side_pos_s, nside, key_a,
                               key_b, pos_a, pos_b, key_bits, sort_temp, temp_bytes))
    {
      skey = key_b;
      spos = pos_b;
    }
  } else if (M > 1)
  {
    if (radix_sort_pairs_u32(ctx, key_a, key_b, pos_a, pos_b, M, 0, key_bits, sort_temp, temp_bytes))
    {
      skey = key_b;
      spos = pos_b;
    }
  }
  // the sorted positions are sa[0..M)
  if (M > 0 && spos != sa)
  {
    ESA_LAUNCH(ctx, ESA_K_MEMOP, 8.0 * M, CUDA_CHECK(cudaMemcpyAsync(sa, spos, M * sizeof(u32), cudaMemcpyDeviceToDevice, s)));
  }

  // ---- 3. bck from the sorted keys
  if (want_bck)
  {
    const u64 nc = ctx->numofcodes;
    ESA_LAUNCH(ctx, ESA_K_BCK, 8.0 * nc,
               (bck_left_kernel<<<div_up_u32(nc, 256), 256, 0, s>>>(skey, M, nc, kdiv, ctx->d_bck)));
    ESA_LAUNCH(ctx, ESA_K_BCK, 20.0 * nc,
               (bck_mid_kernel<<<div_up_u32(nc, 256), 256, 0, s>>>(nc, N, nspec, bck_cnt, ctx->d_bck)));
  }

  // ---- 4. first regroup
  outcome->h0 = h0;
  outcome->rounds = 0;
  outcome->active_sum = 0;
  outcome->active_first = 0;
  outcome->num_sorted = M;
  ctx->isa_valid = true;
  if (M == 0)
  {
    ctx->sched = DepthSched{h0, 0, 0};
    return;
  }
  RefinePlan plan = make_refine_plan(ctx, bits_for(N - 1), h0);
  const u32 rank_shift = plan.rank_shift;
  const RegroupScratch rs = alloc_regroup_scratch(ar, M);
  u64 m = run_regroup_summary<u32, true>(ctx, rs, d_scalars, skey, nullptr, M, 0u, 1u);
  // region Y: fresh buffers for the first doubling input; their ping-pong partners alias region X, which is
  // dead once the first regroup has consumed the sorted 32-bit keys
  const u64 mcap = m + 64;
  u64 *dk_in = ar.alloc<u64>(mcap);
  u32 *dp_in = ar.alloc<u32>(mcap), *slots_cur = ar.alloc<u32>(mcap);
  // rank scatter of the first regroup: deferred when text rounds follow (or nothing follows): the rank list in
  // suffix-array order is kept, the N-wide scatter happens only if a doubling round needs rank[].  Otherwise
  // two-level, and its scratch costs nothing -- the rank list in suffix-array order goes into the position buffer the
  // sort did not end in, the partitioned copy into the two key buffers, which are dead once regroup_apply has read
  // the sorted keys
  plan_deferred_scatter(ctx, plan, M, m);
  u32 *part_rank = nullptr, *part_tp = nullptr, *part_tv = nullptr;
  if (!plan.pending && M >= PARTITIONED_SCATTER_MIN && rank_shift > SCATTER_WINDOW_BITS)
  {
    part_rank = spos == sa ? pos_b : ar.alloc<u32>(Mal);  // the position buffer the sort did not end in
    part_tp = key_a;
    part_tv = key_b;
  }
  run_regroup_apply<u32, true>(ctx, rs, m, skey, spos, nullptr, M, sa, rank, dk_in, dp_in, slots_cur, 0u,
                               plan.shift_of(1), nullptr, 0u, plan.pending ? plan.rank_sa : part_rank, 0u, 1u, nullptr);
  outcome->active_first = m;
  if (part_rank != nullptr)
  {
    scatter_ranks_partitioned(ctx, spos, part_rank, M, rank_shift, rank, part_tp, part_tv, sort_temp, temp_bytes);
  }
  // partners carved out of region X (key_a .. pos_b are contiguous allocations of 4*Mal words = 16*Mal bytes)
  const u64 mal = m + 64;
  u64 *dk_out = reinterpret_cast<u64 *>(key_a);          // 8*mal bytes  <= 8*Mal  (key_a + key_b)
  u32 *dp_out = ar.alloc<u32>(mal);                       // (pos_a is the suffix array itself)
  u32 *slots_next = pos_b;                                // 4*mal bytes

  // ---- 5. refinement rounds
  u64 h = h0;
  static const bool round_log = getenv("ESA_ROUND_LOG") != nullptr;  // tuning aid: one stderr line per refinement round
  while (m > 0)
  {
    const u32 round = ++outcome->rounds;
    const bool text_round = plan.is_text(round);
    const u32 cur_shift = plan.shift_of(round);
    struct timespec rl0;
    if (round_log)
    {
      CUDA_CHECK(cudaStreamSynchronize(s));
      clock_gettime(CLOCK_MONOTONIC, &rl0);
    }
    outcome->active_sum += m;
    ctx->lcp_work_estimate += m * (h / 32 + 1);
    if (text_round)
    {
      ESA_LAUNCH(ctx, ESA_K_GATHER_RANK, 20.0 * m,
                 (text_key_kernel<<<div_up_u32(m, 256), 256, 0, s>>>(dk_in, dp_in, m, h, plan.tk)));
    } else
    {
      flush_deferred_scatter(ctx, plan, sa, M, rank, sort_temp, temp_bytes);
      ESA_LAUNCH(ctx, ESA_K_GATHER_RANK, 24.0 * m,
                 (gather_rank_kernel<<<div_up_u32(m, 256), 256, 0, s>>>(dk_in, dp_in, m, h, rank)));
    }
    u64 *ks = dk_in, *ko = dk_out;
    u32 *ps = dp_in, *po = dp_out;
    if (m > 1 && doubling_sort(ctx, dk_in, dk_out, dp_in, dp_out, m, cur_shift, cur_shift + rank_shift, sort_temp,
                               temp_bytes))
    {
      ks = dk_out;
      ko = dk_in;
      ps = dp_out;
      po = dp_in;
    }
    const size_t scatter_mark = ar.mark();
    const ScatterScratch sc = plan.pending ? ScatterScratch() : round_scatter_scratch(ar, m, rank_shift);
    const u64 m_next = run_regroup<u64, false>(ctx, rs, d_scalars, ks, ps, slots_cur, m, sa, rank, ko, po,
                                               slots_next, cur_shift, plan.shift_of(round + 1), ctx->d_hround, round,
                                               sc.rank_out, 0u, plan.mask_of(round),
                                               plan.pending ? plan.rank_sa : nullptr);
    if (sc.rank_out != nullptr)
    {
      scatter_ranks_partitioned(ctx, ps, sc.rank_out, m, rank_shift, rank, sc.tp, sc.tv, sort_temp, temp_bytes);
    }
    ar.rewind(scatter_mark);
    if (round_log)
    {
      struct timespec rl1;
      CUDA_CHECK(cudaStreamSynchronize(s));
      clock_gettime(CLOCK_MONOTONIC, &rl1);
      fprintf(stderr, "ESA_ROUND_LOG round %u %s h=%llu m=%llu next=%llu spilled=%llu  %.3f ms\n", round,
              text_round ? "text" : "rank", (unsigned long long) h, (unsigned long long) m, (unsigned long long) m_next,
              (unsigned long long) ctx->h_scalars[8],
              (rl1.tv_sec - rl0.tv_sec) * 1e3 + (rl1.tv_nsec - rl0.tv_nsec) * 1e-6);
    }
    dk_in = ko;
    dp_in = po;
    dk_out = ks;
    dp_out = ps;
    u32 *t = slots_cur;
    slots_cur = slots_next;
    slots_next = t;
    m = m_next;
    h = text_round ? h + plan.tks.len : h * 2;
    if (outcome->rounds > 64)
    {
      esa_throw(ESA_ERR_CUDA, "prefix doubling did not converge");
    }
  }
  ctx->isa_valid = !plan.pending;  // a sort that ended inside the text rounds never needed rank[]: see ensure_isa
}

// the inverse suffix array for its consumers (plcp, sti1, .sti, esa_check) when the sorter did not leave it behind
void ensure_isa(esa_ctx *ctx)
{
  if (ctx->isa_valid || ctx->d_sa == nullptr || ctx->d_isa == nullptr)
  {
    return;
  }
  const u64 N = ctx->n + 1;
  Arena &ar = ctx->arena;
  const size_t mark = ar.mark();
  const u32 pos_bits = bits_for(N - 1);
  bool done = false;
  if (N >= PARTITIONED_SCATTER_MIN && pos_bits > SCATTER_WINDOW_BITS)
  {
    const size_t need = 12 * (N + 64) + radix_sort_temp_bytes(N, 4) + ((size_t) 8 << 20);
    if (ar.high - ar.top > need)
    {
      u32 *val = ar.alloc<u32>(N + 64), *tp = ar.alloc<u32>(N + 64), *tv = ar.alloc<u32>(N + 64);
      const size_t tb = radix_sort_temp_bytes(N, 4);
      void *temp = ar.alloc<char>(tb);
      ESA_LAUNCH(ctx, ESA_K_REGROUP, 4.0 * N, (iota_kernel<<<div_up_u32(N, 256), 256, 0, ctx->stream>>>(val, N)));
      done = scatter_ranks_partitioned(ctx, ctx->d_sa, val, N, pos_bits, ctx->d_isa, tp, tv, temp, tb);
    }
  }
  if (!done)
  {
    ESA_LAUNCH(ctx, ESA_K_REGROUP, 12.0 * N,
               (scatter_iota_kernel<<<div_up_u32(N, 256), 256, 0, ctx->stream>>>(ctx->d_sa, N, ctx->d_isa)));
  }
  ar.rewind(mark);
  ctx->isa_valid = true;
}

// =========================================================================================================
// Sharded build: rank g of G builds the suffix-array range that belongs to a contiguous
// range of "bins" (the top 8 bits of the initial sort key, i.e. the first symbols of the suffix).
//
//   1. every rank classifies ITS SLICE of text positions (the text itself is replicated);
//   2. splitter pass: per-bin counts, summed over the ranks, cut into G ranges of about equal size;
//   3. all-to-all: (key, position) pairs travel to the owner of their bin; suffixes with a special inside the
//      h0-prefix ("side" suffixes, few) are all-gathered so that their (h0-j, position) order can be established
//      exactly as on one GPU; special-initial suffixes are gathered to the last rank (tail of the table);
//   4. each rank sorts its range: 32-bit LSD sort, regroup, doubling rounds.  rank[] is a dense array by text
//      position of which a rank only ever fills and reads the entries of the suffixes it owns; the doubling key
//      rank[s+h] of a suffix owned elsewhere is fetched from its owner: requests are partitioned by owner (the
//      owner of s+h follows from the replicated text), exchanged, answered by a gather, and returned.
//   All per-rank work is O(N/G) except the text upload; the exchange volume is 8 bytes per suffix once plus 8 bytes per
//   active suffix and round.
#include "comm.cuh"

namespace
{

constexpr u32 SPECIAL_KEY = 0xffffffffu;

struct OwnerLut
{
  u8 owner[256];  // bin -> rank, non-decreasing
  u32 shift;      // bin = key >> shift
  u32 last;       // rank that holds the special-initial suffixes (G-1)
};

// initial sort key of the suffix at position q, recomputed from the text: the same value classify_emit_kernel
// produced for q on whichever rank classified it (padded h-mer code << 1 | special-inside flag)
__device__ __forceinline__ u32 initial_key_at(const u8 *__restrict__ tis, u64 q, const KmerParams &kp)
{
  u32 code = 0;
  for (u32 i = 0; i < kp.h; i++)
  {
    const u32 a = tis[q + i];  // tis is padded with 255 behind position n
    if (a >= ESA_WILDCARD)
    {
      if (i == 0)
      {
        return SPECIAL_KEY;
      }
      for (u32 j = i; j < kp.h; j++)
      {
        code = code * kp.sigma + (kp.sigma - 1);
      }
      return (code << 1) | 1u;
    }
    code = code * kp.sigma + a;
  }
  return code << 1;
}

// the same for a 4-letter alphabet from the 2-bit packed text: one 64-bit window and its special mask instead of h
// byte loads (the packed window holds symbol i in bits [2i, 2i+2); the code wants symbol 0 in its top digit)
__device__ __forceinline__ u32 initial_key_at_packed(const TextPacked2 &text, u64 q, u32 h)
{
  const TextPacked2::Win tw = text.win(q);  // 28 symbols >= h
  const u64 w = tw.w;
  const u32 m = tw.m;
  const u32 d = m ? (u32) (__ffs((int) m) - 1) : 32u;
  if (d == 0)
  {
    return SPECIAL_KEY;
  }
  u64 r = __brevll(w);
  r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);
  u32 code = (u32) (r >> (64 - 2 * h));
  if (d < h)
  {
    code |= (1u << (2 * (h - d))) - 1;  // padded with the largest letter behind the special
    return (code << 1) | 1u;
  }
  return code << 1;
}

struct KeyFromText
{
  const u8 *tis;
  TextPacked2 packed;  // packed.pk == nullptr: byte path
  KmerParams kp;
  __device__ __forceinline__ u32 operator()(u64 q) const
  {
    return packed.pk != nullptr ? initial_key_at_packed(packed, q, kp.h) : initial_key_at(tis, q, kp);
  }
};

__device__ __forceinline__ u32 owner_of_key(u32 key, const OwnerLut &lut)
{
  return key == SPECIAL_KEY ? lut.last : (u32) lut.owner[(key >> lut.shift) & 0xffu];
}

// sharded builds, 16-symbol DNA keys: raw padded code -> (code - base) << 1 | flag, the form the owner sorts
__global__ void rebase_keys_kernel(u32 *__restrict__ keys, u64 count, u32 base, u32 flag)
{
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count)
  {
    keys[t] = ((keys[t] - base) << 1) | flag;
  }
}

__global__ void bin_hist_kernel(const u32 *__restrict__ keys, u64 count, u32 shift, unsigned long long *__restrict__ hist)
{
  __shared__ u32 sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (u64) gridDim.x * blockDim.x)
  {
    atomicAdd(&sh[(keys[i] >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x] != 0)
  {
    atomicAdd(&hist[threadIdx.x], (unsigned long long) sh[threadIdx.x]);
  }
}

// owner of every key of a list (side suffixes), with per-owner counts
__global__ void key_owner_kernel(const u32 *__restrict__ keys, u64 count, OwnerLut lut, u32 *__restrict__ owner,
                                 unsigned long long *__restrict__ owner_count, bool raw_codes)
{
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count)
  {
    // raw_codes: side suffixes never start with a special, and 0xffffffff is an ordinary 16-symbol code
    const u32 o = raw_codes ? (u32) lut.owner[(keys[t] >> lut.shift) & 0xffu] : owner_of_key(keys[t], lut);
    owner[t] = o;
    atomicAdd(&owner_count[o], 1ull);
  }
}

// special-initial suffixes: their final slots are the tail of the table, in position order
__global__ void spec_tail_kernel(const u32 *__restrict__ spec, u64 count, u64 first_local, u32 slot_base,
                                 u32 *__restrict__ sa, u32 *__restrict__ rank)
{
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count)
  {
    const u32 p = spec[t];
    sa[first_local + t] = p;
    rank[p] = slot_base + (u32) (first_local + t);
  }
}

// doubling, step 1: tag every active suffix with the owner of its h-successor (low key bits) and count per owner
__global__ void __launch_bounds__(256)
owner_tag_kernel(u32 *__restrict__ tag, u32 *__restrict__ idx, const u32 *__restrict__ pos, u64 m, u64 h,
                 KeyFromText keyof, OwnerLut lut, unsigned long long *__restrict__ owner_count)
{
  __shared__ u32 sh[ESA_MAX_RANKS];
  if (threadIdx.x < ESA_MAX_RANKS)
  {
    sh[threadIdx.x] = 0;
  }
  __syncthreads();
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < m)
  {
    const u32 o = owner_of_key(keyof((u64) pos[t] + h), lut);
    tag[t] = o;
    idx[t] = (u32) t;
    atomicAdd(&sh[o], 1u);
  }
  __syncthreads();
  if (threadIdx.x < ESA_MAX_RANKS && sh[threadIdx.x] != 0)
  {
    atomicAdd(&owner_count[threadIdx.x], (unsigned long long) sh[threadIdx.x]);
  }
}

// doubling, step 2 (owner side): answer[j] = rank of the suffix h symbols behind the requested one
__global__ void serve_rank_kernel(const u32 *__restrict__ req_pos, u64 count, u64 h, const u32 *__restrict__ rank,
                                  u32 *__restrict__ answer)
{
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count)
  {
    answer[t] = rank[(u64) req_pos[t] + h];
  }
}

// doubling, step 1b: the requests in owner order (the key/position lists themselves keep their group order)
__global__ void make_request_kernel(const u32 *__restrict__ idx, const u32 *__restrict__ pos, u64 m,
                                    u32 *__restrict__ req)
{
  const u64 j = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m)
  {
    req[j] = pos[idx[j]];
  }
}

// doubling, step 3: the answers go into the low key bits of the suffixes that asked
__global__ void merge_answer_kernel(u64 *__restrict__ key, const u32 *__restrict__ idx,
                                    const u32 *__restrict__ answer, u64 m)
{
  const u64 j = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m)
  {
    key[idx[j]] |= (u64) answer[j];
  }
}

// doubling with peer memory: the rank of the h-successor is loaded straight from its owner's rank array over NVLink
struct PeerRanks
{
  const u32 *rank[ESA_MAX_RANKS];
};
__global__ void __launch_bounds__(256)
gather_rank_peer_kernel(u64 *__restrict__ key, const u32 *__restrict__ pos, u64 m, u64 h, KeyFromText keyof,
                        OwnerLut lut, PeerRanks peers)
{
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < m)
  {
    const u64 q = (u64) pos[t] + h;
    const u32 o = owner_of_key(keyof(q), lut);
    key[t] |= (u64) __ldcv(peers.rank[o] + q);  // never cached: the owner rewrites it between rounds
  }
}

u32 grid_for(u64 count, u32 block) { return (u32) div_up_u64(count, block); }

}  // namespace

// Host logic of the exchange step.  hist[r*stride + b] = number of full suffixes of bin b (top 8 bits of the initial key)
// that rank r classified.  Bins are dealt to the ranks in order, a bin going to the rank in whose quota (total / nranks)
// its midpoint falls, so owners never decrease and a bin is never split (equal keys stay on one rank).  An owner's
// receive buffer is laid out bin-major, then by sender: start[b] is where rank `me` must put its run of bin b.
extern "C" int esa_shard_plan(const uint64_t *hist, size_t stride, int nranks, int me, uint8_t owner[256],
                              uint64_t *send_cnt, uint64_t *recv_cnt, uint32_t start[256])
{
  if (hist == nullptr || nranks < 1 || nranks > ESA_MAX_RANKS || me < 0 || me >= nranks || stride < 256)
  {
    return ESA_ERR_ARGUMENT;
  }
  u64 global[256], total = 0, acc = 0, tail = 0;
  for (int b = 0; b < 256; b++)
  {
    global[b] = 0;
    for (int r = 0; r < nranks; r++)
    {
      global[b] += hist[(size_t) r * stride + b];
    }
    total += global[b];
  }
  // rows of at least 258 words also carry, behind the bins, the rank's suffixes with a special inside the initial key
  // ([256], spread over the ranges like everything else) and its special-initial suffixes ([257]): the latter all live
  // in the tail of the table, i.e. in the LAST range, which therefore takes fewer bins (5 % N in a human-like sequence:
  // without this the last of eight ranks held 56 M entries against 35 M on the others)
  if (stride >= 258)
  {
    for (int r = 0; r < nranks; r++)
    {
      tail += hist[(size_t) r * stride + 257];
    }
  }
  const u64 grand = total + tail;
  int cur = 0;
  for (int b = 0; b < 256; b++)
  {
    const u64 midpoint = acc + global[b] / 2;
    while (cur + 1 < nranks && grand > 0 && midpoint * (u64) nranks >= (u64) (cur + 1) * grand)
    {
      cur++;
    }
    owner[b] = (u8) cur;
    acc += global[b];
  }
  u64 run_of_owner[ESA_MAX_RANKS] = {0};
  for (int r = 0; r < nranks; r++)
  {
    send_cnt[r] = recv_cnt[r] = 0;
  }
  for (int b = 0; b < 256; b++)
  {
    const int o = owner[b];
    send_cnt[o] += hist[(size_t) me * stride + b];
    u64 before_me = 0;
    for (int r = 0; r < nranks; r++)
    {
      if (o == me)
      {
        recv_cnt[r] += hist[(size_t) r * stride + b];
      }
      if (r < me)
      {
        before_me += hist[(size_t) r * stride + b];
      }
    }
    start[b] = (u32) (run_of_owner[o] + before_me);
    run_of_owner[o] += global[b];
  }
  return ESA_OK;
}

// WIDE keys (DNA, three or more ranks): a rank's range covers only a part of the code space, so its keys are stored
// relative to the first code of the range -- (code - base) << 1 | flag -- and 16 symbols fit where the absolute key
// holds 15.  The round that follows the initial sort then starts from a quarter of the collisions (800 Mbp on eight
// ranks: 17 % of the suffixes active instead of 52 %).  Needs every range narrower than 2^31 codes, i.e. fewer than 128
// of the 256 first-symbols bins per rank; when the plan says otherwise (extremely skewed inputs) the function returns
// false before anything was exchanged and the caller repeats it with absolute 15-symbol keys.
static bool suffix_sort_sharded_impl(esa_ctx *ctx, esa_comm *comm, u32 prefixlength, bool want_bck, bool want_hround,
                                     SortOutcome *outcome, bool wide)
{
  const int G = comm->nranks, me = comm->rank;
  const u64 n = ctx->n, N = n + 1;
  const u32 sigma = ctx->sigma;
  cudaStream_t s = ctx->stream;
  Arena &ar = ctx->arena;
  const size_t entry_mark = ar.mark();
  const u32 h0_text = choose_h0(sigma);                       // symbols of the absolute key (owner look-ups from the text)
  const KmerParams kp_text = make_kmer_params(sigma, h0_text);
  const u32 key_bits_text = bits_for(((u64) kp_text.all_pad << 1) | 1);
  const u32 h0 = wide ? 16u : h0_text;
  const KmerParams kp = make_kmer_params(sigma, h0);
  const u32 key_shl = wide ? 0u : 1u;                          // classification emits raw 16-symbol codes when wide
  const u32 key_bits = wide ? 32u : key_bits_text;             // of the keys as classified
  u32 *rank = ctx->d_isa;  // dense by text position; only the entries of suffixes this rank owns are defined
  u64 all[ESA_MAX_RANKS * (ESA_MAX_RANKS + 4)];
  u64 mine[ESA_MAX_RANKS + 4];

  u64 *d_scalars = ar.alloc<u64>(16);
  unsigned long long *d_counts = (unsigned long long *) ar.alloc<u64>(256);
  ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(d_scalars, 0, 16 * sizeof(u64), s)));
  ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(d_counts, 0, 256 * sizeof(u64), s)));

  // ---- 1. classification of my slice of text positions (slices are KS_TILE aligned for the 16-byte loads)
  const u64 slice = div_up_u64(div_up_u64(N, (u64) G), KS_TILE) * KS_TILE;
  const u64 p_begin = min(N, (u64) me * slice), p_end = min(N, (u64) (me + 1) * slice);
  const u32 ktiles = (u32) div_up_u64(p_end - p_begin, KS_TILE);
  // DNA over peer memory: the full suffixes are never emitted -- the splitter histogram and the partition pass take key
  // and position from the packed text (radix_sort.cu: radix_bin_hist_dna, radix_partition_remote_dna), as on one GPU
  static const bool no_text_pass = getenv("ESA_NO_TEXT_PASS") != nullptr;
  const bool text_remote = sigma == 4 && ctx->d_pk != nullptr && comm->direct && !no_text_pass;
  u32 *tile_full = nullptr, *tile_side = nullptr, *tile_spec = nullptr;
  u32 *key_l = nullptr, *pos_l = nullptr, *key_t = nullptr, *pos_t = nullptr;
  u32 *lside_key = nullptr, *lside_pos = nullptr, *lside_dj = nullptr, *lspec = nullptr;
  u64 nfull = 0, nside = 0, nspec = 0;
  if (text_remote)
  {
    static_assert(CD_TILE == KS_TILE, "slices are aligned to both tile sizes");
    const u32 dtiles = (u32) div_up_u64(p_end - p_begin, CD_TILE);
    u32 gshift = 0;
    while (((u64) 1 << (2 * gshift)) < dtiles)
    {
      gshift++;
    }
    u32 *tile_cnt = ar.alloc<u32>(dtiles + 1);
    unsigned long long *group_cnt = reinterpret_cast<unsigned long long *>(ar.alloc<u64>((dtiles >> gshift) + 2));
    if (dtiles > 0)
    {
      ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(group_cnt, 0, ((dtiles >> gshift) + 2) * sizeof(u64), s)));
      ESA_LAUNCH(ctx, ESA_K_CLASSIFY, (double) (p_end - p_begin) / 8,
                 (classify_dna_count_kernel<<<dtiles, CD_TILE / 32, 0, s>>>(ctx->d_spm, p_begin, p_end, h0, gshift, tile_cnt,
                                                                            group_cnt,
                                                                            reinterpret_cast<unsigned long long *>(d_scalars + 1))));
      read_back_u64(ctx, d_scalars, 4);
      nside = ctx->h_scalars[1], nspec = ctx->h_scalars[2];
      nfull = (p_end - p_begin) - nside - nspec;
    }
    lside_key = ar.alloc<u32>(nside + 1), lside_pos = ar.alloc<u32>(nside + 1), lside_dj = ar.alloc<u32>(nside + 1);
    lspec = ar.alloc<u32>(nspec + 1);
    if (dtiles > 0)
    {
      const TextPacked2 text{ctx->d_pk, ctx->d_spm, ctx->d_tis};
      ESA_LAUNCH(ctx, ESA_K_CLASSIFY, 0.125 * (double) (p_end - p_begin) + 12.0 * nside + 4.0 * nspec,
                 (classify_dna_emit_kernel<<<dtiles, CD_THREADS, 0, s>>>(text, p_begin, p_end, h0, 0, nullptr, nullptr, lside_key,
                                                                          lside_pos, lside_dj, nullptr, nullptr, lspec, gshift,
                                                                          tile_cnt, group_cnt, key_shl)));
    }
  } else
  {
    tile_full = ar.alloc<u32>(ktiles + 1), tile_side = ar.alloc<u32>(ktiles + 1), tile_spec = ar.alloc<u32>(ktiles + 1);
    if (ktiles > 0)
    {
      ESA_LAUNCH(ctx, ESA_K_CLASSIFY, (double) (p_end - p_begin),
                 (classify_count_kernel<<<ktiles, KS_THREADS, 0, s>>>(ctx->d_tis, n, p_begin, p_end, kp, tile_full,
                                                                       tile_side, tile_spec)));
      ESA_LAUNCH(ctx, ESA_K_MISC, 0,
                 (tile_exclusive_sum_kernel<<<3, SCAN1_THREADS, 0, s>>>(tile_full, tile_side, tile_spec, ktiles,
                                                                         d_scalars, 0)));
      read_back_u64(ctx, d_scalars, 4);
      nfull = ctx->h_scalars[0], nside = ctx->h_scalars[1], nspec = ctx->h_scalars[2];
    }
    if (nfull + nside + nspec != p_end - p_begin)
    {
      esa_throw(ESA_ERR_CUDA, "classification lost suffixes on rank %d", me);
    }
    key_l = ar.alloc<u32>(nfull + 64), pos_l = ar.alloc<u32>(nfull + 64);
    key_t = ar.alloc<u32>(nfull + 64), pos_t = ar.alloc<u32>(nfull + 64);
    lside_key = ar.alloc<u32>(nside + 1), lside_pos = ar.alloc<u32>(nside + 1), lside_dj = ar.alloc<u32>(nside + 1);
    lspec = ar.alloc<u32>(nspec + 1);
    if (ktiles > 0)
    {
      ESA_LAUNCH(ctx, ESA_K_CLASSIFY, (double) (p_end - p_begin) + 8.0 * (nfull + nside) + 4.0 * nspec,
                 (classify_emit_kernel<<<ktiles, KS_THREADS, 0, s>>>(ctx->d_tis, n, p_begin, p_end, kp, tile_full,
                                                                      tile_side, tile_spec, 0, key_l, pos_l,
                                                                      lside_key, lside_pos, lside_dj, nullptr, nullptr,
                                                                      lspec, key_shl)));
    }
  }

  // ---- 2. splitter pass: bins of the top 8 key bits (= the first symbols, whichever key width)
  OwnerLut lut;
  lut.shift = key_bits > 8 ? key_bits - 8 : 0;
  lut.last = (u32) (G - 1);
  u64 local_hist[256 + 2];
  if (text_remote)
  {
    u32 *d_hist32 = reinterpret_cast<u32 *>(d_counts);  // (zeroed above)
    u32 h32[256];
    radix_bin_hist_dna(ctx, ctx->d_pk, ctx->d_spm, ctx->d_tis, p_begin, p_end, h0, key_shl, lut.shift, d_hist32);
    CUDA_CHECK(cudaMemcpyAsync(h32, d_hist32, sizeof h32, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    for (int b = 0; b < 256; b++)
    {
      local_hist[b] = h32[b];
    }
  } else
  {
    if (nfull > 0)
    {
      const u32 grid = (u32) min((u64) ctx->num_sms * 8, div_up_u64(nfull, 256));
      ESA_LAUNCH(ctx, ESA_K_EXCHANGE, 4.0 * nfull,
                 (bin_hist_kernel<<<grid, 256, 0, s>>>(key_l, nfull, lut.shift, d_counts)));
    }
    CUDA_CHECK(cudaMemcpyAsync(local_hist, d_counts, 256 * sizeof(u64), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
  local_hist[256] = nside;
  local_hist[257] = nspec;
  // every rank learns every rank's bin counts: the splitters, all send/receive counts and, for the fused
  // partition + all-to-all, the exact place of every (sender, bin) run in its owner's buffer follow from this matrix
  constexpr size_t HROW = 258;
  std::vector<u64> hist_all((size_t) G * HROW);
  comm_exchange_host_u64(ctx, comm, local_hist, HROW, hist_all.data());
  // splitters, counts per (sender, destination) and the place of my runs in the owners' buffers (host logic,
  // exported as esa_shard_plan for the CPU tests)
  u64 global_hist[256];
  for (int b = 0; b < 256; b++)
  {
    global_hist[b] = 0;
    for (int r = 0; r < G; r++)
    {
      global_hist[b] += hist_all[(size_t) r * HROW + b];
    }
  }
  u64 send_cnt[ESA_MAX_RANKS] = {0}, send_off[ESA_MAX_RANKS], recv_cnt[ESA_MAX_RANKS] = {0}, recv_off[ESA_MAX_RANKS];
  u32 h_start[256];
  esa_shard_plan(hist_all.data(), HROW, G, me, lut.owner, send_cnt, recv_cnt, h_start);
  // first code of every rank's range and its width in bins (wide keys)
  u32 key_base[ESA_MAX_RANKS] = {0}, range_bins[ESA_MAX_RANKS] = {0};
  {
    bool seen[ESA_MAX_RANKS] = {false};
    for (int b = 0; b < 256; b++)
    {
      const int o = lut.owner[b];
      if (!seen[o])
      {
        seen[o] = true;
        key_base[o] = wide ? (u32) b << 24 : 0u;
      }
      range_bins[o]++;
    }
    if (wide)
    {
      for (int r = 0; r < G; r++)
      {
        if (range_bins[r] >= 128)
        {
          ar.rewind(entry_mark);
          return false;  // the same decision on every rank: the plan is a function of the exchanged histogram
        }
      }
    }
  }
  const u32 sort_key_bits = wide ? bits_for(((((u64) range_bins[me]) << 24) << 1) | 1) : key_bits;
  OwnerLut lut_text = lut;  // owner of a suffix from the TEXT (doubling rounds): absolute keys of h0_text symbols
  lut_text.shift = key_bits_text > 8 ? key_bits_text - 8 : 0;
  u64 side_cnt[ESA_MAX_RANKS], side_off[ESA_MAX_RANKS], spec_cnt[ESA_MAX_RANKS], spec_off[ESA_MAX_RANKS];
  u64 nside_tot = 0, nspec_tot = 0, mfull = 0, off = 0;
  for (int r = 0; r < G; r++)
  {
    send_off[r] = off;
    off += send_cnt[r];
    recv_off[r] = mfull;
    mfull += recv_cnt[r];
    side_cnt[r] = hist_all[(size_t) r * HROW + 256];
    side_off[r] = nside_tot;
    nside_tot += side_cnt[r];
    spec_cnt[r] = hist_all[(size_t) r * HROW + 257];
    spec_off[r] = nspec_tot;
    nspec_tot += spec_cnt[r];
  }

  // ---- 3b. side suffixes: all-gather (rank order = position order), stable sort by h0-j, keep mine
  const size_t temp_bytes = max(max(radix_sort_temp_bytes(mfull + nside_tot + 64, 4), radix_sort_temp_bytes(mfull + nside_tot + 64, 8)),
                                radix_sort_temp_bytes((p_end - p_begin) + 16384, 4));
  void *sort_temp = ar.alloc<char>(temp_bytes);
  u32 *gs_key = ar.alloc<u32>(nside_tot + 1), *gs_pos = ar.alloc<u32>(nside_tot + 1), *gs_dj = ar.alloc<u32>(nside_tot + 1);
  u32 *gs_tmp = ar.alloc<u32>(nside_tot + 1), *gs_key2 = ar.alloc<u32>(nside_tot + 1), *gs_pos2 = ar.alloc<u32>(nside_tot + 1);
  u32 *gs_own = ar.alloc<u32>(nside_tot + 1);
  u64 nside_mine = 0, side_first = 0;
  u32 *my_side_key = gs_key, *my_side_pos = gs_pos;
  u32 kdiv = 1;
  u32 *bck_cnt = nullptr;
  if (want_bck)
  {
    for (u32 i = prefixlength; i < h0; i++)
    {
      kdiv *= sigma;
    }
    bck_cnt = ar.alloc<u32>(ctx->numofcodes);
    ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(bck_cnt, 0, ctx->numofcodes * sizeof(u32), s)));
  }
  if (nside_tot > 0)
  {
    comm_allgatherv_u32(ctx, comm, lside_key, gs_key, side_off, side_cnt);
    comm_allgatherv_u32(ctx, comm, lside_pos, gs_pos, side_off, side_cnt);
    comm_allgatherv_u32(ctx, comm, lside_dj, gs_dj, side_off, side_cnt);
    if (want_bck)
    {
      ESA_LAUNCH(ctx, ESA_K_BCK, 8.0 * nside_tot,
                 (side_bucket_count_kernel<<<grid_for(nside_tot, 256), 256, 0, s>>>(gs_key, gs_dj, nside_tot, h0,
                                                                                     prefixlength, kdiv, bck_cnt, key_shl)));
    }
    const u32 *sk = gs_key, *sp = gs_pos;
    if (nside_tot > 1)
    {
      const int w = radix_sort_pairs_u32(ctx, gs_dj, gs_tmp, gs_pos, gs_pos2, nside_tot, 0, 8, sort_temp, temp_bytes);
      const int w2 = radix_sort_pairs_u32(ctx, gs_dj, gs_tmp, gs_key, gs_key2, nside_tot, 0, 8, sort_temp, temp_bytes);
      if (w != 1 || w2 != 1)
      {
        esa_throw(ESA_ERR_CUDA, "internal: side sort expected exactly one pass");
      }
      sk = gs_key2;
      sp = gs_pos2;
    }
    // owner of every side suffix; then a stable partition by owner
    ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(d_counts, 0, ESA_MAX_RANKS * sizeof(u64), s)));
    ESA_LAUNCH(ctx, ESA_K_EXCHANGE, 8.0 * nside_tot,
               (key_owner_kernel<<<grid_for(nside_tot, 256), 256, 0, s>>>(sk, nside_tot, lut, gs_own, d_counts, wide)));
    u64 own_cnt[ESA_MAX_RANKS];
    CUDA_CHECK(cudaMemcpyAsync(own_cnt, d_counts, sizeof own_cnt, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    for (int r = 0; r < me; r++)
    {
      side_first += own_cnt[r];
    }
    nside_mine = own_cnt[me];
    if (nside_tot > 1)
    {
      // sk/sp live in gs_key2/gs_pos2 -> partitioned copies go back into gs_key/gs_pos
      const int w = radix_sort_pairs_u32(ctx, gs_own, gs_tmp, const_cast<u32 *>(sp), gs_pos, nside_tot, 0, 8, sort_temp, temp_bytes);
      const int w2 = radix_sort_pairs_u32(ctx, gs_own, gs_tmp, const_cast<u32 *>(sk), gs_key, nside_tot, 0, 8, sort_temp, temp_bytes);
      if (w != 1 || w2 != 1)
      {
        esa_throw(ESA_ERR_CUDA, "internal: side partition expected exactly one pass");
      }
    }
    my_side_key = gs_key + side_first;
    my_side_pos = gs_pos + side_first;
  }

  // ---- 3c. all-to-all of the full suffixes: stable partition by bin, then one send per destination
  const u64 M = mfull + nside_mine;  // suffixes of my range that take part in the sort
  const u64 Mal = M + 64;
  u32 *key_a = ar.alloc<u32>(Mal), *key_b = ar.alloc<u32>(Mal), *pos_a = ar.alloc<u32>(Mal), *pos_b = ar.alloc<u32>(Mal);
  if (comm->direct)
  {
    // fused partition + all-to-all: one pass over my (key, position) pairs whose output runs land directly in the
    // owners' key_a / pos_a over NVLink.  Layout inside an owner: by sender, each sender's pairs in its own order
    // (= text position order); the owner's LSD sort does not care.
    u64 my_off[2] = {(u64) ((char *) key_a - ar.base), (u64) ((char *) pos_a - ar.base)}, offs[2 * ESA_MAX_RANKS];
    comm_exchange_host_u64(ctx, comm, my_off, 2, offs);  // also: every rank has carved its receive buffers
    RadixRemoteDst h_dst;
    memset(&h_dst, 0, sizeof h_dst);
    for (int r = 0; r < G; r++)
    {
      h_dst.keys[r] = comm_peer_ptr<u32>(ctx, comm, r, offs[2 * r]);
      h_dst.vals[r] = comm_peer_ptr<u32>(ctx, comm, r, offs[2 * r + 1]);
    }
    memcpy(h_dst.owner, lut.owner, 256);
    for (int r = 0; r < G; r++)
    {
      h_dst.key_sub[r] = key_base[r];  // wide keys: stored relative to the owner's first code, flag bit 0
    }
    h_dst.key_shl = wide ? 1u : 0u;
    // the pass partitions by owner: rank r's runs lie in sender order inside its buffers
    u32 owner_start[256] = {0};
    for (int o = 0; o < G; o++)
    {
      u64 before_me = 0;
      for (int r = 0; r < me; r++)
      {
        for (int b = 0; b < 256; b++)
        {
          before_me += lut.owner[b] == o ? hist_all[(size_t) r * HROW + b] : 0;
        }
      }
      owner_start[o] = (u32) before_me;
    }
    RadixRemoteDst *d_dst = reinterpret_cast<RadixRemoteDst *>(ar.alloc<char>(sizeof(RadixRemoteDst)));
    u32 *d_start = ar.alloc<u32>(256);
    CUDA_CHECK(cudaMemcpyAsync(d_dst, &h_dst, sizeof h_dst, cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(d_start, owner_start, sizeof owner_start, cudaMemcpyHostToDevice, s));
    if (text_remote)
    {
      radix_partition_remote_dna(ctx, ctx->d_pk, ctx->d_spm, ctx->d_tis, p_begin, p_end, h0, key_shl, nfull, lut.shift, d_dst,
                                 d_start, sort_temp, temp_bytes);
    } else
    {
      radix_partition_remote_u32(ctx, key_l, pos_l, nfull, lut.shift, key_bits > lut.shift ? key_bits : lut.shift + 1,
                                 d_dst, d_start, sort_temp, temp_bytes);
    }
    comm->bytes_sent += 8 * (nfull - send_cnt[me]);
    CUDA_CHECK(cudaStreamSynchronize(s));  // my stores are done; the exchange below tells me when everybody's are
  } else
  {
    const u32 *ks = key_l, *ps = pos_l;
    if (nfull > 1 && radix_sort_pairs_u32(ctx, key_l, key_t, pos_l, pos_t, nfull, lut.shift, key_bits, sort_temp, temp_bytes))
    {
      ks = key_t;
      ps = pos_t;
    }
    comm_alltoallv_u32(ctx, comm, ks, send_off, send_cnt, key_a, recv_off, recv_cnt);
    comm_alltoallv_u32(ctx, comm, ps, send_off, send_cnt, pos_a, recv_off, recv_cnt);
    if (wide && mfull > 0)
    {
      ESA_LAUNCH(ctx, ESA_K_EXCHANGE, 8.0 * mfull,
                 (rebase_keys_kernel<<<grid_for(mfull, 256), 256, 0, s>>>(key_a, mfull, key_base[me], 0u)));
    }
  }
  if (nside_mine > 0)
  {
    CUDA_CHECK(cudaMemcpyAsync(key_a + mfull, my_side_key, nside_mine * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(pos_a + mfull, my_side_pos, nside_mine * 4, cudaMemcpyDeviceToDevice, s));
    if (wide)
    {
      ESA_LAUNCH(ctx, ESA_K_EXCHANGE, 8.0 * nside_mine,
                 (rebase_keys_kernel<<<grid_for(nside_mine, 256), 256, 0, s>>>(key_a + mfull, nside_mine, key_base[me], 1u)));
    }
  }

  // ---- 3d. shape of the table: my range, the tail of special-initial suffixes on the last rank
  mine[0] = M;
  comm_exchange_host_u64(ctx, comm, mine, 1, all);
  u64 slot_base64 = 0, Mtot = 0;
  for (int r = 0; r < G; r++)
  {
    if (r < me)
    {
      slot_base64 += all[r];
    }
    Mtot += all[r];
  }
  if (Mtot + nspec_tot != N)
  {
    esa_throw(ESA_ERR_COMM, "exchange lost suffixes: %llu + %llu != %llu", (unsigned long long) Mtot,
              (unsigned long long) nspec_tot, (unsigned long long) N);
  }
  const u32 slot_base = (u32) slot_base64;
  const u64 shard_count = M + (me == G - 1 ? nspec_tot : 0);
  ctx->shard_first = slot_base64;
  ctx->shard_count = shard_count;
  u32 *sa = ctx->d_sa = ar.alloc_high<u32>(shard_count + 64);
  if (want_hround)
  {
    ctx->d_hround = ar.alloc_high<u8>(shard_count + 64);
    ESA_LAUNCH(ctx, ESA_K_MEMOP, (double) shard_count, CUDA_CHECK(cudaMemsetAsync(ctx->d_hround, 0, shard_count + 64, s)));
  }
  if (nspec_tot > 0)
  {
    u32 *gspec = me == G - 1 ? ar.alloc<u32>(nspec_tot + 1) : nullptr;
    comm_gatherv_u32(ctx, comm, G - 1, lspec, gspec, spec_off, spec_cnt);
    if (me == G - 1)
    {
      ESA_LAUNCH(ctx, ESA_K_EXCHANGE, 12.0 * nspec_tot,
                 (spec_tail_kernel<<<grid_for(nspec_tot, 256), 256, 0, s>>>(gspec, nspec_tot, M, slot_base, sa, rank)));
    }
  }

  // ---- 4. the initial sort of my range
  u32 *skey = key_a, *spos = pos_a;
  if (M > 1 && radix_sort_pairs_u32(ctx, key_a, key_b, pos_a, pos_b, M, 0, sort_key_bits, sort_temp, temp_bytes))
  {
    skey = key_b;
    spos = pos_b;
  }
  if (M > 0)
  {
    ESA_LAUNCH(ctx, ESA_K_MEMOP, 8.0 * M, CUDA_CHECK(cudaMemcpyAsync(sa, spos, M * sizeof(u32), cudaMemcpyDeviceToDevice, s)));
  }

  // ---- bck: lower bounds inside my range, summed over the ranks
  if (want_bck)
  {
    const u64 nc = ctx->numofcodes;
    ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(ctx->d_bck, 0, 2 * nc * sizeof(u64), s)));
    ESA_LAUNCH(ctx, ESA_K_BCK, 8.0 * nc,
               (bck_left_kernel<<<grid_for(nc, 256), 256, 0, s>>>(skey, M, nc, kdiv, ctx->d_bck, (u64) key_base[me])));
    const size_t mark = ar.mark();
    u64 *tmp = ar.alloc<u64>(2 * nc);
    comm_allreduce_u64(ctx, comm, ctx->d_bck, tmp, 2 * nc, false);
    CUDA_CHECK(cudaStreamSynchronize(s));
    ar.rewind(mark);
    ESA_LAUNCH(ctx, ESA_K_BCK, 20.0 * nc,
               (bck_mid_kernel<<<grid_for(nc, 256), 256, 0, s>>>(nc, N, nspec_tot, bck_cnt, ctx->d_bck)));
  }

  // ---- first regroup
  outcome->h0 = h0;
  outcome->rounds = 0;
  outcome->active_sum = 0;
  outcome->active_first = 0;
  outcome->num_sorted = M;
  ctx->isa_valid = false;  // rank[] of a sharded build only ever holds the entries of the suffixes a rank owns
  RefinePlan plan = make_refine_plan(ctx, max(bits_for(N - 1), 8u), h0);
  const u32 rank_shift = plan.rank_shift;
  const u32 owner_bits = bits_for((u64) G - 1);
  const RegroupScratch rs = alloc_regroup_scratch(ar, M);
  u64 m = 0;
  if (M > 0)
  {
    m = run_regroup_summary<u32, true>(ctx, rs, d_scalars, skey, nullptr, M, slot_base, 1u);
  }
  const u64 mcap = m + 64;
  u64 *dk_in = ar.alloc<u64>(mcap);
  u32 *dp_in = ar.alloc<u32>(mcap), *slots_cur = ar.alloc<u32>(mcap);
  // the scatter of the first regroup is deferred like on one GPU; whether a doubling round follows the text rounds
  // is a property of all ranks together (a rank without active suffixes still owns ranks others will ask for)
  if (plan.text_rounds > 0)
  {
    plan_deferred_scatter(ctx, plan, M, m);
  }
  u32 *part_rank = nullptr, *part_tp = nullptr, *part_tv = nullptr;
  if (!plan.pending && M >= PARTITIONED_SCATTER_MIN && rank_shift > SCATTER_WINDOW_BITS)
  {
    part_rank = spos == pos_a ? pos_b : pos_a;  // scratch of the first scatter lives in the dead sort buffers
    part_tp = key_a;
    part_tv = key_b;
  }
  if (M > 0)
  {
    run_regroup_apply<u32, true>(ctx, rs, m, skey, spos, nullptr, M, sa, rank, dk_in, dp_in, slots_cur, 0u,
                                 plan.shift_of(1), nullptr, 0u, plan.pending ? plan.rank_sa : part_rank, slot_base, 1u,
                                 nullptr);
    if (part_rank != nullptr)
    {
      scatter_ranks_partitioned(ctx, spos, part_rank, M, rank_shift, rank, part_tp, part_tv, sort_temp, temp_bytes);
    }
  }
  outcome->active_first = m;
  // ping-pong partners of the doubling buffers alias the 32-bit sort region (dead after the first regroup)
  u64 *dk_out = reinterpret_cast<u64 *>(key_a);
  u32 *dp_out = pos_a;
  u32 *slots_next = pos_b;

  // ---- 5. refinement rounds, all ranks in lock step until no rank has an active suffix left
  KeyFromText keyof;
  keyof.tis = ctx->d_tis;
  keyof.packed = TextPacked2{sigma == 4 ? ctx->d_pk : nullptr, ctx->d_spm, ctx->d_tis};
  keyof.kp = kp_text;
  // the sort of one round on this rank (the secondary keys are in place): afterwards (dk_in, dp_in) is sorted
  auto sort_round = [&](u32 round) {
    const u32 cur_shift = plan.shift_of(round);
    if (m > 1 && doubling_sort(ctx, dk_in, dk_out, dp_in, dp_out, m, cur_shift, cur_shift + rank_shift, sort_temp,
                               temp_bytes))
    {
      u64 *tk = dk_in;
      dk_in = dk_out;
      dk_out = tk;
      u32 *tp = dp_in;
      dp_in = dp_out;
      dp_out = tp;
    }
  };
  // new ranks, suffix array entries and the next active list of one round on this rank; returns its size
  auto regroup_round = [&](u32 round) -> u64 {
    const size_t scatter_mark = ar.mark();
    const ScatterScratch sc = plan.pending ? ScatterScratch() : round_scatter_scratch(ar, m, rank_shift);
    const u64 m_next = run_regroup<u64, false>(ctx, rs, d_scalars, dk_in, dp_in, slots_cur, m, sa, rank,
                                               dk_out, dp_out, slots_next, plan.shift_of(round),
                                               plan.shift_of(round + 1), ctx->d_hround, round, sc.rank_out, slot_base,
                                               plan.mask_of(round), plan.pending ? plan.rank_sa : nullptr);
    if (sc.rank_out != nullptr)
    {
      scatter_ranks_partitioned(ctx, dp_in, sc.rank_out, m, rank_shift, rank, sc.tp, sc.tv, sort_temp, temp_bytes);
    }
    ar.rewind(scatter_mark);
    u64 *tk = dk_in;
    dk_in = dk_out;
    dk_out = tk;
    u32 *tp = dp_in;
    dp_in = dp_out;
    dp_out = tp;
    u32 *t = slots_cur;
    slots_cur = slots_next;
    slots_next = t;
    return m_next;
  };
  // text rounds: the secondary keys come from the replicated text -- no communication but the active counts
  u64 h = h0;
  u64 m_world = 0;
  {
    u64 world[ESA_MAX_RANKS];
    comm_exchange_host_u64(ctx, comm, &m, 1, world);
    for (int r = 0; r < G; r++)
    {
      m_world += world[r];
    }
  }
  while (m_world > 0 && plan.is_text(outcome->rounds + 1))
  {
    const u32 round = ++outcome->rounds;
    outcome->active_sum += m;
    ctx->lcp_work_estimate += m * (h / 32 + 1);
    u64 m_next = 0;
    if (m > 0)
    {
      ESA_LAUNCH(ctx, ESA_K_GATHER_RANK, 20.0 * m,
                 (text_key_kernel<<<grid_for(m, 256), 256, 0, s>>>(dk_in, dp_in, m, h, plan.tk)));
      sort_round(round);
      m_next = regroup_round(round);
    }
    u64 world[ESA_MAX_RANKS];
    comm_exchange_host_u64(ctx, comm, &m_next, 1, world);
    m_world = 0;
    for (int r = 0; r < G; r++)
    {
      m_world += world[r];
    }
    m = m_next;
    h += plan.tks.len;
  }
  if (m_world == 0)
  {
    return true;
  }
  // a doubling round follows: rank[] is needed now
  flush_deferred_scatter(ctx, plan, sa, M, rank, sort_temp, temp_bytes);
  if (comm->direct && comm->direct_loads)
  {
    // peer-memory variant: two barriers per round keep the remote loads of rank[] apart from the owners' updates
    u64 my_off = (u64) ((char *) rank - ar.base), offs[ESA_MAX_RANKS];
    CUDA_CHECK(cudaStreamSynchronize(s));  // my ranks are written
    u64 pair[2] = {my_off, m}, pairs[2 * ESA_MAX_RANKS];
    comm_exchange_host_u64(ctx, comm, pair, 2, pairs);
    PeerRanks peers;
    m_world = 0;
    for (int r = 0; r < G; r++)
    {
      offs[r] = pairs[2 * r];
      peers.rank[r] = comm_peer_ptr<const u32>(ctx, comm, r, offs[r]);
      m_world += pairs[2 * r + 1];
    }
    while (m_world > 0)
    {
      const u32 round = ++outcome->rounds;
      outcome->active_sum += m;
      ctx->lcp_work_estimate += m * (h / 32 + 1);
      if (outcome->rounds > 64)
      {
        esa_throw(ESA_ERR_CUDA, "prefix doubling did not converge");
      }
      u64 m_next = 0;
      if (m > 0)
      {
        ESA_LAUNCH(ctx, ESA_K_GATHER_RANK, 24.0 * m,
                   (gather_rank_peer_kernel<<<grid_for(m, 256), 256, 0, s>>>(dk_in, dp_in, m, h, keyof, lut_text, peers)));
        comm->bytes_sent += 4 * m;  // upper bound of what crossed NVLink for me (own range included)
        sort_round(round);
      }
      // barrier 1: every rank has finished reading rank[] (its gather precedes this point on its stream)
      CUDA_CHECK(cudaStreamSynchronize(s));
      {
        u64 zero = 0, dummy[ESA_MAX_RANKS];
        comm_exchange_host_u64(ctx, comm, &zero, 1, dummy);
      }
      if (m > 0)
      {
        m_next = regroup_round(round);  // updates rank[]; synchronises the stream (read-back of the active count)
      }
      // barrier 2 (carries the active counts): every rank has finished updating rank[] before anybody gathers again
      u64 world_next[ESA_MAX_RANKS], sum_next = 0;
      CUDA_CHECK(cudaStreamSynchronize(s));
      comm_exchange_host_u64(ctx, comm, &m_next, 1, world_next);
      for (int r = 0; r < G; r++)
      {
        sum_next += world_next[r];
      }
      m = m_next;
      m_world = sum_next;
      h *= 2;
    }
    return true;
  }
  while (true)
  {
    u64 req_cnt[ESA_MAX_RANKS] = {0}, req_off[ESA_MAX_RANKS], in_cnt[ESA_MAX_RANKS], in_off[ESA_MAX_RANKS];
    const size_t round_mark = ar.mark();
    // requests travel in owner order; the key / position lists themselves stay in group order (doubling_sort relies
    // on it), so the partition works on (owner tag, list index) pairs
    u32 *tag_a = ar.alloc<u32>(m + 4), *tag_b = ar.alloc<u32>(m + 4), *idx_a = ar.alloc<u32>(m + 4),
        *idx_b = ar.alloc<u32>(m + 4), *req_out = ar.alloc<u32>(m + 4);
    if (m > 0)
    {
      ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(d_counts, 0, ESA_MAX_RANKS * sizeof(u64), s)));
      ESA_LAUNCH(ctx, ESA_K_EXCHANGE, 16.0 * m,
                 (owner_tag_kernel<<<grid_for(m, 256), 256, 0, s>>>(tag_a, idx_a, dp_in, m, h, keyof, lut_text, d_counts)));
      CUDA_CHECK(cudaMemcpyAsync(req_cnt, d_counts, G * sizeof(u64), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
    }
    for (int r = 0; r < G; r++)
    {
      mine[r] = req_cnt[r];
    }
    mine[G] = m;
    comm_exchange_host_u64(ctx, comm, mine, (size_t) G + 1, all);
    u64 in_tot = 0, o2 = 0;
    m_world = 0;
    for (int r = 0; r < G; r++)
    {
      m_world += all[(size_t) r * (G + 1) + G];
      in_cnt[r] = all[(size_t) r * (G + 1) + me];
      in_off[r] = in_tot;
      in_tot += in_cnt[r];
      req_off[r] = o2;
      o2 += req_cnt[r];
    }
    if (m_world == 0)
    {
      break;
    }
    const u32 round = ++outcome->rounds;
    outcome->active_sum += m;
    ctx->lcp_work_estimate += m * (h / 32 + 1);
    if (outcome->rounds > 64)
    {
      esa_throw(ESA_ERR_CUDA, "prefix doubling did not converge");
    }
    // partition my requests by owner (stable, one pass over the tag bits)
    const u32 *idx = idx_a;
    if (m > 1 && radix_sort_pairs_u32(ctx, tag_a, tag_b, idx_a, idx_b, m, 0, owner_bits, sort_temp, temp_bytes))
    {
      idx = idx_b;
    }
    if (m > 0)
    {
      ESA_LAUNCH(ctx, ESA_K_EXCHANGE, 12.0 * m,
                 (make_request_kernel<<<grid_for(m, 256), 256, 0, s>>>(idx, dp_in, m, req_out)));
    }
    u32 *req_in = ar.alloc<u32>(in_tot + 1), *ans_out = ar.alloc<u32>(in_tot + 1), *ans_in = ar.alloc<u32>(m + 1);
    comm_alltoallv_u32(ctx, comm, req_out, req_off, req_cnt, req_in, in_off, in_cnt);
    if (in_tot > 0)
    {
      ESA_LAUNCH(ctx, ESA_K_GATHER_RANK, 12.0 * in_tot,
                 (serve_rank_kernel<<<grid_for(in_tot, 256), 256, 0, s>>>(req_in, in_tot, h, rank, ans_out)));
    }
    comm_alltoallv_u32(ctx, comm, ans_out, in_off, in_cnt, ans_in, req_off, req_cnt);
    u64 m_next = 0;
    if (m > 0)
    {
      ESA_LAUNCH(ctx, ESA_K_EXCHANGE, 20.0 * m,
                 (merge_answer_kernel<<<grid_for(m, 256), 256, 0, s>>>(dk_in, idx, ans_in, m)));
      ar.rewind(round_mark);  // the round's exchange buffers are dead (later users are stream ordered)
      sort_round(round);
      m_next = regroup_round(round);
    } else
    {
      ar.rewind(round_mark);
    }
    m = m_next;
    h *= 2;
  }
  return true;
}

void suffix_sort_sharded(esa_ctx *ctx, esa_comm *comm, u32 prefixlength, bool want_bck, bool want_hround,
                         SortOutcome *outcome)
{
  static const bool no_wide = getenv("ESA_NO_WIDE_KEY") != nullptr;
  // (two ranks: each range is half of the code space, 2^31 codes, one bit too many)
  const bool try_wide = ctx->sigma == 4 && comm->nranks >= 3 && !no_wide;
  if (try_wide && suffix_sort_sharded_impl(ctx, comm, prefixlength, want_bck, want_hround, outcome, true))
  {
    return;
  }
  suffix_sort_sharded_impl(ctx, comm, prefixlength, want_bck, want_hround, outcome, false);
}
