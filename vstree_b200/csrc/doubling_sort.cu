// doubling_sort.cu -- the sort step of one refinement round.
//
// Input: the active suffixes in suffix-array order, key = rank(s) << shift | secondary key (rank(s+h) in a doubling
// round, the code of the next symbols in a text round).  Because the list is in
// suffix-array order the high key part is already sorted: groups (equal rank(s)) are contiguous, and only the order
// INSIDE each group is open.  A global LSD radix sort ignores that and pays 7 memory sweeps of 24 bytes per pair for a
// 54-bit key.  Most groups are small (random DNA: 2-3 suffixes, 50 diverged copies of a repeat: about 50), so:
//   plan     one read of the keys: a group is sorted LOCALLY when it lies inside one tile of 2048 pairs and has at most
//            128 members, otherwise it is SPILLED; per-tile spill counts.
//   local    one read + one write of the pairs: every tile ranks the members of its local groups against each other
//            in shared memory (all pairs; at most 128 comparisons per pair) and writes them in order; spilled pairs
//            are copied through.
//   spill    the spilled pairs are compacted (order preserved), sorted with the global radix sort and written back to
//            the places they came from -- their groups are contiguous and in rank order, so the sorted run fits the
//            same index set.
// When more than half of the pairs would spill (huge groups: the first rounds of highly repetitive inputs) the
// global sort takes everything, as before.  Every path is STABLE: equal keys keep their order, which is text order
// from the initial sort on -- the text rounds rely on it (equal keys with a special in their window are final).
#include "scan.cuh"

namespace
{

constexpr int DS_THREADS = 256;
constexpr int DS_ITEMS = 8;
constexpr int DS_TILE = DS_THREADS * DS_ITEMS;
constexpr u32 DS_CAP = 128;  // largest group ranked by all-pairs comparison

// shared per-tile group structure: local group index of every pair and the first slot of every local group
struct TileGroups
{
  u32 gidx[DS_TILE];       // local group index
  u32 start[DS_TILE + 1];  // first local slot of group g; start[ngroups] = tile_count
};

// loads the tile's keys into skey (blocked order) and fills tg; returns through out params whether the first / last
// local group continues in the neighbouring tiles.  Every thread of the block must call it.
__device__ __forceinline__ void ds_groups(const u64 *__restrict__ keys, u64 m, u64 tile_base, u32 tile_count, u32 shift,
                                          u64 *skey, TileGroups &tg, u32 *scratch, u32 &ngroups, bool &first_open,
                                          bool &last_open)
{
  for (int l = threadIdx.x; l < DS_TILE; l += DS_THREADS)
  {
    skey[l] = (u32) l < tile_count ? keys[tile_base + l] : ~(u64) 0;
  }
  __syncthreads();
  const int b0 = threadIdx.x * DS_ITEMS;
  u32 heads = 0, nh = 0;
#pragma unroll
  for (int i = 0; i < DS_ITEMS; i++)
  {
    const int l = b0 + i;
    const bool head = (u32) l < tile_count && (l == 0 || (skey[l] >> shift) != (skey[l - 1] >> shift));
    heads |= head ? (1u << i) : 0u;
    nh += head;
  }
  u32 total;
  u32 g = block_exclusive_sum<u32, DS_THREADS>(nh, scratch, &total);
#pragma unroll
  for (int i = 0; i < DS_ITEMS; i++)
  {
    const int l = b0 + i;
    if ((heads >> i) & 1u)
    {
      tg.start[g] = (u32) l;
      g++;
    }
    tg.gidx[l] = g - 1;  // g >= 1 for every valid slot: slot 0 is a head
  }
  if (threadIdx.x == 0)
  {
    tg.start[total] = tile_count;
  }
  ngroups = total;
  first_open = tile_base > 0 && tile_count > 0 && (keys[tile_base - 1] >> shift) == (skey[0] >> shift);
  last_open = tile_base + tile_count < m && tile_count > 0
              && (keys[tile_base + tile_count] >> shift) == (skey[tile_count - 1] >> shift);
  __syncthreads();
}

__device__ __forceinline__ bool ds_spilled(const TileGroups &tg, u32 l, u32 ngroups, bool first_open, bool last_open)
{
  const u32 g = tg.gidx[l];
  const u32 size = tg.start[g + 1] - tg.start[g];
  return size > DS_CAP || (g == 0 && first_open) || (g + 1 == ngroups && last_open);
}

__global__ void __launch_bounds__(DS_THREADS)
ds_plan_kernel(const u64 *__restrict__ keys, u64 m, u32 shift, u8 *__restrict__ flags, u32 *__restrict__ tile_spill)
{
  __shared__ u64 skey[DS_TILE];
  __shared__ TileGroups tg;
  __shared__ u32 scratch[DS_THREADS / 32 + 1];
  __shared__ u32 red[DS_THREADS / 32];
  const u64 tile_base = (u64) blockIdx.x * DS_TILE;
  const u32 tile_count = (u32) min((u64) DS_TILE, m - tile_base);
  u32 ngroups;
  bool first_open, last_open;
  ds_groups(keys, m, tile_base, tile_count, shift, skey, tg, scratch, ngroups, first_open, last_open);
  u32 cnt = 0;
  for (u32 l = threadIdx.x; l < tile_count; l += DS_THREADS)
  {
    const bool sp = ds_spilled(tg, l, ngroups, first_open, last_open);
    flags[tile_base + l] = sp ? 1 : 0;
    cnt += sp;
  }
  cnt = warp_reduce_sum(cnt);
  if ((threadIdx.x & 31) == 0)
  {
    red[threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    u32 s = 0;
    for (int w = 0; w < DS_THREADS / 32; w++)
    {
      s += red[w];
    }
    tile_spill[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(DS_THREADS)
ds_local_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ pos, u64 m, u32 shift, u64 *__restrict__ kout,
                u32 *__restrict__ pout)
{
  __shared__ u64 skey[DS_TILE];
  __shared__ TileGroups tg;
  __shared__ u32 scratch[DS_THREADS / 32 + 1];
  const u64 tile_base = (u64) blockIdx.x * DS_TILE;
  const u32 tile_count = (u32) min((u64) DS_TILE, m - tile_base);
  u32 ngroups;
  bool first_open, last_open;
  ds_groups(keys, m, tile_base, tile_count, shift, skey, tg, scratch, ngroups, first_open, last_open);
  for (u32 l = threadIdx.x; l < tile_count; l += DS_THREADS)
  {
    const u64 k = skey[l];
    u32 dst = l;
    if (!ds_spilled(tg, l, ngroups, first_open, last_open))
    {
      const u32 g = tg.gidx[l];
      const u32 gs = tg.start[g], ge = tg.start[g + 1];
      u32 r = 0;
      for (u32 j = gs; j < ge; j++)
      {
        const u64 kj = skey[j];
        r += (kj < k) || (kj == k && j < l);
      }
      dst = gs + r;
    }
    kout[tile_base + dst] = k;
    pout[tile_base + dst] = pos[tile_base + l];
  }
}

// spilled pairs, in index order, to compact arrays (tile_off = exclusive sums of tile_spill)
__global__ void __launch_bounds__(DS_THREADS)
ds_collect_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ pos, const u8 *__restrict__ flags, u64 m,
                  const u32 *__restrict__ tile_off, u64 *__restrict__ skeys, u32 *__restrict__ spos,
                  u32 *__restrict__ sidx)
{
  __shared__ u32 scratch[DS_THREADS / 32 + 1];
  const u64 tile_base = (u64) blockIdx.x * DS_TILE;
  const u32 tile_count = (u32) min((u64) DS_TILE, m - tile_base);
  const int b0 = threadIdx.x * DS_ITEMS;
  u32 f = 0, n = 0;
#pragma unroll
  for (int i = 0; i < DS_ITEMS; i++)
  {
    const u32 l = (u32) (b0 + i);
    if (l < tile_count && flags[tile_base + l])
    {
      f |= 1u << i;
      n++;
    }
  }
  u32 o = tile_off[blockIdx.x] + block_exclusive_sum<u32, DS_THREADS>(n, scratch, nullptr);
#pragma unroll
  for (int i = 0; i < DS_ITEMS; i++)
  {
    if ((f >> i) & 1u)
    {
      const u64 t = tile_base + b0 + i;
      skeys[o] = keys[t];
      spos[o] = pos[t];
      sidx[o] = (u32) t;
      o++;
    }
  }
}

__global__ void ds_scatter_back_kernel(const u64 *__restrict__ skeys, const u32 *__restrict__ spos,
                                       const u32 *__restrict__ sidx, u64 count, u64 *__restrict__ kout,
                                       u32 *__restrict__ pout)
{
  const u64 j = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (j < count)
  {
    const u32 t = sidx[j];
    kout[t] = skeys[j];
    pout[t] = spos[j];
  }
}

// single CTA exclusive sum over the tile counts; total -> *total_out
__global__ void __launch_bounds__(1024) ds_scan_kernel(u32 *counts, u32 ntiles, u64 *total_out)
{
  __shared__ u64 scratch[1024 / 32 + 1];
  u64 carry = 0;
  for (u32 b = 0; b < ntiles; b += 1024)
  {
    const u32 i = b + threadIdx.x;
    const u64 v = i < ntiles ? counts[i] : 0;
    u64 total;
    const u64 ex = block_exclusive_sum<u64, 1024>(v, scratch, &total);
    if (i < ntiles)
    {
      counts[i] = (u32) (carry + ex);
    }
    carry += total;
  }
  if (threadIdx.x == 0)
  {
    *total_out = carry;
  }
}

}  // namespace

// Sorts (keys_a, pos_a)[0..m) by key bits [0, end_bit) given that the list is grouped by key >> shift (rank rounds:
// end_bit = 2*shift; text rounds: shift + width of the text key).  Stable.
// Returns 0 when the result is in (keys_a, pos_a), 1 when it is in (keys_b, pos_b).
int doubling_sort(esa_ctx *ctx, u64 *keys_a, u64 *keys_b, u32 *pos_a, u32 *pos_b, u64 m, u32 shift, u32 end_bit,
                  void *sort_temp, size_t temp_bytes)
{
  static const bool disabled = getenv("ESA_NO_LOCAL_SORT") != nullptr;
  if (m <= 1)
  {
    return 0;
  }
  if (disabled || m < 4096 || ctx->arena.high - ctx->arena.top < 16 * m + ((size_t) 4 << 20))
  {
    return radix_sort_pairs_u64(ctx, keys_a, keys_b, pos_a, pos_b, m, 0, end_bit, sort_temp, temp_bytes);
  }
  cudaStream_t s = ctx->stream;
  Arena &ar = ctx->arena;
  const size_t mark = ar.mark();
  const u32 tiles = (u32) div_up_u64(m, DS_TILE);
  u8 *flags = ar.alloc<u8>(m + 16);
  u32 *tile_spill = ar.alloc<u32>(tiles + 1);
  u64 *d_total = ar.alloc<u64>(1);
  ESA_LAUNCH(ctx, ESA_K_REGROUP, 9.0 * m,
             (ds_plan_kernel<<<tiles, DS_THREADS, 0, s>>>(keys_a, m, shift, flags, tile_spill)));
  ESA_LAUNCH(ctx, ESA_K_MISC, 0, (ds_scan_kernel<<<1, 1024, 0, s>>>(tile_spill, tiles, d_total)));
  read_back_u64(ctx, d_total, 1, 8);
  const u64 nspill = ctx->h_scalars[8];
  if (2 * nspill > m)
  {
    ar.rewind(mark);
    return radix_sort_pairs_u64(ctx, keys_a, keys_b, pos_a, pos_b, m, 0, end_bit, sort_temp, temp_bytes);
  }
  ESA_LAUNCH(ctx, ESA_K_RADIX_PASS, 24.0 * m,
             (ds_local_kernel<<<tiles, DS_THREADS, 0, s>>>(keys_a, pos_a, m, shift, keys_b, pos_b)));
  if (nspill > 0)
  {
    u64 *sk_a = ar.alloc<u64>(nspill + 8), *sk_b = ar.alloc<u64>(nspill + 8);
    u32 *sp_a = ar.alloc<u32>(nspill + 8), *sp_b = ar.alloc<u32>(nspill + 8), *sidx = ar.alloc<u32>(nspill + 8);
    ESA_LAUNCH(ctx, ESA_K_REGROUP, (double) m + 28.0 * nspill,
               (ds_collect_kernel<<<tiles, DS_THREADS, 0, s>>>(keys_a, pos_a, flags, m, tile_spill, sk_a, sp_a, sidx)));
    const int w = radix_sort_pairs_u64(ctx, sk_a, sk_b, sp_a, sp_b, nspill, 0, end_bit, sort_temp, temp_bytes);
    ESA_LAUNCH(ctx, ESA_K_REGROUP, 28.0 * nspill,
               (ds_scatter_back_kernel<<<div_up_u32(nspill, 256), 256, 0, s>>>(w ? sk_b : sk_a, w ? sp_b : sp_a, sidx,
                                                                               nspill, keys_b, pos_b)));
  }
  ar.rewind(mark);
  return 1;
}
