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
//   5. doubling rounds on the ACTIVE set only: sort by (rank[s], rank[s+h]) as one 64-bit key, regroup, h *= 2.
#include "kmer_scan.cuh"
#include "scan.cuh"

void radix_sort_init();

namespace
{

constexpr u32 NONE32 = 0xffffffffu;

// =========================================================================================================
// 1. classification
// scalars (device u64[16]): [0] nfull [1] nside [2] nspec [3] num_sorted M (= nfull + nside) [4] active count
// [5] carry (unused) [6] longest ...

__global__ void __launch_bounds__(KS_THREADS)
classify_count_kernel(const u8 *__restrict__ tis, u64 n, KmerParams kp, u32 *__restrict__ tile_full,
                      u32 *__restrict__ tile_side, u32 *__restrict__ tile_spec)
{
  __shared__ __align__(16) u8 stext[KS_TILE + KS_HALO];
  __shared__ u32 red[3][KS_THREADS / 32];
  const u64 tile_base = (u64) blockIdx.x * KS_TILE;
  const u64 N = n + 1;
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

// In-place exclusive sums of up to four per-tile count arrays, single CTA; totals go to scalars[first_scalar+a].
constexpr int SCAN1_THREADS = 1024;
__global__ void __launch_bounds__(SCAN1_THREADS)
tile_exclusive_sum_kernel(u32 *a0, u32 *a1, u32 *a2, u32 ntiles, u64 *scalars, int first_scalar)
{
  __shared__ u64 scratch[SCAN1_THREADS / 32 + 1];
  u32 *arrs[3] = {a0, a1, a2};
  for (int a = 0; a < 3; a++)
  {
    u32 *arr = arrs[a];
    if (arr == nullptr)
    {
      continue;
    }
    u64 carry = 0;
    for (u32 b = 0; b < ntiles; b += SCAN1_THREADS)
    {
      const u32 i = b + threadIdx.x;
      const u64 v = i < ntiles ? arr[i] : 0;
      u64 total;
      const u64 ex = block_exclusive_sum<u64, SCAN1_THREADS>(v, scratch, &total);
      if (i < ntiles)
      {
        arr[i] = (u32) (carry + ex);  // offsets stay below 2^32 because N does
      }
      carry += total;
    }
    if (threadIdx.x == 0)
    {
      scalars[first_scalar + a] = carry;
    }
  }
  if (threadIdx.x == 0 && first_scalar == 0)
  {
    scalars[3] = scalars[0] + scalars[1];
  }
}

__global__ void __launch_bounds__(KS_THREADS)
classify_emit_kernel(const u8 *__restrict__ tis, u64 n, KmerParams kp, const u32 *__restrict__ tile_full,
                     const u32 *__restrict__ tile_side, const u32 *__restrict__ tile_spec,
                     const u64 *__restrict__ scalars, u32 *__restrict__ main_key, u32 *__restrict__ main_pos,
                     u32 *__restrict__ side_key, u32 *__restrict__ side_pos, u32 *__restrict__ side_dj,
                     u32 *__restrict__ sa, u32 *__restrict__ rank)
{
  __shared__ __align__(16) u8 stext[KS_TILE + KS_HALO];
  // staging for coalesced output; index e is stored at e + e/32 so that the 16-strided per-thread writes hit
  // 32 different banks
  __shared__ u32 skey[KS_TILE + KS_TILE / 32];
  __shared__ u32 spos[KS_TILE + KS_TILE / 32];
  __shared__ u32 scratch[KS_THREADS / 32 + 1];
  const u64 tile_base = (u64) blockIdx.x * KS_TILE;
  const u64 N = n + 1;
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
  const u64 M = scalars[3];
#pragma unroll
  for (int i = 0; i < KS_ITEMS; i++)
  {
    const u64 p = tile_base + lo + i;
    if (p < N)
    {
      if (dist[i] == 0)
      {
        const u64 slot = M + off_spec + rp;  // tail of the table, position order
        sa[slot] = (u32) p;
        rank[p] = (u32) slot;
        rp++;
      } else if (dist[i] < kp.h)
      {
        const u64 o = off_side + rs;
        side_key[o] = (code[i] << 1) | 1u;
        side_pos[o] = (u32) p;
        side_dj[o] = kp.h - dist[i];
        rs++;
      } else
      {
        skey[rf + (rf >> 5)] = code[i] << 1;
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

// per-bucket count of suffixes with a special inside the k-prefix (side suffixes with j < k)
__global__ void side_bucket_count_kernel(const u32 *__restrict__ side_key, const u32 *__restrict__ side_dj, u64 nside,
                                         u32 h0, u32 k, u32 kdiv, u32 *__restrict__ cnt)
{
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nside)
  {
    const u32 j = h0 - side_dj[t];
    if (j < k)
    {
      atomicAdd(&cnt[(side_key[t] >> 1) / kdiv], 1u);
    }
  }
}

// =========================================================================================================
// 3. bck
__global__ void bck_left_kernel(const u32 *__restrict__ sorted_key, u64 m, u64 numofcodes, u32 kdiv,
                                u64 *__restrict__ bck)
{
  const u64 c = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= numofcodes)
  {
    return;
  }
  const u64 target = (c * kdiv) << 1;  // first key of the k-mer c; may be 2^32 for c == numofcodes only
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
                                         u64 tile_base, KeyT *skey /* [RG_TILE+2] */, u32 *sslot, u8 *sflag)
{
#pragma unroll
  for (int j = 0; j < RG_ITEMS; j++)
  {
    const int l = j * RG_THREADS + threadIdx.x;
    const u64 t = tile_base + l;
    if (t < m)
    {
      skey[l + 1] = keys[t];
      sslot[l] = FIRST ? (u32) t : slots[t];
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
      if (FIRST)
      {
        // a key with the flag bit is a suffix with a special inside its h0-prefix: unique by construction
        head = head || ((k & 1) != 0);
        next_head = next_head || ((t + 1 < m) && ((skey[l + 2] & 1) != 0));
      }
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
                       u32 *__restrict__ tile_maxrank, u32 *__restrict__ tile_active)
{
  __shared__ KeyT skey[RG_TILE + 2];
  __shared__ u32 sslot[RG_TILE];
  __shared__ __align__(8) u8 sflag[RG_TILE];
  __shared__ u32 red[2][RG_THREADS / 32];
  const u64 tile_base = (u64) blockIdx.x * RG_TILE;
  rg_flags<KeyT, FIRST>(keys, slots, m, tile_base, skey, sslot, sflag);
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

// single CTA: tile_maxrank -> exclusive running max (the rank carried into the tile),
// tile_active -> exclusive sum (first output slot of the tile); scalars[4] = total active
__global__ void __launch_bounds__(SCAN1_THREADS)
regroup_scan_kernel(u32 *tile_maxrank, u32 *tile_active, u32 ntiles, u64 *scalars)
{
  __shared__ u64 scratch[SCAN1_THREADS / 32 + 1];
  __shared__ u32 mscratch[SCAN1_THREADS / 32];
  __shared__ u32 sincl[SCAN1_THREADS];
  u64 carry_sum = 0;
  u32 carry_max = 0;
  for (u32 b = 0; b < ntiles; b += SCAN1_THREADS)
  {
    const u32 i = b + threadIdx.x;
    const u32 mv = i < ntiles ? tile_maxrank[i] : 0;
    const u64 cv = i < ntiles ? tile_active[i] : 0;
    u64 total;
    const u64 ex = block_exclusive_sum<u64, SCAN1_THREADS>(cv, scratch, &total);
    const u32 incl = block_inclusive_max<u32, SCAN1_THREADS>(mv, mscratch);
    sincl[threadIdx.x] = incl;
    __syncthreads();
    const u32 exmax = max(carry_max, threadIdx.x > 0 ? sincl[threadIdx.x - 1] : 0u);
    const u32 chunk_max = sincl[SCAN1_THREADS - 1];
    __syncthreads();
    if (i < ntiles)
    {
      tile_maxrank[i] = exmax;
      tile_active[i] = (u32) (carry_sum + ex);
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
                     u32 *__restrict__ next_slots, u32 rank_shift, u8 *__restrict__ hround, u32 round,
                     u32 *__restrict__ rank_out /* if set: rank_out[t] = rank instead of the scatter rank[pos] */)
{
  __shared__ KeyT skey[RG_TILE + 2];
  __shared__ u32 sslot[RG_TILE];
  __shared__ __align__(8) u8 sflag[RG_TILE];
  __shared__ u32 srank[RG_TILE];
  __shared__ u32 sdest[RG_TILE];
  __shared__ u32 scratch[RG_THREADS / 32 + 1];
  __shared__ u32 sincl[RG_THREADS];
  const u64 tile_base = (u64) blockIdx.x * RG_TILE;
  rg_flags<KeyT, FIRST>(keys, slots, m, tile_base, skey, sslot, sflag);
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
        sa[slot] = p;  // FIRST: the sorted position buffer already is sa[0..m)
      }
      if (rank_out != nullptr)
      {
        rank_out[t] = r;
      } else
      {
        rank[p] = r;
      }
      if (d != NONE32)
      {
        next_key[d] = (u64) r << rank_shift;
        next_pos[d] = p;
        next_slots[d] = slot;
      }
      if (!FIRST && hround != nullptr && t > 0 && (sflag[l] & 1))
      {
        // a head whose predecessor belongs to the same old group was split off in this round: the two
        // suffixes agree on the depth of the round (lcp lower bound for the direct lcp kernel)
        if ((u64) (skey[l + 1] >> rank_shift) == (u64) (skey[l] >> rank_shift))
        {
          hround[slot] = (u8) round;
        }
      }
    }
  }
}

// 5. doubling key: low bits = rank of the suffix h symbols further
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

u64 read_scalar(esa_ctx *ctx, const u64 *d_scalars, int idx)
{
  CUDA_CHECK(cudaMemcpyAsync(ctx->h_scalars, d_scalars + idx, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return ctx->h_scalars[0];
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
  const u32 ktiles = (u32) div_up_u64(N, KS_TILE);
  u32 *tile_full = ar.alloc<u32>(ktiles), *tile_side = ar.alloc<u32>(ktiles), *tile_spec = ar.alloc<u32>(ktiles);
  ESA_LAUNCH(ctx, ESA_K_CLASSIFY, (double) N,
             (classify_count_kernel<<<ktiles, KS_THREADS, 0, s>>>(ctx->d_tis, n, kp, tile_full, tile_side, tile_spec)));
  ESA_LAUNCH(ctx, ESA_K_MISC, 0,
             (tile_exclusive_sum_kernel<<<1, SCAN1_THREADS, 0, s>>>(tile_full, tile_side, tile_spec, ktiles, d_scalars,
                                                                     0)));
  CUDA_CHECK(cudaMemcpyAsync(ctx->h_scalars, d_scalars, 4 * sizeof(u64), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  const u64 nfull = ctx->h_scalars[0], nside = ctx->h_scalars[1], nspec = ctx->h_scalars[2];
  const u64 M = nfull + nside;  // suffixes that take part in the sort
  if (M + nspec != N)
  {
    esa_throw(ESA_ERR_CUDA, "classification lost suffixes: %llu + %llu + %llu != %llu", (unsigned long long) nfull,
              (unsigned long long) nside, (unsigned long long) nspec, (unsigned long long) N);
  }

  // region X: ping-pong buffers of the initial 32-bit sort (16 bytes per sorted suffix)
  const u64 Mal = M + 64;
  u32 *key_a = ar.alloc<u32>(Mal), *key_b = ar.alloc<u32>(Mal), *pos_a = ar.alloc<u32>(Mal), *pos_b = ar.alloc<u32>(Mal);
  u32 *side_key = ar.alloc<u32>(nside + 1), *side_pos = ar.alloc<u32>(nside + 1), *side_dj = ar.alloc<u32>(nside + 1),
      *side_tmp = ar.alloc<u32>(nside + 1);
  const size_t temp_bytes = max(radix_sort_temp_bytes(M, 4), radix_sort_temp_bytes(M, 8));
  void *sort_temp = ar.alloc<char>(temp_bytes);

  ESA_LAUNCH(ctx, ESA_K_CLASSIFY, (double) N + 8.0 * M + 8.0 * nspec,
             (classify_emit_kernel<<<ktiles, KS_THREADS, 0, s>>>(ctx->d_tis, n, kp, tile_full, tile_side, tile_spec,
                                                                  d_scalars, key_a, pos_a, side_key, side_pos, side_dj,
                                                                  sa, rank)));
  // side suffixes: stable sort by h0-j, written straight behind the full ones
  if (nside == 1)
  {
    CUDA_CHECK(cudaMemcpyAsync(key_a + nfull, side_key, 4, cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(pos_a + nfull, side_pos, 4, cudaMemcpyDeviceToDevice, s));
  } else if (nside > 1)
  {
    int w = radix_sort_pairs_u32(ctx, side_dj, side_tmp, side_pos, pos_a + nfull, nside, 0, 8, sort_temp, temp_bytes);
    int w2 = radix_sort_pairs_u32(ctx, side_dj, side_tmp, side_key, key_a + nfull, nside, 0, 8, sort_temp, temp_bytes);
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
  if (M > 1)
  {
    if (radix_sort_pairs_u32(ctx, key_a, key_b, pos_a, pos_b, M, 0, key_bits, sort_temp, temp_bytes))
    {
      skey = key_b;
      spos = pos_b;
    }
  }
  // the sorted positions are sa[0..M)
  if (M > 0)
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
  if (M == 0)
  {
    return;
  }
  const u32 rank_shift = bits_for(N - 1);  // doubling key = rank << rank_shift | rank of the suffix h further
  u32 rtiles = (u32) div_up_u64(M, RG_TILE);
  u32 *tile_maxrank = ar.alloc<u32>(rtiles), *tile_active = ar.alloc<u32>(rtiles);
  ESA_LAUNCH(ctx, ESA_K_REGROUP, 4.0 * M,
             (regroup_summary_kernel<u32, true><<<rtiles, RG_THREADS, 0, s>>>(skey, nullptr, M, tile_maxrank,
                                                                                tile_active)));
  ESA_LAUNCH(ctx, ESA_K_MISC, 0,
             (regroup_scan_kernel<<<1, SCAN1_THREADS, 0, s>>>(tile_maxrank, tile_active, rtiles, d_scalars)));
  u64 m = read_scalar(ctx, d_scalars, 4);
  outcome->active_first = m;

  // region Y: fresh buffers for the first doubling input; their ping-pong partners alias region X, which is
  // dead once the first regroup has consumed the sorted 32-bit keys
  const u64 mal = m + 64;
  u64 *dk_in = ar.alloc<u64>(mal);
  u32 *dp_in = ar.alloc<u32>(mal), *slots_cur = ar.alloc<u32>(mal);
  // scratch of the two-level rank scatter (rank in suffix-array order + the partitioned copy of (pos, rank));
  // skipped when the arena is too tight, then ranks are scattered directly
  u32 *part_rank = nullptr, *part_tp = nullptr, *part_tv = nullptr;
  if (M >= PARTITIONED_SCATTER_MIN && rank_shift > SCATTER_WINDOW_BITS && ar.capacity - ar.top > 12 * (M + 64) + (1 << 20))
  {
    part_rank = ar.alloc<u32>(M + 64);
    part_tp = ar.alloc<u32>(M + 64);
    part_tv = ar.alloc<u32>(M + 64);
  }
  ESA_LAUNCH(ctx, ESA_K_REGROUP, 8.0 * M + 4.0 * M + 16.0 * m,
             (regroup_apply_kernel<u32, true><<<rtiles, RG_THREADS, 0, s>>>(skey, spos, nullptr, M, tile_maxrank,
                                                                              tile_active, sa, rank, dk_in, dp_in,
                                                                              slots_cur, rank_shift, nullptr, 0u,
                                                                              part_rank)));
  if (part_rank != nullptr)
  {
    scatter_ranks_partitioned(ctx, spos, part_rank, M, rank_shift, rank, part_tp, part_tv, sort_temp, temp_bytes);
  }
  if (m == 0)
  {
    return;
  }
  // partners carved out of region X (key_a .. pos_b are contiguous allocations of 4*Mal words = 16*Mal bytes)
  u64 *dk_out = reinterpret_cast<u64 *>(key_a);          // 8*mal bytes  <= 8*Mal  (key_a + key_b)
  u32 *dp_out = pos_a;                                    // 4*mal bytes
  u32 *slots_next = pos_b;                                // 4*mal bytes

  // ---- 5. doubling
  u64 h = h0;
  while (m > 0)
  {
    outcome->rounds++;
    outcome->active_sum += m;
    ctx->lcp_work_estimate += m * (h / 32 + 1);
    ESA_LAUNCH(ctx, ESA_K_GATHER_RANK, 24.0 * m,
               (gather_rank_kernel<<<div_up_u32(m, 256), 256, 0, s>>>(dk_in, dp_in, m, h, rank)));
    u64 *ks = dk_in, *ko = dk_out;
    u32 *ps = dp_in, *po = dp_out;
    if (m > 1 && radix_sort_pairs_u64(ctx, dk_in, dk_out, dp_in, dp_out, m, 0, 2 * rank_shift, sort_temp, temp_bytes))
    {
      ks = dk_out;
      ko = dk_in;
      ps = dp_out;
      po = dp_in;
    }
    rtiles = (u32) div_up_u64(m, RG_TILE);
    ESA_LAUNCH(ctx, ESA_K_REGROUP, 12.0 * m,
               (regroup_summary_kernel<u64, false><<<rtiles, RG_THREADS, 0, s>>>(ks, slots_cur, m, tile_maxrank,
                                                                                   tile_active)));
    ESA_LAUNCH(ctx, ESA_K_MISC, 0,
               (regroup_scan_kernel<<<1, SCAN1_THREADS, 0, s>>>(tile_maxrank, tile_active, rtiles, d_scalars)));
    const u64 m_next = read_scalar(ctx, d_scalars, 4);
    const bool two_level = part_rank != nullptr && m >= PARTITIONED_SCATTER_MIN;
    ESA_LAUNCH(ctx, ESA_K_REGROUP, 24.0 * m + 16.0 * m_next,
               (regroup_apply_kernel<u64, false><<<rtiles, RG_THREADS, 0, s>>>(ks, ps, slots_cur, m, tile_maxrank,
                                                                                 tile_active, sa, rank, ko, po,
                                                                                 slots_next, rank_shift,
                                                                                 ctx->d_hround, outcome->rounds,
                                                                                 two_level ? part_rank : nullptr)));
    if (two_level)
    {
      scatter_ranks_partitioned(ctx, ps, part_rank, m, rank_shift, rank, part_tp, part_tv, sort_temp, temp_bytes);
    }
    dk_in = ko;
    dp_in = po;
    dk_out = ks;
    dp_out = ps;
    u32 *t = slots_cur;
    slots_cur = slots_next;
    slots_next = t;
    m = m_next;
    h *= 2;
    if (outcome->rounds > 64)
    {
      esa_throw(ESA_ERR_CUDA, "prefix doubling did not converge");
    }
  }
}
