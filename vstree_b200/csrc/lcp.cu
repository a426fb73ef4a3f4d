// lcp.cu -- permuted LCP array in text order, V[i] = i + PLCP[i].
//
// The reference gets its lcp values as a by-product of the multikey quicksort plus direct comparisons at bucket
// borders (src/Mkvtree/bese.c:91-103,692-699,827-846).  Here they come from the finished suffix array:
// with phi(i) = sa[isa[i]-1], PLCP[i] = lcp(i, phi(i)).  A position is REDUCIBLE when the symbols in front of
// i and phi(i) are the same letter; then phi(i-1) = phi(i)-1 and PLCP[i] = PLCP[i-1]-1, i.e. V[i] = V[i-1].
// Only irreducible positions are compared symbol by symbol (their lcp values sum to O(n log n)), every
// reducible one copies the value of the nearest irreducible position on its left with one "last valid" scan.
// Specials never match (bese.c:26-55), so a comparison stops at the first special on either side; the text
// buffer is padded with 255 behind position n, which doubles as the sentinel.
#include "comm.cuh"
#include "scan.cuh"
#include "textcmp.cuh"

namespace
{

constexpr u32 MARK = 0xffffffffu;
constexpr int PL_THREADS = 256;
constexpr int PL_ITEMS = 8;
constexpr int PL_TILE = PL_THREADS * PL_ITEMS;
constexpr int SCAN1_THREADS = 1024;

template <typename Text>
__global__ void __launch_bounds__(PL_THREADS)
plcp_irreducible_kernel(Text text, u64 n, const u32 *__restrict__ sa, const u32 *__restrict__ isa,
                        u32 *__restrict__ v, u32 *__restrict__ tile_last)
{
  __shared__ u32 red_idx[PL_THREADS / 32];
  __shared__ u32 red_val[PL_THREADS / 32];
  const u64 N = n + 1;
  const u64 tile_base = (u64) blockIdx.x * PL_TILE;
  u32 last_idx = 0, last_val = MARK;  // last valid (irreducible) item of this thread: local index + 1, its value
  // the dependent gathers are issued loop by loop so that PL_ITEMS of them are in flight per thread
  u32 rr[PL_ITEMS], jj[PL_ITEMS], cc[PL_ITEMS];
#pragma unroll
  for (int j = 0; j < PL_ITEMS; j++)
  {
    const u64 i = tile_base + j * PL_THREADS + threadIdx.x;
    rr[j] = i < N ? isa[i] : 0u;
  }
#pragma unroll
  for (int j = 0; j < PL_ITEMS; j++)
  {
    jj[j] = rr[j] > 0 ? sa[rr[j] - 1] : 0u;
  }
#pragma unroll
  for (int j = 0; j < PL_ITEMS; j++)
  {
    const u64 i = tile_base + j * PL_THREADS + threadIdx.x;
    const u32 ci = i < N ? text.before(i) : 255u;
    const u32 cj = rr[j] > 0 ? text.before(jj[j]) : 255u;
    cc[j] = (ci == cj && ci < ESA_WILDCARD) ? 1u : 0u;  // reducible
  }
#pragma unroll
  for (int j = 0; j < PL_ITEMS; j++)
  {
    const int l = j * PL_THREADS + threadIdx.x;
    const u64 i = tile_base + l;
    if (i < N)
    {
      u32 val;
      if (rr[j] == 0)
      {
        val = (u32) i;  // lcp of the first suffix is 0 (bese.c:639)
      } else if (cc[j])
      {
        val = MARK;
      } else
      {
        val = (u32) i + text.lcp(i, jj[j], 0);
      }
      v[i] = val;
      if (val != MARK)
      {
        last_idx = (u32) l + 1;
        last_val = val;
      }
    }
  }
  // tile summary: value of the last irreducible position of the tile (MARK if none)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    const u32 oi = __shfl_down_sync(FULL_MASK, last_idx, o);
    const u32 ov = __shfl_down_sync(FULL_MASK, last_val, o);
    if (oi > last_idx)
    {
      last_idx = oi;
      last_val = ov;
    }
  }
  if (lane == 0)
  {
    red_idx[warp] = last_idx;
    red_val[warp] = last_val;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    u32 bi = 0, bv = MARK;
    for (int w = 0; w < PL_THREADS / 32; w++)
    {
      if (red_idx[w] > bi)
      {
        bi = red_idx[w];
        bv = red_val[w];
      }
    }
    tile_last[blockIdx.x] = bv;
  }
}

// single CTA: tile_last[t] := value of the last irreducible position in tiles 0..t-1 (exclusive "last valid")
__global__ void __launch_bounds__(SCAN1_THREADS) last_valid_scan_kernel(u32 *tile_last, u32 ntiles)
{
  __shared__ u32 mscratch[SCAN1_THREADS / 32];
  __shared__ u32 sval[SCAN1_THREADS];
  __shared__ u32 sincl[SCAN1_THREADS];
  u32 carry = MARK;
  for (u32 b = 0; b < ntiles; b += SCAN1_THREADS)
  {
    const u32 i = b + threadIdx.x;
    const u32 val = i < ntiles ? tile_last[i] : MARK;
    sval[threadIdx.x] = val;
    const u32 idx = val != MARK ? threadIdx.x + 1 : 0u;
    const u32 incl = block_inclusive_max<u32, SCAN1_THREADS>(idx, mscratch);
    sincl[threadIdx.x] = incl;
    __syncthreads();
    const u32 ex = threadIdx.x > 0 ? sincl[threadIdx.x - 1] : 0u;
    const u32 out = ex > 0 ? sval[ex - 1] : carry;
    const u32 last = sincl[SCAN1_THREADS - 1];
    const u32 new_carry = last > 0 ? sval[last - 1] : carry;
    __syncthreads();
    if (i < ntiles)
    {
      tile_last[i] = out;
    }
    carry = new_carry;
  }
}

__global__ void __launch_bounds__(PL_THREADS)
plcp_fill_kernel(u32 *__restrict__ v, u64 N, const u32 *__restrict__ tile_carry, const u8 *__restrict__ tile_flag = nullptr)
{
  __shared__ u32 sv[PL_TILE];
  __shared__ u32 scratch[PL_THREADS / 32];
  __shared__ u32 sincl[PL_THREADS];
  if (tile_flag != nullptr && tile_flag[blockIdx.x] == 0)
  {
    return;  // no deep position in this tile (hybrid lcp): nothing to fill
  }
  const u64 tile_base = (u64) blockIdx.x * PL_TILE;
#pragma unroll
  for (int j = 0; j < PL_ITEMS; j++)
  {
    const int l = j * PL_THREADS + threadIdx.x;
    const u64 i = tile_base + l;
    sv[l] = i < N ? v[i] : MARK;
  }
  __syncthreads();
  const int b0 = threadIdx.x * PL_ITEMS;
  u32 last = 0;
#pragma unroll
  for (int i = 0; i < PL_ITEMS; i++)
  {
    if (sv[b0 + i] != MARK)
    {
      last = (u32) (b0 + i) + 1;
    }
  }
  const u32 incl = block_inclusive_max<u32, PL_THREADS>(last, scratch);
  sincl[threadIdx.x] = incl;
  __syncthreads();
  const u32 ex = threadIdx.x > 0 ? sincl[threadIdx.x - 1] : 0u;
  // sv entries that are valid are never overwritten below, so reading sv[ex-1] races with nothing
  u32 cur = ex > 0 ? sv[ex - 1] : tile_carry[blockIdx.x];
#pragma unroll
  for (int i = 0; i < PL_ITEMS; i++)
  {
    const u32 x = sv[b0 + i];
    if (x != MARK)
    {
      cur = x;
    } else
    {
      sv[b0 + i] = cur;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < PL_ITEMS; j++)
  {
    const int l = j * PL_THREADS + threadIdx.x;
    const u64 i = tile_base + l;
    if (i < N)
    {
      v[i] = sv[l];
    }
  }
}

// value of the last irreducible position of every tile (MARK if none), for the text-order fill
__global__ void __launch_bounds__(PL_THREADS)
plcp_tile_last_kernel(const u32 *__restrict__ v, u64 N, u32 *__restrict__ tile_last, const u8 *__restrict__ tile_flag = nullptr)
{
  __shared__ u32 red_idx[PL_THREADS / 32];
  __shared__ u32 red_val[PL_THREADS / 32];
  if (tile_flag != nullptr && tile_flag[blockIdx.x] == 0)
  {
    if (threadIdx.x == 0)
    {
      tile_last[blockIdx.x] = MARK;
    }
    return;
  }
  const u64 tile_base = (u64) blockIdx.x * PL_TILE;
  u32 last_idx = 0, last_val = MARK;
#pragma unroll
  for (int j = 0; j < PL_ITEMS; j++)
  {
    const int l = j * PL_THREADS + threadIdx.x;
    const u64 i = tile_base + l;
    const u32 val = i < N ? v[i] : MARK;
    if (val != MARK)
    {
      last_idx = (u32) l + 1;
      last_val = val;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    const u32 oi = __shfl_down_sync(FULL_MASK, last_idx, o);
    const u32 ov = __shfl_down_sync(FULL_MASK, last_val, o);
    if (oi > last_idx)
    {
      last_idx = oi;
      last_val = ov;
    }
  }
  if (lane == 0)
  {
    red_idx[warp] = last_idx;
    red_val[warp] = last_val;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    u32 bi = 0, bv = MARK;
    for (int w = 0; w < PL_THREADS / 32; w++)
    {
      if (red_idx[w] > bi)
      {
        bi = red_idx[w];
        bv = red_val[w];
      }
    }
    tile_last[blockIdx.x] = bv;
  }
}


// ---- hybrid lcp: the text-order machinery for the DEEP entries only ----
// A slot that a refinement round t >= deep_round split from its predecessor has an lcp value of at least
// round_start_depth(t) symbols; comparing such pairs directly costs up to that many symbols per entry (long exact tandem
// arrays: hundreds of thousands), so they take their value from V[i] = i + PLCP[i] instead.  The set of deep positions is
// closed under "left neighbour of a reducible member" (PLCP[i-1] = PLCP[i] + 1 is deeper still), so the last-valid fill
// works on the deep stretches alone: no memset of V, no work in tiles without a deep position, and only the
// IRREDUCIBLE deep positions are ever compared.  Everything else keeps the direct comparison of lcp_direct.cu, whose
// cost per entry is now bounded by the depth of round deep_round.
// list (sharded builds): (position, value) pairs of this rank's deep slots, for the other ranks.
template <typename Text>
__global__ void __launch_bounds__(256)
plcp_deep_kernel(Text text, u64 count, const u32 *__restrict__ sa, const u8 *__restrict__ hround, DepthSched sched,
                 u32 deep_round, u32 halo_prev, u32 *__restrict__ v, u8 *__restrict__ tile_flag, u32 *__restrict__ list,
                 unsigned long long *__restrict__ list_count)
{
  const u64 r = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  const u32 hr = r < count ? hround[r] : 0u;
  const bool deep = hr >= deep_round;
  u32 i = 0, val = MARK;
  if (deep)
  {
    i = sa[r];
    const u32 j = r > 0 ? sa[r - 1] : halo_prev;  // hr > 0: the slot has a predecessor in its old group
    const u32 ci = text.before(i), cj = text.before(j);
    if (!(ci == cj && ci < ESA_WILDCARD))
    {
      val = i + text.lcp(i, j, round_start_depth(sched, hr));
    }
    v[i] = val;
    tile_flag[i / PL_TILE] = 1;
  }
  if (list != nullptr)
  {
    const u32 bal = __ballot_sync(FULL_MASK, deep);
    if (bal != 0)
    {
      const int lane = threadIdx.x & 31, leader = __ffs((int) bal) - 1;
      unsigned long long base = 0;
      if (lane == leader)
      {
        base = atomicAdd(list_count, (unsigned long long) __popc(bal));
      }
      base = __shfl_sync(FULL_MASK, base, leader);
      if (deep)
      {
        const u64 o = base + __popc(bal & ((1u << lane) - 1));
        list[2 * o] = i;
        list[2 * o + 1] = val;
      }
    }
  }
}

__global__ void plcp_deep_scatter_kernel(const u32 *__restrict__ list, u64 pairs, u32 *__restrict__ v, u8 *__restrict__ tile_flag)
{
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < pairs)
  {
    const u32 i = list[2 * t];
    v[i] = list[2 * t + 1];
    tile_flag[i / PL_TILE] = 1;
  }
}

}  // namespace

// V[i] for every deep position i (see above); comm == nullptr on one GPU.  d_v needs no initialisation.
void compute_plcp_deep(esa_ctx *ctx, esa_comm *comm, u32 *d_v, u32 halo_prev, u32 deep_round)
{
  const u64 N = ctx->n + 1;
  cudaStream_t s = ctx->stream;
  Arena &ar = ctx->arena;
  const u32 tiles = (u32) div_up_u64(N, PL_TILE);
  u32 *tile_last = ar.alloc<u32>(tiles);
  u8 *tile_flag = ar.alloc<u8>(tiles + 16);
  unsigned long long *d_cnt = (unsigned long long *) ar.alloc<u64>(2);
  ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(tile_flag, 0, tiles + 16, s)));
  ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(u64), s)));
  const u64 count = ctx->shard_count;
  const bool sharded = comm != nullptr && comm->nranks > 1;
  // sharded: my pairs go to a list sized for the worst case of this range; it lives at the top of the scratch region
  const size_t mark = ar.mark();
  u32 *list = sharded ? ar.alloc<u32>(2 * count + 2) : nullptr;
  if (count > 0)
  {
    const u32 grid = (u32) div_up_u64(count, 256);
    if (ctx->d_pk != nullptr)
    {
      TextPacked2 text{ctx->d_pk, ctx->d_spm, ctx->d_tis};
      ESA_LAUNCH(ctx, ESA_K_PLCP, 5.0 * count,
                 (plcp_deep_kernel<TextPacked2><<<grid, 256, 0, s>>>(text, count, ctx->d_sa, ctx->d_hround, ctx->sched,
                                                                      deep_round, halo_prev, d_v, tile_flag, list, d_cnt)));
    } else if (ctx->d_pk5 != nullptr)
    {
      TextPacked5 text{ctx->d_pk5};
      ESA_LAUNCH(ctx, ESA_K_PLCP, 5.0 * count,
                 (plcp_deep_kernel<TextPacked5><<<grid, 256, 0, s>>>(text, count, ctx->d_sa, ctx->d_hround, ctx->sched,
                                                                      deep_round, halo_prev, d_v, tile_flag, list, d_cnt)));
    } else
    {
      TextBytes text{ctx->d_tis};
      ESA_LAUNCH(ctx, ESA_K_PLCP, 5.0 * count,
                 (plcp_deep_kernel<TextBytes><<<grid, 256, 0, s>>>(text, count, ctx->d_sa, ctx->d_hround, ctx->sched,
                                                                    deep_round, halo_prev, d_v, tile_flag, list, d_cnt)));
    }
  }
  if (sharded)
  {
    read_back_u64(ctx, d_cnt, 1, 8);
    const u64 mine = ctx->h_scalars[8];
    u64 all[ESA_MAX_RANKS], off[ESA_MAX_RANKS], cnt[ESA_MAX_RANKS], total = 0;
    comm_exchange_host_u64(ctx, comm, &mine, 1, all);
    for (int r = 0; r < comm->nranks; r++)
    {
      off[r] = 2 * total;
      cnt[r] = 2 * all[r];
      total += all[r];
    }
    if (total > 0)
    {
      u32 *gathered = ar.alloc<u32>(2 * total + 2);
      comm_allgatherv_u32(ctx, comm, list, gathered, off, cnt);
      ESA_LAUNCH(ctx, ESA_K_PLCP, 12.0 * total,
                 (plcp_deep_scatter_kernel<<<(u32) div_up_u64(total, 256), 256, 0, s>>>(gathered, total, d_v, tile_flag)));
    }
  }
  ESA_LAUNCH(ctx, ESA_K_PLCP, (double) tiles, (plcp_tile_last_kernel<<<tiles, PL_THREADS, 0, s>>>(d_v, N, tile_last, tile_flag)));
  ESA_LAUNCH(ctx, ESA_K_MISC, 0, (last_valid_scan_kernel<<<1, SCAN1_THREADS, 0, s>>>(tile_last, tiles)));
  ESA_LAUNCH(ctx, ESA_K_PLCP, (double) tiles, (plcp_fill_kernel<<<tiles, PL_THREADS, 0, s>>>(d_v, N, tile_last, tile_flag)));
  CUDA_CHECK(cudaStreamSynchronize(s));  // the list buffers are released below
  ar.rewind(mark);
}

namespace
{
}  // namespace

void compute_plcp(esa_ctx *ctx, u32 *d_v)
{
  const u64 n = ctx->n, N = n + 1;
  cudaStream_t s = ctx->stream;
  ensure_isa(ctx);
  const u32 tiles = (u32) div_up_u64(N, PL_TILE);
  u32 *tile_last = ctx->arena.alloc<u32>(tiles);
  // algorithmic bytes: isa 4 + sa gather 4 + two symbols 2 + V write 4, plus the compared windows (data dependent)
  if (ctx->d_pk != nullptr)
  {
    TextPacked2 text{ctx->d_pk, ctx->d_spm, ctx->d_tis};
    ESA_LAUNCH(ctx, ESA_K_PLCP, 14.0 * N,
               (plcp_irreducible_kernel<TextPacked2><<<tiles, PL_THREADS, 0, s>>>(text, n, ctx->d_sa, ctx->d_isa, d_v,
                                                                                   tile_last)));
  } else if (ctx->d_pk5 != nullptr)
  {
    TextPacked5 text{ctx->d_pk5};
    ESA_LAUNCH(ctx, ESA_K_PLCP, 14.0 * N,
               (plcp_irreducible_kernel<TextPacked5><<<tiles, PL_THREADS, 0, s>>>(text, n, ctx->d_sa, ctx->d_isa, d_v,
                                                                                   tile_last)));
  } else
  {
    TextBytes text{ctx->d_tis};
    ESA_LAUNCH(ctx, ESA_K_PLCP, 14.0 * N,
               (plcp_irreducible_kernel<TextBytes><<<tiles, PL_THREADS, 0, s>>>(text, n, ctx->d_sa, ctx->d_isa, d_v,
                                                                                 tile_last)));
  }
  ESA_LAUNCH(ctx, ESA_K_MISC, 0, (last_valid_scan_kernel<<<1, SCAN1_THREADS, 0, s>>>(tile_last, tiles)));
  ESA_LAUNCH(ctx, ESA_K_PLCP, 8.0 * N, (plcp_fill_kernel<<<tiles, PL_THREADS, 0, s>>>(d_v, N, tile_last)));
}
