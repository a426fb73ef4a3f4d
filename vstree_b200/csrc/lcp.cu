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
#include "scan.cuh"

namespace
{

constexpr u32 MARK = 0xffffffffu;
constexpr int PL_THREADS = 256;
constexpr int PL_ITEMS = 8;
constexpr int PL_TILE = PL_THREADS * PL_ITEMS;
constexpr int SCAN1_THREADS = 1024;

__device__ __forceinline__ u64 load_window8(const u8 *__restrict__ tis, u64 p)
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

// number of equal leading letters of the suffixes at i and j (i != j)
__device__ __forceinline__ u32 lcp_direct(const u8 *__restrict__ tis, u64 i, u64 j)
{
  u32 l = 0;
  while (true)
  {
    const u64 a = load_window8(tis, i + l), b = load_window8(tis, j + l);
    const u64 x = a ^ b;
    // bytes of a that are special (254/255): zero bytes of y; equal specials do not match either
    const u64 y = ~(a | 0x0101010101010101ull);
    const u64 z = (y - 0x0101010101010101ull) & ~y & 0x8080808080808080ull;
    const u32 stop_mis = x ? (u32) (__ffsll((long long) x) - 1) >> 3 : 8u;
    const u32 stop_sp = z ? (u32) (__ffsll((long long) z) - 1) >> 3 : 8u;
    const u32 stop = min(stop_mis, stop_sp);
    l += stop;
    if (stop < 8)
    {
      return l;
    }
  }
}

__global__ void __launch_bounds__(PL_THREADS)
plcp_irreducible_kernel(const u8 *__restrict__ tis, u64 n, const u32 *__restrict__ sa, const u32 *__restrict__ isa,
                        u32 *__restrict__ v, u32 *__restrict__ tile_last)
{
  __shared__ u32 red_idx[PL_THREADS / 32];
  __shared__ u32 red_val[PL_THREADS / 32];
  const u64 N = n + 1;
  const u64 tile_base = (u64) blockIdx.x * PL_TILE;
  u32 last_idx = 0, last_val = MARK;  // last valid (irreducible) item of this thread: local index + 1, its value
#pragma unroll
  for (int j = 0; j < PL_ITEMS; j++)
  {
    const int l = j * PL_THREADS + threadIdx.x;
    const u64 i = tile_base + l;
    if (i < N)
    {
      const u32 r = isa[i];
      u32 val;
      if (r == 0)
      {
        val = (u32) i;  // lcp of the first suffix is 0 (bese.c:639)
      } else
      {
        const u64 jpos = sa[r - 1];
        const u32 ci = i > 0 ? tis[i - 1] : 255u;
        const u32 cj = jpos > 0 ? tis[jpos - 1] : 255u;
        if (ci == cj && ci < ESA_WILDCARD)
        {
          val = MARK;
        } else
        {
          val = (u32) i + lcp_direct(tis, i, jpos);
        }
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
plcp_fill_kernel(u32 *__restrict__ v, u64 N, const u32 *__restrict__ tile_carry)
{
  __shared__ u32 sv[PL_TILE];
  __shared__ u32 scratch[PL_THREADS / 32];
  __shared__ u32 sincl[PL_THREADS];
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

}  // namespace

void compute_plcp(esa_ctx *ctx, u32 *d_v)
{
  const u64 n = ctx->n, N = n + 1;
  cudaStream_t s = ctx->stream;
  const u32 tiles = (u32) div_up_u64(N, PL_TILE);
  u32 *tile_last = ctx->arena.alloc<u32>(tiles);
  // algorithmic bytes: isa 4 + sa gather 4 + two symbols 2 + V write 4, plus the compared windows (data dependent)
  ESA_LAUNCH(ctx, ESA_K_PLCP, 14.0 * N,
             (plcp_irreducible_kernel<<<tiles, PL_THREADS, 0, s>>>(ctx->d_tis, n, ctx->d_sa, ctx->d_isa, d_v, tile_last)));
  ESA_LAUNCH(ctx, ESA_K_MISC, 0, (last_valid_scan_kernel<<<1, SCAN1_THREADS, 0, s>>>(tile_last, tiles)));
  ESA_LAUNCH(ctx, ESA_K_PLCP, 8.0 * N, (plcp_fill_kernel<<<tiles, PL_THREADS, 0, s>>>(d_v, N, tile_last)));
}
