// skp.cu -- the skip table (.skp): skp[s] = (first index i > s with lcp[i] < lcp[s]) - 1, or n when there is none.
//
// Restates makeskiptable (reference: src/kurtz/mkskip.c:53-91), a sequential stack walk over the lcp table: an
// entry (depth = lcp[s], step = s) is popped by the first later lcp value that is strictly smaller, and receives the
// index in front of it; entries that survive get totallength.  That is the "next strictly smaller value" problem.
// Here every entry searches on its own in a min-pyramid over the exact 32-bit lcp values (fan-out 16: level k+1
// holds the minimum of 16 consecutive entries of level k): walk right through the siblings, climb at a group
// border, and once a block with a smaller minimum is met descend to its first smaller leaf.  Lcp tables are locally
// rough (the answer is usually a few entries away), so most searches end inside level 0.
#include "common.cuh"

namespace
{

constexpr u32 SKP_FAN = 16;
constexpr int SKP_MAX_LEVELS = 10;  // 16^9 > 2^32

struct Pyramid
{
  const u32 *level[SKP_MAX_LEVELS];
  u64 size[SKP_MAX_LEVELS];
  int top;  // index of the highest level
};

// exact lcp values: the clamped byte, or for 255 the value behind it (plcp array of the text-order path, lcpfull of
// the direct path)
__global__ void lcp32_kernel(u64 N, const u8 *__restrict__ lcp8, const u32 *__restrict__ sa,
                             const u32 *__restrict__ v, const u32 *__restrict__ lcpfull, u32 *__restrict__ out)
{
  const u64 r = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (r < N)
  {
    u32 l = lcp8[r];
    if (l == 255u)
    {
      l = v != nullptr ? v[sa[r]] - sa[r] : lcpfull[r];
    }
    out[r] = l;
  }
}

__global__ void min_level_kernel(const u32 *__restrict__ in, u64 in_size, u32 *__restrict__ out, u64 out_size)
{
  const u64 j = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (j < out_size)
  {
    u32 m = 0xffffffffu;
    const u64 b = j * SKP_FAN;
#pragma unroll
    for (u32 c = 0; c < SKP_FAN; c++)
    {
      if (b + c < in_size)
      {
        m = min(m, in[b + c]);
      }
    }
    out[j] = m;
  }
}

__global__ void __launch_bounds__(256) skip_kernel(Pyramid py, u64 N, u64 n, u64 *__restrict__ skp)
{
  const u64 r = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N)
  {
    return;
  }
  const u32 v = py.level[0][r];
  u64 result = n;
  if (v > 0)  // nothing is smaller than 0: such entries survive the whole walk (mkskip.c:84-89)
  {
    int k = 0;
    u64 p = r + 1;
    while (true)
    {
      if (p >= py.size[k])
      {
        break;  // ran off the table
      }
      if (py.level[k][p] < v)
      {
        // descend to the first smaller leaf
        while (k > 0)
        {
          k--;
          p *= SKP_FAN;
          while (py.level[k][p] >= v)
          {
            p++;
          }
        }
        result = p - 1;
        break;
      }
      p++;
      if (p % SKP_FAN == 0 && k < py.top)
      {
        p /= SKP_FAN;
        k++;
      }
    }
  }
  skp[r] = result;
}

}  // namespace

// fills ctx->d_skp from the tables of the finished build: lcp8 plus the source of the exact large values
void compute_skp(esa_ctx *ctx, const u32 *d_v, const u32 *d_lcpfull)
{
  const u64 n = ctx->n, N = n + 1;
  cudaStream_t s = ctx->stream;
  Arena &ar = ctx->arena;
  ctx->d_skp = ar.alloc<u64>(N + 4);
  u32 *base = ar.alloc<u32>(N + 64);
  ESA_LAUNCH(ctx, ESA_K_FINALIZE, 9.0 * N,
             (lcp32_kernel<<<div_up_u32(N, 256), 256, 0, s>>>(N, ctx->d_lcp, ctx->d_sa, d_v, d_lcpfull, base)));
  Pyramid py;
  memset(&py, 0, sizeof py);
  py.level[0] = base;
  py.size[0] = N;
  py.top = 0;
  while (py.size[py.top] > SKP_FAN && py.top + 1 < SKP_MAX_LEVELS)
  {
    const u64 in_size = py.size[py.top], out_size = div_up_u64(in_size, SKP_FAN);
    u32 *out = ar.alloc<u32>(out_size + 64);
    ESA_LAUNCH(ctx, ESA_K_FINALIZE, 4.0 * in_size,
               (min_level_kernel<<<div_up_u32(out_size, 256), 256, 0, s>>>(py.level[py.top], in_size, out, out_size)));
    py.top++;
    py.level[py.top] = out;
    py.size[py.top] = out_size;
  }
  ESA_LAUNCH(ctx, ESA_K_FINALIZE, 12.0 * N, (skip_kernel<<<div_up_u32(N, 256), 256, 0, s>>>(py, N, n, ctx->d_skp)));
}
