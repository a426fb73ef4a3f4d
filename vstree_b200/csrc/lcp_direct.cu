// lcp_direct.cu -- lcp / bwt / 64-bit suftab in ONE sweep over the suffix array, with the lcp values taken by
// direct comparison of neighbouring suffixes.
//
// The reference also compares neighbours directly at bucket borders (GETLCP, src/Mkvtree/bese.c:91-95) and gets the
// rest from the depth at which its multikey quicksort split a group (SETLCP1, bese.c:827-846).  The same idea drives
// this kernel: the prefix-doubling sorter records, per suffix-array slot, the round in which the slot became a group
// head (ctx->d_hround).  Round t split groups that agreed on round_start_depth(t) symbols (h0*2^(t-1) for pure doubling),
// so lcp[r] lies in [h, 2h): the comparison starts at offset h and needs at most h/32 + 1 packed windows.
// Everything is read in suffix-array order (coalesced), the only random accesses go to the text, which for DNA is
// 2-bit packed and stays L2-resident up to ~300 Mbp.  No inverse suffix array and no plcp array are needed.
// For inputs whose deep groups are huge (long exact tandem arrays) the host picks the text-order plcp path
// (lcp.cu) instead: its cost does not depend on the lcp values.
#include "llv.cuh"
#include "textcmp.cuh"

namespace
{

// one thread per record of the packed text (layout: textcmp.cuh): 60 symbols starting one in front of position 32w,
// the 8 summary bits, and the 32-bit special mask of positions [32w, 32w + 32)
__global__ void __launch_bounds__(256)
pack2_kernel(const u8 *__restrict__ tis, u64 n, u64 padded_len, u64 words, u64 *__restrict__ pk,
             u32 *__restrict__ spm)
{
  const u64 w = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= words)
  {
    return;
  }
  const u64 p0 = w * 32;
  u64 lo = 0, hi = 0;
  u32 mask = 0, summary = 0;
  // record symbol 0 = the symbol in front of p0
  {
    const bool sp = p0 == 0 || p0 - 1 >= n || tis[p0 - 1] >= ESA_WILDCARD;
    if (sp)
    {
      summary |= 1u;
    } else
    {
      lo |= (u64) (tis[p0 - 1] & 3u);
    }
  }
  if (p0 + 64 <= padded_len)
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(tis + p0);
    const uint4 a = src[0], b = src[1], c = src[2], d = src[3];
    const u32 q[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll
    for (int i = 0; i < 59; i++)
    {
      const u32 ch = (q[i >> 2] >> ((i & 3) * 8)) & 0xffu;
      const bool sp = ch >= ESA_WILDCARD || p0 + i >= n;
      const int j = i + 1;  // record symbol index
      if (sp)
      {
        summary |= 1u << (j >> 3);
        if (i < 32)
        {
          mask |= 1u << i;
        }
      } else if (j < 32)
      {
        lo |= (u64) (ch & 3u) << (2 * j);
      } else
      {
        hi |= (u64) (ch & 3u) << (2 * (j - 32));
      }
    }
  } else
  {
    for (int i = 0; i < 59; i++)
    {
      const u64 pos = p0 + i;
      const u32 ch = pos < n ? tis[pos] : 255u;
      const bool sp = ch >= ESA_WILDCARD;
      const int j = i + 1;
      if (sp)
      {
        summary |= 1u << (j >> 3);
        if (i < 32)
        {
          mask |= 1u << i;
        }
      } else if (j < 32)
      {
        lo |= (u64) (ch & 3u) << (2 * j);
      } else
      {
        hi |= (u64) (ch & 3u) << (2 * (j - 32));
      }
    }
  }
  pk[2 * w] = lo;
  pk[2 * w + 1] = hi | ((u64) summary << 56);
  spm[w] = mask;
}

// one thread per 5-bit record (layout: textcmp.cuh, TextPacked5): the 24 symbols starting one in front of position 16w
__global__ void __launch_bounds__(256)
pack5_kernel(const u8 *__restrict__ tis, u64 n, u64 words, u64 *__restrict__ pk)
{
  const u64 w = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= words)
  {
    return;
  }
  const u64 p0 = w * 16;
  u64 lo = 0, hi = 0;
#pragma unroll
  for (int j = 0; j < 24; j++)
  {
    const u64 pos = p0 + j;  // text position + 1
    u32 c = TextPacked5::CODE_SEPARATOR;
    if (pos >= 1 && pos - 1 < n)
    {
      const u32 a = tis[pos - 1];
      c = a < ESA_WILDCARD ? a : (a == ESA_WILDCARD ? TextPacked5::CODE_WILDCARD : TextPacked5::CODE_SEPARATOR);
    }
    const int sh = 5 * j;
    if (sh < 64)
    {
      lo |= (u64) c << sh;
      if (sh > 59)
      {
        hi |= (u64) c >> (64 - sh);
      }
    } else
    {
      hi |= (u64) c << (sh - 64);
    }
  }
  pk[2 * w] = lo;
  pk[2 * w + 1] = hi;
}

// scalars: [0] maxbranchdepth [1] longest [2] nllv
template <typename Text>
__global__ void __launch_bounds__(FN_THREADS)
finalize_direct_kernel(Text text, const u8 *__restrict__ tis, u64 N, const u32 *__restrict__ sa,
                       const u8 *__restrict__ hround, DepthSched sched, bool want_lcp, u8 *__restrict__ lcp8,
                       u32 *__restrict__ lcpfull, u8 *__restrict__ bwt, u64 *__restrict__ suf64,
                       u32 *__restrict__ tile_large, unsigned long long *__restrict__ scalars, u64 index_base,
                       u32 halo_prev, const u32 *__restrict__ deep_v, u32 deep_round)
{
  // index_base: global suffix-array index of local entry 0 (0 on one GPU); halo_prev: the suffix in front of the
  // shard (last entry of the previous shard), used for the lcp of local entry 0
  __shared__ u32 red[2][FN_THREADS / 32];
  const u64 tile_base = (u64) blockIdx.x * FN_TILE;
  u32 nlarge = 0, mx = 0;
#pragma unroll
  for (int j = 0; j < FN_ITEMS / 4; j++)
  {
    const u64 r0 = tile_base + ((u64) j * FN_THREADS + threadIdx.x) * 4;
    if (r0 >= N)
    {
      continue;
    }
    u32 s[5];  // s[0] = sa[r0-1], s[1..4] = sa[r0..r0+3]
    u32 hr[4] = {0, 0, 0, 0};
    const int cnt = (int) min((u64) 4, N - r0);
    s[0] = r0 > 0 ? sa[r0 - 1] : halo_prev;
    if (cnt == 4)
    {
      const uint4 s4 = *reinterpret_cast<const uint4 *>(sa + r0);
      s[1] = s4.x;
      s[2] = s4.y;
      s[3] = s4.z;
      s[4] = s4.w;
      if (want_lcp)
      {
        const uchar4 h4 = *reinterpret_cast<const uchar4 *>(hround + r0);
        hr[0] = h4.x;
        hr[1] = h4.y;
        hr[2] = h4.z;
        hr[3] = h4.w;
      }
    } else
    {
      for (int q = 0; q < 4; q++)
      {
        s[q + 1] = q < cnt ? sa[r0 + q] : 0u;
        hr[q] = (want_lcp && q < cnt) ? hround[r0 + q] : 0u;
      }
    }
    u32 l[4] = {0, 0, 0, 0};
    u8 bb[4] = {0, 0, 0, 0};
    if (want_lcp || bwt != nullptr)
    {
      // every suffix start is probed once (first window, special mask, preceding symbol): five independent
      // probes per thread serve four comparisons and four bwt symbols
      typename Text::Probe pr[5];
#pragma unroll
      for (int q = 0; q < 5; q++)
      {
        pr[q] = text.probe(s[q]);
      }
      // (taking pr[0] from the neighbouring lane's pr[4] by shuffle saves a fifth of the probes but was slower: the
      // one lane per warp that still has to probe runs after the shuffles, 1.33 -> 1.39 ms per 100 Mbp)
#pragma unroll
      for (int q = 0; q < 4; q++)
      {
        if (q < cnt)
        {
          bb[q] = s[q + 1] > 0 ? (u8) pr[q + 1].before : (u8) ESA_UNDEFBWTCHAR;
        }
        if (want_lcp && q < cnt && hr[q] >= deep_round)
        {
          l[q] = deep_v[s[q + 1]] - s[q + 1];  // deep slot: the text-order value (lcp.cu, compute_plcp_deep)
        } else if (want_lcp && q < cnt && index_base + r0 + q > 0)
        {
          const u32 lb = hr[q] ? round_start_depth(sched, hr[q]) : 0u;
          // pairs split by a doubling round agree on lb symbols: start there (at most lb/WINDOW + 1 steps)
          const u32 stop = lb == 0 ? Text::probe_stop(pr[q], pr[q + 1]) : text.step((u64) s[q] + lb, (u64) s[q + 1] + lb);
          l[q] = lb + stop;
          if (stop == Text::WINDOW)
          {
            l[q] = text.lcp(s[q], s[q + 1], l[q]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
    {
      if (q < cnt && s[q + 1] == 0)
      {
        scalars[1] = index_base + r0 + q;
      }
      if (want_lcp && q < cnt)
      {
        nlarge += l[q] >= 255u;
        mx = max(mx, l[q]);
        if (l[q] >= 255u)
        {
          lcpfull[r0 + q] = l[q];
        }
      }
    }
    if (cnt == 4)
    {
      if (bwt != nullptr)
      {
        *reinterpret_cast<uchar4 *>(bwt + r0) = make_uchar4(bb[0], bb[1], bb[2], bb[3]);
      }
      if (want_lcp)
      {
        *reinterpret_cast<uchar4 *>(lcp8 + r0) = make_uchar4((u8) min(l[0], 255u), (u8) min(l[1], 255u),
                                                             (u8) min(l[2], 255u), (u8) min(l[3], 255u));
      }
      if (suf64 != nullptr)
      {
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(suf64 + r0);
        o[0] = make_ulonglong2(s[1], s[2]);
        o[1] = make_ulonglong2(s[3], s[4]);
      }
    } else
    {
      for (int q = 0; q < cnt; q++)
      {
        if (bwt != nullptr)
        {
          bwt[r0 + q] = bb[q];
        }
        if (want_lcp)
        {
          lcp8[r0 + q] = (u8) min(l[q], 255u);
        }
        if (suf64 != nullptr)
        {
          suf64[r0 + q] = s[q + 1];
        }
      }
    }
  }
  if (!want_lcp)
  {
    return;
  }
  nlarge = warp_reduce_sum(nlarge);
  mx = warp_reduce_max(mx);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
  {
    red[0][warp] = nlarge;
    red[1][warp] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    u32 a = 0, b = 0;
    for (int w = 0; w < FN_THREADS / 32; w++)
    {
      a += red[0][w];
      b = max(b, red[1][w]);
    }
    tile_large[blockIdx.x] = a;
    if (b > 0)
    {
      atomicMax(&scalars[0], (unsigned long long) b);
    }
  }
}

}  // namespace

void pack_text(esa_ctx *ctx)
{
  const u64 n = ctx->n, N = n + 1;
  const u64 words = (N + 31) / 32 + 4;
  const u64 padded = ((n + 15) & ~(u64) 15) + 64;
  ctx->d_pk = ctx->arena.alloc<u64>(2 * words);  // 16-byte records
  ctx->d_spm = ctx->arena.alloc<u32>(words);
  pack2_kernel<<<div_up_u32(words, 256), 256, 0, ctx->stream>>>(ctx->d_tis, n, padded, words, ctx->d_pk, ctx->d_spm);
  CUDA_CHECK(cudaGetLastError());
}

void pack_text5(esa_ctx *ctx)
{
  const u64 n = ctx->n, N = n + 1;
  const u64 words = (N + 15) / 16 + 4;
  ctx->d_pk5 = ctx->arena.alloc<u64>(2 * words);  // 16-byte records
  pack5_kernel<<<div_up_u32(words, 256), 256, 0, ctx->stream>>>(ctx->d_tis, n, words, ctx->d_pk5);
  CUDA_CHECK(cudaGetLastError());
}

// lcp (when demanded) by direct comparison; fills d_lcp, d_llv, d_bwt, d_suf64, maxbranchdepth, longest, nllv
void finalize_tables_direct(esa_ctx *ctx, u32 demand, u64 count, u64 index_base, u32 halo_prev, const u32 *deep_v,
                            u32 deep_round)
{
  // count entries of ctx->d_sa (the whole table on one GPU, this rank's shard otherwise)
  const u64 N = count;
  cudaStream_t s = ctx->stream;
  Arena &ar = ctx->arena;
  const bool want_lcp = (demand & ESA_LCPTAB) != 0;
  unsigned long long *d_scalars = (unsigned long long *) ar.alloc<u64>(8);
  ESA_LAUNCH(ctx, ESA_K_MEMOP, 0, CUDA_CHECK(cudaMemsetAsync(d_scalars, 0, 8 * sizeof(u64), s)));
  if (demand & ESA_SUFTAB)
  {
    ctx->d_suf64 = ar.alloc<u64>(N + 4);
  }
  if (demand & ESA_BWTTAB)
  {
    ctx->d_bwt = ar.alloc<u8>(N + 4);
  }
  u32 *lcpfull = nullptr;
  if (want_lcp)
  {
    ctx->d_lcp = ar.alloc<u8>(N + 4);
    lcpfull = ar.alloc<u32>(N + 4);
    ctx->d_lcpfull = lcpfull;
  }
  const u32 tiles = (u32) div_up_u64(N, FN_TILE);
  u32 *tile_large = ar.alloc<u32>(tiles);
  // algorithmic bytes: sa 4 (+8 widened) + hround 1 + lcp8 1 + bwt 1, plus one packed window per side and lcp
  const double bytes = (double) N * (4.0 + ((demand & ESA_SUFTAB) ? 8.0 : 0.0) + ((demand & ESA_BWTTAB) ? 1.0 : 0.0)
                                     + (want_lcp ? 2.0 + 16.0 : 0.0));
  if (ctx->d_pk != nullptr)
  {
    TextPacked2 text{ctx->d_pk, ctx->d_spm, ctx->d_tis};
    ESA_LAUNCH(ctx, ESA_K_FINALIZE, bytes,
               (finalize_direct_kernel<TextPacked2><<<tiles, FN_THREADS, 0, s>>>(
                   text, ctx->d_tis, N, ctx->d_sa, ctx->d_hround, ctx->sched, want_lcp, ctx->d_lcp, lcpfull, ctx->d_bwt,
                   ctx->d_suf64, tile_large, d_scalars, index_base, halo_prev, deep_v, deep_v != nullptr ? deep_round : 255u)));
  } else if (ctx->d_pk5 != nullptr)
  {
    TextPacked5 text{ctx->d_pk5};
    ESA_LAUNCH(ctx, ESA_K_FINALIZE, bytes,
               (finalize_direct_kernel<TextPacked5><<<tiles, FN_THREADS, 0, s>>>(
                   text, ctx->d_tis, N, ctx->d_sa, ctx->d_hround, ctx->sched, want_lcp, ctx->d_lcp, lcpfull, ctx->d_bwt,
                   ctx->d_suf64, tile_large, d_scalars, index_base, halo_prev, deep_v, deep_v != nullptr ? deep_round : 255u)));
  } else
  {
    TextBytes text{ctx->d_tis};
    ESA_LAUNCH(ctx, ESA_K_FINALIZE, bytes,
               (finalize_direct_kernel<TextBytes><<<tiles, FN_THREADS, 0, s>>>(
                   text, ctx->d_tis, N, ctx->d_sa, ctx->d_hround, ctx->sched, want_lcp, ctx->d_lcp, lcpfull, ctx->d_bwt,
                   ctx->d_suf64, tile_large, d_scalars, index_base, halo_prev, deep_v, deep_v != nullptr ? deep_round : 255u)));
  }
  if (want_lcp)
  {
    ESA_LAUNCH(ctx, ESA_K_MISC, 0, (large_scan_kernel<<<1, LLV_SCAN_THREADS, 0, s>>>(tile_large, tiles, d_scalars)));
  }
  read_back_u64(ctx, d_scalars, 3);
  ctx->maxbranchdepth = ctx->h_scalars[0];
  ctx->longest = ctx->h_scalars[1];
  ctx->nllv = want_lcp ? ctx->h_scalars[2] : 0;
  if (want_lcp && ctx->nllv > 0)
  {
    ctx->d_llv = ar.alloc<u64>(2 * ctx->nllv);
    ESA_LAUNCH(ctx, ESA_K_FINALIZE, (double) N + 20.0 * ctx->nllv,
               (llv_emit_kernel<<<tiles, FN_THREADS, 0, s>>>(N, ctx->d_sa, nullptr, lcpfull, ctx->d_lcp, tile_large,
                                                              ctx->nllv, ctx->d_llv, index_base)));
  }
}
