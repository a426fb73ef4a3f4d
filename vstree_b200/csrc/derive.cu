// derive.cu -- tables derived from the suffix array in one sweep each:
//   finalize:  lcp8 / llv / maxbranchdepth (outlcpsubtab, src/Mkvtree/bese.c:533-591), bwt
//              (encodeburrowswheeler, src/kurtz/bwtcode.c:293-311), longest (determinelongest,
//              src/Mkvtree/mkvprocess.c:857-873) and the 64-bit suftab the file format wants (types.h:41-49)
//   sti1:      makestitab1 (src/Mkvtree/mkvprocess.c:583-643) restated per text position:
//              sti1[s] = 0 if a special lies inside the k-prefix of s, else min(255, isa[s] - bck.left[kmer(s)])
//   checks:    the reference's DEBUG invariants as device-side property checks.
#include "comm.cuh"
#include "kmer_scan.cuh"
#include "llv.cuh"
#include "scan.cuh"

namespace
{


// scalars: [0] maxbranchdepth [1] longest [2] nllv
__global__ void __launch_bounds__(FN_THREADS)
finalize_kernel(const u8 *__restrict__ tis, u64 N, const u32 *__restrict__ sa, const u32 *__restrict__ v,
                u8 *__restrict__ lcp8, u8 *__restrict__ bwt, u64 *__restrict__ suf64, u32 *__restrict__ tile_large,
                unsigned long long *__restrict__ scalars, u64 index_base /* global index of local entry 0 */)
{
  __shared__ u32 red[2][FN_THREADS / 32];
  const u64 tile_base = (u64) blockIdx.x * FN_TILE;
  u32 nlarge = 0, mx = 0;
  // four consecutive ranks per thread and step: 16-byte loads of sa, four independent gathers in flight,
  // 4-byte stores of lcp8/bwt and 32-byte stores of the widened suftab
#pragma unroll
  for (int j = 0; j < FN_ITEMS / 4; j++)
  {
    const u64 r0 = tile_base + ((u64) j * FN_THREADS + threadIdx.x) * 4;
    if (r0 + 4 <= N)
    {
      const uint4 s4 = *reinterpret_cast<const uint4 *>(sa + r0);
      const u32 s[4] = {s4.x, s4.y, s4.z, s4.w};
      u32 vv[4] = {0, 0, 0, 0};
      u8 bb[4] = {0, 0, 0, 0};
      if (v != nullptr)
      {
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
          vv[q] = v[s[q]];
        }
      }
      if (bwt != nullptr)
      {
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
          bb[q] = s[q] > 0 ? tis[s[q] - 1] : (u8) ESA_UNDEFBWTCHAR;
        }
        *reinterpret_cast<uchar4 *>(bwt + r0) = make_uchar4(bb[0], bb[1], bb[2], bb[3]);
      }
      if (suf64 != nullptr)
      {
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(suf64 + r0);
        o[0] = make_ulonglong2(s[0], s[1]);
        o[1] = make_ulonglong2(s[2], s[3]);
      }
#pragma unroll
      for (int q = 0; q < 4; q++)
      {
        if (s[q] == 0)
        {
          scalars[1] = index_base + r0 + q;
        }
      }
      if (v != nullptr)
      {
        u8 l8[4];
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
          const u32 l = vv[q] - s[q];
          l8[q] = (u8) min(l, 255u);
          nlarge += l >= 255u;
          mx = max(mx, l);
        }
        *reinterpret_cast<uchar4 *>(lcp8 + r0) = make_uchar4(l8[0], l8[1], l8[2], l8[3]);
      }
    } else
    {
      for (u64 r = r0; r < N; r++)
      {
        const u32 s = sa[r];
        if (suf64 != nullptr)
        {
          suf64[r] = s;
        }
        if (s == 0)
        {
          scalars[1] = index_base + r;
        }
        if (bwt != nullptr)
        {
          bwt[r] = s > 0 ? tis[s - 1] : (u8) ESA_UNDEFBWTCHAR;
        }
        if (v != nullptr)
        {
          const u32 l = v[s] - s;
          lcp8[r] = (u8) min(l, 255u);
          nlarge += l >= 255u;
          mx = max(mx, l);
        }
      }
    }
  }
  if (v == nullptr)
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

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KS_THREADS)
sti1_kernel(const u8 *__restrict__ tis, u64 n, KmerParams kp, const u32 *__restrict__ isa,
            const u64 *__restrict__ bck, u8 *__restrict__ sti1)
{
  __shared__ __align__(16) u8 stext[KS_TILE + KS_HALO];
  __shared__ __align__(16) u8 sout[KS_TILE];
  const u64 tile_base = (u64) blockIdx.x * KS_TILE;
  const u64 N = n + 1;
  ks_load_tile(tis, n, tile_base, stext);
  __syncthreads();
  u32 code[KS_ITEMS], dist[KS_ITEMS];
  const int lo = threadIdx.x * KS_ITEMS;
  ks_scan_thread(stext, lo, kp, code, dist);
#pragma unroll
  for (int i = 0; i < KS_ITEMS; i++)
  {
    const u64 p = tile_base + lo + i;
    u8 out = 0;
    if (p < N && dist[i] >= kp.h)
    {
      const u64 off = (u64) isa[p] - bck[2 * (u64) code[i]];
      out = (u8) (off < 255 ? off : 255);
    }
    sout[lo + i] = out;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < KS_TILE; l += KS_THREADS)
  {
    const u64 p = tile_base + l;
    if (p < N)
    {
      sti1[p] = sout[l];
    }
  }
}

__global__ void widen_kernel(const u32 *__restrict__ in, u64 *__restrict__ out, u64 N)
{
  const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N)
  {
    out[i] = in[i];
  }
}

// ---------------------------------------------------------------------------------------------------------
// checks.  errors: u32[8] counters, one per property.
__device__ __forceinline__ u64 sym_at(const u8 *tis, u64 n, u64 p)
{
  return (p < n && tis[p] < ESA_WILDCARD) ? (u64) tis[p] : p + 256;
}

// scal (u64): [0] largest lcp value seen, [1] number of values >= 255
__global__ void check_kernel(const u8 *__restrict__ tis, u64 n, const u32 *__restrict__ sa,
                             const u32 *__restrict__ isa, const u32 *__restrict__ v, const u8 *__restrict__ lcp8,
                             const u8 *__restrict__ bwt, const u64 *__restrict__ suf64, const u64 *__restrict__ llv,
                             u64 nllv, u32 *__restrict__ errors, unsigned long long *__restrict__ scal)
{
  const u64 N = n + 1;
  const u64 r = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N)
  {
    return;
  }
  const u64 s = sa[r];
  // 1: permutation (checksuftab, bese.c:355-451): isa inverts sa
  if (s > n || isa[s] != r)
  {
    atomicAdd(&errors[1], 1u);
    return;
  }
  if (suf64 != nullptr && suf64[r] != s)
  {
    atomicAdd(&errors[1], 1u);
  }
  // 4: bwt against the text (checksufencoding, cprsuf.c:66-95)
  if (bwt != nullptr && bwt[r] != (s > 0 ? tis[s - 1] : (u8) ESA_UNDEFBWTCHAR))
  {
    atomicAdd(&errors[4], 1u);
  }
  if (r == 0)
  {
    if (lcp8 != nullptr && lcp8[0] != 0)
    {
      atomicAdd(&errors[3], 1u);
    }
    return;
  }
  // 2 + 3: order of adjacent suffixes and the lcp value, by direct comparison (checklcpsubtab, bese.c:454-528)
  const u64 p = sa[r - 1];
  u64 l = 0;
  while (sym_at(tis, n, p + l) == sym_at(tis, n, s + l))
  {
    l++;
  }
  if (sym_at(tis, n, p + l) > sym_at(tis, n, s + l))
  {
    atomicAdd(&errors[2], 1u);
  }
  if (lcp8 != nullptr && lcp8[r] != (u8) (l < 255 ? l : 255))
  {
    atomicAdd(&errors[3], 1u);
  }
  if (v != nullptr && (u64) (v[s] - (u32) s) != l)
  {
    atomicAdd(&errors[3], 1u);
  }
  if (lcp8 != nullptr)
  {
    // 6: the EXACT value of a large entry is the one the .llv pair of index r holds (outlcpsubtab, bese.c:557-566;
    // checklcpsubtab compares exact values, bese.c:454-528)
    if (l >= 255)
    {
      atomicAdd(&scal[1], 1ull);
      u64 lo = 0, hi = nllv;
      while (lo < hi)
      {
        const u64 mid = (lo + hi) >> 1;
        if (llv[2 * mid] < r)
        {
          lo = mid + 1;
        } else
        {
          hi = mid;
        }
      }
      if (lo >= nllv || llv[2 * lo] != r || llv[2 * lo + 1] != l)
      {
        atomicAdd(&errors[6], 1u);
      }
    }
    if (l > *(volatile unsigned long long *) &scal[0])
    {
      atomicMax(&scal[0], (unsigned long long) l);
    }
  }
}

// 6 (continued): .llv indices strictly increase and every value is >= 255
__global__ void check_llv_kernel(const u64 *__restrict__ llv, u64 nllv, u64 N, u32 *__restrict__ errors)
{
  const u64 j = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nllv)
  {
    return;
  }
  if (llv[2 * j] >= N || llv[2 * j + 1] < 255 || (j > 0 && llv[2 * (j - 1)] >= llv[2 * j]))
  {
    atomicAdd(&errors[6], 1u);
  }
}

// 7: the bucket table against the suffix array (makebcktab, mkvprocess.c:251-327): left is monotone from 0,
// left <= mid <= next left (N behind the last code), ...
__global__ void check_bck_codes_kernel(u64 numofcodes, const u64 *__restrict__ bck, u64 N, u32 *__restrict__ errors)
{
  const u64 c = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= numofcodes)
  {
    return;
  }
  const u64 left = bck[2 * c], mid = bck[2 * c + 1], right = c + 1 < numofcodes ? bck[2 * (c + 1)] : N;
  if ((c == 0 && left != 0) || left > mid || mid > right)
  {
    atomicAdd(&errors[7], 1u);
  }
}
// ... and every suffix lies inside the bucket of its padded k-mer code (ppsort.c:134-153: the code is padded with the
// largest letter behind a special), in front of mid exactly when no special lies inside its k-prefix
__global__ void check_bck_suffix_kernel(const u8 *__restrict__ tis, u64 n, const u32 *__restrict__ sa, u32 sigma, u32 k,
                                        u64 numofcodes, const u64 *__restrict__ bck, u32 *__restrict__ errors)
{
  const u64 N = n + 1;
  const u64 r = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N)
  {
    return;
  }
  const u64 s = sa[r];
  u64 code = 0;
  bool full = true;
  for (u32 i = 0; i < k; i++)
  {
    const u32 a = (full && s + i < n) ? tis[s + i] : 255u;
    if (a >= ESA_WILDCARD)
    {
      full = false;
    }
    code = code * sigma + (full ? a : sigma - 1);
  }
  const u64 left = bck[2 * code], mid = bck[2 * code + 1], right = code + 1 < numofcodes ? bck[2 * (code + 1)] : N;
  if (r < left || r >= right || (full != (r < mid)))
  {
    atomicAdd(&errors[7], 1u);
  }
}

// 5: bck.left + sti1 == isa (saturated) inside [left, mid) (checkstitab1, mkvprocess.c:672-763)
__global__ void check_sti1_kernel(u64 numofcodes, const u64 *__restrict__ bck, const u32 *__restrict__ sa,
                                  const u8 *__restrict__ sti1, u64 N, u32 *__restrict__ errors)
{
  const u64 c = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= numofcodes)
  {
    return;
  }
  const u64 left = bck[2 * c], mid = bck[2 * c + 1];
  if (left > mid || mid > N)
  {
    atomicAdd(&errors[5], 1u);
    return;
  }
  for (u64 i = left; i < mid; i++)
  {
    const u64 off = i - left;
    if (sti1[sa[i]] != (u8) (off < 255 ? off : 255))
    {
      atomicAdd(&errors[5], 1u);
      return;
    }
  }
}

}  // namespace

void finalize_tables(esa_ctx *ctx, u32 demand, const u32 *d_v, u64 count, u64 index_base)
{
  // count entries of ctx->d_sa (the whole table on one GPU, this rank's range otherwise); d_v is the complete
  // text-order plcp array in both cases
  const u64 N = count;
  cudaStream_t s = ctx->stream;
  Arena &ar = ctx->arena;
  const bool want_lcp = d_v != nullptr;
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
  if (want_lcp)
  {
    ctx->d_lcp = ar.alloc<u8>(N + 4);
  }
  const u32 tiles = (u32) div_up_u64(N, FN_TILE);
  u32 *tile_large = ar.alloc<u32>(tiles);
  const double bytes = (double) N * (4.0 + ((demand & ESA_SUFTAB) ? 8.0 : 0.0) + ((demand & ESA_BWTTAB) ? 2.0 : 0.0)
                                     + (want_lcp ? 5.0 : 0.0));
  ESA_LAUNCH(ctx, ESA_K_FINALIZE, bytes,
             (finalize_kernel<<<tiles, FN_THREADS, 0, s>>>(ctx->d_tis, N, ctx->d_sa, d_v, ctx->d_lcp, ctx->d_bwt,
                                                            ctx->d_suf64, tile_large, d_scalars, index_base)));
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
    ESA_LAUNCH(ctx, ESA_K_FINALIZE, (double) N + 24.0 * ctx->nllv,
               (llv_emit_kernel<<<tiles, FN_THREADS, 0, s>>>(N, ctx->d_sa, d_v, nullptr, ctx->d_lcp, tile_large,
                                                              ctx->nllv, ctx->d_llv, index_base)));
  }
}

namespace
{
__global__ void readback_kernel(u64 *__restrict__ host_dst, const u64 *__restrict__ src, u32 count)
{
  if (threadIdx.x < count)
  {
    host_dst[threadIdx.x] = src[threadIdx.x];
  }
  __threadfence_system();
}
__global__ void readback32_kernel(u64 *__restrict__ host_dst, const u32 *__restrict__ src)
{
  host_dst[0] = src[0];
  __threadfence_system();
}
}  // namespace

void read_back_u64(esa_ctx *ctx, const void *d_src, u32 count, u32 h_offset)
{
  if (count == 0 || count + h_offset > 64)
  {
    esa_throw(ESA_ERR_ARGUMENT, "internal: read-back of %u words at %u", count, h_offset);
  }
  readback_kernel<<<1, 64, 0, ctx->stream>>>(ctx->h_scalars + h_offset, (const u64 *) d_src, count);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

u32 read_back_u32(esa_ctx *ctx, const u32 *d_src)
{
  readback32_kernel<<<1, 1, 0, ctx->stream>>>(ctx->h_scalars + 63, d_src);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return (u32) ctx->h_scalars[63];
}

void compute_sti1(esa_ctx *ctx, u32 prefixlength)
{
  const u64 N = ctx->n + 1;
  ensure_isa(ctx);
  ctx->d_sti1 = ctx->arena.alloc<u8>(N + KS_TILE);
  const KmerParams kp = make_kmer_params(ctx->sigma, prefixlength);
  const u32 tiles = (u32) div_up_u64(N, KS_TILE);
  ESA_LAUNCH(ctx, ESA_K_STI1, 14.0 * N,
             (sti1_kernel<<<tiles, KS_THREADS, 0, ctx->stream>>>(ctx->d_tis, ctx->n, kp, ctx->d_isa, ctx->d_bck,
                                                                  ctx->d_sti1)));
}

// ---- sharded builds: the tables that are indexed by TEXT position (.sti1, .sti) ----
// A rank knows the values of the suffixes of its suffix-array range, which lie anywhere in the text: every rank writes
// them into a zeroed table of the whole text, and a reduce-scatter (sum: every position is written by exactly one rank)
// leaves rank g with the text slice [g * slice, (g + 1) * slice) -- concatenated in rank order, the reference's files.
// makestitab1 (src/Mkvtree/mkvprocess.c:583-643) per suffix-array entry: 0 when a special lies inside the k-prefix,
// else min(255, index - left border of the k-mer's bucket); full inverse: mksti.c:15-34.
namespace
{
template <typename T>
__global__ void sti_sharded_kernel(const u8 *__restrict__ tis, const u32 *__restrict__ sa, u64 count, u64 index_base,
                                   u32 sigma, u32 k, const u64 *__restrict__ bck, T *__restrict__ full)
{
  const u64 r = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= count)
  {
    return;
  }
  const u64 s = sa[r], R = index_base + r;
  if (sizeof(T) == 4)
  {
    full[s] = (T) R;  // inverse suffix array
    return;
  }
  u64 code = 0;
  for (u32 i = 0; i < k; i++)
  {
    const u32 a = tis[s + i];  // padded with 255 behind position n
    if (a >= ESA_WILDCARD)
    {
      return;  // special inside the prefix: 0
    }
    code = code * sigma + a;
  }
  const u64 off = R - bck[2 * code];
  full[s] = (T) (off < 255 ? off : 255);
}
}  // namespace

void compute_sti_sharded(esa_ctx *ctx, esa_comm *comm, u32 prefixlength, u32 demand)
{
  const u64 N = ctx->n + 1;
  const int G = comm->nranks, me = comm->rank;
  cudaStream_t s = ctx->stream;
  Arena &ar = ctx->arena;
  const u64 slice = div_up_u64(div_up_u64(N, (u64) G), 64) * 64;  // text positions per rank
  ctx->text_first = (u64) me * slice < N ? (u64) me * slice : N;
  ctx->text_count = ((u64) (me + 1) * slice < N ? (u64) (me + 1) * slice : N) - ctx->text_first;
  const u32 grid = (u32) div_up_u64(ctx->shard_count > 0 ? ctx->shard_count : 1, 256);
  if (demand & ESA_STI1TAB)
  {
    ctx->d_sti1 = ar.alloc<u8>(slice + 64);
    const size_t mark = ar.mark();
    u8 *full = ar.alloc<u8>(slice * G + 64);
    ESA_LAUNCH(ctx, ESA_K_MEMOP, (double) slice * G, CUDA_CHECK(cudaMemsetAsync(full, 0, slice * G, s)));
    if (ctx->shard_count > 0)
    {
      ESA_LAUNCH(ctx, ESA_K_STI1, 14.0 * ctx->shard_count,
                 (sti_sharded_kernel<u8><<<grid, 256, 0, s>>>(ctx->d_tis, ctx->d_sa, ctx->shard_count, ctx->shard_first,
                                                               ctx->sigma, prefixlength, ctx->d_bck, full)));
    }
    comm_reduce_scatter_sum_u64(ctx, comm, (const u64 *) full, (u64 *) ctx->d_sti1, slice / 8);
    CUDA_CHECK(cudaStreamSynchronize(s));
    ar.rewind(mark);
  }
  if (demand & ESA_STITAB)
  {
    ctx->d_sti64 = ar.alloc<u64>(slice + 8);
    const size_t mark = ar.mark();
    u32 *full = ar.alloc<u32>(slice * G + 16), *mine = ar.alloc<u32>(slice + 16);
    ESA_LAUNCH(ctx, ESA_K_MEMOP, 4.0 * slice * G, CUDA_CHECK(cudaMemsetAsync(full, 0, 4 * slice * G, s)));
    if (ctx->shard_count > 0)
    {
      ESA_LAUNCH(ctx, ESA_K_FINALIZE, 8.0 * ctx->shard_count,
                 (sti_sharded_kernel<u32><<<grid, 256, 0, s>>>(ctx->d_tis, ctx->d_sa, ctx->shard_count, ctx->shard_first,
                                                                ctx->sigma, prefixlength, ctx->d_bck, full)));
    }
    comm_reduce_scatter_sum_u64(ctx, comm, (const u64 *) full, (u64 *) mine, slice / 2);
    ESA_LAUNCH(ctx, ESA_K_FINALIZE, 12.0 * slice,
               (widen_kernel<<<div_up_u32(slice, 256), 256, 0, s>>>(mine, ctx->d_sti64, slice)));
    CUDA_CHECK(cudaStreamSynchronize(s));
    ar.rewind(mark);
  }
}

void widen_isa(esa_ctx *ctx)
{
  const u64 N = ctx->n + 1;
  ensure_isa(ctx);
  ctx->d_sti64 = ctx->arena.alloc<u64>(N);
  ESA_LAUNCH(ctx, ESA_K_FINALIZE, 12.0 * N,
             (widen_kernel<<<div_up_u32(N, 256), 256, 0, ctx->stream>>>(ctx->d_isa, ctx->d_sti64, N)));
}

int run_checks(esa_ctx *ctx, std::string &msg)
{
  const u64 N = ctx->n + 1;
  cudaStream_t s = ctx->stream;
  ensure_isa(ctx);
  const size_t mark = ctx->arena.mark();
  u32 *d_err = ctx->arena.alloc<u32>(8);
  unsigned long long *d_scal = (unsigned long long *) ctx->arena.alloc<u64>(4);
  CUDA_CHECK(cudaMemsetAsync(d_err, 0, 8 * sizeof(u32), s));
  CUDA_CHECK(cudaMemsetAsync(d_scal, 0, 4 * sizeof(u64), s));
  const bool lcp_resident = ctx->d_lcp != nullptr && (ctx->demand & ESA_LCPTAB);
  check_kernel<<<div_up_u32(N, 256), 256, 0, s>>>(ctx->d_tis, ctx->n, ctx->d_sa, ctx->d_isa, ctx->d_plcpv,
                                                   lcp_resident ? ctx->d_lcp : nullptr, ctx->d_bwt, ctx->d_suf64,
                                                   ctx->d_llv, ctx->d_llv != nullptr ? ctx->nllv : 0, d_err, d_scal);
  CUDA_CHECK(cudaGetLastError());
  if (lcp_resident && ctx->d_llv != nullptr && ctx->nllv > 0)
  {
    check_llv_kernel<<<div_up_u32(ctx->nllv, 256), 256, 0, s>>>(ctx->d_llv, ctx->nllv, N, d_err);
    CUDA_CHECK(cudaGetLastError());
  }
  if (ctx->d_bck != nullptr && ctx->prefixlength > 0)
  {
    check_bck_codes_kernel<<<div_up_u32(ctx->numofcodes, 256), 256, 0, s>>>(ctx->numofcodes, ctx->d_bck, N, d_err);
    CUDA_CHECK(cudaGetLastError());
    check_bck_suffix_kernel<<<div_up_u32(N, 256), 256, 0, s>>>(ctx->d_tis, ctx->n, ctx->d_sa, ctx->sigma,
                                                                ctx->prefixlength, ctx->numofcodes, ctx->d_bck, d_err);
    CUDA_CHECK(cudaGetLastError());
  }
  if (ctx->d_sti1 != nullptr && ctx->d_bck != nullptr)
  {
    check_sti1_kernel<<<div_up_u32(ctx->numofcodes, 256), 256, 0, s>>>(ctx->numofcodes, ctx->d_bck, ctx->d_sa,
                                                                        ctx->d_sti1, N, d_err);
    CUDA_CHECK(cudaGetLastError());
  }
  u32 h_err[8];
  u64 h_scal[4];
  CUDA_CHECK(cudaMemcpyAsync(h_err, d_err, sizeof h_err, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaMemcpyAsync(h_scal, d_scal, sizeof h_scal, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  ctx->arena.rewind(mark);
  if (lcp_resident)
  {
    // the scalars the .prj file reports (maxbranchdepth, largelcpvalues) against the recomputed lcp values
    if (h_scal[0] != ctx->maxbranchdepth || h_scal[1] != ctx->nllv)
    {
      h_err[6]++;
    }
  }
  static const char *names[8] = {"", "suftab is not a permutation / isa does not invert it",
                                 "adjacent suffixes out of order", "lcp value wrong", "bwt does not match the text",
                                 "bck.left + sti1 != inverse suftab",
                                 "exact lcp values: llv pairs / largelcpvalues / maxbranchdepth wrong",
                                 "bck does not delimit the buckets of the suffix array"};
  for (int i = 1; i <= 7; i++)
  {
    if (h_err[i] != 0)
    {
      char buf[256];
      snprintf(buf, sizeof buf, "check %d failed (%s): %u violations", i, names[i], h_err[i]);
      msg = buf;
      return i;
    }
  }
  return 0;
}
