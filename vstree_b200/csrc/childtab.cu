// childtab.cu -- the child table of the enhanced suffix array (.cld, 3 bytes per entry, and .cld1, 1 byte per entry)
// from the lcp table: SURVEY.md 8f row 4.  Replaces makecldtabandoutput (reference: src/Mkvtree/mkcld.mn.c:13-47) =
// mkcldtabnextlIndex + mkcldtabupdown + compresscldtab (src/kurtz/mkcld.c:41-284).  The per-entry definition and the
// searches are in childtab_core.cuh; this file holds the kernels around them:
//
//   widen_lcp_kernel      lcp bytes -> 32-bit values, count of the saturated ones       1 B read, 4 B written per entry
//   scatter_llv_kernel    the exact values >= 255 from the (index, value) pairs         16 B per large value
//   min_level_kernel      one level of the min-pyramid                                  4 (1 + 1/16) B per entry and level
//   child_near_kernel     up / down / nextlIndex of every entry a bounded scan in a shared-memory window can settle;
//                         the 3-byte records leave as whole words                       4 B read, 3 B written
//   child_far_kernel      the pyramid search for the entries the scan left open         a few 64-byte groups per level
//   cld1_kernel           the one-byte table from the three-field table                 ~ 11 B read, 1 B written
//
// HBM-bound integer work over tables of n + 1 entries; the searches touch neighbours of the entry almost always (the
// pyramid levels above level 0 together are 1/15 of the table and stay in L2).
#include "common.cuh"
#include "childtab_core.cuh"

namespace
{

constexpr u32 CLD_THREADS = 256;
constexpr u32 CLD_NEAR = 32;  // entries on either side that phase 1 of child_kernel looks at

// 32-bit copy of the lcp bytes (the reference's walks start from depth 0 at index 0 whatever lcptab[0] holds,
// mkcld.c:54,113) and the number of saturated bytes: each of them needs exactly one large-value pair
__global__ void widen_lcp_kernel(u64 N, const u8 *__restrict__ lcp8, u32 *__restrict__ out,
                                 unsigned long long *__restrict__ saturated)
{
  const u64 r = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  u32 v = 0;
  if (r < N)
  {
    v = r == 0 ? 0u : (u32) lcp8[r];
    out[r] = v;
  }
  const u32 m = __ballot_sync(0xffffffffu, v == 255u);
  if ((threadIdx.x & 31) == 0 && m != 0)
  {
    atomicAdd(saturated, (unsigned long long) __popc(m));
  }
}

// pairs are (index, value) of the entries whose byte is 255, increasing index (virtualdef.h:203, bese.c:553-566)
__global__ void scatter_llv_kernel(u64 nllv, const u64 *__restrict__ llv, u64 N, u32 *__restrict__ out,
                                   u32 *__restrict__ bad)
{
  const u64 t = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nllv)
  {
    const u64 idx = llv[2 * t], val = llv[2 * t + 1];
    if (idx == 0 || idx >= N || val < 255 || val > 0xfffffffeull || out[idx] != 255u)
    {
      atomicAdd(bad, 1u);  // not a large-value list of this lcp table
    } else
    {
      out[idx] = (u32) val;
    }
  }
}

__global__ void min_level_kernel(const u32 *__restrict__ in, u64 in_size, u32 *__restrict__ out, u64 out_size)
{
  const u64 j = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (j < out_size)
  {
    u32 m = 0xffffffffu;
    const u64 b = j * cld::FAN;
#pragma unroll
    for (u32 c = 0; c < cld::FAN; c++)
    {
      if (b + c < in_size)
      {
        m = min(m, in[b + c]);
      }
    }
    out[j] = m;
  }
}

// Two kernels.  child_near_kernel: every thread scans at most CLD_NEAR entries on either side of its entry in a
// shared-memory window (child_entry_near); on real lcp tables that settles more than nine entries in ten.  The entries it
// cannot settle are appended to a queue.  child_far_kernel: the queued entries, spread evenly over a full grid, run the
// search on the pyramid (child_entry) -- so a far-reaching entry neither holds up the 31 other lanes of its warp nor
// costs the common case its registers.
__global__ void __launch_bounds__(CLD_THREADS) child_near_kernel(cld::Pyramid py, u8 *__restrict__ out,
                                                                 u32 *__restrict__ far_queue, u32 *__restrict__ far_count,
                                                                 bool use_near)
{
  __shared__ __align__(16) u8 rec[CLD_THREADS * 3];
  __shared__ u32 win[CLD_THREADS + 2 * CLD_NEAR];
  __shared__ u32 warp_open[CLD_THREADS / 32];
  __shared__ u32 queue_base;
  const u64 N = py.size[0];
  const u64 first = (u64) blockIdx.x * CLD_THREADS;
  const u64 w0 = first >= CLD_NEAR ? first - CLD_NEAR : 0;
  const u64 wend = min(N, first + CLD_THREADS + CLD_NEAR);
  const u32 wn = (u32) (wend - w0);
  for (u32 t = threadIdx.x; t < wn; t += CLD_THREADS)
  {
    win[t] = py.level[0][w0 + t];
  }
  __syncthreads();
  const u64 i = first + threadIdx.x;
  cld::Child c;
  c.up = c.down = c.nextl = 0;
  bool open = false;  // left to child_far_kernel
  if (i < N)
  {
#ifdef CLD_NEAR32
    open = !(use_near && cld::child_entry_near32(win, wn, wend == N, (u32) (i - w0), CLD_NEAR, c));
#else
    open = !(use_near && cld::child_entry_near(win, w0, wn, N, i, CLD_NEAR, c));
#endif
    if (open)
    {
      c.up = c.down = c.nextl = 0;
    }
  }
  rec[3 * threadIdx.x] = c.up;
  rec[3 * threadIdx.x + 1] = c.down;
  rec[3 * threadIdx.x + 2] = c.nextl;
  // the open entries of the block go to the queue in index order; every step of the compaction runs with the whole
  // warp converged (ballot, one count per warp, one atomic per block)
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 ballot = __ballot_sync(0xffffffffu, open);
  if (lane == 0)
  {
    warp_open[warp] = __popc(ballot);
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    u32 total = 0;
    for (u32 w = 0; w < CLD_THREADS / 32; w++)
    {
      const u32 count = warp_open[w];
      warp_open[w] = total;
      total += count;
    }
    queue_base = total > 0 ? atomicAdd(far_count, total) : 0;
  }
  __syncthreads();
  if (open)
  {
    far_queue[queue_base + warp_open[warp] + __popc(ballot & ((1u << lane) - 1u))] = (u32) i;
  }
  // first * 3 is a multiple of 4 (CLD_THREADS * 3 = 768): whole words while they are inside the table, bytes at its end
  const u64 base = first * 3, total = N * 3;
  if (threadIdx.x < CLD_THREADS * 3 / 4)
  {
    const u64 o = base + 4ull * threadIdx.x;
    if (o + 4 <= total)
    {
      *reinterpret_cast<u32 *>(out + o) = reinterpret_cast<const u32 *>(rec)[threadIdx.x];
    } else
    {
      for (u64 b = o; b < total; b++)
      {
        out[b] = rec[b - base];
      }
    }
  }
}

__global__ void __launch_bounds__(CLD_THREADS) child_far_kernel(cld::Pyramid py, const u32 *__restrict__ far_queue,
                                                                const u32 *__restrict__ far_count, u8 *__restrict__ out)
{
  const u64 count = *far_count;
  for (u64 t = (u64) blockIdx.x * CLD_THREADS + threadIdx.x; t < count; t += (u64) gridDim.x * CLD_THREADS)
  {
    const u64 e = far_queue[t];
    const cld::Child c = cld::child_entry(py, e);
    out[3 * e] = c.up;
    out[3 * e + 1] = c.down;
    out[3 * e + 2] = c.nextl;
  }
}

__global__ void __launch_bounds__(CLD_THREADS) cld1_kernel(const u8 *__restrict__ cldtab, const u32 *__restrict__ L,
                                                           u64 N, u8 *__restrict__ out)
{
  const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N)
  {
    out[i] = cld::cld1_entry(cldtab, L, i, N - 1);
  }
}

struct DeviceBuf  // plain allocations: the call must not disturb the arena of a resident build
{
  void *p = nullptr;
  DeviceBuf() = default;
  DeviceBuf(const DeviceBuf &) = delete;
  DeviceBuf &operator=(const DeviceBuf &) = delete;
  void alloc(size_t bytes) { CUDA_CHECK(cudaMalloc(&p, bytes + 256)); }
  template <typename T> T *as() const { return static_cast<T *>(p); }
  ~DeviceBuf()
  {
    if (p != nullptr)
    {
      cudaFree(p);
    }
  }
};

inline u32 blocks_for(u64 count) { return (u32) div_up_u64(count, CLD_THREADS); }

}  // namespace

// d_lcp8 / d_llv: tables on the device (the resident ones of the last build, or uploaded copies of an index on file);
// h_cld (3 N bytes) and h_cld1 (N bytes, may be null) are the caller's host buffers.
void compute_childtab(esa_ctx *ctx, const u8 *d_lcp8, const u64 *d_llv, u64 nllv, u64 n, u8 *h_cld, u8 *h_cld1)
{
  const u64 N = n + 1;
  cudaStream_t s = ctx->stream;
  DeviceBuf levels, cldtab, cld1, flags, far;
  // level sizes: N, ceil(N / 16), ... until at most 16 entries
  u64 sizes[cld::MAX_LEVELS], offs[cld::MAX_LEVELS], total = 0;
  int top = 0;
  sizes[0] = N;
  while (sizes[top] > cld::FAN && top + 1 < cld::MAX_LEVELS)
  {
    sizes[top + 1] = div_up_u64(sizes[top], cld::FAN);
    top++;
  }
  for (int k = 0; k <= top; k++)
  {
    offs[k] = total;
    total += (sizes[k] + 63) & ~(u64) 63;
  }
  levels.alloc(total * sizeof(u32));
  cldtab.alloc(3 * N + 16);
  far.alloc(N * sizeof(u32));  // worst case: the bounded scan settles nothing
  flags.alloc(64);
  CUDA_CHECK(cudaMemsetAsync(flags.p, 0, 64, s));
  u32 *L = levels.as<u32>();
  // ESA_CHILDTAB_TIMING=1: device times of the three stages on stderr (tools/childtab_probe.py)
  const bool timing = getenv("ESA_CHILDTAB_TIMING") != nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  struct EventGuard
  {
    cudaEvent_t *e;
    ~EventGuard()
    {
      for (int i = 0; i < 4; i++)
      {
        if (e[i] != nullptr)
        {
          cudaEventDestroy(e[i]);
        }
      }
    }
  } event_guard{ev};
  if (timing)
  {
    for (int i = 0; i < 4; i++)
    {
      CUDA_CHECK(cudaEventCreate(&ev[i]));
    }
    CUDA_CHECK(cudaEventRecord(ev[0], s));
  }
  widen_lcp_kernel<<<blocks_for(N), CLD_THREADS, 0, s>>>(N, d_lcp8, L, flags.as<unsigned long long>() + 1);
  CUDA_CHECK(cudaGetLastError());
  if (nllv > 0)
  {
    scatter_llv_kernel<<<blocks_for(nllv), CLD_THREADS, 0, s>>>(nllv, d_llv, N, L, flags.as<u32>());
    CUDA_CHECK(cudaGetLastError());
  }
  cld::Pyramid py;
  memset(&py, 0, sizeof py);
  py.top = top;
  for (int k = 0; k <= top; k++)
  {
    py.level[k] = L + offs[k];
    py.size[k] = sizes[k];
  }
  for (int k = 0; k < top; k++)
  {
    min_level_kernel<<<blocks_for(sizes[k + 1]), CLD_THREADS, 0, s>>>(py.level[k], sizes[k], L + offs[k + 1],
                                                                     sizes[k + 1]);
    CUDA_CHECK(cudaGetLastError());
  }
  if (h_cld1 != nullptr)
  {
    cld1.alloc(N);
  }
  if (timing)
  {
    CUDA_CHECK(cudaEventRecord(ev[1], s));
  }
  // ESA_CHILDTAB_NO_NEAR=1 (tests): every entry takes the pyramid search
  const bool use_near = getenv("ESA_CHILDTAB_NO_NEAR") == nullptr;
  child_near_kernel<<<blocks_for(N), CLD_THREADS, 0, s>>>(py, cldtab.as<u8>(), far.as<u32>(), flags.as<u32>() + 8,
                                                          use_near);
  CUDA_CHECK(cudaGetLastError());
  child_far_kernel<<<(u32) ctx->num_sms * 6, CLD_THREADS, 0, s>>>(py, far.as<u32>(), flags.as<u32>() + 8,
                                                                  cldtab.as<u8>());
  CUDA_CHECK(cudaGetLastError());
  if (timing)
  {
    CUDA_CHECK(cudaEventRecord(ev[2], s));
  }
  if (h_cld1 != nullptr)
  {
    cld1_kernel<<<blocks_for(N), CLD_THREADS, 0, s>>>(cldtab.as<u8>(), L, N, cld1.as<u8>());
    CUDA_CHECK(cudaGetLastError());
  }
  if (timing)
  {
    CUDA_CHECK(cudaEventRecord(ev[3], s));
  }
  u64 h_flags[5] = {0, 0, 0, 0, 0};  // [0] misplaced pairs (u32), [1] saturated bytes, [4] queued far entries (u32)
  CUDA_CHECK(cudaMemcpyAsync(h_flags, flags.p, sizeof h_flags, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  if ((u32) h_flags[0] != 0 || h_flags[1] != nllv)
  {
    esa_throw(ESA_ERR_ARGUMENT,
              "the large lcp values do not belong to this lcp table: %llu entries are 255, %llu pairs given, %u of them "
              "out of place",
              (unsigned long long) h_flags[1], (unsigned long long) nllv, (unsigned) (u32) h_flags[0]);
  }
  if (timing)
  {
    float a = 0, b = 0, c = 0;
    CUDA_CHECK(cudaEventElapsedTime(&a, ev[0], ev[1]));
    CUDA_CHECK(cudaEventElapsedTime(&b, ev[1], ev[2]));
    CUDA_CHECK(cudaEventElapsedTime(&c, ev[2], ev[3]));
    fprintf(stderr, "childtab: entries %llu large %llu far %u  widen+pyramid %.3f ms  child kernels %.3f ms  cld1_kernel %.3f ms\n",
            (unsigned long long) N, (unsigned long long) nllv, (unsigned) (u32) h_flags[4], a, b, c);
  }
  CUDA_CHECK(cudaMemcpyAsync(h_cld, cldtab.p, 3 * N, cudaMemcpyDeviceToHost, s));
  if (h_cld1 != nullptr)
  {
    CUDA_CHECK(cudaMemcpyAsync(h_cld1, cld1.p, N, cudaMemcpyDeviceToHost, s));
  }
  CUDA_CHECK(cudaStreamSynchronize(s));
}
