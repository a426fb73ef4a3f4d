// common.cuh -- context, device arena, launch accounting and error plumbing shared by all translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/vstree_b200.h"

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;

#define ESA_NUM_SMS_FALLBACK 148

// ---------------------------------------------------------------------------------------------------------
// Errors: C++ exceptions are used inside the library only and are caught at the extern "C" boundary.
struct EsaError
{
  int code;
  std::string msg;
};

[[noreturn]] inline void esa_throw(int code, const char *fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw EsaError{code, std::string(buf)};
}

#define CUDA_CHECK(call)                                                                                    \
  do                                                                                                        \
  {                                                                                                         \
    cudaError_t err__ = (call);                                                                             \
    if (err__ != cudaSuccess)                                                                               \
    {                                                                                                       \
      esa_throw(err__ == cudaErrorMemoryAllocation ? ESA_ERR_MEMORY : ESA_ERR_CUDA, "%s failed at %s:%d: %s", \
                #call, __FILE__, __LINE__, cudaGetErrorString(err__));                                      \
    }                                                                                                       \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// Device arena: one big allocation per context, carved by a bump pointer.  Builds reuse it, so the timed
// region of a steady-state build contains no cudaMalloc.
struct Arena
{
  char *base = nullptr;
  size_t capacity = 0;
  size_t top = 0;
  size_t peak = 0;

  void reserve(size_t bytes)
  {
    if (bytes <= capacity)
    {
      return;
    }
    release();
    bytes = (bytes + 255) & ~(size_t) 255;
    CUDA_CHECK(cudaMalloc((void **) &base, bytes));
    capacity = bytes;
    high = capacity;
  }
  void release()
  {
    if (base != nullptr)
    {
      cudaFree(base);
    }
    base = nullptr;
    capacity = top = high = 0;
  }
  void reset()
  {
    top = 0;
    high = capacity;
  }
  // second bump pointer from the high end: tables whose size is only known in the middle of a build (a shard of
  // the suffix array after the exchange) but that must survive the rewinds of the scratch region
  size_t high = 0;
  template <typename T> T *alloc_high(size_t count)
  {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t) 255;
    if (high < bytes || high - bytes < top)
    {
      esa_throw(ESA_ERR_MEMORY, "device arena exhausted (persistent side): need %zu bytes, %zu free", bytes,
                high > top ? high - top : (size_t) 0);
    }
    high -= bytes;
    if (capacity - high + top > peak)
    {
      peak = capacity - high + top;
    }
    return (T *) (base + high);
  }
  size_t mark() const { return top; }
  void rewind(size_t m) { top = m; }
  template <typename T> T *alloc(size_t count)
  {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t) 255;
    if (top + bytes > high)
    {
      esa_throw(ESA_ERR_MEMORY, "device arena exhausted: need %zu more bytes, capacity %zu", top + bytes - high,
                capacity);
    }
    T *p = (T *) (base + top);
    top += bytes;
    if (top + (capacity - high) > peak)
    {
      peak = top + (capacity - high);
    }
    return p;
  }
};

// ---------------------------------------------------------------------------------------------------------
// Per-family kernel timing with CUDA events on the launching stream.
struct KernelProfiler
{
  struct Span
  {
    int kind;
    cudaEvent_t a, b;
    double bytes;
  };
  std::vector<cudaEvent_t> pool;
  std::vector<Span> spans;
  size_t next = 0;
  bool enabled = true;
  u64 launches[ESA_K_NUM] = {0};

  cudaEvent_t get()
  {
    if (next == pool.size())
    {
      cudaEvent_t e;
      CUDA_CHECK(cudaEventCreate(&e));
      pool.push_back(e);
    }
    return pool[next++];
  }
  void reset()
  {
    next = 0;
    spans.clear();
    memset(launches, 0, sizeof launches);
  }
  void destroy()
  {
    for (cudaEvent_t e : pool)
    {
      cudaEventDestroy(e);
    }
    pool.clear();
  }
};

// Depth schedule of the refinement rounds after the initial h0-mer sort: rounds 1..text_rounds extend the order by
// text_len symbols read from the text itself (no rank array needed), later rounds double.  Round t starts from
// groups that agree on round_start_depth(t) symbols (the lcp lower bound of a slot split off in round t).
struct DepthSched
{
  u32 h0 = 0, text_len = 0, text_rounds = 0;
};
__host__ __device__ __forceinline__ u32 round_start_depth(const DepthSched &ds, u32 round /* >= 1 */)
{
  const u32 t = round - 1;
  if (t <= ds.text_rounds)
  {
    return ds.h0 + t * ds.text_len;
  }
  const u32 sh = t - ds.text_rounds;
  const u64 d = (u64) (ds.h0 + ds.text_rounds * ds.text_len) << (sh < 32 ? sh : 32);
  return d > 0xffffffffull ? 0xffffffffu : (u32) d;
}

struct esa_comm;
struct esa_ctx
{
  int device = -1;
  int num_sms = ESA_NUM_SMS_FALLBACK;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;       // compute stream: every kernel of a build is launched here
  cudaStream_t copy_stream = nullptr;  // D2H overlap
  Arena arena;
  KernelProfiler prof;

  // resident input
  u8 *d_tis = nullptr;   // arena; n symbols followed by >= 64 bytes of 255
  u64 *d_pk = nullptr;   // 2-bit packed text (sigma <= 4 only), see textcmp.cuh
  u64 *d_pk5 = nullptr;  // 5-bit packed text (5 <= sigma <= 30: protein), see textcmp.cuh
  u32 *d_spm = nullptr;  // special mask of the packed text
  u8 *d_hround = nullptr;  // [N] doubling round in which slot r became a group head (lcp lower bound), 0 = initial sort
  DepthSched sched;        // depth schedule of the last sort (turns d_hround entries into lcp lower bounds)
  bool isa_valid = false;  // d_isa holds the inverse suffix array (a sort that ends inside the text rounds never
                           // scatters ranks; consumers call ensure_isa)
  u64 lcp_work_estimate = 0;  // sum over rounds of active * depth / 32 (window compares of the direct lcp path)
  int lcp_mode = 0;           // 0 auto, 1 direct (suffix-array order), 2 plcp (text order)
  u64 n = 0;
  u32 sigma = 0;
  bool uploaded = false;

  // resident outputs of the last esa_construct (device pointers into the arena)
  bool constructed = false;
  u32 demand = 0, prefixlength = 0;
  u64 numofcodes = 0;
  u32 *d_sa = nullptr;     // [N] suffix array, 32-bit
  u32 *d_isa = nullptr;    // [N] inverse suffix array
  u64 *d_suf64 = nullptr;  // [N]
  u8 *d_lcp = nullptr;     // [N]
  u8 *d_bwt = nullptr;     // [N]
  u8 *d_sti1 = nullptr;    // [N]
  u64 *d_bck = nullptr;    // [2*numofcodes]
  u64 *d_llv = nullptr;    // [2*nllv]
  u64 *d_sti64 = nullptr;  // [N]
  u64 *d_skp = nullptr;    // [N]
  u32 *d_plcpv = nullptr;  // [N] V[i] = i + PLCP[i]
  u32 *d_deepv = nullptr;  // [N] hybrid lcp: V[i] for the deep positions only (undefined elsewhere)
  u32 *d_lcpfull = nullptr;  // direct lcp path: exact values of the entries whose lcp8 is 255
  u64 nllv = 0, maxbranchdepth = 0, longest = 0;
  // sharded build: this rank holds suffix-array entries [shard_first, shard_first + shard_count) in d_sa, d_suf64,
  // d_lcp, d_bwt, d_hround (all indexed from 0); one GPU: [0, n+1)
  u64 shard_first = 0, shard_count = 0;
  // tables indexed by TEXT position (sti1, sti): the whole text on one GPU, this rank's slice in a sharded build
  u64 text_first = 0, text_count = 0;
  size_t arena_after_upload = 0;
  struct esa_comm *comm = nullptr;  // attached communicator (sharded builds), see comm.cuh

  // small pinned scratch for scalar read-backs
  u64 *h_scalars = nullptr;  // 64 words
};

// ---------------------------------------------------------------------------------------------------------
// Launch bracket: counts the launch, optionally wraps it in an event pair, checks the launch error.
struct LaunchScope
{
  esa_ctx *ctx;
  int kind;
  cudaEvent_t a = nullptr, b = nullptr;
  double bytes;
  LaunchScope(esa_ctx *c, int k, double algorithmic_bytes) : ctx(c), kind(k), bytes(algorithmic_bytes)
  {
    ctx->prof.launches[kind]++;
    if (ctx->prof.enabled)
    {
      a = ctx->prof.get();
      b = ctx->prof.get();
      CUDA_CHECK(cudaEventRecord(a, ctx->stream));
    }
  }
  void done()
  {
    CUDA_CHECK(cudaGetLastError());
    if (ctx->prof.enabled)
    {
      CUDA_CHECK(cudaEventRecord(b, ctx->stream));
      ctx->prof.spans.push_back({kind, a, b, bytes});
    }
  }
};

#define ESA_LAUNCH(ctx, kind, bytes, ...)            \
  do                                                 \
  {                                                  \
    LaunchScope scope__((ctx), (kind), (double) (bytes)); \
    __VA_ARGS__;                                     \
    scope__.done();                                  \
  } while (0)

// Small device -> host read-backs (counts, maxima) go through a one-warp kernel that stores into the MAPPED pinned
// words ctx->h_scalars, not through the copy engine: a 24-byte cudaMemcpyAsync queues behind the hundreds of megabytes
// the fused download of esa_build has in flight on that engine and stalled the lcp/bwt sweep for 8 ms.
// Copies count 64-bit words from d_src to ctx->h_scalars + h_offset and synchronises the stream.
void read_back_u64(esa_ctx *ctx, const void *d_src, u32 count, u32 h_offset = 0);
u32 read_back_u32(esa_ctx *ctx, const u32 *d_src);  // one 32-bit word, same path

static inline u32 div_up_u32(u64 a, u64 b) { return (u32) ((a + b - 1) / b); }
static inline u64 div_up_u64(u64 a, u64 b) { return (a + b - 1) / b; }

// number of bits needed to represent values in [0, v]
static inline u32 bits_for(u64 v)
{
  u32 b = 0;
  while (v > 0)
  {
    b++;
    v >>= 1;
  }
  return b == 0 ? 1 : b;
}

// ---------------------------------------------------------------------------------------------------------
// stage entry points (one per translation unit)

// radix_sort.cu
struct RadixSortTemp
{
  u32 *hist;     // [parts][passes][256]
  u32 *base;     // [parts][passes][256]
  u32 *status;   // [tiles_per_part_max][256]
  u32 *counter;  // [passes*parts] dynamic tile tickets
  u64 n_capacity;
};
size_t radix_sort_temp_bytes(u64 n, u32 key_bytes);
// Sorts (keys, vals) by key bits [begin_bit, end_bit).  Buffers ping-pong; returns 0 when the result is in
// (keys_a, vals_a) and 1 when it is in (keys_b, vals_b).  vals may be NULL (keys only).
int radix_sort_pairs_u32(esa_ctx *ctx, u32 *keys_a, u32 *keys_b, u32 *vals_a, u32 *vals_b, u64 n, u32 begin_bit,
                         u32 end_bit, void *temp, size_t temp_bytes);
int radix_sort_pairs_u64(esa_ctx *ctx, u64 *keys_a, u64 *keys_b, u32 *vals_a, u32 *vals_b, u64 n, u32 begin_bit,
                         u32 end_bit, void *temp, size_t temp_bytes);

// initial sort of a DNA text with the keys of pass 1 taken from the packed text (radix_sort.cu); same return convention
int radix_sort_dna_initial(esa_ctx *ctx, const u64 *d_pk, const u32 *d_spm, const u8 *d_tis, u64 N, u32 h, u64 nfull,
                           const u32 *side_key, const u32 *side_pos, u64 nside, u32 *keys_a, u32 *keys_b, u32 *vals_a,
                           u32 *vals_b, u32 key_bits, void *temp, size_t temp_bytes);

// the partition step of the sharded build fused with its all-to-all (stores into peer memory); see radix_sort.cu
struct RadixRemoteDst
{
  u32 *keys[64];   // receive buffers of every rank, as mapped into this rank's address space
  u32 *vals[64];
  u8 owner[256];   // digit -> rank
  u32 key_sub[64]; // the key stored at owner r is (key - key_sub[r]) << key_shl (identity: 0 and 0)
  u32 key_shl;
};
void radix_partition_remote_u32(esa_ctx *ctx, const u32 *keys, const u32 *vals, u64 n, u32 begin_bit, u32 end_bit,
                                const RadixRemoteDst *d_remote, const u32 *d_start, void *temp, size_t temp_bytes);

// the same two steps for a DNA slice with the keys taken from the packed text (radix_sort.cu)
void radix_bin_hist_dna(esa_ctx *ctx, const u64 *d_pk, const u32 *d_spm, const u8 *d_tis, u64 p_begin, u64 p_end, u32 h,
                        u32 key_shl, u32 bin_shift, u32 *d_hist256);
void radix_partition_remote_dna(esa_ctx *ctx, const u64 *d_pk, const u32 *d_spm, const u8 *d_tis, u64 p_begin, u64 p_end,
                                u32 h, u32 key_shl, u64 nfull, u32 bin_shift, const RadixRemoteDst *d_remote,
                                const u32 *d_start, void *temp, size_t temp_bytes);

// doubling_sort.cu: the sort of one refinement round (list grouped by key >> shift; sorts by key bits [0, end_bit),
// end_bit > shift); same return convention as radix_sort_pairs_u64
int doubling_sort(esa_ctx *ctx, u64 *keys_a, u64 *keys_b, u32 *pos_a, u32 *pos_b, u64 m, u32 shift, u32 end_bit,
                  void *sort_temp, size_t temp_bytes);

// suffix_sort.cu: fills ctx->d_sa, ctx->d_isa (and ctx->d_bck when demanded) ; returns stats pieces
struct SortOutcome
{
  u32 h0;
  u32 rounds;
  u64 active_sum, active_first;
  u64 num_sorted;   // suffixes that went through the sort (N - specials - 1)
};
void suffix_sort(esa_ctx *ctx, u32 prefixlength, bool want_bck, SortOutcome *outcome, size_t scratch_mark);
// d_isa[d_sa[r]] = r for the whole table unless the sort left it valid already (single-GPU contexts only)
void ensure_isa(esa_ctx *ctx);
// sharded variant (one rank of a communicator); sets ctx->d_sa / d_hround (shard-local, arena high side),
// ctx->shard_first / shard_count, the replicated ctx->d_isa and ctx->d_bck
void suffix_sort_sharded(esa_ctx *ctx, esa_comm *comm, u32 prefixlength, bool want_bck, bool want_hround,
                         SortOutcome *outcome);

// lcp.cu: V[i] = i + PLCP[i] for all text positions (ctx->d_plcpv)
void compute_plcp(esa_ctx *ctx, u32 *d_v);
// hybrid lcp: V[i] only for the positions whose slot was split off in a refinement round >= deep_round (comm may be null)
void compute_plcp_deep(esa_ctx *ctx, esa_comm *comm, u32 *d_v, u32 halo_prev, u32 deep_round);

// derive.cu
void pack_text(esa_ctx *ctx);  // fills ctx->d_pk / d_spm from d_tis (sigma <= 4)
void pack_text5(esa_ctx *ctx);  // fills ctx->d_pk5 from d_tis (5 <= sigma <= 30)
void finalize_tables(esa_ctx *ctx, u32 demand, const u32 *d_v, u64 count, u64 index_base);
// lcp_direct.cu; deep_v / deep_round: slots split off in a round >= deep_round take their lcp from deep_v (hybrid)
void finalize_tables_direct(esa_ctx *ctx, u32 demand, u64 count, u64 index_base, u32 halo_prev, const u32 *deep_v = nullptr,
                            u32 deep_round = 255);
void compute_sti1(esa_ctx *ctx, u32 prefixlength);
void compute_sti_sharded(esa_ctx *ctx, esa_comm *comm, u32 prefixlength, u32 demand);  // .sti1 / .sti of a sharded build
void compute_skp(esa_ctx *ctx, const u32 *d_v, const u32 *d_lcpfull);  // skp.cu
// childtab.cu: the child table from lcp bytes + large-value pairs on the device into the caller's host buffers
void compute_childtab(esa_ctx *ctx, const u8 *d_lcp8, const u64 *d_llv, u64 nllv, u64 n, u8 *h_cld, u8 *h_cld1);
int run_checks(esa_ctx *ctx, std::string &msg);
