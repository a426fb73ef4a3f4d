// api.cu -- the extern "C" boundary declared in include/vstree_b200.h.
//
// Mirrors the call sequence of the reference's mkvtreeprocess (src/Mkvtree/mkvprocess.c:976-1060):
// suffix sort (+ inline lcp) -> llv -> bck -> longest -> suf -> sti1 -> bwt, but every table is produced on the
// device; the host only moves bytes.
#include <emmintrin.h>
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <unistd.h>
#include <stdlib.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <time.h>

#include "comm.cuh"

void radix_sort_init();
void widen_isa(esa_ctx *ctx);

namespace
{

void set_err(char *errbuf, size_t errlen, const char *msg)
{
  if (errbuf != nullptr && errlen > 0)
  {
    snprintf(errbuf, errlen, "%s", msg);
  }
}

template <typename F> int guarded(char *errbuf, size_t errlen, F &&f)
{
  try
  {
    f();
    return ESA_OK;
  } catch (const EsaError &e)
  {
    set_err(errbuf, errlen, e.msg.c_str());
    return e.code;
  } catch (const std::exception &e)
  {
    set_err(errbuf, errlen, e.what());
    return ESA_ERR_MEMORY;
  }
}

// pinned host blocks handed out in esa_result; recycled between builds so that a steady-state build does not pin
struct PinnedPool
{
  struct Block
  {
    void *p;
    size_t bytes;
    bool used;
  };
  std::vector<Block> blocks;
  // ESA_PAGEABLE_RESULTS=1 (set by one-shot processes such as the mkvtree driver): ordinary page-aligned memory instead
  // of pinned memory.  Pinning the 1.3 GB of tables of a 100 Mbp index costs 0.3 to 3 s of a fresh process (measured
  // on the B200 box, erratic), the staged copy into pageable memory about 0.1 s.  A resident context that builds many
  // indexes keeps the default: pinned once, reused, and required by the chunked download that overlaps the build.
  bool pageable = false;
  void *allocate(size_t bytes)
  {
    void *p = nullptr;
    if (pageable)
    {
      if (posix_memalign(&p, 4096, (bytes + 4095) & ~(size_t) 4095) != 0)
      {
        esa_throw(ESA_ERR_MEMORY, "cannot allocate %zu bytes of host memory", bytes);
      }
    } else
    {
      CUDA_CHECK(cudaMallocHost(&p, bytes));
    }
    return p;
  }
  void release(void *p)
  {
    if (pageable)
    {
      free(p);
    } else
    {
      cudaFreeHost(p);
    }
  }
  void *get(size_t bytes)
  {
    if (bytes == 0)
    {
      bytes = 1;
    }
    Block *best = nullptr;
    for (Block &b : blocks)
    {
      if (!b.used && b.bytes >= bytes && (best == nullptr || b.bytes < best->bytes))
      {
        best = &b;
      }
    }
    if (best != nullptr && best->bytes <= 2 * bytes + 4096)
    {
      best->used = true;
      return best->p;
    }
    void *p = allocate(bytes);
    blocks.push_back({p, bytes, true});
    return p;
  }
  void put(void *p)
  {
    for (Block &b : blocks)
    {
      if (b.p == p)
      {
        b.used = false;
        return;
      }
    }
  }
  void trim()  // drop unused blocks
  {
    std::vector<Block> keep;
    for (Block &b : blocks)
    {
      if (b.used)
      {
        keep.push_back(b);
      } else
      {
        release(b.p);
      }
    }
    blocks.swap(keep);
  }
  void destroy()
  {
    for (Block &b : blocks)
    {
      release(b.p);
    }
    blocks.clear();
  }
};

// ---------------------------------------------------------------------------------------------------------
// esa_build only: the 64-bit suftab is the bulk of the device-to-host traffic (8 of 10 bytes per suffix), and PCIe,
// not HBM, bounds the end-to-end time.  So the one-call path ships the 32-bit suffix array as soon as the sort is
// done (copy stream, in chunks, overlapping the lcp/bwt sweep) and widens it to the reference's 8-byte Uint on host
// cores while later chunks are still in flight.  The staged API (esa_construct/esa_download) keeps the widened table
// on the device.
struct HostWiden
{
  bool active = false;
  u32 *staging = nullptr;  // pinned: the 32-bit entries as they arrive
  u64 *dst = nullptr;      // pinned: the table handed out in esa_result
  u64 count = 0, chunk = 0;
  u32 nchunks = 0;
  int nthreads_last = 0;  // workers of the last fused download (0: the widened table crossed PCIe)
  // Progress of the chunked copy without any CUDA call on the worker side: behind chunk c the copy stream copies the
  // word c+1 from a device table into *progress (pinned).  DMA writes of one stream land in order, so a worker that
  // reads progress > c finds chunk c complete.  (Workers that waited in cudaEventSynchronize contended with the
  // launching thread for the driver's locks: the lcp/bwt sweep of a fused build took 14 ms instead of 7.)
  u32 *progress = nullptr;  // pinned word
  u32 *d_ticks = nullptr;   // device: d_ticks[i] = i + 1
  static constexpr u32 MAX_CHUNKS = 1024;
  // persistent workers: created at the first fused download, parked on cv_job between builds
  std::vector<pthread_t> threads;
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  pthread_cond_t cv_job = PTHREAD_COND_INITIALIZER, cv_done = PTHREAD_COND_INITIALIZER;
  u64 job_id = 0;  // under mu
  bool stop = false;
  // Tickets carry the job they belong to: next_ticket = job_id << 32 | next chunk, taken with compare-and-swap, so a
  // worker that is still leaving the previous job can neither take a chunk of the next one with stale parameters nor
  // add a stray completion (a plain fetch_add did both in a narrow window between two fused builds).
  std::atomic<u64> next_ticket{0};
  std::atomic<u32> chunks_done{0};
  std::atomic<bool> abort{false};
};

// what one job looks like to a worker: copied under mu together with the job id
struct WidenJob
{
  u64 id;
  u32 *staging;
  u64 *dst;
  u64 count, chunk;
  u32 nchunks;
};

void widen_range(const u32 *src, u64 *dst, u64 cnt)
{
  u64 i = 0;
  const __m128i zero = _mm_setzero_si128();
  for (; i + 4 <= cnt; i += 4)
  {
    const __m128i x = _mm_load_si128((const __m128i *) (src + i));
    _mm_stream_si128((__m128i *) (dst + i), _mm_unpacklo_epi32(x, zero));  // streaming stores: no read for ownership
    _mm_stream_si128((__m128i *) (dst + i + 2), _mm_unpackhi_epi32(x, zero));
  }
  for (; i < cnt; i++)
  {
    dst[i] = src[i];
  }
  _mm_sfence();
}

// takes chunks off job j until none is left; called by the workers and, at the end, by the building thread
void widen_drain(HostWiden *w, const WidenJob &j)
{
  while (true)
  {
    u64 t = w->next_ticket.load();
    if ((t >> 32) != (j.id & 0xffffffffull) || (u32) t >= j.nchunks)
    {
      return;  // the job is used up, or a newer one has started (its parameters are not mine)
    }
    if (!w->next_ticket.compare_exchange_weak(t, t + 1))
    {
      continue;
    }
    const u32 c = (u32) t;
    u32 spins = 0;
    while (__atomic_load_n(w->progress, __ATOMIC_ACQUIRE) <= c && !w->abort.load())
    {
      if (++spins < 256)
      {
        _mm_pause();
      } else
      {
        spins = 0;
        struct timespec ts = {0, 20000};
        nanosleep(&ts, nullptr);
      }
    }
    if (!w->abort.load())
    {
      const u64 b = (u64) c * j.chunk, e = std::min(j.count, b + j.chunk);
      widen_range(j.staging + b, j.dst + b, e - b);
    }
    // a ticket of job j is completed before finish_host_widen(j) returns, so this never lands in a later job
    if (w->chunks_done.fetch_add(1) + 1 == j.nchunks)
    {
      pthread_mutex_lock(&w->mu);
      pthread_cond_broadcast(&w->cv_done);
      pthread_mutex_unlock(&w->mu);
    }
  }
}

WidenJob widen_job_locked(const HostWiden *w)  // caller holds mu (or is the building thread that wrote the fields)
{
  return WidenJob{w->job_id, w->staging, w->dst, w->count, w->chunk, w->nchunks};
}

void *widen_worker(void *p)
{
  HostWiden *w = (HostWiden *) p;
  u64 seen = 0;
  while (true)
  {
    pthread_mutex_lock(&w->mu);
    while (!w->stop && w->job_id == seen)
    {
      pthread_cond_wait(&w->cv_job, &w->mu);
    }
    const bool stop = w->stop;
    const WidenJob job = widen_job_locked(w);
    seen = w->job_id;
    pthread_mutex_unlock(&w->mu);
    if (stop)
    {
      return nullptr;
    }
    widen_drain(w, job);
  }
}

struct CtxExtra
{
  PinnedPool pool;
  HostWiden widen;
  bool fuse_download = false;  // set by esa_build around do_construct
  esa_result_alloc_fn user_alloc = nullptr;  // esa_set_result_allocator
  void *user_alloc_arg = nullptr;
  u64 h2d_bytes = 0;           // what the last upload moved over PCIe on this rank
  cudaEvent_t ev_sorted = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr, ev_e = nullptr;
};

// esa_ctx is declared in common.cuh; the api-private part hangs off a side table keyed by the ctx pointer
struct FullCtx : esa_ctx
{
  CtxExtra x;
};

size_t arena_estimate(u64 n, int nranks)
{
  const u64 N = n + 1;
  if (nranks <= 1)
  {
    // tis 1 + packed text 0.4 + sa 4 + isa 4 + round table 1 + bck worst case 5 + sort regions X and Y 32 + two-level
    // scatter and doubling-round scratch up to 16: 64 bytes per suffix when the device has them, never less than 48
    size_t free_b = 0, total_b = 0;
    const size_t want = (size_t) (64 * N) + ((size_t) 256 << 20), floor_b = (size_t) (48 * N) + ((size_t) 256 << 20);
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
    {
      cudaGetLastError();
      return want;
    }
    const size_t cap = (size_t) (0.92 * (double) total_b);
    return want <= cap ? want : (floor_b <= cap ? cap : floor_b);
  }
  // sharded: replicated text 1 + packed text 0.6 + dense rank array 4 + text-order lcp array of the deep slots 4 + bck
  // (with all-reduce scratch) 2.5, plus the
  // per-range share of everything else with head room for uneven ranges (ESA_SHARD_SLACK percent, default 60)
  u64 slack = 60;
  if (const char *e = getenv("ESA_SHARD_SLACK"))
  {
    slack = (u64) atoi(e);
  }
  const u64 share = N / (u64) nranks + 1;
  return (size_t) (12 * N) + (size_t) (56 * share * (100 + slack) / 100) + ((size_t) 256 << 20);
}

bool is_sharded(const esa_ctx *ctx) { return ctx->comm != nullptr && ctx->comm->nranks > 1; }

void collect_kernel_stats(esa_ctx *ctx, esa_stats *st)
{
  for (int k = 0; k < ESA_K_NUM; k++)
  {
    st->kernel[k].launches = ctx->prof.launches[k];
    st->kernel[k].ms = 0;
    st->kernel[k].bytes = 0;
    if (k != ESA_K_MEMOP)
    {
      st->launches += ctx->prof.launches[k];
    }
  }
  for (const KernelProfiler::Span &sp : ctx->prof.spans)
  {
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, sp.a, sp.b));
    st->kernel[sp.kind].ms += ms;
    st->kernel[sp.kind].bytes += sp.bytes;
  }
}

void do_upload(FullCtx *ctx, const u8 *tis, u64 n, u32 numofchars)
{
  if (tis == nullptr && n > 0)
  {
    esa_throw(ESA_ERR_ARGUMENT, "tis is NULL");
  }
  if (numofchars == 0 || numofchars > 253)
  {
    esa_throw(ESA_ERR_ARGUMENT, "alphabet size %u out of range 1..253", numofchars);
  }
  if (n + 1 >= ((u64) 1 << 32) - 1)
  {
    esa_throw(ESA_ERR_TOOLARGE, "sequence of length %llu needs more than 32-bit indices; not supported",
              (unsigned long long) n);
  }
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ctx->uploaded = ctx->constructed = false;
  const size_t need = arena_estimate(n, ctx->comm != nullptr ? ctx->comm->nranks : 1);
  if (is_sharded(ctx))
  {
    // collective: when any rank's arena has to move, every rank drops its peer mappings first, then all map again
    esa_comm *comm = ctx->comm;
    const u64 token = (u64) (uintptr_t) ctx->arena.base ^ ((u64) ctx->arena.capacity << 1);
    u64 moved = (need > ctx->arena.capacity || token != comm->mapped_token || !comm->direct) ? 1 : 0, all[ESA_MAX_RANKS];
    comm_exchange_host_u64(ctx, comm, &moved, 1, all);
    if (const char *e = getenv("ESA_DEBUG_FAIL_UPLOAD_RANK"))  // test hook: one rank fails inside the collective upload
    {
      if (atoi(e) == comm->rank)
      {
        esa_throw(ESA_ERR_MEMORY, "injected upload failure on rank %d (ESA_DEBUG_FAIL_UPLOAD_RANK)", comm->rank);
      }
    }
    bool any = false;
    for (int r = 0; r < comm->nranks; r++)
    {
      any = any || all[r] != 0;
    }
    if (any)
    {
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      comm_unmap_arenas(ctx, comm);
      comm_exchange_host_u64(ctx, comm, &moved, 1, all);  // nobody holds a mapping any more: arenas may be freed
      ctx->arena.reserve(need);
      comm_map_arenas(ctx, comm);
    }
  } else
  {
    ctx->arena.reserve(need);
  }
  ctx->arena.reset();
  u64 padded = ((n + 15) & ~(u64) 15) + 64;
  if (is_sharded(ctx) && n > 0)
  {
    // the text is replicated, but it crosses PCIe only once per box: every rank uploads its slice and the slices are
    // all-gathered between the devices (NVLink)
    esa_comm *comm = ctx->comm;
    const u64 slice = div_up_u64(div_up_u64(n, (u64) comm->nranks), 256) * 256;
    padded = std::max(padded, slice * comm->nranks + 64);
    ctx->d_tis = ctx->arena.alloc<u8>(padded);
    const u64 b = std::min(n, (u64) comm->rank * slice), e = std::min(n, (u64) (comm->rank + 1) * slice);
    if (e > b)
    {
      CUDA_CHECK(cudaMemcpyAsync(ctx->d_tis + b, tis + b, e - b, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->x.h2d_bytes = e - b;
    ctx->prof.enabled = false;
    comm_allgather_bytes(ctx, comm, ctx->d_tis, slice);
    CUDA_CHECK(cudaMemsetAsync(ctx->d_tis + n, 255, padded - n, ctx->stream));
  } else
  {
    ctx->d_tis = ctx->arena.alloc<u8>(padded);
    ctx->x.h2d_bytes = n;
    CUDA_CHECK(cudaMemsetAsync(ctx->d_tis + n, 255, padded - n, ctx->stream));
    if (n > 0)
    {
      CUDA_CHECK(cudaMemcpyAsync(ctx->d_tis, tis, n, cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  ctx->n = n;
  ctx->sigma = numofchars;
  ctx->arena_after_upload = ctx->arena.mark();
  ctx->uploaded = true;
}

// host threads available to one rank for widening; measured on the B200 box (32 cores, 100 M entries): 2 threads
// 37.8 ms end to end, 4: 27.8, 8: 24.5, 12: 24.1, against 30.6 ms when the widened table crosses PCIe -- so the
// fused path is only taken with at least 6 threads per rank (ESA_WIDEN_THREADS overrides, 0 disables)
// CPUs this process may actually use: the affinity mask, capped by the cgroup CPU quota (a container on a large host
// sees every core in sysconf(_SC_NPROCESSORS_ONLN) but may be throttled to a few)
int usable_cpus()
{
  long cpus = sysconf(_SC_NPROCESSORS_ONLN);
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof set, &set) == 0)
  {
    const int c = CPU_COUNT(&set);
    if (c > 0 && c < cpus)
    {
      cpus = c;
    }
  }
  if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r"))  // cgroup v2: "<quota|max> <period>"
  {
    char q[32];
    long period = 0;
    if (fscanf(f, "%31s %ld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0)
    {
      const long lim = (atol(q) + period - 1) / period;
      if (lim > 0 && lim < cpus)
      {
        cpus = lim;
      }
    }
    fclose(f);
  } else if (FILE *g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r"))  // cgroup v1
  {
    long quota = -1, period = 0;
    if (fscanf(g, "%ld", &quota) != 1)
    {
      quota = -1;
    }
    fclose(g);
    if (FILE *h = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r"))
    {
      if (fscanf(h, "%ld", &period) != 1)
      {
        period = 0;
      }
      fclose(h);
    }
    if (quota > 0 && period > 0 && (quota + period - 1) / period < cpus)
    {
      cpus = (quota + period - 1) / period;
    }
  }
  return (int) std::max<long>(1, cpus);
}

int widen_threads(const FullCtx *ctx)
{
  if (const char *e = getenv("ESA_WIDEN_THREADS"))
  {
    return atoi(e);
  }
  // one CPU per rank stays free for the thread that drives the build; fewer than 4 workers lose against shipping
  // the widened table (measured above)
  const int ranks = ctx->comm != nullptr ? ctx->comm->nranks : 1;
  const int per_rank = usable_cpus() / ranks - 1;
  return per_rank >= 4 ? std::min(12, per_rank) : 0;
}

void start_host_widen(FullCtx *ctx)
{
  HostWiden &w = ctx->x.widen;
  w.count = ctx->shard_count;
  w.staging = (u32 *) ctx->x.pool.get((w.count + 4) * sizeof(u32));
  w.dst = (u64 *) ctx->x.pool.get((w.count + 4) * sizeof(u64));
  w.chunk = std::max<u64>((u64) 1 << 21, div_up_u64(div_up_u64(w.count, 192), 4) * 4);  // multiples of 4 entries
  w.nchunks = (u32) div_up_u64(w.count, w.chunk);
  if (w.nchunks > HostWiden::MAX_CHUNKS)
  {
    w.chunk = div_up_u64(div_up_u64(w.count, HostWiden::MAX_CHUNKS), 4) * 4;
    w.nchunks = (u32) div_up_u64(w.count, w.chunk);
  }
  if (w.progress == nullptr)
  {
    CUDA_CHECK(cudaMallocHost((void **) &w.progress, 64));
    CUDA_CHECK(cudaMalloc((void **) &w.d_ticks, HostWiden::MAX_CHUNKS * sizeof(u32)));
    std::vector<u32> ticks(HostWiden::MAX_CHUNKS);
    for (u32 i = 0; i < HostWiden::MAX_CHUNKS; i++)
    {
      ticks[i] = i + 1;
    }
    CUDA_CHECK(cudaMemcpy(w.d_ticks, ticks.data(), ticks.size() * sizeof(u32), cudaMemcpyHostToDevice));
  }
  if (w.threads.empty())
  {
    const int want = std::max(1, widen_threads(ctx));
    for (int t = 0; t < want; t++)
    {
      pthread_t th;
      if (pthread_create(&th, nullptr, widen_worker, &w) != 0)
      {
        break;  // fewer workers than planned: the building thread drains what is left in finish_host_widen
      }
      w.threads.push_back(th);
    }
  }
  w.nthreads_last = (int) w.threads.size();
  __atomic_store_n(w.progress, 0u, __ATOMIC_RELEASE);
  w.chunks_done.store(0);
  w.abort.store(false);
  CUDA_CHECK(cudaEventRecord(ctx->x.ev_sorted, ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->x.ev_sorted, 0));
  w.active = true;
  for (u32 c = 0; c < w.nchunks; c++)
  {
    const u64 b = (u64) c * w.chunk, e = std::min(w.count, b + w.chunk);
    CUDA_CHECK(cudaMemcpyAsync(w.staging + b, ctx->d_sa + b, (e - b) * sizeof(u32), cudaMemcpyDeviceToHost,
                               ctx->copy_stream));
    CUDA_CHECK(cudaMemcpyAsync(w.progress, w.d_ticks + c, sizeof(u32), cudaMemcpyDeviceToHost, ctx->copy_stream));
  }
  pthread_mutex_lock(&w.mu);
  w.job_id++;
  w.next_ticket.store((w.job_id & 0xffffffffull) << 32);  // tickets of this job open only now, parameters in place
  pthread_cond_broadcast(&w.cv_job);
  pthread_mutex_unlock(&w.mu);
}

// waits for the copy and the workers; returns the widened table (owned by the pinned pool like every result table)
u64 *finish_host_widen(FullCtx *ctx, bool keep)
{
  HostWiden &w = ctx->x.widen;
  if (!w.active)
  {
    return nullptr;
  }
  if (cudaStreamSynchronize(ctx->copy_stream) != cudaSuccess)
  {
    w.abort.store(true);  // the copy died: nobody waits for chunks that will never arrive
    keep = false;
  }
  widen_drain(&w, widen_job_locked(&w));  // the building thread takes what the workers have not started yet
  pthread_mutex_lock(&w.mu);
  while (w.chunks_done.load() < w.nchunks)
  {
    struct timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    ts.tv_nsec += 2000000;
    if (ts.tv_nsec >= 1000000000)
    {
      ts.tv_sec++;
      ts.tv_nsec -= 1000000000;
    }
    pthread_cond_timedwait(&w.cv_done, &w.mu, &ts);
  }
  pthread_mutex_unlock(&w.mu);
  w.active = false;
  ctx->x.pool.put(w.staging);
  w.staging = nullptr;
  u64 *r = w.dst;
  w.dst = nullptr;
  if (!keep)
  {
    ctx->x.pool.put(r);
    return nullptr;
  }
  return r;
}

// esa_shutdown: parks no more
void destroy_host_widen(FullCtx *ctx)
{
  HostWiden &w = ctx->x.widen;
  pthread_mutex_lock(&w.mu);
  w.stop = true;
  pthread_cond_broadcast(&w.cv_job);
  pthread_mutex_unlock(&w.mu);
  for (pthread_t th : w.threads)
  {
    pthread_join(th, nullptr);
  }
  w.threads.clear();
  if (w.progress != nullptr)
  {
    cudaFreeHost(w.progress);
    cudaFree(w.d_ticks);
    w.progress = nullptr;
    w.d_ticks = nullptr;
  }
}

void do_construct(FullCtx *ctx, u32 prefixlength, u32 demand, esa_stats *stats)
{
  if (!ctx->uploaded)
  {
    esa_throw(ESA_ERR_STATE, "esa_construct called before esa_upload");
  }
  CUDA_CHECK(cudaSetDevice(ctx->device));
  if (demand & ESA_STI1TAB)
  {
    demand |= ESA_LCPTAB | ESA_BCKTAB;  // reference: mkvtree.c:376
  }
  if (demand & ESA_SKPTAB)
  {
    demand |= ESA_LCPTAB;  // the skip table is derived from the lcp table (mkvprocess.c:767-815)
  }
  if ((demand & (ESA_BCKTAB | ESA_STI1TAB)) && prefixlength == 0)
  {
    esa_throw(ESA_ERR_ARGUMENT, "option -bck requires option -pl");  // reference: mkvtree.c:470-473
  }
  const u64 n = ctx->n, N = n + 1;
  if (prefixlength > 0)
  {
    const u32 maxk = esa_maximal_prefixlength(ctx->sigma, n);
    if (prefixlength > maxk)
    {
      // reference: src/kurtz/detpfxlen.c:95-98
      esa_throw(ESA_ERR_PREFIXLEN,
                "prefix length %lu is too large, maximal prefix length for this input size and alphabet size is %lu",
                (unsigned long) prefixlength, (unsigned long) maxk);
    }
  }
  cudaStream_t s = ctx->stream;
  Arena &ar = ctx->arena;
  ar.rewind(ctx->arena_after_upload);
  ctx->constructed = false;
  ctx->d_suf64 = nullptr;
  ctx->d_lcp = ctx->d_bwt = ctx->d_sti1 = nullptr;
  ctx->d_bck = ctx->d_llv = ctx->d_sti64 = ctx->d_skp = nullptr;
  ctx->d_plcpv = ctx->d_lcpfull = ctx->d_deepv = nullptr;
  ctx->nllv = ctx->maxbranchdepth = ctx->longest = 0;
  ctx->prof.reset();
  ctx->prof.enabled = stats != nullptr;
  if (ctx->comm != nullptr)
  {
    ctx->comm->bytes_sent = 0;
  }

  const bool sharded = is_sharded(ctx);
  if (sharded && (demand & ESA_SKPTAB))
  {
    esa_throw(ESA_ERR_ARGUMENT, "the skip table needs the whole lcp table on one device: not available in a sharded build");
  }
  ctx->shard_first = 0;
  ctx->shard_count = N;
  ctx->text_first = 0;
  ctx->text_count = N;
  CUDA_CHECK(cudaEventRecord(ctx->x.ev_a, s));
  ctx->d_sa = sharded ? nullptr : ar.alloc<u32>(N + 64);
  ctx->d_isa = ar.alloc<u32>(N + 64);
  ctx->prefixlength = prefixlength;
  ctx->demand = demand;
  ctx->numofcodes = prefixlength > 0 ? esa_numofcodes(ctx->sigma, prefixlength) : 0;
  const bool want_bck = (demand & ESA_BCKTAB) != 0;
  if (want_bck)
  {
    ctx->d_bck = ar.alloc<u64>(2 * ctx->numofcodes);
  }
  const bool want_lcp = (demand & ESA_LCPTAB) != 0;
  ctx->d_pk = nullptr;
  ctx->d_pk5 = nullptr;
  ctx->d_spm = nullptr;
  ctx->d_hround = nullptr;
  ctx->lcp_work_estimate = 0;
  if (ctx->sigma <= 4)
  {
    pack_text(ctx);  // 2-bit text + special mask: text keys of the sorter, suffix-start accesses of lcp and bwt
  } else if (ctx->sigma <= 30 && getenv("ESA_NO_PACK5") == nullptr)
  {
    pack_text5(ctx);  // 5-bit records (protein): one load per suffix start instead of three
  }
  if (want_lcp && ctx->lcp_mode != 2 && !sharded)
  {
    ctx->d_hround = ar.alloc<u8>(N + 64);
    ESA_LAUNCH(ctx, ESA_K_MEMOP, (double) N, CUDA_CHECK(cudaMemsetAsync(ctx->d_hround, 0, N + 64, s)));
  }
  const size_t scratch_mark = ar.mark();
  SortOutcome so;
  if (sharded)
  {
    ar.high = ar.capacity;
    suffix_sort_sharded(ctx, ctx->comm, prefixlength, want_bck, want_lcp, &so);
  } else
  {
    suffix_sort(ctx, prefixlength, want_bck, &so, scratch_mark);
  }
  CUDA_CHECK(cudaEventRecord(ctx->x.ev_b, s));
  ar.rewind(scratch_mark);  // everything the sorter carved is dead now
  u32 table_demand = demand;  // what the sweeps over the suffix array still have to produce on the device
  if (ctx->x.fuse_download && !ctx->x.pool.pageable && ctx->x.user_alloc == nullptr && (demand & ESA_SUFTAB)
      && ctx->shard_count >= ((u64) 1 << 22)
      && widen_threads(ctx) > 0)
  {
    start_host_widen(ctx);
    table_demand &= ~(u32) ESA_SUFTAB;
  }
  // lcp path: direct comparison in suffix-array order unless the deep groups are so large that the work bound
  // sum(active * depth / 32) explodes (long exact tandem arrays); then the text-order plcp path, whose cost does
  // not depend on the lcp values
  bool use_plcp = false;
  u32 halo = 0;
  if (sharded)
  {
    // halo: the suffix in front of my range is the last entry of the nearest non-empty range before mine; the lcp
    // path is chosen on the work estimate of all ranks together
    esa_comm *comm = ctx->comm;
    u64 mine[3] = {ctx->shard_count, 0, ctx->lcp_work_estimate}, all[3 * ESA_MAX_RANKS];
    if (ctx->shard_count > 0)
    {
      mine[1] = read_back_u32(ctx, ctx->d_sa + ctx->shard_count - 1);
    }
    comm_exchange_host_u64(ctx, comm, mine, 3, all);
    u64 work = 0;
    for (int r = 0; r < comm->nranks; r++)
    {
      if (r < comm->rank && all[3 * r] > 0)
      {
        halo = (u32) all[3 * r + 1];
      }
      work += all[3 * r + 2];
    }
    ctx->lcp_work_estimate = work;
  }
  // lcp: direct comparison in suffix-array order (lcp_direct.cu) for every slot except the DEEP ones -- slots that a
  // refinement round with a start depth of at least ESA_LCP_DEEP symbols (default 1024) split from their predecessor.
  // Those take their value from the text-order array V, computed for the deep stretches of the text only (lcp.cu,
  // compute_plcp_deep): the cost of a slot no longer grows with its lcp value (long exact tandem arrays), and inputs
  // without deep slots never see the text-order machinery.  ESA_LCP_MODE=plcp keeps the full text-order path of round 1
  // (one GPU; tests compare the paths), ESA_LCP_MODE=direct compares everything directly.
  u32 deep_round = 255;
  if (want_lcp && ctx->lcp_mode != 1)
  {
    static const u32 deep_lb = [] {
      const char *e = getenv("ESA_LCP_DEEP");
      return (u32) (e != nullptr ? atoi(e) : 1024);
    }();
    const u32 want_depth = (ctx->lcp_mode == 2 && sharded) ? 1u : deep_lb;
    for (u32 t = 1; t <= so.rounds && t < 255; t++)
    {
      if (round_start_depth(ctx->sched, t) >= want_depth)
      {
        deep_round = t;
        break;
      }
    }
  }
  if (want_lcp)
  {
    use_plcp = ctx->lcp_mode == 2 && !sharded;
  }
  ctx->d_deepv = nullptr;
  if (use_plcp)
  {
    ctx->d_plcpv = ar.alloc<u32>(N + 64);
    compute_plcp(ctx, ctx->d_plcpv);
  } else if (want_lcp && deep_round != 255)
  {
    ctx->d_deepv = ar.alloc<u32>(N + 64);
    compute_plcp_deep(ctx, sharded ? ctx->comm : nullptr, ctx->d_deepv, halo, deep_round);
  }
  CUDA_CHECK(cudaEventRecord(ctx->x.ev_c, s));
  if (!sharded)
  {
    if (use_plcp)
    {
      finalize_tables(ctx, table_demand, ctx->d_plcpv, N, 0);
    } else
    {
      finalize_tables_direct(ctx, table_demand, N, 0, 0, ctx->d_deepv, deep_round);
    }
  } else
  {
    if (ctx->shard_count == 0)
    {
      ctx->d_suf64 = (table_demand & ESA_SUFTAB) ? ar.alloc<u64>(4) : nullptr;
      ctx->d_bwt = (demand & ESA_BWTTAB) ? ar.alloc<u8>(4) : nullptr;
      ctx->d_lcp = want_lcp ? ar.alloc<u8>(4) : nullptr;
    } else
    {
      finalize_tables_direct(ctx, table_demand, ctx->shard_count, ctx->shard_first, halo, ctx->d_deepv, deep_round);
    }
    // scalars of the whole table
    u64 sc[2] = {ctx->maxbranchdepth, ctx->longest}, all[2 * ESA_MAX_RANKS];
    comm_exchange_host_u64(ctx, ctx->comm, sc, 2, all);
    for (int r = 0; r < ctx->comm->nranks; r++)
    {
      ctx->maxbranchdepth = std::max(ctx->maxbranchdepth, all[2 * r]);
      ctx->longest = std::max(ctx->longest, all[2 * r + 1]);
    }
  }
  if (sharded && (demand & (ESA_STI1TAB | ESA_STITAB)))
  {
    compute_sti_sharded(ctx, ctx->comm, prefixlength, demand);
  } else
  {
    if (demand & ESA_STI1TAB)
    {
      compute_sti1(ctx, prefixlength);
    }
    if (demand & ESA_STITAB)
    {
      widen_isa(ctx);
    }
  }
  if (demand & ESA_SKPTAB)
  {
    compute_skp(ctx, use_plcp ? ctx->d_plcpv : nullptr, ctx->d_lcpfull);
  }
  CUDA_CHECK(cudaEventRecord(ctx->x.ev_d, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  ctx->constructed = true;
  if (stats != nullptr)
  {
    memset(stats, 0, sizeof *stats);
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->x.ev_a, ctx->x.ev_d));
    stats->ms_total = ms;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->x.ev_a, ctx->x.ev_b));
    stats->ms_initial_sort = ms;  // refined below: split into initial sort and doubling by kernel family
    CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->x.ev_b, ctx->x.ev_c));
    stats->ms_lcp = ms;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->x.ev_c, ctx->x.ev_d));
    stats->ms_derive = ms;
    stats->h0 = so.h0;
    stats->rounds = so.rounds;
    stats->lcp_path = want_lcp ? (use_plcp ? 2 : (ctx->d_deepv != nullptr ? 3 : 1)) : 0;
    stats->active_sum = so.active_sum;
    stats->active_first = so.active_first;
    stats->device_bytes_peak = ar.peak;
    stats->shard_first = ctx->shard_first;
    stats->shard_count = ctx->shard_count;
    stats->nllv = ctx->nllv;
    stats->maxbranchdepth = ctx->maxbranchdepth;
    stats->num_sorted = so.num_sorted;
    stats->comm_bytes_sent = ctx->comm != nullptr ? ctx->comm->bytes_sent : 0;
    stats->comm_direct = (is_sharded(ctx) && ctx->comm->direct) ? 1 : 0;
    collect_kernel_stats(ctx, stats);
    stats->ms_doubling = 0;
  }
}

template <typename T> T *download_table(FullCtx *ctx, const T *d, u64 count, u32 table = 0)
{
  T *h = nullptr;
  if (ctx->x.user_alloc != nullptr && table != 0)
  {
    h = (T *) ctx->x.user_alloc(ctx->x.user_alloc_arg, table, count * sizeof(T));  // the caller's memory (may decline)
  }
  if (h == nullptr)
  {
    h = (T *) ctx->x.pool.get(count * sizeof(T));
  }
  if (count > 0)
  {
    CUDA_CHECK(cudaMemcpyAsync(h, d, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  }
  return h;
}

u64 do_download(FullCtx *ctx, esa_result *out)
{
  if (!ctx->constructed)
  {
    esa_throw(ESA_ERR_STATE, "esa_download called before esa_construct");
  }
  CUDA_CHECK(cudaSetDevice(ctx->device));
  memset(out, 0, sizeof *out);
  const u64 N = ctx->shard_count;  // the whole table on one GPU, this rank's range in a sharded build
  u64 bytes = 0;
  out->n = ctx->n;
  out->shard_first = ctx->shard_first;
  out->shard_count = ctx->shard_count;
  out->numofcodes = ctx->numofcodes;
  out->nllv = ctx->nllv;
  out->maxbranchdepth = ctx->maxbranchdepth;
  out->longest = ctx->longest;
  if (ctx->x.widen.active)
  {
    bytes += 4 * N;  // what crossed PCIe for this table
  } else if (ctx->d_suf64 != nullptr)
  {
    out->suftab = download_table(ctx, ctx->d_suf64, N, ESA_SUFTAB);
    bytes += 8 * N;
  }
  if (ctx->d_lcp != nullptr && (ctx->demand & ESA_LCPTAB))
  {
    out->lcptab = download_table(ctx, ctx->d_lcp, N);
    out->llv = (esa_pairuint *) download_table(ctx, ctx->d_llv, 2 * ctx->nllv);
    bytes += N + 16 * ctx->nllv;
  }
  if (ctx->d_bwt != nullptr)
  {
    out->bwttab = download_table(ctx, ctx->d_bwt, N);
    bytes += N;
  }
  if (ctx->d_bck != nullptr)
  {
    out->bcktab = download_table(ctx, ctx->d_bck, 2 * ctx->numofcodes);
    bytes += 16 * ctx->numofcodes;
  }
  out->text_first = ctx->text_first;
  out->text_count = ctx->text_count;
  if (ctx->d_sti1 != nullptr)
  {
    out->stitab1 = download_table(ctx, ctx->d_sti1, ctx->text_count);
    bytes += ctx->text_count;
  }
  if (ctx->d_sti64 != nullptr)
  {
    out->stitab = download_table(ctx, ctx->d_sti64, ctx->text_count);
    bytes += 8 * ctx->text_count;
  }
  if (ctx->d_skp != nullptr)
  {
    out->skiptab = download_table(ctx, ctx->d_skp, N);
    bytes += 8 * N;
  }
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (ctx->x.widen.active)
  {
    out->suftab = finish_host_widen(ctx, true);
  }
  return bytes;
}

}  // namespace

// =========================================================================================================
extern "C" {

int esa_init(esa_ctx **out, int device, char *errbuf, size_t errlen)
{
  if (out == nullptr)
  {
    set_err(errbuf, errlen, "ctx pointer is NULL");
    return ESA_ERR_ARGUMENT;
  }
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
  {
    char buf[256];
    snprintf(buf, sizeof buf, "no usable CUDA device (%s); this library has no CPU fallback",
             e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    set_err(errbuf, errlen, buf);
    return ESA_ERR_NODEVICE;
  }
  if (device < 0 || device >= count)
  {
    set_err(errbuf, errlen, "device index out of range");
    return ESA_ERR_NODEVICE;
  }
  FullCtx *ctx = new FullCtx();
  int rc = guarded(errbuf, errlen, [&]() {
    CUDA_CHECK(cudaSetDevice(device));
    ctx->device = device;
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    if (prop.major < 10)
    {
      esa_throw(ESA_ERR_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                prop.major, prop.minor);
    }
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaMallocHost((void **) &ctx->h_scalars, 64 * sizeof(u64)));
    for (cudaEvent_t *ev : {&ctx->x.ev_a, &ctx->x.ev_b, &ctx->x.ev_c, &ctx->x.ev_d, &ctx->x.ev_e, &ctx->x.ev_sorted})
    {
      CUDA_CHECK(cudaEventCreate(ev));
    }
    radix_sort_init();
    if (const char *pg = getenv("ESA_PAGEABLE_RESULTS"))
    {
      ctx->x.pool.pageable = atoi(pg) != 0;
    }
    if (const char *m = getenv("ESA_LCP_MODE"))
    {
      ctx->lcp_mode = strcmp(m, "direct") == 0 ? 1 : strcmp(m, "plcp") == 0 ? 2 : 0;
    }
    // the scatter/gather kernels (inverse suffix array, phi, plcp-to-lcp) touch 4 bytes per 32-byte sector:
    // keep L2 from fetching 64/128-byte neighbourhoods for them (ncu: 99 B of DRAM reads per gathered word)
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    cudaGetLastError();  // a refused hint is not an error
  });
  if (rc != ESA_OK)
  {
    delete ctx;
    return rc;
  }
  *out = ctx;
  return ESA_OK;
}

void esa_shutdown(esa_ctx *c)
{
  if (c == nullptr)
  {
    return;
  }
  FullCtx *ctx = static_cast<FullCtx *>(c);
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  comm_destroy(ctx->comm);
  ctx->comm = nullptr;
  ctx->arena.release();
  ctx->prof.destroy();
  ctx->x.pool.destroy();
  finish_host_widen(ctx, false);
  destroy_host_widen(ctx);
  for (cudaEvent_t ev : {ctx->x.ev_a, ctx->x.ev_b, ctx->x.ev_c, ctx->x.ev_d, ctx->x.ev_e, ctx->x.ev_sorted})
  {
    if (ev != nullptr)
    {
      cudaEventDestroy(ev);
    }
  }
  if (ctx->h_scalars != nullptr)
  {
    cudaFreeHost(ctx->h_scalars);
  }
  if (ctx->stream != nullptr)
  {
    cudaStreamDestroy(ctx->stream);
  }
  if (ctx->copy_stream != nullptr)
  {
    cudaStreamDestroy(ctx->copy_stream);
  }
  delete ctx;
}

int esa_upload(esa_ctx *c, const uint8_t *tis, uint64_t n, uint32_t numofchars, char *errbuf, size_t errlen)
{
  if (c == nullptr)
  {
    set_err(errbuf, errlen, "ctx is NULL");
    return ESA_ERR_ARGUMENT;
  }
  FullCtx *ctx = static_cast<FullCtx *>(c);
  const int rc = guarded(errbuf, errlen, [&]() {
    do_upload(ctx, tis, n, numofchars);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  });
  if (rc != ESA_OK && is_sharded(ctx))
  {
    comm_abort(ctx->comm, rc);  // a sharded upload is collective: the other ranks must not wait forever
  }
  return rc;
}

int esa_construct(esa_ctx *c, uint32_t prefixlength, uint32_t demand, esa_stats *stats, char *errbuf, size_t errlen)
{
  if (c == nullptr)
  {
    set_err(errbuf, errlen, "ctx is NULL");
    return ESA_ERR_ARGUMENT;
  }
  FullCtx *ctx = static_cast<FullCtx *>(c);
  const int rc = guarded(errbuf, errlen, [&]() { do_construct(ctx, prefixlength, demand, stats); });
  if (rc != ESA_OK && is_sharded(ctx))
  {
    comm_abort(ctx->comm, rc);  // the other ranks must not wait for this one forever
  }
  return rc;
}

int esa_download(esa_ctx *c, esa_result *out, char *errbuf, size_t errlen)
{
  if (c == nullptr || out == nullptr)
  {
    set_err(errbuf, errlen, "ctx or result is NULL");
    return ESA_ERR_ARGUMENT;
  }
  FullCtx *ctx = static_cast<FullCtx *>(c);
  int rc = guarded(errbuf, errlen, [&]() { do_download(ctx, out); });
  if (rc != ESA_OK)
  {
    esa_free(c, out);
  }
  return rc;
}

int esa_build(esa_ctx *c, const uint8_t *tis, uint64_t n, uint32_t numofchars, uint32_t prefixlength, uint32_t demand,
              esa_result *out, esa_stats *stats, char *errbuf, size_t errlen)
{
  if (c == nullptr || out == nullptr)
  {
    set_err(errbuf, errlen, "ctx or result is NULL");
    return ESA_ERR_ARGUMENT;
  }
  FullCtx *ctx = static_cast<FullCtx *>(c);
  memset(out, 0, sizeof *out);
  int rc = guarded(errbuf, errlen, [&]() {
    cudaStream_t s = ctx->stream;
    CUDA_CHECK(cudaSetDevice(ctx->device));
    CUDA_CHECK(cudaEventRecord(ctx->x.ev_e, s));
    try
    {
      // the upload of a sharded context is collective too (arena mapping, all-gather of the slices): a rank that
      // fails in it must wake the others exactly like one that fails in the construction
      do_upload(ctx, tis, n, numofchars);
      ctx->x.fuse_download = getenv("ESA_NO_HOST_WIDEN") == nullptr;
      do_construct(ctx, prefixlength, demand, stats);  // records ev_a .. ev_d, synchronises
      ctx->x.fuse_download = false;
    } catch (const EsaError &e)
    {
      ctx->x.fuse_download = false;
      finish_host_widen(ctx, false);
      if (is_sharded(ctx))
      {
        comm_abort(ctx->comm, e.code);
      }
      throw;
    } catch (...)
    {
      ctx->x.fuse_download = false;
      finish_host_widen(ctx, false);
      if (is_sharded(ctx))
      {
        comm_abort(ctx->comm, ESA_ERR_MEMORY);
      }
      throw;
    }
    const u64 d2h = do_download(ctx, out);
    if (stats != nullptr)
    {
      float ms = 0;
      CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->x.ev_e, ctx->x.ev_a));
      stats->ms_h2d = ms;
      CUDA_CHECK(cudaEventRecord(ctx->x.ev_e, s));
      CUDA_CHECK(cudaEventSynchronize(ctx->x.ev_e));
      CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->x.ev_d, ctx->x.ev_e));
      stats->ms_d2h = ms;
      stats->h2d_bytes = ctx->x.h2d_bytes;
      stats->d2h_bytes = d2h;
      stats->host_cpus = (uint32_t) usable_cpus();
      stats->host_widen_threads = (uint32_t) ctx->x.widen.nthreads_last;
    }
  });
  if (rc != ESA_OK)
  {
    finish_host_widen(ctx, false);
    esa_free(c, out);
  }
  return rc;
}

void esa_free(esa_ctx *c, esa_result *out)
{
  if (c == nullptr || out == nullptr)
  {
    return;
  }
  FullCtx *ctx = static_cast<FullCtx *>(c);
  void *ptrs[] = {out->suftab, out->lcptab, out->llv, out->bwttab, out->bcktab, out->stitab1, out->stitab, out->skiptab};
  for (void *p : ptrs)
  {
    if (p != nullptr)
    {
      ctx->x.pool.put(p);
    }
  }
  memset(out, 0, sizeof *out);
}

int esa_check(esa_ctx *c, char *errbuf, size_t errlen)
{
  if (c == nullptr)
  {
    set_err(errbuf, errlen, "ctx is NULL");
    return ESA_ERR_ARGUMENT;
  }
  FullCtx *ctx = static_cast<FullCtx *>(c);
  int bad = 0;
  std::string msg;
  int rc = guarded(errbuf, errlen, [&]() {
    if (!ctx->constructed)
    {
      esa_throw(ESA_ERR_STATE, "esa_check called before esa_construct");
    }
    if (is_sharded(ctx))
    {
      esa_throw(ESA_ERR_STATE, "esa_check needs the whole table on one device: not available on a sharded context");
    }
    CUDA_CHECK(cudaSetDevice(ctx->device));
    bad = run_checks(ctx, msg);
  });
  if (rc != ESA_OK)
  {
    return rc;
  }
  if (bad != 0)
  {
    set_err(errbuf, errlen, msg.c_str());
    return ESA_ERR_CHECK;
  }
  return ESA_OK;
}

int esa_childtab(esa_ctx *c, const uint8_t *lcptab, const esa_pairuint *llv, uint64_t nllv, uint64_t totallength,
                 uint8_t *cldtab, uint8_t *cldtab1, char *errbuf, size_t errlen)
{
  if (c == nullptr || cldtab == nullptr)
  {
    set_err(errbuf, errlen, c == nullptr ? "ctx is NULL" : "cldtab is NULL");
    return ESA_ERR_ARGUMENT;
  }
  FullCtx *ctx = static_cast<FullCtx *>(c);
  return guarded(errbuf, errlen, [&]() {
    CUDA_CHECK(cudaSetDevice(ctx->device));
    if (lcptab == nullptr)
    {
      if (!ctx->constructed || ctx->d_lcp == nullptr || !(ctx->demand & ESA_LCPTAB))
      {
        esa_throw(ESA_ERR_STATE, "esa_childtab without tables needs a finished build that demanded the lcp table");
      }
      if (is_sharded(ctx))
      {
        esa_throw(ESA_ERR_STATE, "the child table needs the whole lcp table on one device: not available on a sharded context");
      }
      compute_childtab(ctx, ctx->d_lcp, ctx->d_llv, ctx->nllv, ctx->n, cldtab, cldtab1);
      return;
    }
    if (totallength >= 0xfffffffeull)
    {
      esa_throw(ESA_ERR_TOOLARGE, "totallength %llu does not fit the 32-bit device index", (unsigned long long) totallength);
    }
    if (nllv > 0 && llv == nullptr)
    {
      esa_throw(ESA_ERR_ARGUMENT, "%llu large lcp values announced, none given", (unsigned long long) nllv);
    }
    const u64 N = totallength + 1;
    u8 *d_lcp8 = nullptr;
    u64 *d_pairs = nullptr;
    struct Guard
    {
      u8 *&a;
      u64 *&b;
      ~Guard()
      {
        cudaFree(a);
        cudaFree(b);
      }
    } guard{d_lcp8, d_pairs};
    CUDA_CHECK(cudaMalloc((void **) &d_lcp8, N + 256));
    CUDA_CHECK(cudaMemcpyAsync(d_lcp8, lcptab, N, cudaMemcpyHostToDevice, ctx->stream));
    if (nllv > 0)
    {
      CUDA_CHECK(cudaMalloc((void **) &d_pairs, 16 * nllv));
      CUDA_CHECK(cudaMemcpyAsync(d_pairs, llv, 16 * nllv, cudaMemcpyHostToDevice, ctx->stream));
    }
    compute_childtab(ctx, d_lcp8, d_pairs, nllv, totallength, cldtab, cldtab1);
  });
}

int esa_comm_unique_id(uint8_t id[ESA_COMM_ID_BYTES], char *errbuf, size_t errlen)
{
  if (id == nullptr)
  {
    set_err(errbuf, errlen, "id is NULL");
    return ESA_ERR_ARGUMENT;
  }
  return guarded(errbuf, errlen, [&]() { comm_nccl_unique_id(id); });
}

int esa_comm_attach_nccl(esa_ctx *c, const uint8_t id[ESA_COMM_ID_BYTES], int nranks, int rank, char *errbuf,
                         size_t errlen)
{
  if (c == nullptr || id == nullptr)
  {
    set_err(errbuf, errlen, "ctx or id is NULL");
    return ESA_ERR_ARGUMENT;
  }
  return guarded(errbuf, errlen, [&]() {
    comm_destroy(c->comm);
    c->comm = nullptr;
    c->uploaded = c->constructed = false;
    c->comm = comm_nccl_attach(c, id, nranks, rank);
  });
}

int esa_comm_attach_group(esa_ctx *c, esa_group *group, int rank, char *errbuf, size_t errlen)
{
  if (c == nullptr || group == nullptr)
  {
    set_err(errbuf, errlen, "ctx or group is NULL");
    return ESA_ERR_ARGUMENT;
  }
  return guarded(errbuf, errlen, [&]() {
    comm_destroy(c->comm);
    c->comm = nullptr;
    c->uploaded = c->constructed = false;
    CUDA_CHECK(cudaSetDevice(c->device));
    // ranks on other devices are read with peer copies and, for the small reductions, direct loads
    int count = 0;
    CUDA_CHECK(cudaGetDeviceCount(&count));
    for (int d = 0; d < count; d++)
    {
      int can = 0;
      if (d != c->device && cudaDeviceCanAccessPeer(&can, c->device, d) == cudaSuccess && can)
      {
        cudaDeviceEnablePeerAccess(d, 0);
      }
    }
    cudaGetLastError();  // "already enabled" is fine
    c->comm = comm_group_attach(c, group, rank);
  });
}

void esa_comm_detach(esa_ctx *c)
{
  if (c != nullptr)
  {
    comm_destroy(c->comm);
    c->comm = nullptr;
    c->uploaded = c->constructed = false;
  }
}

int esa_debug_radix_sort(esa_ctx *c, int key_bytes, void *keys, uint32_t *vals, uint64_t n, uint32_t begin_bit,
                         uint32_t end_bit, char *errbuf, size_t errlen)
{
  if (c == nullptr || keys == nullptr || (key_bytes != 4 && key_bytes != 8))
  {
    set_err(errbuf, errlen, "bad argument");
    return ESA_ERR_ARGUMENT;
  }
  FullCtx *ctx = static_cast<FullCtx *>(c);
  return guarded(errbuf, errlen, [&]() {
    CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    ctx->uploaded = ctx->constructed = false;
    const size_t tb = radix_sort_temp_bytes(n, (u32) key_bytes);
    ctx->arena.reserve(2 * n * key_bytes + 2 * n * 4 + tb + (1 << 20));
    ctx->arena.reset();
    ctx->prof.reset();
    ctx->prof.enabled = false;
    char *ka = ctx->arena.alloc<char>(n * key_bytes + 16), *kb = ctx->arena.alloc<char>(n * key_bytes + 16);
    u32 *va = nullptr, *vb = nullptr;
    if (vals != nullptr)
    {
      va = ctx->arena.alloc<u32>(n + 4);
      vb = ctx->arena.alloc<u32>(n + 4);
      CUDA_CHECK(cudaMemcpyAsync(va, vals, n * 4, cudaMemcpyHostToDevice, s));
    }
    void *temp = ctx->arena.alloc<char>(tb);
    CUDA_CHECK(cudaMemcpyAsync(ka, keys, n * key_bytes, cudaMemcpyHostToDevice, s));
    int w = key_bytes == 4 ? radix_sort_pairs_u32(ctx, (u32 *) ka, (u32 *) kb, va, vb, n, begin_bit, end_bit, temp, tb)
                           : radix_sort_pairs_u64(ctx, (u64 *) ka, (u64 *) kb, va, vb, n, begin_bit, end_bit, temp, tb);
    CUDA_CHECK(cudaMemcpyAsync(keys, w ? kb : ka, n * key_bytes, cudaMemcpyDeviceToHost, s));
    if (vals != nullptr)
    {
      CUDA_CHECK(cudaMemcpyAsync(vals, w ? vb : va, n * 4, cudaMemcpyDeviceToHost, s));
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
  });
}

int esa_set_result_allocator(esa_ctx *c, esa_result_alloc_fn fn, void *user)
{
  if (c == nullptr)
  {
    return ESA_ERR_ARGUMENT;
  }
  FullCtx *ctx = static_cast<FullCtx *>(c);
  ctx->x.user_alloc = fn;
  ctx->x.user_alloc_arg = user;
  return ESA_OK;
}

int esa_debug_corrupt(esa_ctx *c, uint32_t table, uint64_t index, uint64_t xor_mask, char *errbuf, size_t errlen)
{
  if (c == nullptr)
  {
    set_err(errbuf, errlen, "ctx is NULL");
    return ESA_ERR_ARGUMENT;
  }
  FullCtx *ctx = static_cast<FullCtx *>(c);
  return guarded(errbuf, errlen, [&]() {
    if (!ctx->constructed)
    {
      esa_throw(ESA_ERR_STATE, "esa_debug_corrupt called before esa_construct");
    }
    CUDA_CHECK(cudaSetDevice(ctx->device));
    const u64 N = ctx->shard_count;
    char *base = nullptr;
    size_t width = 0;
    u64 count = 0;
    switch (table)
    {
      case ESA_SUFTAB: base = (char *) ctx->d_suf64, width = 8, count = N; break;
      case ESA_LCPTAB: base = (char *) ctx->d_lcp, width = 1, count = N; break;
      case ESA_BWTTAB: base = (char *) ctx->d_bwt, width = 1, count = N; break;
      case ESA_BCKTAB: base = (char *) ctx->d_bck, width = 8, count = 2 * ctx->numofcodes; break;
      case ESA_STI1TAB: base = (char *) ctx->d_sti1, width = 1, count = N; break;
      case ESA_DEBUG_LLVTAB: base = (char *) ctx->d_llv, width = 8, count = 2 * ctx->nllv; break;
      default: esa_throw(ESA_ERR_ARGUMENT, "esa_debug_corrupt: unknown table 0x%x", table);
    }
    if (base == nullptr || index >= count)
    {
      esa_throw(ESA_ERR_ARGUMENT, "esa_debug_corrupt: table 0x%x is not resident or index out of range", table);
    }
    u64 word = 0;
    CUDA_CHECK(cudaMemcpy(&word, base + index * width, width, cudaMemcpyDeviceToHost));
    word ^= xor_mask;
    CUDA_CHECK(cudaMemcpy(base + index * width, &word, width, cudaMemcpyHostToDevice));
  });
}

uint64_t esa_numofcodes(uint32_t numofchars, uint32_t prefixlength)
{
  return (uint64_t) pow((double) numofchars, (double) prefixlength);
}

static uint64_t logalphasize(uint64_t numofchars, double value)
{
  if (value <= (double) numofchars)
  {
    return 1;
  }
  return (uint64_t) floor(log(value) / log((double) numofchars));
}

uint32_t esa_recommended_prefixlength(uint32_t numofchars, uint64_t totallength)
{
  const uint64_t k = logalphasize(numofchars, (double) totallength / 16);
  return (uint32_t) (k == 0 ? 1 : k);
}

uint32_t esa_maximal_prefixlength(uint32_t numofchars, uint64_t totallength)
{
  const unsigned prefixlenbits = 4;
  uint64_t maxk = logalphasize(numofchars, (double) totallength / (16 / 4));
  uint64_t t = logalphasize(numofchars, (double) ((((uint64_t) 1) << (64 - prefixlenbits)) - 1));
  maxk = std::min(maxk, t);
  t = (((uint64_t) 1) << prefixlenbits) - 1;
  maxk = std::min(maxk, t);
  return (uint32_t) maxk;
}

const char *esa_version(void) { return "vstree_b200 0.1.0 sm_100a"; }

}  // extern "C"
