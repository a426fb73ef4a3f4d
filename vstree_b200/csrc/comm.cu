// comm.cu -- see comm.cuh.  Two transports behind one set of collectives.
#include "comm.cuh"

#include <dlfcn.h>
#include <pthread.h>
#include <unistd.h>
#include <nccl.h>  // types and prototypes only: the library is resolved with dlopen, never linked

namespace
{

// ---------------------------------------------------------------------------------------------------------
// NCCL through dlopen
struct NcclApi
{
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;  // optional
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};

NcclApi g_nccl;
pthread_mutex_t g_nccl_lock = PTHREAD_MUTEX_INITIALIZER;

template <typename F> void nccl_sym(F &f, const char *name)
{
  f = (F) dlsym(g_nccl.handle, name);
  if (f == nullptr)
  {
    esa_throw(ESA_ERR_COMM, "libnccl does not export %s", name);
  }
}

NcclApi &nccl()
{
  pthread_mutex_lock(&g_nccl_lock);
  if (g_nccl.handle == nullptr)
  {
    // a process that already carries an NCCL (PyTorch) must share it: RTLD_NOLOAD finds the loaded copy by soname
    const char *names[] = {getenv("ESA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names)
    {
      if (nm != nullptr && h == nullptr)
      {
        h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
      }
    }
    for (const char *nm : names)
    {
      if (nm != nullptr && h == nullptr)
      {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      }
    }
    if (h == nullptr)
    {
      pthread_mutex_unlock(&g_nccl_lock);
      esa_throw(ESA_ERR_COMM, "cannot load libnccl.so.2 (%s); set ESA_NCCL_LIB to its path", dlerror());
    }
    g_nccl.handle = h;
    try
    {
      nccl_sym(g_nccl.GetUniqueId, "ncclGetUniqueId");
      nccl_sym(g_nccl.CommInitRank, "ncclCommInitRank");
      nccl_sym(g_nccl.CommDestroy, "ncclCommDestroy");
      nccl_sym(g_nccl.GetErrorString, "ncclGetErrorString");
      nccl_sym(g_nccl.AllReduce, "ncclAllReduce");
      nccl_sym(g_nccl.AllGather, "ncclAllGather");
      nccl_sym(g_nccl.ReduceScatter, "ncclReduceScatter");
      nccl_sym(g_nccl.Send, "ncclSend");
      nccl_sym(g_nccl.Recv, "ncclRecv");
      nccl_sym(g_nccl.GroupStart, "ncclGroupStart");
      nccl_sym(g_nccl.GroupEnd, "ncclGroupEnd");
      g_nccl.CommAbort = (ncclResult_t (*)(ncclComm_t)) dlsym(g_nccl.handle, "ncclCommAbort");
    } catch (...)
    {
      g_nccl.handle = nullptr;
      pthread_mutex_unlock(&g_nccl_lock);
      throw;
    }
  }
  pthread_mutex_unlock(&g_nccl_lock);
  return g_nccl;
}

#define NCCL_CHECK(call)                                                                                        \
  do                                                                                                            \
  {                                                                                                             \
    ncclResult_t res__ = (call);                                                                                \
    if (res__ != ncclSuccess)                                                                                   \
    {                                                                                                           \
      esa_throw(ESA_ERR_COMM, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, nccl().GetErrorString(res__)); \
    }                                                                                                           \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// in-process group: a barrier that can be aborted by a failing rank
struct GroupSync
{
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  pthread_cond_t cv = PTHREAD_COND_INITIALIZER;
  int waiting = 0;
  u64 generation = 0;
};

}  // namespace

namespace
{

GroupSync *gs(esa_group *g) { return static_cast<GroupSync *>(g->sync); }

void group_barrier(esa_group *g)
{
  GroupSync *s = gs(g);
  pthread_mutex_lock(&s->mu);
  if (!g->failed)
  {
    const u64 gen = s->generation;
    if (++s->waiting == g->nranks)
    {
      s->waiting = 0;
      s->generation++;
      pthread_cond_broadcast(&s->cv);
    } else
    {
      while (gen == s->generation && !g->failed)
      {
        pthread_cond_wait(&s->cv, &s->mu);
      }
    }
  }
  const int failed = g->failed;
  pthread_mutex_unlock(&s->mu);
  if (failed)
  {
    esa_throw(ESA_ERR_COMM, "another rank of the group failed; build abandoned");
  }
}

struct PeerPtrs
{
  const u64 *p[ESA_MAX_RANKS];
};

__global__ void group_reduce_kernel(PeerPtrs peers, int nranks, u64 *__restrict__ out, u64 count, bool use_max)
{
  for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (u64) gridDim.x * blockDim.x)
  {
    u64 acc = peers.p[0][i];
    for (int r = 1; r < nranks; r++)
    {
      const u64 v = peers.p[r][i];
      acc = use_max ? (v > acc ? v : acc) : acc + v;
    }
    out[i] = acc;
  }
}

// publish -> barrier -> pull -> barrier.  The pulled buffers must be complete: every rank drains its stream first.
template <typename Pull> void group_round(esa_ctx *ctx, esa_comm *comm, const void *ptr, const u64 *off, const u64 *cnt, Pull &&pull)
{
  esa_group *g = comm->group;
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  esa_group::Slot &me = g->slot[comm->rank];
  me.ptr = ptr;
  me.off = off;
  me.cnt = cnt;
  group_barrier(g);
  pull(g);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  group_barrier(g);
}

void copy_words(esa_ctx *ctx, u32 *dst, const u32 *src, u64 words)
{
  if (words > 0)
  {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, words * sizeof(u32), cudaMemcpyDefault, ctx->stream));
  }
}

}  // namespace

// =========================================================================================================
// setup

void comm_nccl_unique_id(u8 id[ESA_COMM_ID_BYTES])
{
  static_assert(sizeof(ncclUniqueId) <= ESA_COMM_ID_BYTES, "ESA_COMM_ID_BYTES too small");
  ncclUniqueId uid;
  NCCL_CHECK(nccl().GetUniqueId(&uid));
  memset(id, 0, ESA_COMM_ID_BYTES);
  memcpy(id, &uid, sizeof uid);
}

esa_comm *comm_nccl_attach(esa_ctx *ctx, const u8 id[ESA_COMM_ID_BYTES], int nranks, int rank)
{
  if (nranks < 1 || nranks > ESA_MAX_RANKS || rank < 0 || rank >= nranks)
  {
    esa_throw(ESA_ERR_ARGUMENT, "bad communicator shape: rank %d of %d", rank, nranks);
  }
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  ncclComm_t c = nullptr;
  NCCL_CHECK(nccl().CommInitRank(&c, nranks, uid, rank));
  esa_comm *comm = new esa_comm();
  comm->rank = rank;
  comm->nranks = nranks;
  comm->nccl = c;
  comm->xchg_bytes = (size_t) 1 << 20;
  CUDA_CHECK(cudaMalloc((void **) &comm->d_xchg, comm->xchg_bytes));
  return comm;
}

esa_comm *comm_group_attach(esa_ctx *ctx, esa_group *g, int rank)
{
  if (g == nullptr || rank < 0 || rank >= g->nranks)
  {
    esa_throw(ESA_ERR_ARGUMENT, "bad group rank %d", rank);
  }
  g->slot[rank].device = ctx->device;
  esa_comm *comm = new esa_comm();
  comm->rank = rank;
  comm->nranks = g->nranks;
  comm->group = g;
  return comm;
}

void comm_unmap_arenas(esa_ctx *ctx, esa_comm *comm)
{
  if (comm == nullptr)
  {
    return;
  }
  if (ctx != nullptr)
  {
    cudaSetDevice(ctx->device);
  }
  for (int r = 0; r < ESA_MAX_RANKS; r++)
  {
    if (comm->ipc_opened[r] != nullptr)
    {
      cudaIpcCloseMemHandle(comm->ipc_opened[r]);
      comm->ipc_opened[r] = nullptr;
    }
    comm->peer_base[r] = nullptr;
  }
  cudaGetLastError();
  comm->direct = false;
  comm->direct_loads = false;
}

void comm_map_arenas(esa_ctx *ctx, esa_comm *comm)
{
  comm_unmap_arenas(ctx, comm);
  if (comm->nranks <= 1)
  {
    return;
  }
  bool want = true;
  if (const char *e = getenv("ESA_COMM_DIRECT"))
  {
    want = atoi(e) != 0;
  }
  // words per rank: [0] usable flag, [1] base pointer value, [2] capacity, [3] process id, [4..11] IPC handle
  constexpr size_t ROW = 12;
  u64 mine[ROW] = {0}, all[ROW * ESA_MAX_RANKS];
  static_assert(sizeof(cudaIpcMemHandle_t) <= 8 * sizeof(u64), "IPC handle size");
  mine[1] = (u64) (uintptr_t) ctx->arena.base;
  mine[2] = ctx->arena.capacity;
  mine[3] = (u64) getpid();
  bool ok = want && ctx->arena.base != nullptr;
  if (ok && comm->group == nullptr)
  {
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, ctx->arena.base) == cudaSuccess)
    {
      memcpy(&mine[4], &h, sizeof h);
    } else
    {
      cudaGetLastError();
      ok = false;
    }
  }
  mine[0] = ok ? 1 : 0;
  comm_exchange_host_u64(ctx, comm, mine, ROW, all);
  bool everyone = true;
  for (int r = 0; r < comm->nranks; r++)
  {
    everyone = everyone && all[r * ROW] != 0;
  }
  u64 mapped = 0;
  if (everyone)
  {
    mapped = 1;
    for (int r = 0; r < comm->nranks; r++)
    {
      if (r == comm->rank)
      {
        comm->peer_base[r] = ctx->arena.base;
      } else if (all[r * ROW + 3] == (u64) getpid())
      {
        comm->peer_base[r] = (char *) (uintptr_t) all[r * ROW + 1];  // same process: the pointer itself
      } else
      {
        cudaIpcMemHandle_t h;
        memcpy(&h, &all[r * ROW + 4], sizeof h);
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        {
          cudaGetLastError();
          mapped = 0;
          break;
        }
        comm->ipc_opened[r] = p;
        comm->peer_base[r] = (char *) p;
      }
    }
  }
  // all or nothing: a rank that could not map a peer makes everybody use the message-passing paths
  u64 votes[ESA_MAX_RANKS];
  comm_exchange_host_u64(ctx, comm, &mapped, 1, votes);
  bool agreed = true;
  for (int r = 0; r < comm->nranks; r++)
  {
    agreed = agreed && votes[r] != 0;
  }
  if (!agreed)
  {
    comm_unmap_arenas(ctx, comm);
    return;
  }
  comm->direct = true;
  comm->direct_loads = true;
  for (int r = 0; r < comm->nranks; r++)
  {
    comm->direct_loads = comm->direct_loads && all[r * ROW + 3] == (u64) getpid();
  }
  if (const char *e = getenv("ESA_COMM_DIRECT_LOADS"))
  {
    comm->direct_loads = atoi(e) != 0;
  }
  comm->mapped_token = (u64) (uintptr_t) ctx->arena.base ^ ((u64) ctx->arena.capacity << 1);
}

// A rank that failed inside a collective call leaves the others behind.  In-process group: wake them (they fail with
// ESA_ERR_COMM at their next barrier).  NCCL: abort this rank's communicator so that its own queued operations and
// its later esa_shutdown cannot block; the peers' pending NCCL operations end when their NCCL notices the loss of
// this rank (connection reset when the process exits) -- the launcher tears the job down, as for any NCCL program.
void comm_abort(esa_comm *comm, int code)
{
  if (comm == nullptr)
  {
    return;
  }
  if (comm->group != nullptr)
  {
    esa_group_abort(comm->group);
  }
  // argument / state errors are raised before the first collective of a call, identically on every rank: the
  // communicator stays usable.  Run-time failures (CUDA, memory, transport) leave the ranks out of step.
  const bool runtime_failure = code == ESA_ERR_CUDA || code == ESA_ERR_MEMORY || code == ESA_ERR_COMM;
  if (comm->nccl != nullptr && comm->nranks > 1 && runtime_failure)
  {
    NcclApi &api = nccl();
    if (api.CommAbort != nullptr)
    {
      api.CommAbort((ncclComm_t) comm->nccl);
    }
    comm->nccl = nullptr;
    comm->aborted = true;
  }
}

void comm_destroy(esa_comm *comm)
{
  if (comm == nullptr)
  {
    return;
  }
  comm_unmap_arenas(nullptr, comm);
  if (comm->nccl != nullptr)
  {
    nccl().CommDestroy((ncclComm_t) comm->nccl);
  }
  if (comm->d_xchg != nullptr)
  {
    cudaFree(comm->d_xchg);
  }
  delete comm;
}

extern "C" {

int esa_group_create(esa_group **out, int nranks, char *errbuf, size_t errlen)
{
  if (out == nullptr || nranks < 1 || nranks > ESA_MAX_RANKS)
  {
    if (errbuf != nullptr && errlen > 0)
    {
      snprintf(errbuf, errlen, "group size %d out of range 1..%d", nranks, ESA_MAX_RANKS);
    }
    return ESA_ERR_ARGUMENT;
  }
  esa_group *g = new esa_group();
  g->nranks = nranks;
  g->sync = new GroupSync();
  *out = g;
  return ESA_OK;
}

void esa_group_destroy(esa_group *g)
{
  if (g != nullptr)
  {
    delete gs(g);
    delete g;
  }
}

/* marks the group failed and wakes every rank that waits in a collective */
void esa_group_abort(esa_group *g)
{
  if (g == nullptr)
  {
    return;
  }
  GroupSync *s = gs(g);
  pthread_mutex_lock(&s->mu);
  g->failed = 1;
  pthread_cond_broadcast(&s->cv);
  pthread_mutex_unlock(&s->mu);
}

/* after every rank has returned from a failed collective call: makes the group usable again */
void esa_group_reset(esa_group *g)
{
  if (g == nullptr)
  {
    return;
  }
  GroupSync *s = gs(g);
  pthread_mutex_lock(&s->mu);
  g->failed = 0;
  s->waiting = 0;
  s->generation++;
  pthread_mutex_unlock(&s->mu);
}

}  // extern "C"

// =========================================================================================================
// collectives

void comm_exchange_host_u64(esa_ctx *ctx, esa_comm *comm, const u64 *mine, size_t count, u64 *all)
{
  if (comm->nranks == 1)
  {
    memcpy(all, mine, count * sizeof(u64));
    return;
  }
  if (comm->aborted)
  {
    esa_throw(ESA_ERR_COMM, "the communicator was aborted after an earlier failure: attach a new one");
  }
  if (comm->group != nullptr)
  {
    esa_group *g = comm->group;
    g->slot[comm->rank].host.assign(mine, mine + count);
    group_barrier(g);
    for (int r = 0; r < comm->nranks; r++)
    {
      memcpy(all + (size_t) r * count, g->slot[r].host.data(), count * sizeof(u64));
    }
    group_barrier(g);
    return;
  }
  // small device staging owned by the communicator (the arena may not exist yet, or may be about to move)
  if ((count + count * comm->nranks) * sizeof(u64) > comm->xchg_bytes)
  {
    esa_throw(ESA_ERR_COMM, "host value exchange of %zu words exceeds the staging buffer", count);
  }
  u64 *d_send = comm->d_xchg, *d_recv = comm->d_xchg + count;
  CUDA_CHECK(cudaMemcpyAsync(d_send, mine, count * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
  ESA_LAUNCH(ctx, ESA_K_COMM, 0,
             NCCL_CHECK(nccl().AllGather(d_send, d_recv, count, ncclUint64, (ncclComm_t) comm->nccl, ctx->stream)));
  CUDA_CHECK(cudaMemcpyAsync(all, d_recv, count * comm->nranks * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

void comm_allreduce_u64(esa_ctx *ctx, esa_comm *comm, u64 *d_buf, u64 *d_tmp, size_t count, bool max_instead_of_sum)
{
  if (comm->nranks == 1 || count == 0)
  {
    return;
  }
  comm->bytes_sent += (u64) count * 8;
  if (comm->group != nullptr)
  {
    group_round(ctx, comm, d_buf, nullptr, nullptr, [&](esa_group *g) {
      PeerPtrs pp;
      for (int r = 0; r < comm->nranks; r++)
      {
        pp.p[r] = (const u64 *) g->slot[r].ptr;
      }
      const u32 grid = (u32) min((u64) ctx->num_sms * 8, div_up_u64(count, 256));
      ESA_LAUNCH(ctx, ESA_K_COMM, 8.0 * count * comm->nranks,
                 (group_reduce_kernel<<<grid, 256, 0, ctx->stream>>>(pp, comm->nranks, d_tmp, count, max_instead_of_sum)));
    });
    CUDA_CHECK(cudaMemcpyAsync(d_buf, d_tmp, count * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    return;
  }
  ESA_LAUNCH(ctx, ESA_K_COMM, 8.0 * count,
             NCCL_CHECK(nccl().AllReduce(d_buf, d_buf, count, ncclUint64, max_instead_of_sum ? ncclMax : ncclSum,
                                         (ncclComm_t) comm->nccl, ctx->stream)));
}

namespace
{
// out[i] = sum over ranks of their word (first + i)
__global__ void group_reduce_slice_kernel(PeerPtrs peers, int nranks, u64 first, u64 *__restrict__ out, u64 count)
{
  for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (u64) gridDim.x * blockDim.x)
  {
    u64 acc = 0;
    for (int r = 0; r < nranks; r++)
    {
      acc += __ldcv(peers.p[r] + first + i);
    }
    out[i] = acc;
  }
}
}  // namespace

void comm_reduce_scatter_sum_u64(esa_ctx *ctx, esa_comm *comm, const u64 *d_full, u64 *d_out, size_t slice_words)
{
  if (slice_words == 0)
  {
    return;
  }
  const int G = comm->nranks, me = comm->rank;
  if (G == 1)
  {
    CUDA_CHECK(cudaMemcpyAsync(d_out, d_full, slice_words * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    return;
  }
  comm->bytes_sent += (u64) slice_words * 8 * (G - 1);
  if (comm->group != nullptr)
  {
    group_round(ctx, comm, d_full, nullptr, nullptr, [&](esa_group *g) {
      PeerPtrs pp;
      for (int r = 0; r < G; r++)
      {
        pp.p[r] = (const u64 *) g->slot[r].ptr;
      }
      const u32 grid = (u32) min((u64) ctx->num_sms * 8, div_up_u64(slice_words, 256));
      ESA_LAUNCH(ctx, ESA_K_COMM, 8.0 * slice_words * G,
                 (group_reduce_slice_kernel<<<grid, 256, 0, ctx->stream>>>(pp, G, (u64) me * slice_words, d_out, slice_words)));
    });
    return;
  }
  ESA_LAUNCH(ctx, ESA_K_COMM, 8.0 * slice_words * G,
             NCCL_CHECK(nccl().ReduceScatter(d_full, d_out, slice_words, ncclUint64, ncclSum, (ncclComm_t) comm->nccl,
                                             ctx->stream)));
}

void comm_allgather_bytes(esa_ctx *ctx, esa_comm *comm, u8 *d_buf, size_t slice)
{
  if (comm->nranks == 1 || slice == 0)
  {
    return;
  }
  const int G = comm->nranks, me = comm->rank;
  comm->bytes_sent += (u64) slice * (G - 1);
  if (comm->group != nullptr)
  {
    group_round(ctx, comm, d_buf, nullptr, nullptr, [&](esa_group *g) {
      ESA_LAUNCH(ctx, ESA_K_COMM, (double) slice * (G - 1), {
        for (int r = 0; r < G; r++)
        {
          if (r != me)
          {
            CUDA_CHECK(cudaMemcpyAsync(d_buf + (size_t) r * slice, (const u8 *) g->slot[r].ptr + (size_t) r * slice,
                                       slice, cudaMemcpyDefault, ctx->stream));
          }
        }
      });
    });
    return;
  }
  ESA_LAUNCH(ctx, ESA_K_COMM, (double) slice * (G - 1),
             NCCL_CHECK(nccl().AllGather(d_buf + (size_t) me * slice, d_buf, slice, ncclUint8, (ncclComm_t) comm->nccl,
                                         ctx->stream)));
}

void comm_alltoallv_u32(esa_ctx *ctx, esa_comm *comm, const u32 *d_send, const u64 *send_off, const u64 *send_cnt,
                        u32 *d_recv, const u64 *recv_off, const u64 *recv_cnt)
{
  const int me = comm->rank;
  u64 out_words = 0;
  for (int r = 0; r < comm->nranks; r++)
  {
    out_words += r == me ? 0 : send_cnt[r];
  }
  comm->bytes_sent += out_words * 4;
  if (comm->group != nullptr && comm->nranks > 1)
  {
    group_round(ctx, comm, d_send, send_off, send_cnt, [&](esa_group *g) {
      ESA_LAUNCH(ctx, ESA_K_COMM, 4.0 * out_words, {
        for (int r = 0; r < comm->nranks; r++)
        {
          const esa_group::Slot &peer = g->slot[r];
          if (peer.cnt[me] != recv_cnt[r])
          {
            esa_throw(ESA_ERR_COMM, "all-to-all count mismatch: rank %d sends %llu, rank %d expects %llu", r,
                      (unsigned long long) peer.cnt[me], me, (unsigned long long) recv_cnt[r]);
          }
          copy_words(ctx, d_recv + recv_off[r], (const u32 *) peer.ptr + peer.off[me], recv_cnt[r]);
        }
      });
    });
    return;
  }
  ESA_LAUNCH(ctx, ESA_K_COMM, 4.0 * out_words, {
    copy_words(ctx, d_recv + recv_off[me], d_send + send_off[me], send_cnt[me]);
    if (comm->nranks > 1)
    {
      NCCL_CHECK(nccl().GroupStart());
      for (int peer = 0; peer < comm->nranks; peer++)
      {
        if (peer == me)
        {
          continue;
        }
        if (send_cnt[peer] > 0)
        {
          NCCL_CHECK(nccl().Send(d_send + send_off[peer], send_cnt[peer], ncclUint32, peer, (ncclComm_t) comm->nccl,
                                 ctx->stream));
        }
        if (recv_cnt[peer] > 0)
        {
          NCCL_CHECK(nccl().Recv(d_recv + recv_off[peer], recv_cnt[peer], ncclUint32, peer, (ncclComm_t) comm->nccl,
                                 ctx->stream));
        }
      }
      NCCL_CHECK(nccl().GroupEnd());
    }
  });
}

void comm_allgatherv_u32(esa_ctx *ctx, esa_comm *comm, const u32 *d_send, u32 *d_recv, const u64 *off, const u64 *cnt)
{
  const int me = comm->rank;
  comm->bytes_sent += cnt[me] * 4 * (u64) (comm->nranks - 1);
  if (comm->group != nullptr && comm->nranks > 1)
  {
    group_round(ctx, comm, d_send, nullptr, nullptr, [&](esa_group *g) {
      ESA_LAUNCH(ctx, ESA_K_COMM, 4.0 * cnt[me] * (comm->nranks - 1), {
        for (int r = 0; r < comm->nranks; r++)
        {
          if (d_recv + off[r] != (const u32 *) g->slot[r].ptr)
          {
            copy_words(ctx, d_recv + off[r], (const u32 *) g->slot[r].ptr, cnt[r]);
          }
        }
      });
    });
    return;
  }
  ESA_LAUNCH(ctx, ESA_K_COMM, 4.0 * cnt[me] * (comm->nranks - 1), {
    if (d_recv + off[me] != d_send)
    {
      copy_words(ctx, d_recv + off[me], d_send, cnt[me]);
    }
    if (comm->nranks > 1)
    {
      NCCL_CHECK(nccl().GroupStart());
      for (int peer = 0; peer < comm->nranks; peer++)
      {
        if (peer == me)
        {
          continue;
        }
        if (cnt[me] > 0)
        {
          NCCL_CHECK(nccl().Send(d_send, cnt[me], ncclUint32, peer, (ncclComm_t) comm->nccl, ctx->stream));
        }
        if (cnt[peer] > 0)
        {
          NCCL_CHECK(nccl().Recv(d_recv + off[peer], cnt[peer], ncclUint32, peer, (ncclComm_t) comm->nccl, ctx->stream));
        }
      }
      NCCL_CHECK(nccl().GroupEnd());
    }
  });
}

void comm_gatherv_u32(esa_ctx *ctx, esa_comm *comm, int root, const u32 *d_send, u32 *d_recv, const u64 *off,
                      const u64 *cnt)
{
  const int me = comm->rank;
  comm->bytes_sent += me == root ? 0 : cnt[me] * 4;
  if (comm->group != nullptr && comm->nranks > 1)
  {
    group_round(ctx, comm, d_send, nullptr, nullptr, [&](esa_group *g) {
      if (me == root)
      {
        ESA_LAUNCH(ctx, ESA_K_COMM, 0, {
          for (int r = 0; r < comm->nranks; r++)
          {
            copy_words(ctx, d_recv + off[r], (const u32 *) g->slot[r].ptr, cnt[r]);
          }
        });
      }
    });
    return;
  }
  ESA_LAUNCH(ctx, ESA_K_COMM, 4.0 * cnt[me], {
    if (me == root)
    {
      copy_words(ctx, d_recv + off[me], d_send, cnt[me]);
    }
    if (comm->nranks > 1)
    {
      NCCL_CHECK(nccl().GroupStart());
      if (me == root)
      {
        for (int peer = 0; peer < comm->nranks; peer++)
        {
          if (peer != me && cnt[peer] > 0)
          {
            NCCL_CHECK(nccl().Recv(d_recv + off[peer], cnt[peer], ncclUint32, peer, (ncclComm_t) comm->nccl, ctx->stream));
          }
        }
      } else if (cnt[me] > 0)
      {
        NCCL_CHECK(nccl().Send(d_send, cnt[me], ncclUint32, root, (ncclComm_t) comm->nccl, ctx->stream));
      }
      NCCL_CHECK(nccl().GroupEnd());
    }
  });
}
