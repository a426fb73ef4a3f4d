// comm.cuh -- the exchange steps of the sharded build: one rank per GPU, collectives ordered on the rank's compute
// stream.
//
// The reference has no distributed code at all.  The sharded suffix sorter needs
//   * small exchanges of counts (host values of every rank),
//   * one all-to-all of (key, position) pairs after the splitter pass,
//   * per doubling round one all-to-all of rank requests and one of the answers,
//   * all-reduces of the bucket table pieces, a gather of the special-initial suffixes to the last rank and a
//     one-element halo between neighbouring shards.
// Two transports implement them:
//   NCCL   one PROCESS per GPU (torchrun / MPI style launch); libnccl.so.2 is resolved with dlopen at attach time so
//          that the single-GPU library has no link-time dependency on it.  Grouped ncclSend/ncclRecv over NVLink.
//   group  one THREAD per rank inside one process (the mkvtree host driver with several GPUs, and the tests, which
//          run several ranks on ONE device): peers publish their buffers in a shared table and every rank pulls its
//          part with cudaMemcpyAsync (peer-to-peer over NVLink when the ranks sit on different devices).
#pragma once

#include "common.cuh"

constexpr int ESA_MAX_RANKS = 64;

// in-process group (transport "group")
struct esa_group
{
  int nranks = 0;
  void *sync = nullptr;             // abortable barrier state (comm.cu)
  struct Slot
  {
    const void *ptr = nullptr;      // published device buffer
    const u64 *off = nullptr;       // published host offset/count tables (nranks entries)
    const u64 *cnt = nullptr;
    int device = -1;
    std::vector<u64> host;          // published host values
  } slot[ESA_MAX_RANKS];
  volatile int failed = 0;          // a rank hit an error: everybody bails out at the next barrier
};

struct esa_comm
{
  int rank = 0;
  int nranks = 1;
  void *nccl = nullptr;       // ncclComm_t when the transport is NCCL
  esa_group *group = nullptr; // else the in-process group
  u64 bytes_sent = 0;         // payload this rank sent to OTHER ranks
  u64 *d_xchg = nullptr;      // NCCL transport: device staging of the host value exchanges
  size_t xchg_bytes = 0;
  // peer memory: the device arenas of all ranks, addressable from this rank's device (CUDA IPC between processes, plain
  // pointers inside one process).  Kernels then store into / load from other ranks' buffers directly over NVLink.
  bool direct = false;                       // arenas are mapped
  // Random 4-byte LOADS from a peer arena are only used when all ranks live in one process: through a legacy
  // CUDA IPC mapping they ran 85x slower on B200 (261 ms instead of 2.9 ms for 31 M loads over 4 GPUs; small pages in
  // the importer's page tables), while the streaming STORES of the fused partition pass are unaffected.
  bool direct_loads = false;
  char *peer_base[ESA_MAX_RANKS] = {nullptr};
  void *ipc_opened[ESA_MAX_RANKS] = {nullptr};  // mappings to close (NCCL transport)
  u64 mapped_token = 0;                      // identifies the local arena allocation the peers have mapped
  bool aborted = false;                      // NCCL transport: comm_abort ran, every later collective fails
};

// transport setup (comm.cu)
void comm_nccl_unique_id(u8 id[ESA_COMM_ID_BYTES]);
esa_comm *comm_nccl_attach(esa_ctx *ctx, const u8 id[ESA_COMM_ID_BYTES], int nranks, int rank);
esa_comm *comm_group_attach(esa_ctx *ctx, esa_group *g, int rank);
void comm_destroy(esa_comm *comm);
void comm_abort(esa_comm *comm, int code);  // a rank failed inside a collective call: do not leave the others (or itself) hanging

// peer memory (collective): after the call comm->direct tells whether every rank could map every arena;
// ESA_COMM_DIRECT=0 forces the message-passing paths.  Must be re-run (after comm_unmap_arenas) when any arena moved.
void comm_map_arenas(esa_ctx *ctx, esa_comm *comm);
void comm_unmap_arenas(esa_ctx *ctx, esa_comm *comm);
// translate a pointer into this rank's arena to the same offset of rank r's arena
template <typename T> static inline T *comm_peer_ptr(esa_ctx *ctx, esa_comm *comm, int r, u64 byte_offset)
{
  (void) ctx;
  return reinterpret_cast<T *>(comm->peer_base[r] + byte_offset);
}

// ---- collectives; every rank of the communicator must call them in the same order ----
// every rank learns `count` 64-bit host values of every rank: all[r*count + i]; synchronises the stream
void comm_exchange_host_u64(esa_ctx *ctx, esa_comm *comm, const u64 *mine, size_t count, u64 *all);
// in-place sum / max over ranks of `count` 64-bit words in device memory (`d_tmp`: count words of scratch)
void comm_allreduce_u64(esa_ctx *ctx, esa_comm *comm, u64 *d_buf, u64 *d_tmp, size_t count, bool max_instead_of_sum);
// reduce-scatter: d_full holds nranks * slice_words 64-bit words on every rank; rank r receives in d_out the sums over
// the ranks of the words [r * slice_words, (r + 1) * slice_words)
void comm_reduce_scatter_sum_u64(esa_ctx *ctx, esa_comm *comm, const u64 *d_full, u64 *d_out, size_t slice_words);
// in-place all-gather of equal byte slices: rank r's slice lives at d_buf + r*slice on every rank
void comm_allgather_bytes(esa_ctx *ctx, esa_comm *comm, u8 *d_buf, size_t slice);
// variable all-to-all of 32-bit words: counts/offsets are host arrays of nranks entries (in words);
// recv_cnt[r] must equal what rank r sends here
void comm_alltoallv_u32(esa_ctx *ctx, esa_comm *comm, const u32 *d_send, const u64 *send_off, const u64 *send_cnt,
                        u32 *d_recv, const u64 *recv_off, const u64 *recv_cnt);
// variable all-gather of 32-bit words: rank r contributes cnt[r] words, placed at off[r] of d_recv on every rank
void comm_allgatherv_u32(esa_ctx *ctx, esa_comm *comm, const u32 *d_send, u32 *d_recv, const u64 *off,
                         const u64 *cnt);
// gather to one rank (root): rank r sends cnt[r] words, root places them at off[r]
void comm_gatherv_u32(esa_ctx *ctx, esa_comm *comm, int root, const u32 *d_send, u32 *d_recv, const u64 *off,
                      const u64 *cnt);
