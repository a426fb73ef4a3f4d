/*
 * vstree_b200.h -- C ABI of the B200-native enhanced-suffix-array builder (libvstree_b200.so).
 *
 * This is the compute seam inside mkvtree's `mkvtreeprocess` (reference: src/Mkvtree/mkvprocess.c:875-1089):
 * the host driver keeps option parsing, alphabet transformation and all file I/O, and hands the transformed
 * sequence `tis` to this library, which builds the tables on one B200 (sm_100a) with hand-written CUDA kernels.
 * There is no CPU fallback: every entry point that needs the GPU fails with a negative code and a message
 * when no device is usable.
 *
 * Conventions (the reference's own, src/include/errordef.h:45-110): return 0 on success, a negative code on
 * error with a NUL-terminated message in `errbuf` (may be NULL); plain pointers and sizes only; no exceptions
 * cross this boundary.  One build at a time per context; a context is bound to one device.
 *
 * Sizes: n = totallength (letters + wildcards + separators), N = n + 1 (the sentinel suffix at position n is a
 * member of every table).  `Uint` of the reference is 8 bytes, little endian (src/include/types.h:41-49).
 */
#ifndef VSTREE_B200_H
#define VSTREE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Table demand bits; same values as the reference's TISTAB..STITAB (src/include/virtualdef.h:24-64). */
#define ESA_TISTAB  (1u << 0)
#define ESA_OISTAB  (1u << 1)
#define ESA_DESTAB  (1u << 2)
#define ESA_SSPTAB  (1u << 3)
#define ESA_BWTTAB  (1u << 4)   /* .bwt  replaces encodeburrowswheeler, src/kurtz/bwtcode.c:293-311 */
#define ESA_SUFTAB  (1u << 5)   /* .suf  replaces besesuffixsort[special], src/Mkvtree/bese.c:855-1105 */
#define ESA_LCPTAB  (1u << 6)   /* .lcp/.llv  replaces outlcpsubtab, src/Mkvtree/bese.c:533-591 */
#define ESA_BCKTAB  (1u << 7)   /* .bck  replaces bucketsortsuffixes + makebcktab, ppsort.c:83-313, mkvprocess.c:251-327 */
#define ESA_STI1TAB (1u << 8)   /* .sti1 replaces makestitab1, src/Mkvtree/mkvprocess.c:583-643 */
#define ESA_SKPTAB  (1u << 9)   /* .skp  replaces makeskiptable, src/kurtz/mkskip.c:53-91 */
#define ESA_STITAB  (1u << 13)  /* .sti  full inverse suffix array, src/Mkvtree/mksti.c:15-34 */

/* Symbol codes of the transformed sequence (src/include/chardef.h:19-50). */
#define ESA_WILDCARD     254u
#define ESA_SEPARATOR    255u
#define ESA_UNDEFBWTCHAR 253u

/* Error codes. */
#define ESA_OK              0
#define ESA_ERR_ARGUMENT   (-1)   /* inconsistent arguments (e.g. -bck without a prefix length) */
#define ESA_ERR_NODEVICE   (-2)   /* no usable CUDA device: there is no CPU fallback */
#define ESA_ERR_CUDA       (-3)   /* a CUDA call or kernel failed */
#define ESA_ERR_MEMORY     (-4)   /* host or device memory exhausted */
#define ESA_ERR_TOOLARGE   (-5)   /* n + 1 does not fit the 32-bit device index (n < 2^32 - 1 required) */
#define ESA_ERR_PREFIXLEN  (-6)   /* prefix length larger than vm_whatisthemaximalprefixlength allows */
#define ESA_ERR_STATE      (-7)   /* call sequence violated (e.g. construct before upload) */
#define ESA_ERR_CHECK      (-8)   /* esa_check found an inconsistent table */
#define ESA_ERR_COMM       (-9)   /* NCCL / multi-GPU communication failed */

typedef struct esa_ctx esa_ctx;   /* opaque; owns the device, its streams and the device arena */

/* PairUint of the reference (src/include/types.h:289-292): .llv entry. */
typedef struct
{
  uint64_t uint0;   /* index into lcptab */
  uint64_t uint1;   /* the lcp value (>= 255) */
} esa_pairuint;

/*
 * Tables in PINNED HOST memory owned by the library until esa_free().  A pointer is NULL when the table was
 * not demanded.  Mirrors the table members of Virtualtree (src/include/virtualdef.h:186-219).
 */
typedef struct
{
  uint64_t n;                /* totallength */
  uint64_t numofcodes;       /* numofchars^prefixlength (0 when prefixlength == 0) */
  uint64_t *suftab;          /* [N]   suffix array */
  uint8_t *lcptab;           /* [N]   min(lcp,255) */
  esa_pairuint *llv;         /* [nllv] (index,value) for lcp >= 255, increasing index */
  uint64_t nllv;             /* largelcpvalues */
  uint64_t maxbranchdepth;   /* max lcp */
  uint64_t longest;          /* index i with suftab[i] == 0 (determinelongest, mkvprocess.c:857-873) */
  uint8_t *bwttab;           /* [N]   tis[suftab[i]-1], 253 where suftab[i] == 0 */
  uint64_t *bcktab;          /* [2*numofcodes] (left, mid) per k-mer code */
  uint8_t *stitab1;          /* [N]   min(rank - bucket left, 255), by text position */
  uint64_t *stitab;          /* [N]   full inverse suffix array */
  uint64_t *skiptab;         /* [N]   skip table */
  /* sharded builds (esa_construct_sharded): suftab/lcptab/bwttab hold only the suffix-array entries
     [shard_first, shard_first + shard_count) of this rank, llv indices are global; one GPU: 0 and N */
  uint64_t shard_first;
  uint64_t shard_count;
  /* stitab1 / stitab are indexed by TEXT position: sharded builds return the positions [text_first, text_first +
     text_count) of this rank (slices in rank order tile the text); one GPU: 0 and N */
  uint64_t text_first;
  uint64_t text_count;
} esa_result;

/* Kernel families, for esa_stats.kernel[]. */
enum
{
  ESA_K_CLASSIFY = 0,   /* special-distance classification + initial h0-mer keys */
  ESA_K_RADIX_HIST,     /* all-digit histogram of a radix sort */
  ESA_K_RADIX_PASS,     /* one onesweep LSD pass (the dominant kernel) */
  ESA_K_REGROUP,        /* group heads, rank scatter, active-set compaction */
  ESA_K_GATHER_RANK,    /* doubling key gather rank[sa+h] */
  ESA_K_BCK,            /* bucket boundaries */
  ESA_K_PLCP,           /* irreducible lcp values in text order + fill */
  ESA_K_FINALIZE,       /* lcp/llv/bwt/suf widening in suffix-array order */
  ESA_K_STI1,           /* reduced inverse suffix array */
  ESA_K_MISC,           /* scans over tile summaries, small helpers */
  ESA_K_MEMOP,          /* cudaMemsetAsync / device-to-device cudaMemcpyAsync nodes (not kernels of this library) */
  ESA_K_COMM,           /* sharded builds: collectives (NCCL or in-process peer copies) between the ranks */
  ESA_K_EXCHANGE,       /* sharded builds: owner tagging, request serving and answer merging around the collectives */
  ESA_K_RADIX_PASS_TEXT,/* first pass of the initial DNA sort: keys and positions computed from the packed text inside the pass */
  ESA_K_NUM
};

typedef struct
{
  uint64_t launches;         /* kernel launches of this family */
  double ms;                 /* summed device time (CUDA events on the launching stream) */
  double bytes;              /* summed ALGORITHMIC bytes (DESIGN.md lists the per-unit figure of each family) */
} esa_kernel_stat;

typedef struct
{
  double ms_total;           /* device time of esa_construct (CUDA events) */
  double ms_initial_sort;    /* classification + h0-mer radix sort + first regroup */
  double ms_doubling;        /* all prefix-doubling rounds */
  double ms_lcp;             /* PLCP + fill */
  double ms_derive;          /* finalize (lcp8, llv, bwt, suf64), bck, sti1 */
  uint32_t h0;               /* symbols per initial key */
  uint32_t rounds;           /* prefix-doubling rounds after the initial sort */
  uint32_t lcp_path;         /* 0 no lcp, 1 direct comparison in suffix-array order, 2 text-order plcp, 3 direct + text-order for the deep slots */
  uint32_t comm_direct;      /* sharded builds: 1 when the ranks exchanged through mapped peer memory, 0 through messages */
  uint64_t active_sum;       /* sum over rounds of the active-set sizes */
  uint64_t active_first;     /* active-set size of round 1 */
  uint64_t launches;         /* kernel launches of this library inside esa_construct (ESA_K_MEMOP excluded) */
  uint64_t device_bytes_peak;/* arena high-water mark */
  esa_kernel_stat kernel[ESA_K_NUM];
  /* end-to-end legs, filled by esa_build (0 otherwise) */
  double ms_h2d, ms_d2h;
  uint64_t h2d_bytes, d2h_bytes;
  /* sharded builds (0 otherwise) */
  uint64_t shard_first, shard_count; /* this rank's slice of the suffix array */
  uint64_t comm_bytes_sent;          /* payload this rank sent to other ranks */
  /* esa_build: host side of the fused download */
  uint32_t host_cpus;                /* CPUs usable by this process (affinity mask and cgroup quota) */
  uint32_t host_widen_threads;       /* host threads that widened the 32-bit suffix array (0: widened on the device) */
  /* scalars of the tables left resident (this rank's range in a sharded build) */
  uint64_t nllv;                     /* lcp values >= 255 (largelcpvalues) */
  uint64_t maxbranchdepth;
  uint64_t num_sorted;               /* suffixes that went through the sort (N - special-initial ones) */
} esa_stats;

/* ---- lifetime ---- */

/* Bind a context to CUDA device `device` (>= 0).  Fails with ESA_ERR_NODEVICE when there is none. */
int esa_init(esa_ctx **ctx, int device, char *errbuf, size_t errlen);
void esa_shutdown(esa_ctx *ctx);

/* ---- the one-call seam: host buffers in, pinned host tables out ----
 * Replaces besesuffixsort[special] + makelargelcp + makebcktab + determinelongest + makestitab1 + makebwttab
 * (src/Mkvtree/mkvprocess.c:976-1060).  `tis` holds n symbols (codes < numofchars, 254, 255) in ordinary host
 * memory.  prefixlength 0 means "no bucket table" (then ESA_BCKTAB/ESA_STI1TAB are an argument error, as
 * mkvtree.c:470-473).  ESA_STI1TAB implies lcp and bck are computed (mkvtree.c:376).  `stats` may be NULL. */
int esa_build(esa_ctx *ctx, const uint8_t *tis, uint64_t n, uint32_t numofchars, uint32_t prefixlength,
              uint32_t demand, esa_result *out, esa_stats *stats, char *errbuf, size_t errlen);
void esa_free(esa_ctx *ctx, esa_result *out);

/* Optional: where the result tables go.  By default esa_build / esa_download return tables in pinned memory owned by the
 * library.  With an allocator set, the library asks it for the memory of a table (table = one ESA_*TAB bit; bytes = size
 * of the table) and copies the table there instead; NULL from the allocator means "use your own buffer".  The memory
 * stays the caller's (esa_free leaves it alone).  Meant for one-shot processes that want a table straight in its final
 * place -- the mkvtree driver hands out the mapped pages of the .suf file, so 8 of the 10 output bytes per suffix are
 * never copied on the host.  Ordinary (pageable) memory is fine; the chunked, overlapped download of esa_build needs
 * pinned memory and is not used for such a table. */
typedef void *(*esa_result_alloc_fn)(void *user, uint32_t table, uint64_t bytes);
int esa_set_result_allocator(esa_ctx *ctx, esa_result_alloc_fn fn, void *user);

/* ---- the same work in three stages, so that the device-resident part can be timed alone ---- */
int esa_upload(esa_ctx *ctx, const uint8_t *tis, uint64_t n, uint32_t numofchars, char *errbuf, size_t errlen);
int esa_construct(esa_ctx *ctx, uint32_t prefixlength, uint32_t demand, esa_stats *stats, char *errbuf,
                  size_t errlen);
int esa_download(esa_ctx *ctx, esa_result *out, char *errbuf, size_t errlen);

/* ---- sharded builds: several GPUs build ONE index ----
 * The suffix array is cut into contiguous ranges along first-symbols bins of the initial sort key; rank r of R builds
 * range r.  Every rank uploads the WHOLE sequence (it is replicated), classifies its slice of text positions, sends each
 * suffix to the rank that owns its bin (all-to-all), sorts its range (prefix doubling; ranks of foreign suffixes are
 * fetched from their owners each round) and derives lcp/llv/bwt for its range.  Attach a communicator to a context,
 * then call esa_upload / esa_construct / esa_download (or esa_build) on EVERY rank: they are collective.  The tables of
 * esa_result then cover suffix-array entries [shard_first, shard_first + shard_count) only (llv indices are global;
 * bcktab, maxbranchdepth and longest are complete on every rank); concatenating the ranks' tables in rank order gives
 * the files the reference writes.  stitab1 / stitab (ESA_STI1TAB / ESA_STITAB) are indexed by text position: every rank
 * returns the slice [text_first, text_first + text_count) of them, again tiling the table in rank order.  ESA_SKPTAB and
 * esa_check need whole tables on one device and are refused on a sharded context.
 *
 * Transport A, one process per GPU (NCCL over NVLink): rank 0 calls esa_comm_unique_id and distributes the id by any
 * means (MPI, torch.distributed, a file); every rank calls esa_comm_attach_nccl.  libnccl.so.2 is loaded with dlopen.
 * Transport B, one thread per rank inside one process: esa_group_create once, esa_comm_attach_group per context. */
#define ESA_COMM_ID_BYTES 128
typedef struct esa_group esa_group;
int esa_comm_unique_id(uint8_t id[ESA_COMM_ID_BYTES], char *errbuf, size_t errlen);
int esa_comm_attach_nccl(esa_ctx *ctx, const uint8_t id[ESA_COMM_ID_BYTES], int nranks, int rank, char *errbuf,
                         size_t errlen);
int esa_group_create(esa_group **group, int nranks, char *errbuf, size_t errlen);
void esa_group_destroy(esa_group *group);
void esa_group_abort(esa_group *group);   /* wakes ranks waiting in a collective; they fail with ESA_ERR_COMM */
void esa_group_reset(esa_group *group);   /* after all ranks returned from a failed call: group usable again */
int esa_comm_attach_group(esa_ctx *ctx, esa_group *group, int rank, char *errbuf, size_t errlen);
void esa_comm_detach(esa_ctx *ctx);
/* Host logic of the exchange step, exposed for tests (no GPU needed): from every rank's per-bin suffix counts
 * hist[r*stride + b] (256 bins = top 8 bits of the initial sort key) derive the bin -> rank table, what rank `me`
 * sends to / receives from every rank, and where its run of every bin starts inside the owner's buffer.  With
 * stride >= 258, hist[r*stride + 257] = special-initial suffixes rank r classified: they all end up in the tail of the
 * table, i.e. in the last rank's range, and count towards its quota. */
int esa_shard_plan(const uint64_t *hist, size_t stride, int nranks, int me, uint8_t owner[256], uint64_t *send_cnt,
                   uint64_t *recv_cnt, uint32_t start[256]);

/* ---- suffix-to-text checks on the device (reference DEBUG checks: bese.c:355-528, cprsuf.c:66-95,
 * mkvprocess.c:672-763).  Runs on the tables left resident by the last esa_construct: permutation, order of
 * neighbours, lcp8, the EXACT values of the .llv pairs with largelcpvalues and maxbranchdepth, bwt, bck against the
 * suffix array, bck.left + sti1.  Returns 0 when consistent, ESA_ERR_CHECK with a message naming the violated
 * property otherwise. ---- */
int esa_check(esa_ctx *ctx, char *errbuf, size_t errlen);

/* ---- the child table (.cld, .cld1) from the lcp table: SURVEY.md 8f row 4 ----
 * Replaces makecldtabandoutput (src/Mkvtree/mkcld.mn.c:13-47) = mkcldtabnextlIndex + mkcldtabupdown + compresscldtab
 * (src/kurtz/mkcld.c:41-284); the reference's mkcld maps SUFTAB | LCPTAB of an index and reads the lcp values only.
 *   lcptab != NULL : host tables as mapped from an index: lcptab[totallength + 1] and the `nllv` large-value pairs
 *                    (the .lcp and .llv files); works on any context, resident builds are left untouched
 *   lcptab == NULL : the lcp table left resident by the last esa_construct / esa_build of this context (it must have
 *                    demanded ESA_LCPTAB; one GPU); llv / nllv / totallength are ignored
 * cldtab  : caller's buffer of 3 * (totallength + 1) bytes, Childinfo {up, down, nextlIndex} per entry
 *           (src/include/virtualdef.h:165-179), the bytes of the .cld file
 * cldtab1 : caller's buffer of totallength + 1 bytes (the .cld1 file) or NULL.  Entries the reference never writes
 *           (the last one, mkcld.c:230-283 stops below totallength) are 0.
 * Fails with ESA_ERR_ARGUMENT when the pairs do not belong to the lcp table (every 255 needs exactly one pair), with
 * ESA_ERR_MEMORY when the device cannot hold the working set: about 12.3 bytes per entry (32-bit lcp values and their
 * min-pyramid 4.3, the records 3, the queue of far-reaching entries 4, .cld1 1) next to whatever a resident build
 * keeps -- a 3 Gbp index therefore needs a context without a resident build of that size. */
int esa_childtab(esa_ctx *ctx, const uint8_t *lcptab, const esa_pairuint *llv, uint64_t nllv, uint64_t totallength,
                 uint8_t *cldtab, uint8_t *cldtab1, char *errbuf, size_t errlen);

/* ---- host-side scalar math that must match the reference bit for bit ---- */
/* vm_determinenumofcodes, src/kurtz-basic/detnumofcodes.c:8-11 */
uint64_t esa_numofcodes(uint32_t numofchars, uint32_t prefixlength);
/* vm_recommendedprefixlength, src/kurtz/detpfxlen.c:52-61 (sizeofbckentry = 16, mkvtree.c:652-658) */
uint32_t esa_recommended_prefixlength(uint32_t numofchars, uint64_t totallength);
/* vm_whatisthemaximalprefixlength, src/kurtz/detpfxlen.c:63-89 (PREFIXLENBITS 4, mkvaux.h:27) */
uint32_t esa_maximal_prefixlength(uint32_t numofchars, uint64_t totallength);

/* ---- testing hook: the LSD radix sort the suffix sorter is built on, on host arrays (in place).
 * key_bytes is 4 or 8; vals may be NULL.  Sorts by key bits [begin_bit, end_bit), stable. ---- */
int esa_debug_radix_sort(esa_ctx *ctx, int key_bytes, void *keys, uint32_t *vals, uint64_t n, uint32_t begin_bit,
                         uint32_t end_bit, char *errbuf, size_t errlen);

/* ---- testing hook for esa_check: XORs `xor_mask` into element `index` of a table left resident on the device by the
 * last esa_construct (table = one ESA_*TAB bit; ESA_DEBUG_LLVTAB addresses the .llv words: 2*j index, 2*j+1 value), so
 * that tests can prove esa_check notices a damaged table (the reference's DEBUG checks abort the same way). ---- */
#define ESA_DEBUG_LLVTAB (1u << 31)
int esa_debug_corrupt(esa_ctx *ctx, uint32_t table, uint64_t index, uint64_t xor_mask, char *errbuf, size_t errlen);

/* Library identification: "vstree_b200 <version> sm_100a". */
const char *esa_version(void);

#ifdef __cplusplus
}
#endif

#endif
