"""ctypes binding of include/vstree_b200.h.  There is no fallback: if the shared library is missing or no GPU is
usable, calls raise."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)

DEMAND_BITS = {"tis": 1 << 0, "ois": 1 << 1, "des": 1 << 2, "ssp": 1 << 3, "bwt": 1 << 4, "suf": 1 << 5,
               "lcp": 1 << 6, "bck": 1 << 7, "sti1": 1 << 8, "skp": 1 << 9, "sti": 1 << 13}
KERNEL_FAMILIES = ["classify", "radix_hist", "radix_pass", "regroup", "gather_rank", "bck", "plcp", "finalize",
                   "sti1", "misc", "memop", "comm", "exchange", "radix_pass_text"]
ERRBUF = 1024


class EsaError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


class KernelStat(ctypes.Structure):
    _fields_ = [("launches", ctypes.c_uint64), ("ms", ctypes.c_double), ("bytes", ctypes.c_double)]


class Stats(ctypes.Structure):
    _fields_ = [("ms_total", ctypes.c_double), ("ms_initial_sort", ctypes.c_double), ("ms_doubling", ctypes.c_double),
                ("ms_lcp", ctypes.c_double), ("ms_derive", ctypes.c_double), ("h0", ctypes.c_uint32),
                ("rounds", ctypes.c_uint32), ("lcp_path", ctypes.c_uint32), ("comm_direct", ctypes.c_uint32), ("active_sum", ctypes.c_uint64), ("active_first", ctypes.c_uint64),
                ("launches", ctypes.c_uint64), ("device_bytes_peak", ctypes.c_uint64),
                ("kernel", KernelStat * len(KERNEL_FAMILIES)), ("ms_h2d", ctypes.c_double),
                ("ms_d2h", ctypes.c_double), ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64),
                ("shard_first", ctypes.c_uint64), ("shard_count", ctypes.c_uint64),
                ("comm_bytes_sent", ctypes.c_uint64), ("host_cpus", ctypes.c_uint32),
                ("host_widen_threads", ctypes.c_uint32), ("nllv", ctypes.c_uint64),
                ("maxbranchdepth", ctypes.c_uint64), ("num_sorted", ctypes.c_uint64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "kernel"}
        d["kernel"] = {name: {"launches": self.kernel[i].launches, "ms": self.kernel[i].ms,
                              "bytes": self.kernel[i].bytes} for i, name in enumerate(KERNEL_FAMILIES)}
        return d


class Result(ctypes.Structure):
    _fields_ = [("n", ctypes.c_uint64), ("numofcodes", ctypes.c_uint64), ("suftab", ctypes.POINTER(ctypes.c_uint64)),
                ("lcptab", ctypes.POINTER(ctypes.c_uint8)), ("llv", ctypes.POINTER(ctypes.c_uint64)),
                ("nllv", ctypes.c_uint64), ("maxbranchdepth", ctypes.c_uint64), ("longest", ctypes.c_uint64),
                ("bwttab", ctypes.POINTER(ctypes.c_uint8)), ("bcktab", ctypes.POINTER(ctypes.c_uint64)),
                ("stitab1", ctypes.POINTER(ctypes.c_uint8)), ("stitab", ctypes.POINTER(ctypes.c_uint64)),
                ("skiptab", ctypes.POINTER(ctypes.c_uint64)), ("shard_first", ctypes.c_uint64),
                ("shard_count", ctypes.c_uint64), ("text_first", ctypes.c_uint64), ("text_count", ctypes.c_uint64)]


def library_path():
    """The product library; VSTREE_B200_LIB overrides the path (tuning experiments with alternative builds)."""
    return os.environ.get("VSTREE_B200_LIB") or os.path.join(_HERE, "lib", "libvstree_b200.so")


_lib = None


def load_library():
    """dlopen the product library; raises if it has not been built (run `make` or __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise EsaError(-2, f"{path} is missing: build it with `make` at the repository root (no fallback exists)")
    lib = ctypes.CDLL(path)
    c = ctypes
    vp, cp, sz = c.c_void_p, c.c_char_p, c.c_size_t
    lib.esa_init.restype = c.c_int
    lib.esa_init.argtypes = [c.POINTER(vp), c.c_int, cp, sz]
    lib.esa_shutdown.restype = None
    lib.esa_shutdown.argtypes = [vp]
    lib.esa_build.restype = c.c_int
    lib.esa_build.argtypes = [vp, vp, c.c_uint64, c.c_uint32, c.c_uint32, c.c_uint32, c.POINTER(Result),
                              c.POINTER(Stats), cp, sz]
    lib.esa_free.restype = None
    lib.esa_free.argtypes = [vp, c.POINTER(Result)]
    lib.esa_upload.restype = c.c_int
    lib.esa_upload.argtypes = [vp, vp, c.c_uint64, c.c_uint32, cp, sz]
    lib.esa_construct.restype = c.c_int
    lib.esa_construct.argtypes = [vp, c.c_uint32, c.c_uint32, c.POINTER(Stats), cp, sz]
    lib.esa_download.restype = c.c_int
    lib.esa_download.argtypes = [vp, c.POINTER(Result), cp, sz]
    lib.esa_check.restype = c.c_int
    lib.esa_check.argtypes = [vp, cp, sz]
    lib.esa_numofcodes.restype = c.c_uint64
    lib.esa_numofcodes.argtypes = [c.c_uint32, c.c_uint32]
    lib.esa_recommended_prefixlength.restype = c.c_uint32
    lib.esa_recommended_prefixlength.argtypes = [c.c_uint32, c.c_uint64]
    lib.esa_maximal_prefixlength.restype = c.c_uint32
    lib.esa_maximal_prefixlength.argtypes = [c.c_uint32, c.c_uint64]
    lib.esa_debug_radix_sort.restype = c.c_int
    lib.esa_debug_radix_sort.argtypes = [vp, c.c_int, vp, vp, c.c_uint64, c.c_uint32, c.c_uint32, cp, sz]
    lib.esa_debug_corrupt.restype = c.c_int
    lib.esa_debug_corrupt.argtypes = [vp, c.c_uint32, c.c_uint64, c.c_uint64, cp, sz]
    lib.esa_childtab.restype = c.c_int
    lib.esa_childtab.argtypes = [vp, vp, vp, c.c_uint64, c.c_uint64, vp, vp, cp, sz]
    lib.esa_version.restype = cp
    lib.esa_version.argtypes = []
    lib.esa_comm_unique_id.restype = c.c_int
    lib.esa_comm_unique_id.argtypes = [vp, cp, sz]
    lib.esa_comm_attach_nccl.restype = c.c_int
    lib.esa_comm_attach_nccl.argtypes = [vp, vp, c.c_int, c.c_int, cp, sz]
    lib.esa_group_create.restype = c.c_int
    lib.esa_group_create.argtypes = [c.POINTER(vp), c.c_int, cp, sz]
    lib.esa_group_destroy.restype = None
    lib.esa_group_destroy.argtypes = [vp]
    lib.esa_group_abort.restype = None
    lib.esa_group_abort.argtypes = [vp]
    lib.esa_group_reset.restype = None
    lib.esa_group_reset.argtypes = [vp]
    lib.esa_comm_attach_group.restype = c.c_int
    lib.esa_comm_attach_group.argtypes = [vp, vp, c.c_int, cp, sz]
    lib.esa_comm_detach.restype = None
    lib.esa_comm_detach.argtypes = [vp]
    _lib = lib
    return lib


def demand_mask(tables):
    m = 0
    for t in tables:
        m |= DEMAND_BITS[t]
    return m


def _as_ptr(tis):
    """Accepts a numpy uint8 array, a torch CPU uint8 tensor (e.g. pinned) or a raw (address, length) tuple."""
    if isinstance(tis, tuple):
        return ctypes.c_void_p(tis[0]), int(tis[1]), None
    if hasattr(tis, "data_ptr"):
        return ctypes.c_void_p(tis.data_ptr()), int(tis.numel()), tis
    arr = np.ascontiguousarray(tis, dtype=np.uint8)
    return ctypes.c_void_p(arr.ctypes.data), int(arr.shape[0]), arr


class Esa:
    """One context = one GPU.  Mirrors the C ABI one to one."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.ctx = ctypes.c_void_p()
        err = ctypes.create_string_buffer(ERRBUF)
        rc = self.lib.esa_init(ctypes.byref(self.ctx), device, err, ERRBUF)
        if rc != 0:
            raise EsaError(rc, err.value.decode())

    def close(self):
        if self.ctx:
            self.lib.esa_shutdown(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, err):
        if rc != 0:
            raise EsaError(rc, err.value.decode())

    @staticmethod
    def _tables(res, copy=True):
        N = res.shard_count  # n + 1 on one GPU, this rank's range of the suffix array in a sharded build

        def arr(ptr, count, dt):
            if not ptr:
                return None
            a = np.ctypeslib.as_array(ptr, shape=(int(count),)) if count else np.zeros(0, dt)
            return a.copy() if copy else a

        return {"suf": arr(res.suftab, N, np.uint64), "lcp": arr(res.lcptab, N, np.uint8),
                "llv": arr(res.llv, 2 * res.nllv, np.uint64) if res.lcptab else None,
                "bwt": arr(res.bwttab, N, np.uint8), "bck": arr(res.bcktab, 2 * res.numofcodes, np.uint64),
                "sti1": arr(res.stitab1, res.text_count, np.uint8), "sti": arr(res.stitab, res.text_count, np.uint64),
                "skp": arr(res.skiptab, N, np.uint64),
                "maxbranchdepth": int(res.maxbranchdepth), "longest": int(res.longest),
                "largelcpvalues": int(res.nllv), "numofcodes": int(res.numofcodes),
                "shard_first": int(res.shard_first), "shard_count": int(res.shard_count),
                "text_first": int(res.text_first), "text_count": int(res.text_count)}

    def build(self, tis, numofchars, prefixlength, tables=("suf", "lcp", "bwt"), want_stats=True, copy=True):
        """esa_build: host buffer in, tables out.  Returns (tables dict, stats dict or None).
        With copy=False the arrays alias the library's pinned buffers until free(tables) is called."""
        ptr, n, keep = _as_ptr(tis)
        res, st = Result(), Stats()
        err = ctypes.create_string_buffer(ERRBUF)
        rc = self.lib.esa_build(self.ctx, ptr, n, numofchars, prefixlength, demand_mask(tables), ctypes.byref(res),
                                ctypes.byref(st) if want_stats else None, err, ERRBUF)
        self._check(rc, err)
        try:
            out = self._tables(res, copy=copy)
        finally:
            if copy:
                self.lib.esa_free(self.ctx, ctypes.byref(res))
        if not copy:
            out["_result"] = res
        return out, (st.as_dict() if want_stats else None)

    def free(self, tables):
        res = tables.pop("_result", None)
        if res is not None:
            self.lib.esa_free(self.ctx, ctypes.byref(res))

    def upload(self, tis, numofchars):
        ptr, n, keep = _as_ptr(tis)
        err = ctypes.create_string_buffer(ERRBUF)
        self._check(self.lib.esa_upload(self.ctx, ptr, n, numofchars, err, ERRBUF), err)

    def construct(self, prefixlength, tables=("suf", "lcp", "bwt"), want_stats=True):
        st = Stats()
        err = ctypes.create_string_buffer(ERRBUF)
        rc = self.lib.esa_construct(self.ctx, prefixlength, demand_mask(tables),
                                    ctypes.byref(st) if want_stats else None, err, ERRBUF)
        self._check(rc, err)
        return st.as_dict() if want_stats else None

    def download(self):
        res = Result()
        err = ctypes.create_string_buffer(ERRBUF)
        self._check(self.lib.esa_download(self.ctx, ctypes.byref(res), err, ERRBUF), err)
        try:
            return self._tables(res)
        finally:
            self.lib.esa_free(self.ctx, ctypes.byref(res))

    def debug_radix_sort(self, keys, vals, begin_bit, end_bit):
        """Testing hook: sorts the numpy arrays in place (keys uint32/uint64, vals uint32 or None)."""
        assert keys.dtype in (np.uint32, np.uint64) and keys.flags.c_contiguous
        err = ctypes.create_string_buffer(ERRBUF)
        rc = self.lib.esa_debug_radix_sort(self.ctx, keys.dtype.itemsize, keys.ctypes.data,
                                           vals.ctypes.data if vals is not None else None, len(keys), begin_bit,
                                           end_bit, err, ERRBUF)
        self._check(rc, err)

    def check(self):
        err = ctypes.create_string_buffer(ERRBUF)
        self._check(self.lib.esa_check(self.ctx, err, ERRBUF), err)

    def childtab(self, lcp=None, llv=None, n=None, cld1=True):
        """esa_childtab: (cld bytes [3 per entry], cld1 bytes or None).  lcp=None: from the lcp table of the last build
        on this context (n = its totallength); else from host tables lcp[n+1] (uint8) and llv (uint64 pairs)."""
        err = ctypes.create_string_buffer(ERRBUF)
        if lcp is None:
            lp, vp_, nllv, n = None, None, 0, int(n)
        else:
            lcp = np.ascontiguousarray(lcp, np.uint8)
            llv = np.ascontiguousarray(np.zeros(0, np.uint64) if llv is None else llv, np.uint64).ravel()
            n, nllv = len(lcp) - 1, len(llv) // 2
            lp = ctypes.c_void_p(lcp.ctypes.data)
            vp_ = ctypes.c_void_p(llv.ctypes.data) if nllv else None
        cld = np.empty(3 * (n + 1), np.uint8)
        c1 = np.empty(n + 1, np.uint8) if cld1 else None
        rc = self.lib.esa_childtab(self.ctx, lp, vp_, nllv, n, ctypes.c_void_p(cld.ctypes.data),
                                   ctypes.c_void_p(c1.ctypes.data) if cld1 else None, err, ERRBUF)
        self._check(rc, err)
        return cld, c1

    def debug_corrupt(self, table, index, xor_mask):
        """Testing hook: damages one element of a resident device table ("llv" addresses the .llv words)."""
        bit = (1 << 31) if table == "llv" else DEMAND_BITS[table]
        err = ctypes.create_string_buffer(ERRBUF)
        self._check(self.lib.esa_debug_corrupt(self.ctx, bit, index, xor_mask, err, ERRBUF), err)

    # ---- sharded builds: attach a communicator, then upload/construct/download/build are collective ----
    def attach_nccl(self, unique_id, nranks, rank):
        """One process per GPU: `unique_id` comes from comm_unique_id() on rank 0 (distribute it yourself)."""
        assert len(unique_id) == COMM_ID_BYTES
        buf = ctypes.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        err = ctypes.create_string_buffer(ERRBUF)
        self._check(self.lib.esa_comm_attach_nccl(self.ctx, buf, nranks, rank, err, ERRBUF), err)

    def detach(self):
        self.lib.esa_comm_detach(self.ctx)


COMM_ID_BYTES = 128


def comm_unique_id():
    lib = load_library()
    buf = ctypes.create_string_buffer(COMM_ID_BYTES)
    err = ctypes.create_string_buffer(ERRBUF)
    rc = lib.esa_comm_unique_id(buf, err, ERRBUF)
    if rc != 0:
        raise EsaError(rc, err.value.decode())
    return buf.raw


def merge_shards(parts):
    """Concatenates the per-rank tables of a sharded build in rank order: the tables of the whole index."""
    parts = sorted(parts, key=lambda t: t["shard_first"])
    pos = 0
    for t in parts:
        assert t["shard_first"] == pos, "ranges of the ranks do not tile the suffix array"
        pos += t["shard_count"]
    out = dict(parts[0])
    for name in ("suf", "lcp", "bwt", "llv"):
        if parts[0].get(name) is not None:
            out[name] = np.concatenate([t[name] for t in parts])
    tpos = 0
    if any(parts[0].get(name) is not None for name in ("sti1", "sti")):
        for t in parts:  # the tables indexed by text position are cut into slices of the text, in rank order as well
            assert t["text_first"] == tpos or t["text_count"] == 0, "text slices of the ranks do not tile the text"
            tpos += t["text_count"]
        for name in ("sti1", "sti"):
            if parts[0].get(name) is not None:
                out[name] = np.concatenate([t[name] for t in parts])
    out["largelcpvalues"] = sum(t["largelcpvalues"] for t in parts)
    out["shard_first"], out["shard_count"] = 0, pos
    out["text_first"], out["text_count"] = 0, tpos
    return out


class EsaGroup:
    """Several ranks inside one process, one thread per rank (transport B of include/vstree_b200.h).  `devices` lists
    the CUDA device of every rank; ranks may share a device (that is how the single-GPU test box exercises the
    sharded path)."""

    def __init__(self, devices):
        import threading
        self._threading = threading
        self.lib = load_library()
        self.nranks = len(devices)
        self.group = ctypes.c_void_p()
        err = ctypes.create_string_buffer(ERRBUF)
        rc = self.lib.esa_group_create(ctypes.byref(self.group), self.nranks, err, ERRBUF)
        if rc != 0:
            raise EsaError(rc, err.value.decode())
        self.ranks = [Esa(d) for d in devices]
        for r, e in enumerate(self.ranks):
            e._check(self.lib.esa_comm_attach_group(e.ctx, self.group, r, err, ERRBUF), err)

    def close(self):
        for e in self.ranks:
            e.close()
        self.ranks = []
        if self.group:
            self.lib.esa_group_destroy(self.group)
            self.group = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _collective(self, fn):
        """Runs fn(rank, esa) on one thread per rank (ctypes releases the GIL inside the library)."""
        results, errors = [None] * self.nranks, [None] * self.nranks

        def work(r):
            try:
                results[r] = fn(r, self.ranks[r])
            except BaseException as exc:  # noqa: BLE001 -- reported below, after waking the other ranks
                errors[r] = exc
                self.lib.esa_group_abort(self.group)

        threads = [self._threading.Thread(target=work, args=(r,)) for r in range(self.nranks)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if any(e is not None for e in errors):
            self.lib.esa_group_reset(self.group)
        for e in errors:
            if e is not None and not (isinstance(e, EsaError) and e.code == -9):
                raise e
        for e in errors:
            if e is not None:
                raise e
        return results

    def build(self, tis, numofchars, prefixlength, tables=("suf", "lcp", "bwt"), merge=True):
        """Collective esa_build.  Returns (tables of the whole index or the per-rank list, per-rank stats)."""
        out = self._collective(lambda r, e: e.build(tis, numofchars, prefixlength, tables))
        parts, stats = [o[0] for o in out], [o[1] for o in out]
        return (merge_shards(parts) if merge else parts), stats

    def upload(self, tis, numofchars):
        self._collective(lambda r, e: e.upload(tis, numofchars))

    def construct(self, prefixlength, tables=("suf", "lcp", "bwt")):
        return self._collective(lambda r, e: e.construct(prefixlength, tables))

    def download(self, merge=True):
        parts = self._collective(lambda r, e: e.download())
        return merge_shards(parts) if merge else parts
