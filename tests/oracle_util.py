"""Test-side access to the checker: the C restatement in oracle/ and the reference binary in oracle/_ref/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes
import os
import shutil
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libesa_oracle.so")
REF_MKVTREE = os.path.join(ORACLE_DIR, "_ref", "mkvtree.x")
REF_MKVTREE_DBG = os.path.join(ORACLE_DIR, "_ref", "mkvtree.dbg.x")

_lib = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def oracle_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
                os.path.join(ORACLE_DIR, "esa_oracle.c")):
            build_oracle()
        lib = ctypes.CDLL(ORACLE_SO)
        u8p, u64p = ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint64)
        lib.oracle_esa_build.restype = ctypes.c_int
        lib.oracle_esa_build.argtypes = [u8p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, u64p, u8p, u64p,
                                         u64p, u64p, u64p, u8p, u64p, u8p]
        lib.oracle_check_tables.restype = ctypes.c_int
        lib.oracle_check_tables.argtypes = [u8p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, u64p, u8p, u64p,
                                            ctypes.c_uint64, u8p, u64p, u8p]
        lib.oracle_skiptable.restype = ctypes.c_int
        lib.oracle_skiptable.argtypes = [u8p, u64p, ctypes.c_uint64, u64p]
        lib.oracle_childtab.restype = ctypes.c_int
        lib.oracle_childtab.argtypes = [u8p, u64p, ctypes.c_uint64, u8p, u8p]
        lib.oracle_numofcodes.restype = ctypes.c_uint64
        lib.oracle_numofcodes.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        lib.oracle_recommendedprefixlength.restype = ctypes.c_uint32
        lib.oracle_recommendedprefixlength.argtypes = [ctypes.c_uint32, ctypes.c_uint64]
        lib.oracle_maximalprefixlength.restype = ctypes.c_uint32
        lib.oracle_maximalprefixlength.argtypes = [ctypes.c_uint32, ctypes.c_uint64]
        _lib = lib
    return _lib


def _p(arr, ctype):
    return arr.ctypes.data_as(ctypes.POINTER(ctype)) if arr is not None else None


def oracle_build(tis, numofchars, prefixlength, bck=True, sti1=True):
    """Run the C restatement.  tis: uint8 array (codes, 254, 255).  Returns a dict of numpy tables."""
    lib = oracle_lib()
    tis = np.ascontiguousarray(tis, dtype=np.uint8)
    n = int(tis.shape[0])
    N = n + 1
    tisp = np.concatenate([tis, np.zeros(1, np.uint8)])  # readable byte behind the text; never inspected
    suf = np.zeros(N, np.uint64)
    lcp = np.zeros(N, np.uint8)
    llv = np.zeros(2 * N, np.uint64)
    bwt = np.zeros(N, np.uint8)
    nllv, maxbd, longest = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
    ncodes = lib.oracle_numofcodes(numofchars, prefixlength) if prefixlength > 0 else 0
    bckt = np.zeros(2 * ncodes, np.uint64) if (bck and prefixlength > 0) else None
    sti = np.zeros(N, np.uint8) if (sti1 and prefixlength > 0) else None
    rc = lib.oracle_esa_build(_p(tisp, ctypes.c_uint8), n, numofchars, prefixlength, _p(suf, ctypes.c_uint64),
                              _p(lcp, ctypes.c_uint8), _p(llv, ctypes.c_uint64), ctypes.byref(nllv),
                              ctypes.byref(maxbd), ctypes.byref(longest), _p(bwt, ctypes.c_uint8),
                              _p(bckt, ctypes.c_uint64), _p(sti, ctypes.c_uint8))
    if rc != 0:
        raise RuntimeError(f"oracle_esa_build failed with {rc}")
    return {"suf": suf, "lcp": lcp, "llv": llv[:2 * nllv.value].copy(), "bwt": bwt, "bck": bckt, "sti1": sti,
            "maxbranchdepth": maxbd.value, "longest": longest.value, "largelcpvalues": nllv.value}


def oracle_skiptable(lcp, llv):
    """Skip table from the oracle's lcp/llv (restates src/kurtz/mkskip.c)."""
    lib = oracle_lib()
    lcp = np.ascontiguousarray(lcp, np.uint8)
    llv = np.ascontiguousarray(np.concatenate([llv, np.zeros(2, np.uint64)]), np.uint64)
    n = len(lcp) - 1
    skp = np.zeros(n + 1, np.uint64)
    rc = lib.oracle_skiptable(_p(lcp, ctypes.c_uint8), _p(llv, ctypes.c_uint64), n, _p(skp, ctypes.c_uint64))
    if rc != 0:
        raise RuntimeError("oracle_skiptable failed")
    return skp


def oracle_childtab(lcp, llv):
    """Child table (.cld records, .cld1 bytes) from lcp/llv (restates src/kurtz/mkcld.c)."""
    lib = oracle_lib()
    lcp = np.ascontiguousarray(lcp, np.uint8)
    llv = np.ascontiguousarray(np.concatenate([np.asarray(llv, np.uint64).ravel(), np.zeros(2, np.uint64)]), np.uint64)
    n = len(lcp) - 1
    cld = np.zeros(3 * (n + 1), np.uint8)
    cld1 = np.zeros(n + 1, np.uint8)
    rc = lib.oracle_childtab(_p(lcp, ctypes.c_uint8), _p(llv, ctypes.c_uint64), n, _p(cld, ctypes.c_uint8),
                             _p(cld1, ctypes.c_uint8))
    if rc != 0:
        raise RuntimeError("oracle_childtab failed")
    return cld, cld1


def oracle_check(tis, numofchars, prefixlength, tables):
    """DEBUG-invariant property check (0 = consistent)."""
    lib = oracle_lib()
    tis = np.ascontiguousarray(tis, dtype=np.uint8)
    n = int(tis.shape[0])
    tisp = np.concatenate([tis, np.zeros(1, np.uint8)])
    llv = tables.get("llv")
    nllv = 0 if llv is None else len(llv) // 2
    if llv is not None and len(llv) == 0:
        llv = np.zeros(2, np.uint64)
    g = lambda k, dt: (np.ascontiguousarray(tables[k], dtype=dt) if tables.get(k) is not None else None)
    suf, lcp, bwt, bck, sti = g("suf", np.uint64), g("lcp", np.uint8), g("bwt", np.uint8), g("bck", np.uint64), g(
        "sti1", np.uint8)
    llv = np.ascontiguousarray(llv, dtype=np.uint64) if llv is not None else None
    return lib.oracle_check_tables(_p(tisp, ctypes.c_uint8), n, numofchars, prefixlength, _p(suf, ctypes.c_uint64),
                                   _p(lcp, ctypes.c_uint8), _p(llv, ctypes.c_uint64), nllv, _p(bwt, ctypes.c_uint8),
                                   _p(bck, ctypes.c_uint64), _p(sti, ctypes.c_uint8))


def have_ref():
    return os.access(REF_MKVTREE, os.X_OK)


TABLE_DTYPES = {"suf": np.uint64, "lcp": np.uint8, "llv": np.uint64, "bwt": np.uint8, "bck": np.uint64,
                "sti1": np.uint8, "tis": np.uint8, "ois": np.uint8, "sds": np.uint64, "ssp": np.uint64, "skp": np.uint64}


def parse_prj(text):
    out = {}
    for line in text.splitlines():
        if "=" in line:
            k, v = line.split("=", 1)
            if k in ("dbfile", "queryfile"):
                out.setdefault(k, []).append(v)
            else:
                out[k] = int(v)
    return out


def run_mkvtree(binary, dbfile, args, indexname="idx", env=None, keepdir=None):
    """Run a mkvtree binary (reference or ours) in a scratch directory; return (returncode, stdout, stderr, files).

    files maps the index file suffix to its raw bytes."""
    work = keepdir or tempfile.mkdtemp(prefix="mkv_")
    try:
        local = os.path.join(work, os.path.basename(dbfile))
        if os.path.abspath(dbfile) != os.path.abspath(local):
            shutil.copyfile(dbfile, local)
        cmd = [binary, "-db", os.path.basename(dbfile), "-indexname", indexname] + list(args)
        e = dict(os.environ)
        if env:
            e.update(env)
        pr = subprocess.run(cmd, cwd=work, capture_output=True, env=e)
        files = {}
        for fn in os.listdir(work):
            if fn.startswith(indexname + "."):
                with open(os.path.join(work, fn), "rb") as f:
                    files[fn[len(indexname) + 1:]] = f.read()
        return pr.returncode, pr.stdout.decode(errors="replace"), pr.stderr.decode(errors="replace"), files
    finally:
        if keepdir is None:
            shutil.rmtree(work, ignore_errors=True)


def tables_from_files(files):
    out = {}
    for k, dt in TABLE_DTYPES.items():
        if k in files:
            out[k] = np.frombuffer(files[k], dtype=dt)
    if "prj" in files:
        out["prj"] = parse_prj(files["prj"].decode())
    return out
