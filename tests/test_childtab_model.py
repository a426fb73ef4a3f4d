"""Child table (.cld / .cld1) on the CPU:

1. the oracle's restatement of the reference's stack walks (oracle/esa_oracle.c: oracle_childtab, after
   src/kurtz/mkcld.c:41-284) against the files the UNMODIFIED reference tools wrote (tests/golden/cld_golden.json, made by
   oracle/_ref/mkvtree.x + oracle/_ref/mkcld.x through tests/golden/make_cld_golden.py);
2. the per-entry functions the CUDA kernels run per thread (vstree_b200/csrc/childtab_core.cuh), compiled for the host
   by g++ (tests/model/cld_model.cpp), against the oracle on lcp-shaped and adversarial arrays, and their three searches
   against brute force.

The GPU tests proper are tests/test_gpu_childtab.py."""
import ctypes
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_util as ou
from golden_inputs import CASES, make_input

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MODEL_SRC = os.path.join(HERE, "model", "cld_model.cpp")
MODEL_SO = os.path.join(ROOT, "oracle", "_build", "libcld_model.so")
CORE = os.path.join(ROOT, "vstree_b200", "csrc", "childtab_core.cuh")

with open(os.path.join(HERE, "golden", "cld_golden.json")) as f:
    GOLDEN = {c["name"]: c for c in json.load(f)["cases"]}


def sha(b):
    return hashlib.sha256(b).hexdigest()[:16]


_model = None


def model():
    global _model
    if _model is None:
        newest = max(os.path.getmtime(MODEL_SRC), os.path.getmtime(CORE))
        if not os.path.exists(MODEL_SO) or os.path.getmtime(MODEL_SO) < newest:
            os.makedirs(os.path.dirname(MODEL_SO), exist_ok=True)
            subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-fPIC", "-shared", "-o", MODEL_SO, MODEL_SRC], check=True)
        lib = ctypes.CDLL(MODEL_SO)
        u8p, u32p, u64p = (ctypes.POINTER(t) for t in (ctypes.c_uint8, ctypes.c_uint32, ctypes.c_uint64))
        lib.cld_model_tables.restype = None
        lib.cld_model_tables.argtypes = [u32p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, u8p, u8p, u64p]
        lib.cld_model_queries.restype = None
        lib.cld_model_queries.argtypes = [u32p, ctypes.c_uint64, u64p, ctypes.c_uint64, u64p]
        _model = lib
    return _model


def model_tables(L, threads=256, near=32, count=None, near32=False):
    """The tables as child_kernel computes them: blocks of `threads` entries, the bounded scan over `near` entries on
    either side first (CLD_THREADS / CLD_NEAR of childtab.cu), the pyramid search for the rest; near=0: pyramid only;
    near32: the -DCLD_NEAR32 experiment variant of the bounded scan."""
    L = np.ascontiguousarray(L, np.uint32)
    N = len(L)
    cld, cld1 = np.zeros(3 * N, np.uint8), np.zeros(N, np.uint8)
    settled = ctypes.c_uint64(0)
    model().cld_model_tables(L.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), N, threads, near, int(near32),
                             cld.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                             cld1.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), ctypes.byref(settled))
    if count is not None:
        count.append(settled.value)
    return cld, cld1


def split_lcp(L):
    """exact values -> (.lcp bytes, .llv pairs) as the index stores them (bese.c:553-566)"""
    L = np.asarray(L, np.uint64)
    big = np.nonzero(L >= 255)[0]
    llv = np.empty(2 * len(big), np.uint64)
    llv[0::2] = big
    llv[1::2] = L[big]
    return np.minimum(L, 255).astype(np.uint8), llv


def exact_lcp(lcp, llv):
    L = np.asarray(lcp, np.uint32).copy()
    pairs = np.asarray(llv, np.uint64).reshape(-1, 2)
    L[pairs[:, 0].astype(np.int64)] = pairs[:, 1].astype(np.uint32)
    return L


def shaped_arrays():
    """name -> exact lcp-like array with L[0] = 0"""
    rng = np.random.default_rng(2024)
    out = {}
    for N in (1, 2, 3, 15, 16, 17, 18, 255, 256, 257, 258, 4095, 4096, 4097, 65537):
        out[f"uniform4_{N}"] = rng.integers(0, 4, N)
        out[f"uniform300_{N}"] = rng.integers(0, 300, N)
    N = 70001
    out["constant0"] = np.zeros(3000, np.int64)
    out["constant7"] = np.full(3000, 7)
    out["rising"] = np.arange(N) % 100000
    out["falling"] = np.arange(N)[::-1].copy()
    out["sawtooth_up_17"] = np.arange(N) % 17
    out["sawtooth_down_300"] = 299 - (np.arange(N) % 300)
    out["plateaus"] = np.repeat(rng.integers(0, 50, N // 700 + 1), 700)[:N]
    out["long_intervals"] = np.where(np.arange(N) % 1000 == 0, 1, 5 + rng.integers(0, 3, N))  # distances >= 255
    out["nested_deep"] = np.minimum(np.arange(N), np.arange(N)[::-1]) // 3  # one pyramid of nested intervals
    out["large_values"] = np.where(rng.random(N) < 0.1, rng.integers(255, 100000, N), rng.integers(0, 40, N))
    out["geometric"] = rng.geometric(0.2, N) - 1
    walk = np.abs(np.cumsum(rng.integers(-3, 4, N)))
    out["walk"] = walk
    for k in out:
        a = np.asarray(out[k], np.int64).copy()
        a[0] = 0
        out[k] = a
    return out


SHAPED = shaped_arrays()


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_childtab_against_reference_files(name):
    """oracle_esa_build's lcp/llv (pinned by tests/test_oracle.py) -> oracle_childtab == what mkcld.x wrote"""
    g = GOLDEN[name]
    tis, sigma = make_input(name)
    _, _, flag, pl = CASES[name]
    k = pl if pl else ou.oracle_lib().oracle_recommendedprefixlength(sigma, len(tis))
    o = ou.oracle_build(tis, sigma, k, bck=False, sti1=False)
    assert sha(o["lcp"].tobytes()) == g["sha"]["lcp"] and sha(o["llv"].tobytes()) == g["sha"]["llv"]
    cld, cld1 = ou.oracle_childtab(o["lcp"], o["llv"])
    assert sha(cld.tobytes()) == g["sha"]["cld"], "the .cld records differ from the reference's file"
    assert sha(cld1.tobytes()) == g["sha"]["cld1"], "the .cld1 bytes differ from the reference's file"
    if "tables" in g:
        assert cld.tolist() == g["tables"]["cld"] and cld1.tolist() == g["tables"]["cld1"]
    # the same tables from the per-entry functions of the kernels
    near = []
    for kw in ({"count": near}, {"near": 0}, {"threads": 16, "near": 3}, {"near32": True}):
        m_cld, m_cld1 = model_tables(exact_lcp(o["lcp"], o["llv"]), **kw)
        assert np.array_equal(m_cld, cld) and np.array_equal(m_cld1, cld1), kw
    if len(tis) >= 20000 and "poly" not in name:
        assert near[0] > 0.8 * len(tis), f"the bounded scan settles only {near[0]} of {len(tis)} entries"


@pytest.mark.parametrize("name", list(SHAPED))
def test_kernel_functions_against_the_stack_walk(name):
    L = SHAPED[name]
    lcp, llv = split_lcp(L)
    cld, cld1 = ou.oracle_childtab(lcp, llv)
    for kw in ({}, {"near": 0}, {"threads": 16, "near": 3}, {"threads": 64, "near": 1}, {"threads": 32, "near": 64},
               {"near32": True}, {"threads": 16, "near": 3, "near32": True}, {"threads": 64, "near": 1, "near32": True}):
        m_cld, m_cld1 = model_tables(L, **kw)
        bad = np.nonzero(m_cld != cld)[0]
        assert len(bad) == 0, f"{name} {kw}: record {bad[0] // 3} field {bad[0] % 3}: {m_cld[bad[0]]} != {cld[bad[0]]}"
        assert np.array_equal(m_cld1, cld1), kw
    c = cld.reshape(-1, 3)
    if name in ("long_intervals", "nested_deep", "falling"):
        assert (c == 255).any(), "the case is meant to reach LARGECHILDVALUE"


@pytest.mark.parametrize("name", ["uniform4_4097", "uniform300_65537", "plateaus", "nested_deep", "walk", "sawtooth_up_17"])
def test_searches_against_brute_force(name):
    L = np.ascontiguousarray(SHAPED[name], np.uint32)
    N = len(L)
    rng = np.random.default_rng(7)
    q = []
    for _ in range(1500):
        i = int(rng.integers(0, N))
        v = int(L[i]) if rng.random() < 0.7 else int(rng.integers(0, int(L.max()) + 2))
        q.append((0, i, v))
        q.append((1, i, v))
        a = int(rng.integers(0, N))
        b = int(min(N - 1, a + (rng.integers(0, 40) if rng.random() < 0.6 else rng.integers(0, N))))
        q.append((2, a, b))
    q += [(2, 0, N - 1), (0, 0, 0), (1, N - 1, 0), (1, 0, 5), (0, N - 1, 5)]
    qa = np.array(q, np.uint64)
    ans = np.zeros(len(q), np.uint64)
    model().cld_model_queries(L.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), N,
                              qa.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(q),
                              ans.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
    NONE = (1 << 64) - 1
    for (kind, a, b), got in zip(q, ans.tolist()):
        if kind == 0:
            hit = np.nonzero(L[a + 1:] <= b)[0]
            want = a + 1 + int(hit[0]) if len(hit) else N
        elif kind == 1:
            hit = np.nonzero(L[:a] <= b)[0]
            want = int(hit[-1]) if len(hit) else NONE
        else:
            want = a + int(np.argmin(L[a:b + 1]))
        assert got == want, (name, kind, a, b, got, want)


MKCLD = os.path.join(ROOT, "vstree_b200", "bin", "mkcld")


def test_mkcld_tool_without_a_gpu(tmp_path):
    """argument errors need no device; a well-formed index does, and the tool says so instead of computing on the CPU"""
    pr = subprocess.run([MKCLD], cwd=tmp_path, capture_output=True, text=True)
    assert pr.returncode == 1 and pr.stderr.startswith(f"Usage: {MKCLD} indexname\n")
    pr = subprocess.run([MKCLD, "nosuch"], cwd=tmp_path, capture_output=True, text=True)
    assert pr.returncode == 1 and pr.stderr == f'{MKCLD}: cannot open file "nosuch.prj": No such file or directory\n'
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lcp, llv = split_lcp(SHAPED["uniform300_4097"])
    (tmp_path / "idx.prj").write_text(f"totallength={len(lcp) - 1}\nlargelcpvalues={len(llv) // 2}\n")
    (tmp_path / "idx.lcp").write_bytes(lcp.tobytes())
    (tmp_path / "idx.llv").write_bytes(llv.tobytes())
    (tmp_path / "idx.suf").write_bytes(b"")
    pr = subprocess.run([MKCLD, "idx"], cwd=tmp_path, capture_output=True, text=True)
    assert pr.returncode == 1 and "no CPU fallback" in pr.stderr and not (tmp_path / "idx.cld").exists()
