"""Sharded builds on the single-GPU test box: several ranks, one thread each, share device 0 through
the in-process transport (include/vstree_b200.h, transport B).  The exchange logic — splitter pass, all-to-all of the
suffixes, rank requests between the ranges, halo, merged scalars — is the same code the NCCL transport drives; the
NCCL path itself is covered by tests/test_gpu_nccl.py on a multi-GPU box."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_util as ou
from golden_inputs import CASES, make_input
from vstree_b200 import Esa, EsaError, synth
from vstree_b200.capi import EsaGroup

pytestmark = pytest.mark.gpu

TABLES = ("suf", "lcp", "llv", "bwt", "bck")
SCALARS = ("longest", "maxbranchdepth", "largelcpvalues")
SHARDED = ("suf", "lcp", "bwt", "bck")


def first_diff(a, b):
    n = min(len(a), len(b))
    d = np.flatnonzero(a[:n] != b[:n])
    if len(d):
        i = int(d[0])
        return f"first difference at {i}: got {a[max(0, i - 2):i + 3].tolist()} want {b[max(0, i - 2):i + 3].tolist()} ({len(d)} differ)"
    return f"lengths differ: {len(a)} vs {len(b)}"


def assert_same(got, want, what):
    for name in TABLES:
        if want.get(name) is None:
            continue
        g, w = got[name], np.asarray(want[name])
        assert g is not None, f"{what}: {name} missing"
        assert len(g) == len(w) and np.array_equal(g, w), f"{what}: {name}: {first_diff(g, w)}"
    for name in SCALARS:
        if name in want:
            assert got[name] == want[name], f"{what}: {name}: got {got[name]} want {want[name]}"


@pytest.fixture(scope="module", params=[2, 3, 8])
def group(request):
    g = EsaGroup([0] * request.param)
    yield g
    g.close()


@pytest.mark.parametrize("name", list(CASES.keys()))
def test_golden_cases_sharded(group, name):
    """Every golden input, cut into 2, 3 and 8 ranges: the concatenated tables are the reference's."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")) as f:
        g = {c["name"]: c for c in json.load(f)["cases"]}[name]
    tis, sigma = make_input(name)
    k = g["prefixlength"]
    got, stats = group.build(tis, sigma, k, tables=SHARDED)
    want = ou.oracle_build(tis, sigma, k)
    assert_same(got, want, f"{name} on {group.nranks} ranks")
    for t in TABLES:
        assert hashlib.sha256(np.ascontiguousarray(got[t]).tobytes()).hexdigest()[:16] == g["sha"][t], t
    assert sum(s["shard_count"] for s in stats) == len(tis) + 1


@pytest.mark.parametrize("seed", range(10))
def test_random_small_inputs_sharded(group, seed):
    rng = np.random.default_rng(5000 + seed)
    n = int(rng.choice([0, 1, 2, 17, 255, 4095, 4096, 4097, 8200, 20000, 65537]))
    sigma = int(rng.choice([2, 4, 5, 20, 200]))
    if n == 0:
        tis = np.zeros(0, np.uint8)
    else:
        tis = rng.integers(0, sigma, size=n, dtype=np.uint8)
        sp = rng.random(n) < rng.choice([0.0, 0.01, 0.3])
        tis[sp] = rng.choice([254, 255], size=int(sp.sum()))
        if rng.random() < 0.5 and n > 40:
            tis[n // 2:n // 2 + n // 3] = tis[:n // 3]
    lib = ou.oracle_lib()
    kmax = lib.oracle_maximalprefixlength(sigma, n)
    k = int(rng.integers(1, max(1, kmax) + 1))
    got, _ = group.build(tis, sigma, k, tables=SHARDED)
    want = ou.oracle_build(tis, sigma, k)
    assert_same(got, want, f"seed {seed} n={n} sigma={sigma} k={k} ranks={group.nranks}")


def test_sharded_equals_single_gpu_on_a_larger_input(group):
    """8 Mbp of human-like DNA (repeats, N runs, several records): ranges of several ranks against one GPU."""
    tis = synth.human_like_dna(8_000_000, seed=5, records=7)
    with Esa(0) as one:
        want, _ = one.build(tis, 4, 9, tables=SHARDED)
    got, stats = group.build(tis, 4, 9, tables=SHARDED)
    assert_same(got, want, f"human-like on {group.nranks} ranks")
    sizes = [s["shard_count"] for s in stats]
    assert max(sizes) < 2.0 * (len(tis) + 1) / group.nranks, f"ranges badly balanced: {sizes}"
    assert all(s["comm_bytes_sent"] > 0 for s in stats)
    want_direct = os.environ.get("ESA_COMM_DIRECT", "1") != "0"
    assert all(bool(s["comm_direct"]) == want_direct for s in stats)


def test_sharded_protein_and_staged_calls(group):
    tis = synth.random_protein(3_000_000, seed=17, records=2000, xfrac=0.0005)
    with Esa(0) as one:
        want, _ = one.build(tis, 20, 4, tables=SHARDED)
    group.upload(tis, 20)
    for _ in range(2):  # same resident text, arenas reused
        stats = group.construct(4, tables=SHARDED)
    got = group.download()
    assert_same(got, want, f"protein on {group.nranks} ranks")
    assert stats[0]["kernel"]["comm"]["launches"] > 0


def test_sharded_tables_by_text_position(group):
    """.sti1 and the full inverse suffix array of a sharded build: slices of the text in rank order."""
    for tis, sigma, k in ((synth.dna_with_specials(70000, seed=23, nfrac=0.004, mean_run=6, records=11), 4, 6),
                          (synth.random_protein(40000, seed=24, records=30, xfrac=0.002), 20, 2),
                          (np.array([0, 1, 2, 3, 254, 0, 1, 2, 255, 0, 1, 2, 3, 255, 254, 254, 0, 1], np.uint8), 4, 1)):
        want = ou.oracle_build(tis, sigma, k)
        got, _ = group.build(tis, sigma, k, tables=("suf", "lcp", "bwt", "bck", "sti1", "sti"))
        assert_same(got, want, f"sigma {sigma} on {group.nranks} ranks")
        assert np.array_equal(got["sti1"], want["sti1"]), first_diff(got["sti1"], want["sti1"])
        inv = np.zeros(len(tis) + 1, np.uint64)
        inv[want["suf"]] = np.arange(len(tis) + 1, dtype=np.uint64)
        assert np.array_equal(got["sti"], inv)


def test_sharded_context_refuses_the_skip_table(group):
    tis = synth.random_dna(50000, seed=1)
    with pytest.raises(EsaError) as ei:
        group.build(tis, 4, 5, tables=("suf", "lcp", "skp"))
    assert "sharded" in str(ei.value)
    # the group survives a refused build
    got, _ = group.build(tis, 4, 5, tables=("suf",))
    assert int(got["suf"].sum()) == 50001 * 50000 // 2


def test_upload_failure_on_one_rank_wakes_the_others(monkeypatch):
    """A rank that fails inside the collective upload must abort the group by itself (esa_upload / esa_build call
    esa_group_abort): the C driver runs one thread per rank and joins them -- nobody may wait forever.  The threads here
    call the C ABI directly, without the Python wrapper's own abort."""
    import ctypes
    import threading
    from vstree_b200.capi import ERRBUF, Result, Stats, load_library
    lib = load_library()
    tis = synth.random_dna(200000, seed=2)
    for call in ("upload", "build"):
        with EsaGroup([0, 0, 0]) as grp:
            monkeypatch.setenv("ESA_DEBUG_FAIL_UPLOAD_RANK", "1")
            codes, msgs = [None] * 3, [None] * 3

            def work(r):
                err = ctypes.create_string_buffer(ERRBUF)
                ptr = ctypes.c_void_p(tis.ctypes.data)
                if call == "upload":
                    codes[r] = lib.esa_upload(grp.ranks[r].ctx, ptr, len(tis), 4, err, ERRBUF)
                else:
                    res, st = Result(), Stats()
                    codes[r] = lib.esa_build(grp.ranks[r].ctx, ptr, len(tis), 4, 5, 1 << 5, ctypes.byref(res),
                                             ctypes.byref(st), err, ERRBUF)
                msgs[r] = err.value.decode()

            threads = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(3)]
            for t in threads:
                t.start()
            for t in threads:
                t.join(timeout=60)
            assert not any(t.is_alive() for t in threads), f"{call}: ranks hang after a failed upload on rank 1"
            assert codes[1] == -4 and "injected" in msgs[1]
            assert codes[0] == -9 and codes[2] == -9, (codes, msgs)
            # the group is usable again after a reset
            monkeypatch.delenv("ESA_DEBUG_FAIL_UPLOAD_RANK")
            lib.esa_group_reset(grp.group)
            got, _ = grp.build(tis, 4, 5, tables=("suf",))
            assert int(got["suf"].sum()) == 200001 * 200000 // 2
