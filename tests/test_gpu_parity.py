"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle (C restatement, committed goldens
made by the reference binary, and the reference binary itself when oracle/_ref travelled to the box)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_util as ou
from golden_inputs import CASES, make_input
from vstree_b200 import Esa, EsaError, synth

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden.json")
ALL = ("suf", "lcp", "bwt", "bck", "sti1")
TABLES = ("suf", "lcp", "llv", "bwt", "bck", "sti1")
SCALARS = ("longest", "maxbranchdepth", "largelcpvalues")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def esa():
    e = Esa(0)
    yield e
    e.close()


def golden():
    with open(GOLDEN) as f:
        return {c["name"]: c for c in json.load(f)["cases"]}


def first_diff(a, b):
    n = min(len(a), len(b))
    d = np.flatnonzero(a[:n] != b[:n])
    if len(d):
        i = int(d[0])
        return f"first difference at {i}: got {a[max(0, i - 2):i + 3].tolist()} want {b[max(0, i - 2):i + 3].tolist()} ({len(d)} differ)"
    return f"lengths differ: {len(a)} vs {len(b)}"


def assert_tables_equal(got, want, what=""):
    for name in TABLES:
        if want.get(name) is None:
            continue
        g, w = got[name], np.asarray(want[name])
        assert g is not None, f"{what}: {name} missing"
        assert len(g) == len(w) and np.array_equal(g, w), f"{what}: {name}: {first_diff(g, w)}"
    for name in SCALARS:
        if name in want:
            assert got[name] == want[name], f"{what}: {name}: got {got[name]} want {want[name]}"


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,bits,n", [(np.uint32, (0, 32), 1), (np.uint32, (0, 32), 1000), (np.uint32, (0, 31), 8192),
                                          (np.uint32, (0, 32), 8193), (np.uint32, (3, 17), 300001),
                                          (np.uint32, (0, 8), 70000), (np.uint32, (0, 32), 3000000),
                                          (np.uint64, (0, 64), 6144), (np.uint64, (0, 54), 500000),
                                          (np.uint64, (5, 50), 6145), (np.uint64, (0, 64), 2500000)])
def test_radix_sort_matches_numpy_stable_sort(esa, dtype, bits, n):
    rng = np.random.default_rng(n)
    hi = np.iinfo(dtype).max
    keys = rng.integers(0, hi, size=n, dtype=dtype, endpoint=True)
    if n > 5000:  # plant heavy duplicates and a constant digit
        keys[: n // 3] &= dtype(0xFF00FF)
    vals = np.arange(n, dtype=np.uint32)
    b, e = bits
    mask = ((1 << (e - b)) - 1) << b
    order = np.argsort((keys & dtype(mask)), kind="stable")
    k2, v2 = keys.copy(), vals.copy()
    esa.debug_radix_sort(k2, v2, b, e)
    assert np.array_equal(v2, order.astype(np.uint32)), first_diff(v2, order.astype(np.uint32))
    assert np.array_equal(k2, keys[order])
    k3 = keys.copy()
    esa.debug_radix_sort(k3, None, b, e)
    assert np.array_equal(k3 & dtype(mask), (keys & dtype(mask))[order])


def test_radix_sort_skewed(esa):
    n = 1 << 20
    keys = np.zeros(n, np.uint32)
    keys[::7] = 1
    keys[5::1001] = 0xFFFFFFFF
    vals = np.arange(n, dtype=np.uint32)
    order = np.argsort(keys, kind="stable")
    esa.debug_radix_sort(keys, vals, 0, 32)
    assert np.array_equal(vals, order.astype(np.uint32))


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(CASES.keys()))
def test_golden_cases(esa, name):
    """Byte parity with the committed goldens (made by the reference binary) and with the C restatement."""
    g = golden()[name]
    tis, sigma = make_input(name)
    k = g["prefixlength"]
    got, stats = esa.build(tis, sigma, k, tables=ALL)
    want = ou.oracle_build(tis, sigma, k)
    assert_tables_equal(got, want, name)
    for t in TABLES:
        assert sha(got[t]) == g["sha"][t], f"{name}: {t} hash differs from the reference golden"
    for sname in SCALARS:
        assert got[sname] == g[sname]
    if "tables" in g:
        for t in TABLES:
            assert got[t].tolist() == g["tables"][t], f"{name}: {t}"
    esa.check()


@pytest.mark.parametrize("name", list(CASES.keys()))
def test_skip_table_golden_cases(esa, name):
    """-skp (src/kurtz/mkskip.c:53-91): byte parity with the reference's .skp and with the C restatement, on both
    lcp paths (the exact large values come from lcpfull on the direct path and from the plcp array otherwise)."""
    g = golden()[name]
    tis, sigma = make_input(name)
    got, _ = esa.build(tis, sigma, g["prefixlength"], tables=("suf", "lcp", "skp"))
    want = ou.oracle_build(tis, sigma, g["prefixlength"])
    skp = ou.oracle_skiptable(want["lcp"], want["llv"])
    assert np.array_equal(got["skp"], skp), first_diff(got["skp"], skp)
    assert sha(got["skp"]) == g["sha"]["skp"]
    assert np.array_equal(got["suf"], want["suf"]) and np.array_equal(got["lcp"], want["lcp"])


def test_skip_table_alone_forces_lcp(esa):
    tis = synth.repeat_rich_dna(300000, seed=5, copies=10, divergence=0.01, tandem_frac=0.05, tandem_len=3000)
    want = ou.oracle_build(tis, 4, 6)
    skp = ou.oracle_skiptable(want["lcp"], want["llv"])
    got, _ = esa.build(tis, 4, 6, tables=("skp",))
    assert np.array_equal(got["skp"], skp) and np.array_equal(got["lcp"], want["lcp"]) and got["suf"] is None


def test_kat_survey(esa):
    tis = np.array([0, 1, 2, 3, 254, 0, 1, 2, 255, 0, 1, 2, 3, 255, 254, 254, 0, 1], np.uint8)
    got, _ = esa.build(tis, 4, 1, tables=ALL)
    assert got["suf"].tolist() == [0, 9, 5, 16, 1, 10, 6, 17, 2, 11, 7, 3, 12, 4, 8, 13, 14, 15, 18]
    assert got["lcp"].tolist() == [0, 4, 3, 2, 0, 3, 2, 1, 0, 2, 1, 0, 1, 0, 0, 0, 0, 0, 0]
    assert got["bwt"].tolist() == [253, 255, 254, 254, 0, 0, 0, 0, 1, 1, 1, 2, 2, 3, 2, 3, 255, 254, 1]
    assert got["bck"].tolist() == [0, 4, 4, 8, 8, 11, 11, 13]
    assert got["sti1"].tolist() == [0, 0, 0, 0, 0, 2, 2, 2, 0, 1, 1, 1, 1, 0, 0, 0, 3, 3, 0]
    assert got["longest"] == 0 and got["maxbranchdepth"] == 4 and got["largelcpvalues"] == 0


@pytest.mark.parametrize("seed", range(12))
def test_random_small_inputs_against_oracle(esa, seed):
    """Ragged sizes around tile borders, every alphabet, dense specials."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([0, 1, 2, 3, 15, 16, 17, 255, 256, 4095, 4096, 4097, 8191, 8192, 8200, 20000, 65537]))
    sigma = int(rng.choice([2, 3, 4, 5, 20, 31, 200]))
    if n == 0:
        tis = np.zeros(0, np.uint8)
    else:
        tis = rng.integers(0, sigma, size=n, dtype=np.uint8)
        sp = rng.random(n) < rng.choice([0.0, 0.01, 0.3])
        tis[sp] = rng.choice([254, 255], size=int(sp.sum()))
        if rng.random() < 0.5 and n > 40:  # long repeats
            tis[n // 2:n // 2 + n // 3] = tis[:n // 3]
    lib = ou.oracle_lib()
    kmax = lib.oracle_maximalprefixlength(sigma, n)
    k = int(rng.integers(1, max(1, kmax) + 1))
    got, _ = esa.build(tis, sigma, k, tables=ALL)
    want = ou.oracle_build(tis, sigma, k)
    assert_tables_equal(got, want, f"seed {seed} n={n} sigma={sigma} k={k}")
    esa.check()


def test_no_prefixlength_global_sort(esa):
    """-pl absent: one global sort, no bck/sti1 (bese.c:1048-1075)."""
    tis = synth.dna_with_specials(30000, seed=77, nfrac=0.01, mean_run=4, records=9)
    got, _ = esa.build(tis, 4, 0, tables=("suf", "lcp", "bwt"))
    want = ou.oracle_build(tis, 4, 0, bck=False, sti1=False)
    assert_tables_equal(got, want, "k=0")
    assert got["bck"] is None and got["sti1"] is None
    with pytest.raises(EsaError) as ei:
        esa.build(tis, 4, 0, tables=("suf", "bck"))
    assert "requires option -pl" in str(ei.value)


def test_prefixlength_too_large_is_the_reference_error(esa):
    tis = synth.random_dna(1000, seed=3)
    with pytest.raises(EsaError) as ei:
        esa.build(tis, 4, 9, tables=("suf",))
    assert "prefix length 9 is too large, maximal prefix length for this input size and alphabet size is 3" in str(
        ei.value)


def test_demand_subsets(esa):
    tis = synth.dna_with_specials(40000, seed=8, nfrac=0.005, mean_run=6, records=5)
    want = ou.oracle_build(tis, 4, 6)
    got, _ = esa.build(tis, 4, 6, tables=("suf",))
    assert np.array_equal(got["suf"], want["suf"]) and got["lcp"] is None and got["bwt"] is None
    got, _ = esa.build(tis, 4, 6, tables=("bwt",))
    assert np.array_equal(got["bwt"], want["bwt"]) and got["suf"] is None
    got, _ = esa.build(tis, 4, 6, tables=("sti1",))  # forces lcp + bck like mkvtree.c:376
    assert np.array_equal(got["sti1"], want["sti1"]) and np.array_equal(got["bck"], want["bck"])
    assert np.array_equal(got["lcp"], want["lcp"])
    got, _ = esa.build(tis, 4, 6, tables=("suf", "sti"))
    inv = np.zeros(len(tis) + 1, np.uint64)
    inv[want["suf"]] = np.arange(len(tis) + 1, dtype=np.uint64)
    assert np.array_equal(got["sti"], inv)


def test_staged_api_and_repeatability(esa):
    tis = synth.random_protein(50000, seed=21, records=30, xfrac=0.001)
    want = ou.oracle_build(tis, 20, 3)
    esa.upload(tis, 20)
    for _ in range(3):  # same resident text, arena reused
        st = esa.construct(3, tables=ALL)
        assert st["launches"] > 0 and st["kernel"]["radix_pass"]["launches"] > 0
        esa.check()
    got = esa.download()
    assert_tables_equal(got, want, "staged")


@pytest.mark.skipif(not ou.have_ref(), reason="oracle/_ref/mkvtree.x did not travel")
@pytest.mark.parametrize("kind", ["dna", "protein"])
def test_against_live_reference_binary(esa, kind, tmp_path):
    """The unmodified reference mkvtree, run on this box, on a FASTA rendering of a seeded input."""
    if kind == "dna":
        tis, sigma, flag = synth.dna_with_specials(150000, seed=31, nfrac=0.003, mean_run=12, records=25), 4, "-dna"
    else:
        tis, sigma, flag = synth.random_protein(120000, seed=32, records=80, xfrac=0.001), 20, "-protein"
    path = str(tmp_path / "in.fna")
    with open(path, "wb") as f:
        f.write(synth.to_fasta(tis, kind))
    rc, out, err, files = ou.run_mkvtree(ou.REF_MKVTREE, path, [flag, "-pl", "-tis", "-suf", "-lcp", "-bwt", "-bck",
                                                                "-sti1"])
    assert rc == 0, err
    ref = ou.tables_from_files(files)
    assert np.array_equal(ref["tis"], tis)
    k = ref["prj"]["prefixlength"]
    got, _ = esa.build(tis, sigma, k, tables=ALL)
    want = {t: ref[t] for t in TABLES}
    want.update({"longest": ref["prj"]["longest"], "maxbranchdepth": ref["prj"]["maxbranchdepth"],
                 "largelcpvalues": ref["prj"]["largelcpvalues"]})
    assert_tables_equal(got, want, f"reference {kind}")


# ---------------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs): too big for the CPU oracle, so size-independent invariants
def _properties(esa, tis, sigma, k, tables=ALL, check=True):
    got, st = esa.build(tis, sigma, k, tables=tables)
    if check:
        esa.check()  # permutation, order, lcp (exact values, llv), bwt, bck, sti1 on the device
    N = len(tis) + 1
    suf = got["suf"]
    assert len(suf) == N
    assert int(suf.sum()) == N * (N - 1) // 2          # checksum of a permutation
    assert suf[N - 1] == N - 1                          # the sentinel is the largest suffix
    assert suf[got["longest"]] == 0
    nspecial = int((tis >= 254).sum())
    tail = suf[N - 1 - nspecial:]
    assert np.all(np.diff(tail.astype(np.int64)) > 0)   # special-initial suffixes in position order
    if got["bck"] is not None:
        bck = got["bck"].reshape(-1, 2)
        assert bck[0, 0] == 0 and np.all(bck[:, 0] <= bck[:, 1]) and np.all(bck[1:, 0] >= bck[:-1, 1])
    if got["lcp"] is not None:
        llv = got["llv"].reshape(-1, 2)
        big = np.flatnonzero(got["lcp"] == 255)
        assert np.array_equal(llv[:, 0], big.astype(np.uint64)) and np.all(llv[:, 1] >= 255)
        assert got["maxbranchdepth"] == (int(llv[:, 1].max()) if len(llv) else int(got["lcp"].max()))
    return got, st


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json configurations at FULL size against the reference: tests/golden/bench_golden.json holds the sha256 of
# every table file the UNMODIFIED reference mkvtree wrote for the same seeded input (tests/golden/make_bench_golden.py,
# run where /root/reference is present).  The file is committed; a missing file or entry FAILS the test.
BENCH_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "bench_golden.json")


def bench_golden(name):
    assert os.path.exists(BENCH_GOLDEN), "tests/golden/bench_golden.json is missing (make_bench_golden.py makes it)"
    with open(BENCH_GOLDEN) as f:
        g = json.load(f)
    assert name in g, f"bench_golden.json has no entry {name}"
    return g[name]


def assert_matches_reference_golden(got, g, what):
    for t in ("suf", "lcp", "llv", "bwt", "bck"):
        h = hashlib.sha256(np.ascontiguousarray(got[t]).tobytes()).hexdigest()
        assert h == g["sha256"][t], f"{what}: {t} differs from the reference's file ({h[:16]} vs {g['sha'][t]})"
    assert got["longest"] == g["prj"]["longest"], what
    assert got["maxbranchdepth"] == g["prj"]["maxbranchdepth"], what
    assert got["largelcpvalues"] == g["prj"]["largelcpvalues"], what


FULL = ("suf", "lcp", "bwt", "bck")


def full_size_case(esa, name, check=True):
    g = bench_golden(name)
    tis = getattr(synth, g["generator"])(g["n"], **g["kwargs"])
    assert hashlib.sha256(tis.tobytes()).hexdigest() == g["tis_sha256"], f"{name}: the generator no longer makes the golden's input"
    got, st = _properties(esa, tis, g["sigma"], g["prefixlength"], tables=FULL, check=check)
    assert_matches_reference_golden(got, g, name)
    return got, st


def test_full_size_c2_100mbp_random_dna(esa):
    """BASELINE.json configs[1]: 100 Mbp random DNA, byte-identical .suf/.lcp/.llv/.bwt/.bck."""
    full_size_case(esa, "c2_100mbp")


def test_full_size_c2_twin_with_specials(esa):
    """the C2 twin with 0.1 % N in runs and 1000 records (specials at scale)."""
    full_size_case(esa, "c2n_100mbp")


def test_full_size_c3_250mbp_repeat_rich(esa):
    """BASELINE.json configs[2]: 250 Mbp repeat-rich DNA, 67 M lcp values >= 255 (exact values in .llv)."""
    got, st = full_size_case(esa, "c3_250mbp")
    assert st["rounds"] >= 5 and got["largelcpvalues"] > 60_000_000


def test_full_size_c4_500m_protein(esa):
    """BASELINE.json configs[3]: 500 M residues, 100 000 records."""
    full_size_case(esa, "c4_500mres")


def test_full_size_c5_sample_30mbp_human_like(esa):
    """configs[4]'s generator at 30 Mbp (satellite arrays: lcp up to 233 222, the plcp path) against the reference."""
    got, st = full_size_case(esa, "c5m_30mbp")
    assert got["maxbranchdepth"] > 100_000


def test_full_size_c5_sample_300mbp_human_like(esa):
    """configs[4]'s generator at 300 Mbp against the reference (about an hour of reference time, made once).  The device
    check is skipped here: its brute-force lcp recomputation needs minutes on the satellite arrays."""
    full_size_case(esa, "c5s_300mbp", check=False)


@pytest.mark.skipif(not ou.have_ref(), reason="oracle/_ref/mkvtree.x did not travel")
def test_c2_against_the_live_reference_binary(esa, tmp_path_factory):
    """The unmodified reference, run on THIS box on the 100 Mbp C2 input (about 10-20 s), table by table."""
    tis = synth.random_dna(100_000_000, seed=42)
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    import tempfile
    with tempfile.TemporaryDirectory(dir=shm) as d:
        path = os.path.join(d, "in.fna")
        with open(path, "wb") as f:
            f.write(synth.to_fasta(tis, "dna", width=1 << 20))
        rc, out, err, files = ou.run_mkvtree(ou.REF_MKVTREE, path, ["-dna", "-pl", "-suf", "-lcp", "-bwt", "-bck"],
                                             keepdir=d)
        assert rc == 0, err
        ref = ou.tables_from_files(files)
        del files
    k = ref["prj"]["prefixlength"]
    assert k == 11
    got, _ = esa.build(tis, 4, k, tables=FULL)
    for t in ("suf", "lcp", "llv", "bwt", "bck"):
        assert len(got[t]) == len(ref[t]) and np.array_equal(got[t], ref[t]), f"{t}: {first_diff(got[t], ref[t])}"
    assert got["longest"] == ref["prj"]["longest"] and got["maxbranchdepth"] == ref["prj"]["maxbranchdepth"]


def test_full_size_repeat_rich_30mbp(esa):
    """Scaled-down C3 (deep LCPs, many doubling rounds), checked on the device."""
    tis = synth.repeat_rich_dna(30_000_000, seed=7, copies=50, divergence=0.005, tandem_frac=0.01)
    got, st = _properties(esa, tis, 4, 10)
    assert st["rounds"] >= 5 and got["largelcpvalues"] > 0


def test_full_size_protein_50m(esa):
    tis = synth.random_protein(50_000_000, seed=11, records=10000, xfrac=0.0005)
    _properties(esa, tis, 20, 5)


# ---------------------------------------------------------------------------------------------------------
# esa_check must FAIL on a damaged table (the reference's DEBUG checks abort the same way, bese.c:355-528)
@pytest.fixture(scope="module")
def resident(esa):
    """A resident build with deep lcp values (llv pairs), bck and sti1."""
    tis = synth.repeat_rich_dna(400_000, seed=3, copies=8, divergence=0.002, tandem_frac=0.05, tandem_len=3000)
    return tis


def _construct(esa, tis, mode=None):
    esa.upload(tis, 4)
    esa.construct(7, tables=("suf", "lcp", "bwt", "bck", "sti1"))
    esa.check()


@pytest.mark.parametrize("table,index,xor,needle", [
    ("suf", 1234, 1 << 3, "permutation"), ("lcp", 4321, 1, "lcp"), ("bwt", 99, 2, "bwt"),
    ("bck", 2 * 100, 1, "bck"), ("bck", 2 * 100 + 1, 1, "bck"), ("sti1", 777, 1, "sti1"),
    ("llv", 1, 1, "llv"), ("llv", 2 * 5 + 1, 1 << 9, "llv"), ("llv", 2 * 7, 1, "llv")])
def test_check_fails_on_a_damaged_table(esa, resident, table, index, xor, needle):
    _construct(esa, resident)
    if table == "sti1":
        # damage an entry the check looks at: a suffix inside [left, mid) of some bucket
        tabs = esa.download()
        index = int(tabs["suf"][int(tabs["bck"][2 * 50])])
    esa.debug_corrupt(table, index, xor)
    with pytest.raises(EsaError) as ei:
        esa.check()
    assert ei.value.code == -8 and needle in str(ei.value), str(ei.value)
    esa.debug_corrupt(table, index, xor)  # undo: consistent again
    esa.check()


def test_check_compares_exact_large_values_on_both_lcp_paths(esa, resident):
    """largelcpvalues and maxbranchdepth are part of the check: a wrong exact value >= 255 is found even though its
    saturated .lcp byte is right."""
    _construct(esa, resident)
    tabs = esa.download()
    assert tabs["largelcpvalues"] > 100 and tabs["maxbranchdepth"] > 255
    j = int(np.argmax(tabs["llv"][1::2]))  # the pair that carries maxbranchdepth
    esa.debug_corrupt("llv", 2 * j + 1, 1 << 20)
    with pytest.raises(EsaError):
        esa.check()
