"""The child table (.cld / .cld1) on the GPU, through the C ABI (esa_childtab) and the mkcld tool,
against: the files the UNMODIFIED reference tools wrote (tests/golden/cld_golden.json: mkvtree.x + mkcld.x, small cases
and BASELINE-size ones), the oracle's restatement of the reference's stack walks (oracle_childtab, pinned on the CPU by
tests/test_childtab_model.py), and the live reference mkcld.x where oracle/_ref travelled to the box."""
import hashlib
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_util as ou
from golden_inputs import CASES, make_input
from test_childtab_model import GOLDEN, SHAPED, exact_lcp, sha, split_lcp
from vstree_b200 import Esa, EsaError, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS_MKVTREE = os.path.join(ROOT, "vstree_b200", "bin", "mkvtree")
OURS_MKCLD = os.path.join(ROOT, "vstree_b200", "bin", "mkcld")
REF_MKCLD = os.path.join(ROOT, "oracle", "_ref", "mkcld.x")
SEAM_MKCLD = os.path.join(ROOT, "oracle", "_ref", "gpu", "mkcld.x")  # reference main + integration/mkcldgpu.c

with open(os.path.join(os.path.dirname(__file__), "golden", "cld_golden.json")) as f:
    FULL = json.load(f).get("full", {})


@pytest.fixture(scope="module")
def esa():
    e = Esa(0)
    yield e
    e.close()


def first_diff(a, b, width=1):
    d = np.flatnonzero(a != b)
    i = int(d[0])
    return f"entry {i // width} (byte {i % width}): got {a[i]} want {b[i]}; {len(d)} bytes differ"


def assert_child_tables(got, want, what):
    (cld, cld1), (wcld, wcld1) = got, want
    assert len(cld) == len(wcld) and np.array_equal(cld, wcld), f"{what}: cld: {first_diff(cld, wcld, 3)}"
    assert len(cld1) == len(wcld1) and np.array_equal(cld1, wcld1), f"{what}: cld1: {first_diff(cld1, wcld1)}"


@pytest.mark.parametrize("name", list(CASES))
def test_resident_build_then_child_table_against_reference_files(esa, name):
    """the whole chain on the GPU: build the index, then the child table from the lcp table left on the device"""
    g = GOLDEN[name]
    tis, sigma = make_input(name)
    _, _, flag, pl = CASES[name]
    k = pl if pl else ou.oracle_lib().oracle_recommendedprefixlength(sigma, len(tis))
    tabs, _ = esa.build(tis, sigma, k, tables=("suf", "lcp"))
    assert sha(tabs["lcp"].tobytes()) == g["sha"]["lcp"] and sha(tabs["llv"].tobytes()) == g["sha"]["llv"]
    cld, cld1 = esa.childtab(n=len(tis))
    assert sha(cld.tobytes()) == g["sha"]["cld"], f"{name}: .cld differs from the reference's file"
    assert sha(cld1.tobytes()) == g["sha"]["cld1"], f"{name}: .cld1 differs from the reference's file"
    if "tables" in g:
        assert cld.tolist() == g["tables"]["cld"] and cld1.tolist() == g["tables"]["cld1"]
    # the same tables from host copies of .lcp / .llv (the mkcld seam), on the same context, and the build is still there
    assert_child_tables(esa.childtab(tabs["lcp"], tabs["llv"]), (cld, cld1), name + " (host tables)")
    assert_child_tables(esa.childtab(n=len(tis)), (cld, cld1), name + " (resident, second call)")
    esa.check()


@pytest.mark.parametrize("name", list(SHAPED))
def test_host_tables_against_the_stack_walk(esa, name):
    """lcp-shaped and adversarial arrays (ties, plateaus, nested intervals, distances >= 255, values >= 255, sizes around
    the powers of the pyramid's fan-out)"""
    lcp, llv = split_lcp(SHAPED[name])
    assert_child_tables(esa.childtab(lcp, llv), ou.oracle_childtab(lcp, llv), name)


@pytest.mark.parametrize("name", ["uniform4_4097", "uniform300_65537", "plateaus", "nested_deep", "long_intervals",
                                  "large_values", "walk"])
def test_pyramid_search_for_every_entry(esa, name, monkeypatch):
    """ESA_CHILDTAB_NO_NEAR=1: the bounded scan is skipped, every entry goes through the queue and child_far_kernel"""
    monkeypatch.setenv("ESA_CHILDTAB_NO_NEAR", "1")
    lcp, llv = split_lcp(SHAPED[name])
    assert_child_tables(esa.childtab(lcp, llv), ou.oracle_childtab(lcp, llv), name + " (pyramid only)")


def test_twenty_million_entries_against_the_stack_walk(esa):
    rng = np.random.default_rng(99)
    N = 20_000_001
    L = rng.geometric(0.08, N) - 1
    L[rng.integers(0, N, N // 50)] = rng.integers(255, 3_000_000, N // 50)   # scattered large values
    L[5_000_000:5_400_000] = 50 + rng.integers(0, 3, 400_000)                 # long intervals: every 300th entry closes
    L[5_000_000:5_400_000:300] = 45                                           # one, distances beyond 255
    L[12_000_000:12_000_600] = np.arange(600) + 300                           # a staircase of large values
    L[0] = 0
    lcp, llv = split_lcp(L)
    got = esa.childtab(lcp, llv)
    assert_child_tables(got, ou.oracle_childtab(lcp, llv), "20 M entries")
    assert (got[0] == 255).sum() > 1000 and len(llv) > 700_000


def test_without_cld1_and_argument_errors(esa):
    lcp, llv = split_lcp(SHAPED["large_values"])
    cld, none = esa.childtab(lcp, llv, cld1=False)
    assert none is None and np.array_equal(cld, ou.oracle_childtab(lcp, llv)[0])
    # pairs that do not belong to the table: one missing, one that points at a small value, one duplicated
    for bad in (llv[2:], np.concatenate([llv, np.array([3, 999], np.uint64)]), np.concatenate([llv[:2], llv])):
        with pytest.raises(EsaError) as e:
            esa.childtab(lcp, bad)
        assert e.value.code == -1 and "large lcp values do not belong" in e.value.message
    with Esa(0) as fresh:
        with pytest.raises(EsaError) as e:
            fresh.childtab(n=10)  # nothing built on this context
        assert e.value.code == -7
        tis = synth.random_dna(5000, seed=3)
        fresh.build(tis, 4, 3, tables=("suf",))  # a build without the lcp table
        with pytest.raises(EsaError) as e:
            fresh.childtab(n=len(tis))
        assert e.value.code == -7


@pytest.mark.parametrize("name", [n for n in ("c2_100mbp", "c5m_30mbp", "c3_250mbp")])
def test_full_size_against_reference_files(esa, name):
    """BASELINE-size indexes: .cld / .cld1 sha256 == the files mkcld.x wrote from the reference's own index"""
    assert name in FULL, f"tests/golden/cld_golden.json has no full-size entry {name} (make_cld_golden.py --full makes it)"
    g = FULL[name]
    with open(os.path.join(os.path.dirname(__file__), "golden", "bench_golden.json")) as f:
        b = json.load(f)[name]
    tis = getattr(synth, b["generator"])(b["n"], **b["kwargs"])
    assert hashlib.sha256(tis.tobytes()).hexdigest() == b["tis_sha256"]
    tabs, _ = esa.build(tis, b["sigma"], b["prefixlength"], tables=("suf", "lcp"), copy=False)
    full = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert full(tabs["lcp"]) == g["sha256"]["lcp"] and full(tabs["llv"]) == g["sha256"]["llv"]
    esa.free(tabs)
    cld, cld1 = esa.childtab(n=b["n"])
    assert full(cld) == g["sha256"]["cld"], f"{name}: .cld differs from the reference's file"
    assert full(cld1) == g["sha256"]["cld1"], f"{name}: .cld1 differs from the reference's file"
    assert int(np.count_nonzero(cld == 255)) == g["large_child_values"]


def run(cmd, cwd):
    pr = subprocess.run(cmd, cwd=cwd, capture_output=True)
    return pr.returncode, pr.stdout.decode(errors="replace"), pr.stderr.decode(errors="replace")


@pytest.mark.parametrize("name", ["kat_survey", "dna_specials_50k", "dna_repeat_200k", "protein_30k", "human_like_400k"])
def test_mkcld_tool(tmp_path, name):
    """`mkcld indexname` on an index written by this repository's mkvtree: the files of the reference's mkcld (golden
    and, when oracle/_ref travelled, the live binary on the same index)"""
    g = GOLDEN[name]
    tis, sigma = make_input(name)
    _, _, flag, pl = CASES[name]
    (tmp_path / "in.fna").write_bytes(synth.to_fasta(tis, "dna" if flag == "-dna" else "protein"))
    rc, out, err = run([OURS_MKVTREE, "-db", "in.fna", "-indexname", "idx", flag, "-pl"] + ([str(pl)] if pl else []) +
                       ["-tis", "-suf", "-lcp"], str(tmp_path))
    assert rc == 0, err
    live = tmp_path / "live"
    if os.access(REF_MKCLD, os.X_OK):
        live.mkdir()
        for fn in os.listdir(tmp_path):
            if fn.startswith("idx.") and os.path.isfile(tmp_path / fn):
                shutil.copy(tmp_path / fn, live / fn)
    rc, out, err = run([OURS_MKCLD, "idx"], str(tmp_path))
    assert rc == 0 and err == "", err
    assert out.startswith("# overall space peak: main=") and out.count("\n") == 1
    cld = np.fromfile(tmp_path / "idx.cld", np.uint8)
    cld1 = np.fromfile(tmp_path / "idx.cld1", np.uint8)
    assert len(cld) == 3 * (len(tis) + 1) and len(cld1) == len(tis) + 1
    assert sha(cld.tobytes()) == g["sha"]["cld"] and sha(cld1.tobytes()) == g["sha"]["cld1"]
    if live.exists():
        rc, rout, rerr = run([REF_MKCLD, "idx"], str(live))
        assert rc == 0, rerr
        assert (live / "idx.cld").read_bytes() == cld.tobytes()
        ref1 = np.fromfile(live / "idx.cld1", np.uint8)
        keep = np.ones(len(ref1), bool)
        keep[g["cld1_unwritten"]] = False  # bytes the reference never writes (make_cld_golden.py)
        assert g["cld1_unwritten_count"] <= 16 and np.array_equal(ref1[keep], cld1[keep])
        assert rout.startswith("# overall space peak: main=") and rout.count("\n") == out.count("\n")
    if os.access(SEAM_MKCLD, os.X_OK):
        # the reference's own mkcld with its three child-table functions replaced at their seam (integration/mkcldgpu.c)
        seam = tmp_path / "seam"
        seam.mkdir()
        for fn in os.listdir(tmp_path):
            if fn.startswith("idx.") and not fn.startswith("idx.cld") and os.path.isfile(tmp_path / fn):
                shutil.copy(tmp_path / fn, seam / fn)
        rc, sout, serr = run([SEAM_MKCLD, "idx"], str(seam))
        assert rc == 0, serr
        assert (seam / "idx.cld").read_bytes() == cld.tobytes() and (seam / "idx.cld1").read_bytes() == cld1.tobytes()


def test_mkcld_tool_errors(tmp_path):
    """usage text, missing and truncated files: messages and exit codes of the reference (mkcld.mn.c:50-66)"""
    rc, out, err = run([OURS_MKCLD], str(tmp_path))
    assert rc == 1 and err == (f"Usage: {OURS_MKCLD} indexname\nsee Vmatch manual at http://www.vmatch.de  for more "
                               "information\n")
    rc, out, err = run([OURS_MKCLD, "nosuch"], str(tmp_path))
    assert rc == 1 and err == f'{OURS_MKCLD}: cannot open file "nosuch.prj": No such file or directory\n'
    tis = synth.random_dna(3000, seed=8)
    (tmp_path / "in.fna").write_bytes(synth.to_fasta(tis, "dna"))
    rc, out, err = run([OURS_MKVTREE, "-db", "in.fna", "-indexname", "idx", "-dna", "-pl", "-suf", "-lcp"], str(tmp_path))
    assert rc == 0, err
    whole = (tmp_path / "idx.lcp").read_bytes()
    (tmp_path / "idx.lcp").write_bytes(whole[:100])
    rc, out, err = run([OURS_MKCLD, "idx"], str(tmp_path))
    assert rc == 1 and err == f"{OURS_MKCLD}: 100 bytes read from idx.lcp, but 3001 expected\n"
    (tmp_path / "idx.lcp").write_bytes(whole)
    os.remove(tmp_path / "idx.suf")
    rc, out, err = run([OURS_MKCLD, "idx"], str(tmp_path))
    assert rc == 1 and 'cannot read file "idx.suf": open failed: No such file or directory' in err
    assert not (tmp_path / "idx.cld").exists()
