"""Pins the CPU restatement (oracle/esa_oracle.c) against the unmodified reference binary (oracle/_ref/mkvtree.x)
and against the committed goldens.  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_util as ou
from vstree_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden.json")

# SURVEY.md 8a known-answer vector (produced by the reference; >s1 ACGTNACG >s2 ACGT >s3 NNAC, -dna -pl 1)
KAT_TIS = [0, 1, 2, 3, 254, 0, 1, 2, 255, 0, 1, 2, 3, 255, 254, 254, 0, 1]
KAT = {
    "suf": [0, 9, 5, 16, 1, 10, 6, 17, 2, 11, 7, 3, 12, 4, 8, 13, 14, 15, 18],
    "lcp": [0, 4, 3, 2, 0, 3, 2, 1, 0, 2, 1, 0, 1, 0, 0, 0, 0, 0, 0],
    "bwt": [253, 255, 254, 254, 0, 0, 0, 0, 1, 1, 1, 2, 2, 3, 2, 3, 255, 254, 1],
    "bck": [0, 4, 4, 8, 8, 11, 11, 13],
    "sti1": [0, 0, 0, 0, 0, 2, 2, 2, 0, 1, 1, 1, 1, 0, 0, 0, 3, 3, 0],
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_kat():
    t = ou.oracle_build(np.array(KAT_TIS, np.uint8), 4, 1)
    for k, v in KAT.items():
        assert t[k].tolist() == v, k
    assert t["longest"] == 0 and t["maxbranchdepth"] == 4 and t["largelcpvalues"] == 0
    assert ou.oracle_check(np.array(KAT_TIS, np.uint8), 4, 1, t) == 0


def test_prefixlength_math():
    lib = ou.oracle_lib()
    # facts printed by the reference's -v transcript for at1MB and BASELINE.md section 2
    assert lib.oracle_recommendedprefixlength(4, 772376) == 7
    assert lib.oracle_maximalprefixlength(4, 772376) == 8
    assert lib.oracle_recommendedprefixlength(4, 100000000) == 11
    assert lib.oracle_recommendedprefixlength(20, 20000000) == 4
    assert lib.oracle_recommendedprefixlength(4, 10) == 1
    assert lib.oracle_numofcodes(4, 13) == 67108864
    assert lib.oracle_numofcodes(20, 5) == 3200000


def test_check_tables_detects_corruption():
    tis = synth.dna_with_specials(3000, seed=5, nfrac=0.01, mean_run=3, records=4)
    t = ou.oracle_build(tis, 4, 3)
    assert ou.oracle_check(tis, 4, 3, t) == 0
    bad = dict(t)
    bad["suf"] = t["suf"].copy()
    bad["suf"][[10, 11]] = bad["suf"][[11, 10]]
    assert ou.oracle_check(tis, 4, 3, bad) != 0
    bad = dict(t)
    bad["lcp"] = t["lcp"].copy()
    bad["lcp"][7] ^= 1
    assert ou.oracle_check(tis, 4, 3, bad) == 3


def _compare_with_reference(fasta_path, alpha_flag, sigma, pl_args, tmp_path):
    args = [alpha_flag, "-pl"] + pl_args + ["-tis", "-suf", "-lcp", "-bwt", "-bck", "-sti1"]
    # -sti1 suppresses nothing here because -suf is given explicitly
    rc, out, err, files = ou.run_mkvtree(ou.REF_MKVTREE, fasta_path, args)
    assert rc == 0, err
    ref = ou.tables_from_files(files)
    k = ref["prj"]["prefixlength"]
    mine = ou.oracle_build(ref["tis"], sigma, k)
    for name in ("suf", "lcp", "llv", "bwt", "bck", "sti1"):
        assert np.array_equal(mine[name], ref[name]), f"{name} differs for {fasta_path}"
    assert mine["longest"] == ref["prj"]["longest"]
    assert mine["maxbranchdepth"] == ref["prj"]["maxbranchdepth"]
    assert mine["largelcpvalues"] == ref["prj"]["largelcpvalues"]
    return ref


REF_FIXTURES = [
    ("testdata/at100K1", "-dna", 4, []),
    ("testdata/at1MB", "-dna", 4, []),
    ("testdata/at1MB", "-dna", 4, ["3"]),
    ("testdata/Grumbach/Wildcards.fna", "-dna", 4, []),
    ("testdata/U89959.fna", "-dna", 4, ["1"]),
    ("testdata/swiss10K", "-protein", 20, []),
    ("testdata/sw100K1", "-protein", 20, ["2"]),
    ("testdata/swiss1MB", "-protein", 20, []),
    ("Mkvtree/DATA/DBdna1", "-dna", 4, ["2"]),
    ("Mkvtree/DATA/DBprot1", "-protein", 20, ["1"]),
]


@pytest.mark.parametrize("rel,flag,sigma,pl", REF_FIXTURES)
def test_against_reference_fixtures(refdata, rel, flag, sigma, pl, tmp_path):
    if not ou.have_ref():
        pytest.skip("oracle/_ref/mkvtree.x not built")
    path = os.path.join(refdata, rel)
    if not os.path.exists(path):
        pytest.skip(f"{rel} not in the reference tree")
    _compare_with_reference(path, flag, sigma, pl, tmp_path)


def test_at1mb_reference_hashes(refdata):
    """The reference build itself still reproduces the hashes recorded in SURVEY.md 8c."""
    if not ou.have_ref():
        pytest.skip("oracle/_ref/mkvtree.x not built")
    rc, out, err, files = ou.run_mkvtree(ou.REF_MKVTREE, os.path.join(refdata, "testdata/at1MB"),
                                         ["-dna", "-pl", "-suf", "-lcp", "-bwt", "-tis", "-ois", "-bck", "-sti1"])
    assert rc == 0, err
    want = {"suf": "d35617c3307679b0", "lcp": "010a6436643f0908", "llv": "5bea62168cc44547",
            "bwt": "8ce08b07227a8a93", "bck": "3d3d5e340324cebc", "sti1": "1b8f0b5f567ebebc",
            "tis": "0d91a7b28cfa0b7a", "ois": "02747886eee026fd"}
    for k, v in want.items():
        assert hashlib.sha256(files[k]).hexdigest()[:16] == v, k


def golden_cases():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", golden_cases() if os.path.exists(GOLDEN) else [], ids=lambda c: c["name"])
def test_against_committed_goldens(case):
    """Goldens were produced by the reference binary (tests/golden/make_golden.py); inputs are regenerated by seed."""
    from golden_inputs import make_input
    tis, sigma = make_input(case["name"])
    assert sha(tis) == case["sha"]["tis"], "synthetic input generator drifted"
    t = ou.oracle_build(tis, sigma, case["prefixlength"])
    for name in ("suf", "lcp", "llv", "bwt", "bck", "sti1"):
        assert sha(t[name]) == case["sha"][name], name
    assert t["longest"] == case["longest"]
    assert t["maxbranchdepth"] == case["maxbranchdepth"]
    assert t["largelcpvalues"] == case["largelcpvalues"]
    # skip table (src/kurtz/mkskip.c), pinned by the reference's own .skp
    skp = ou.oracle_skiptable(t["lcp"], t["llv"])
    assert sha(skp) == case["sha"]["skp"], "skp"
    if "tables" in case:
        assert skp.tolist() == case["tables"]["skp"]


@pytest.mark.parametrize("seed", range(24))
def test_against_reference_on_random_inputs(seed, tmp_path):
    """Seeded random inputs through the UNMODIFIED reference binary: every alphabet flag, dense and sparse specials,
    repeats, identical records (ties that only the position of a separator breaks), every prefix length the input
    admits.  Pins the C restatement where the fixtures and goldens are silent."""
    if not ou.have_ref():
        pytest.skip("oracle/_ref/mkvtree.x not built")
    rng = np.random.default_rng(5000 + seed)
    kind = ["dna", "protein"][seed % 2]
    sigma, flag = (4, "-dna") if kind == "dna" else (20, "-protein")
    n = int(rng.choice([40, 300, 2000, 9000, 40000]))
    shape = seed % 6
    if shape == 0:      # identical records
        rec = rng.integers(0, sigma, size=int(rng.integers(5, 60)), dtype=np.uint8)
        tis = np.concatenate([np.append(rec, np.uint8(255)) for _ in range(max(2, n // (len(rec) + 1)))])[:-1]
    elif shape == 1:    # dense specials
        tis = rng.integers(0, sigma, size=n, dtype=np.uint8)
        sp = rng.random(n) < 0.25
        tis[sp] = rng.choice([254, 255], size=int(sp.sum()), p=[0.8, 0.2])
    elif shape == 2:    # long repeats with a few substitutions
        base = rng.integers(0, sigma, size=max(8, n // 8), dtype=np.uint8)
        tis = np.tile(base, 8)[:n].copy()
        hit = rng.random(len(tis)) < 0.01
        tis[hit] = rng.integers(0, sigma, size=int(hit.sum()), dtype=np.uint8)
    elif shape == 3:    # low complexity
        tis = rng.integers(0, 2, size=n, dtype=np.uint8)
        tis[rng.random(n) < 0.002] = 255
    else:               # plain random with sparse wildcards and separators
        tis = rng.integers(0, sigma, size=n, dtype=np.uint8)
        sp = rng.random(n) < 0.004
        tis[sp] = rng.choice([254, 255], size=int(sp.sum()))
    # no empty records: the reader rejects them
    tis = tis[np.concatenate([[True], ~((tis[1:] == 255) & (tis[:-1] == 255))])]
    while len(tis) and tis[0] == 255:
        tis = tis[1:]
    while len(tis) and tis[-1] == 255:
        tis = tis[:-1]
    if len(tis) < 8:
        pytest.skip("degenerate input")
    path = str(tmp_path / "in.fa")
    with open(path, "wb") as f:
        f.write(synth.to_fasta(tis, kind))
    kmax = int(ou.oracle_lib().oracle_maximalprefixlength(sigma, len(tis)))
    pl = [] if seed % 3 == 0 else [str(int(rng.integers(1, max(1, kmax) + 1)))]
    ref = _compare_with_reference(path, flag, sigma, pl, tmp_path)
    assert np.array_equal(ref["tis"], tis)
