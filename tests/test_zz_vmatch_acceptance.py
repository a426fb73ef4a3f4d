"""Downstream acceptance (north star: "an unmodified vmatch consumes the output"): the UNMODIFIED reference vmatch
(oracle/_ref/vmatch.x, built by oracle/build_ref.sh) is run on an index written by this repository's mkvtree and on the
index the reference mkvtree writes for the same input; its reports have to be the same.  Covers repeats, palindromic
and query matches, maximal unique and supermaximal matches, DNA and protein."""
import os
import subprocess

import pytest

import oracle_util as ou
from vstree_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "vstree_b200", "bin", "mkvtree")
VMATCH = os.path.join(ROOT, "oracle", "_ref", "vmatch.x")

needs_vmatch = pytest.mark.skipif(not (ou.have_ref() and os.path.exists(VMATCH)),
                                  reason="oracle/_ref/vmatch.x not built")

QUERY_DNA = b">q1\nACGTTGCATTGACCATGGTTTACAACGGATCGATCGGATCAGCTAGCTAGCTTTTGA\n>q2\nTTTTTTTTTTTTTTTTTTTTACGT\n"


def build_index(binary, workdir, fasta_name, flag, extra=()):
    pr = subprocess.run([binary, "-db", fasta_name, flag, "-pl", "-allout", "-indexname", "idx"] + list(extra),
                        cwd=workdir, capture_output=True, text=True)
    assert pr.returncode == 0, pr.stderr[-2000:]


def vmatch(workdir, args):
    pr = subprocess.run([VMATCH] + args + ["idx"], cwd=workdir, capture_output=True, text=True)
    lines = [l for l in pr.stdout.splitlines() if not l.startswith("# args=")]  # (names the index by its path)
    return pr.returncode, lines, pr.stderr.replace(workdir, "DIR")


def acceptance(tmp_path, fasta, flag, runs, builder=None, files=None):
    """Index by the reference in one directory, by the product in the other; every vmatch run on both."""
    dirs = {}
    for which, binary in (("ref", ou.REF_MKVTREE), ("ours", builder or OURS)):
        d = tmp_path / which
        d.mkdir()
        (d / "in.fa").write_bytes(fasta)
        for name, content in (files or {}).items():
            (d / name).write_bytes(content)
        build_index(binary, str(d), "in.fa", flag)
        dirs[which] = str(d)
    nonempty = 0
    for args in runs:
        ra = vmatch(dirs["ref"], args)
        rb = vmatch(dirs["ours"], args)
        assert ra[0] == 0, ra[2]
        assert rb == ra, f"vmatch {' '.join(args)}: report on the product's index differs from the reference's"
        nonempty += len(ra[1]) > 1
    assert nonempty > 0, "no vmatch run reported a match: the test inputs are too weak"


DNA_RUNS = [["-l", "30"], ["-l", "20", "-p"], ["-l", "25", "-d", "-p", "-s"], ["-supermax", "-l", "24"],
            ["-l", "40", "-h", "2"], ["-l", "35", "-e", "1"], ["-q", "q.fa", "-l", "12"],
            ["-q", "q.fa", "-mum", "-l", "14"], ["-l", "30", "-best", "5"], ["-tandem", "-l", "12"]]
PROT_RUNS = [["-l", "8"], ["-l", "12", "-h", "1"], ["-supermax", "-l", "7"], ["-q", "q.fa", "-l", "6"]]


def test_harness_on_two_reference_indexes(tmp_path):
    """CPU only: the harness itself, both indexes written by the reference (vmatch.x runs, the inputs yield matches)."""
    if not (ou.have_ref() and os.path.exists(VMATCH)):
        pytest.skip("oracle/_ref/vmatch.x not built")
    tis = synth.repeat_rich_dna(60000, seed=3, copies=12, divergence=0.01, tandem_frac=0.02, tandem_len=400)
    acceptance(tmp_path, synth.to_fasta(tis, "dna"), "-dna", DNA_RUNS, builder=ou.REF_MKVTREE,
               files={"q.fa": QUERY_DNA})


@needs_vmatch
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["repeats", "specials", "human_like"])
def test_vmatch_on_the_products_dna_index(tmp_path, case):
    if case == "repeats":
        tis = synth.repeat_rich_dna(200000, seed=7, copies=20, divergence=0.005, tandem_frac=0.02, tandem_len=1500)
    elif case == "specials":
        tis = synth.dna_with_specials(150000, seed=2, nfrac=0.01, mean_run=5, records=120)
        tis[40000:70000] = tis[5000:35000]       # something to find
    else:
        tis = synth.human_like_dna(400000, seed=2026, records=5)
    acceptance(tmp_path, synth.to_fasta(tis, "dna"), "-dna", DNA_RUNS, files={"q.fa": QUERY_DNA})


@needs_vmatch
@pytest.mark.gpu
def test_vmatch_on_the_products_protein_index(tmp_path):
    tis = synth.random_protein(120000, seed=13, records=300, xfrac=0.0005)
    tis[60000:64000] = tis[1000:5000]
    query = b">pq\n" + bytes(b"LVIFKREDAGSTNQYWPHMC"[c] for c in tis[1200:1260] if c < 20) + b"\n"
    acceptance(tmp_path, synth.to_fasta(tis, "protein"), "-protein", PROT_RUNS, files={"q.fa": query})
