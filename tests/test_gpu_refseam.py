"""The reference's own programs with the suffix sorter replaced AT THE REFERENCE'S SEAM.

oracle/build_ref.sh builds two families from the reference sources: the unmodified tools (oracle/_ref/*.x) and, in
oracle/_ref/gpu/, the same reference objects linked with integration/besegpu.c instead of bese.o / besespecial.o /
ppsort.o -- besesuffixsort[special] (src/Mkvtree/bese.c:855) on top of esa_build.  Every caller of the sorter then runs on
the GPU without a changed line: callmkvtree (mkvtree, mkvram: src/Mkvtree/mkvtree.c:689-707), callmkvtreegeneric
(mkrcidx, mkdna6idx: mkrcidx.c:298-325, mkdna6idx.c:75-127) and completevirtualtree (vmatch -online -q builds one index
per query sequence in memory: mkvtree.c:709-738, src/Vmatch/procmatch.c:68-79).  Both families must write the same bytes
and print the same lines."""
import os
import subprocess

import numpy as np
import pytest

import oracle_util as ou
from vstree_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
GPU = os.path.join(REF, "gpu")
TOOLS = ("mkvtree.x", "mkvram.x", "mkrcidx.x", "mkdna6idx.x", "vmatch.x")

needs_families = pytest.mark.skipif(
    not all(os.access(os.path.join(d, t), os.X_OK) for d in (REF, GPU) for t in TOOLS),
    reason="oracle/_ref and oracle/_ref/gpu did not travel (built where /root/reference is present)")


def run(binary, args, cwd):
    pr = subprocess.run([binary] + args, cwd=cwd, capture_output=True)
    files = {}
    for fn in sorted(os.listdir(cwd)):
        if not fn.endswith((".fna", ".fa")):
            with open(os.path.join(cwd, fn), "rb") as f:
                files[fn] = f.read()
    return pr.returncode, pr.stdout.decode(errors="replace"), pr.stderr.decode(errors="replace"), files


def lines(out):
    """stdout without the space-peak line (allocation totals differ: the GPU family never allocates the sorter's
    scratch) and with the program path neutralised."""
    keep = [l for l in out.splitlines() if "overall space peak" not in l and not l.startswith("# args=")]
    return [l.replace(GPU, "DIR").replace(REF, "DIR") for l in keep]


def both(tmp_path, tool, args, inputs, expect_rc=0):
    res = []
    for fam, d in ((REF, tmp_path / "ref"), (GPU, tmp_path / "gpu")):
        d.mkdir(exist_ok=True)
        for name, content in inputs.items():
            (d / name).write_bytes(content)
        res.append(run(os.path.join(fam, tool), args, str(d)))
    (ra, oa, ea, fa), (rb, ob, eb, fb) = res
    assert ra == expect_rc, ea
    assert rb == ra, f"{tool}: exit code {rb} vs reference {ra}: {eb}"
    assert lines(ob) == lines(oa)
    assert sorted(fa) == sorted(fb), f"files differ: {sorted(fb)} vs reference {sorted(fa)}"
    for k in fa:
        assert fa[k] == fb[k], f"{tool}: {k} differs from the reference family ({len(fb[k])} vs {len(fa[k])} bytes)"
    return res


def dna(n, seed, **kw):
    return synth.to_fasta(synth.dna_with_specials(n, seed=seed, **kw), "dna")


@needs_families
@pytest.mark.parametrize("args", [
    ["-db", "g.fna", "-dna", "-pl", "-suf", "-lcp", "-bwt", "-bck", "-tis", "-ois", "-v"],
    ["-db", "g.fna", "-dna", "-pl", "-sti1", "-v"],
    ["-db", "g.fna", "-dna", "-pl", "-allout"],
    ["-db", "g.fna", "-dna", "-suf", "-lcp", "-skp"],
    ["-db", "g.fna", "-dna", "-pl", "3", "-suf", "-bck", "-v"]])
def test_mkvtree_with_the_sorter_replaced(tmp_path, args):
    both(tmp_path, "mkvtree.x", args, {"g.fna": dna(60000, 5, nfrac=0.004, mean_run=7, records=15)})


@needs_families
def test_mkvtree_protein_and_repeats(tmp_path):
    prot = synth.to_fasta(synth.random_protein(40000, seed=3, records=50, xfrac=0.002), "protein")
    both(tmp_path, "mkvtree.x", ["-db", "p.fna", "-protein", "-pl", "-allout", "-v"], {"p.fna": prot})
    rep = synth.to_fasta(synth.repeat_rich_dna(150000, seed=9, copies=12, divergence=0.004, tandem_frac=0.03,
                                               tandem_len=2500), "dna")
    (tmp_path / "ref").rename(tmp_path / "ref1")
    (tmp_path / "gpu").rename(tmp_path / "gpu1")
    both(tmp_path, "mkvtree.x", ["-db", "r.fna", "-dna", "-pl", "-suf", "-lcp", "-bwt", "-bck", "-sti1"], {"r.fna": rep})


@needs_families
def test_mkvtree_maxdepth_is_refused_not_faked(tmp_path):
    (tmp_path / "g.fna").write_bytes(dna(5000, 1, nfrac=0.0, mean_run=1, records=1))
    rc, out, err, files = run(os.path.join(GPU, "mkvtree.x"), ["-db", "g.fna", "-dna", "-pl", "-suf", "-maxdepth", "10"],
                              str(tmp_path))
    assert rc != 0 and "not available with the GPU suffix sorter" in err


@needs_families
def test_mkvram_callmkvtree_in_memory(tmp_path):
    """callmkvtree(storeonfile = False): every table stays in memory (mkvram.c:23)."""
    both(tmp_path, "mkvram.x", ["-db", "g.fna", "-dna", "-pl", "-suf", "-lcp", "-bwt", "-bck", "-sti1", "-tis", "-v"],
         {"g.fna": dna(30000, 8, nfrac=0.003, mean_run=4, records=9)})


@needs_families
def test_mkrcidx_reverse_complement_index(tmp_path):
    """callmkvtreegeneric(useformkrcidx) + mkvtreeprocess on the sequence with its reverse complement appended."""
    both(tmp_path, "mkrcidx.x", ["-db", "g.fna", "-indexname", "rc", "-v"],
         {"g.fna": dna(25000, 11, nfrac=0.002, mean_run=6, records=7)})


@needs_families
@pytest.mark.parametrize("transnum", ["1", "11"])
def test_mkdna6idx_six_frame_index(tmp_path, transnum):
    """callmkvtreegeneric(useformkdna6idx): the six-frame translation of a DNA file, indexed as protein."""
    both(tmp_path, "mkdna6idx.x", ["-db", "g.fna", "-transnum", transnum, "-indexname", "six", "-tis", "-ois", "-v"],
         {"g.fna": dna(18000, 13, nfrac=0.001, mean_run=3, records=4)})


@needs_families
@pytest.mark.parametrize("vmargs", [["-l", "30"], ["-l", "25", "-h", "1"], ["-l", "20", "-p"], ["-l", "40", "-v"]])
def test_vmatch_online_query_indexes_through_completevirtualtree(tmp_path, vmargs):
    """vmatch -online -q: one in-memory index per query sequence (completevirtualtree), matched against a database index
    on file; the database index itself is written by the family's own mkvtree."""
    db = synth.dna_with_specials(40000, seed=21, nfrac=0.002, mean_run=5, records=6)
    rng = np.random.default_rng(4)
    q = np.concatenate([db[1000:1400], synth.random_dna(300, seed=5), np.array([255], np.uint8), db[22000:22600],
                        synth.random_dna(100, seed=6), np.array([255], np.uint8), db[30000:30500]])
    q[rng.integers(0, len(q), 6)] = 0
    q[(q == 255).nonzero()[0]] = 255
    inputs = {"db.fna": synth.to_fasta(db, "dna"), "q.fna": synth.to_fasta(q, "dna", name="query")}
    outs = []
    for fam, d in ((REF, tmp_path / "ref"), (GPU, tmp_path / "gpu")):
        d.mkdir()
        for name, content in inputs.items():
            (d / name).write_bytes(content)
        pr = subprocess.run([os.path.join(fam, "mkvtree.x"), "-db", "db.fna", "-dna", "-pl", "-allout"], cwd=str(d),
                            capture_output=True)
        assert pr.returncode == 0, pr.stderr
        pr = subprocess.run([os.path.join(fam, "vmatch.x"), "-online", "-q", "q.fna"] + vmargs + ["db.fna"], cwd=str(d),
                            capture_output=True)
        assert pr.returncode == 0, pr.stderr.decode(errors="replace")
        outs.append(lines(pr.stdout.decode(errors="replace").replace(str(d), "CWD")))
    assert outs[0] == outs[1]
    if "-p" not in vmargs:  # (the planted repeats are direct ones)
        assert any(l and not l.startswith("#") for l in outs[0]), "the comparison is vacuous: vmatch reported no match"


@needs_families
def test_the_gpu_family_links_the_product_library_and_no_reference_sorter():
    """The replaced seam is real: the binaries need libvstree_b200.so and contain no multikey quicksort."""
    for t in TOOLS:
        deps = subprocess.run(["readelf", "-d", os.path.join(GPU, t)], capture_output=True, text=True).stdout
        assert "libvstree_b200.so" in deps, t
        syms = subprocess.run(["nm", os.path.join(GPU, t)], capture_output=True, text=True).stdout
        assert "bucketsortsuffixes" not in syms and "bentleysedgewick" not in syms, t
        assert " T besesuffixsortspecial" in syms and " U esa_build" in syms, t
