"""The mkvtree host driver (vstree_b200/bin/mkvtree) against the UNMODIFIED reference binary (oracle/_ref/mkvtree.x):
same command line, same files byte for byte, same -v transcript, same error messages and exit codes.
Cases that only write host-side tables (-tis/-ois) run on the CPU; the others need the GPU."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_util as ou
from golden_inputs import CASES, make_input
from vstree_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "vstree_b200", "bin", "mkvtree")

needs_ref = pytest.mark.skipif(not ou.have_ref(), reason="oracle/_ref/mkvtree.x not built")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(OURS):
        import __graft_entry__
        __graft_entry__.build()


def run(binary, args, cwd):
    pr = subprocess.run([binary] + args, cwd=cwd, capture_output=True)
    files = {}
    for fn in sorted(os.listdir(cwd)):
        if "." in fn and fn.split(".")[-1] in ("prj", "al1", "al2", "tis", "ois", "des", "sds", "ssp", "suf", "lcp",
                                                "llv", "bwt", "bck", "sti1", "skp"):
            with open(os.path.join(cwd, fn), "rb") as f:
                files[fn] = f.read()
    return pr.returncode, pr.stdout.decode(errors="replace"), pr.stderr.decode(errors="replace"), files


def transcript(out):
    """-v output without the space-peak line (its numbers are build specific)."""
    return [l for l in out.splitlines() if not l.startswith("overall space peak:")]


def compare(tmp_path, inputs, args, expect_rc=0):
    """Runs both binaries in twin directories holding the same input files; compares everything."""
    da, db = tmp_path / "ref", tmp_path / "ours"
    for d in (da, db):
        d.mkdir()
        for name, content in inputs.items():
            (d / name).write_bytes(content)
    ra = run(ou.REF_MKVTREE, args, str(da))
    rb = run(OURS, args, str(db))
    for d in (da, db):  # inputs are not part of the comparison
        pass
    assert rb[0] == ra[0], f"exit code {rb[0]} vs reference {ra[0]}: {rb[2]} | {ra[2]}"
    if expect_rc is not None:
        assert ra[0] == expect_rc, ra[2]
    assert transcript(rb[1]) == transcript(ra[1])
    if ra[0] != 0:
        strip = lambda s, b: s.replace(b, "PROG")
        assert strip(rb[2], OURS) == strip(ra[2], ou.REF_MKVTREE)
    fa = {k: v for k, v in ra[3].items() if k not in inputs}
    fb = {k: v for k, v in rb[3].items() if k not in inputs}
    assert sorted(fa) == sorted(fb), f"files differ: {sorted(fb)} vs reference {sorted(fa)}"
    for k in fa:
        assert fa[k] == fb[k], f"{k} differs from the reference ({len(fb[k])} vs {len(fa[k])} bytes)"
    return ra, rb


DNA = (b">seq1 first\nACGTNACGTTTGACCA\nacgtnnacg\n>seq2\nGGGACGTRYACGT\n>s3 third one\nNNACGATCGATCGAT\n"
       + b">s4 long\n" + b"\n".join(bytes(np.random.default_rng(5).choice(list(b"ACGTacgtN"), size=60).tolist())
                                     for _ in range(12)) + b"\n>s5\nACGTACGTACGTACGTACGTTTTTTTTTTTTTTTTTTTTTTTTT\n")
PROT = b">p1 alpha\nLVIFKREDAGSTNQYWPHMC\nXUBZLLV\n>p2\nMKVLAAGIVGLLLA*\n"


# ---------------------------------------------------------------------------------------------------------
# host-side only: no GPU involved
@needs_ref
@pytest.mark.parametrize("args", [
    ["-db", "in.fna", "-dna", "-tis", "-ois", "-v"],
    ["-db", "in.fna", "-dna", "-tis", "-indexname", "myidx"],
    ["-db", "in.fna", "-dna", "-ois", "-pl", "2", "-v"],
    ["-db", "in.fna", "-dna", "-pl", "-tis", "-v"],
    ["-dna", "-v", "-tis", "-db", "in.fna"],
])
def test_host_tables_dna(tmp_path, args):
    compare(tmp_path, {"in.fna": DNA}, args)


@needs_ref
@pytest.mark.parametrize("args", [
    ["-db", "in.fna", "-dna", "-tis", "-ois", "-rev", "-v"],
    ["-db", "in.fna", "-dna", "-tis", "-ois", "-cpl", "-v"],
    ["-db", "in.fna", "-dna", "-tis", "-ois", "-rev", "-cpl", "-v"],
    ["-db", "in.fna", "-dna", "-cpl", "-rev", "-tis", "-indexname", "named", "-v"],
    ["-db", "in.fna", "-dna", "-tis", "-pl", "-rev"],
])
def test_reverse_and_complement_host_tables(tmp_path, args):
    """-rev / -cpl transform the sequence only (.ois, .des, .sds, .ssp and the project counts stay) and name the index
    <name>.rev / .cpl / .rcp (src/Mkvtree/mkvinput.c:273-309, mkvtree.c:143-161)."""
    ra, rb = compare(tmp_path, {"in.fna": DNA}, args)
    assert any(k.split(".")[-2] in ("rev", "cpl", "rcp") for k in rb[3])


@needs_ref
def test_reverse_of_odd_and_single_symbol_inputs(tmp_path):
    compare(tmp_path, {"a.fna": b">x\nACGTN\n"}, ["-db", "a.fna", "-dna", "-tis", "-rev", "-cpl"])
    (tmp_path / "two").mkdir()
    compare(tmp_path / "two", {"a.fna": b">x\nG\n"}, ["-db", "a.fna", "-dna", "-tis", "-cpl"])


@needs_ref
@pytest.mark.parametrize("extra", [["-cpl"], ["-rev", "-cpl"]])
def test_complement_of_protein_is_the_reference_error(tmp_path, extra):
    compare(tmp_path, {"in.faa": PROT}, ["-db", "in.faa", "-protein", "-tis"] + extra, expect_rc=1)


@needs_ref
def test_reverse_of_protein(tmp_path):
    compare(tmp_path, {"in.faa": PROT}, ["-db", "in.faa", "-protein", "-tis", "-ois", "-rev", "-v"])


QUERY = b">q1 query\nGGGTTTACGTNNAC\n>q2\nACNNA\n>q3 last\nTTTTACGT\n"


@needs_ref
@pytest.mark.parametrize("args", [
    ["-db", "in.fna", "-q", "q.fna", "-dna", "-tis", "-ois", "-indexname", "both", "-v"],
    ["-q", "q.fna", "q2.fna", "-db", "in.fna", "b.fna", "-dna", "-tis", "-ois", "-pl", "-indexname", "both", "-v"],
    ["-db", "in.fna", "-q", "q.fna", "-dna", "-tis", "-rev", "-indexname", "both"],
])
def test_query_files_follow_the_database_files(tmp_path, args):
    """-q: the query sequences are appended behind the database sequences; the project file names the files
    `queryfile=` and splits numofsequences (src/Mkvtree/mkvtree.c:438-452,502-515; mkvprocess.c:403-503)."""
    ra, rb = compare(tmp_path, {"in.fna": DNA, "q.fna": QUERY, "q2.fna": b">only\nACGTACGTTT\n",
                                "b.fna": b">b1\nTTGGCCAA\n>b2\nNNNACGT\n"}, args)
    prj = [v for k, v in rb[3].items() if k.endswith(".prj")][0].decode()
    assert "queryfile=" in prj and "numofquerysequences=0" not in prj


@needs_ref
@pytest.mark.parametrize("args", [
    ["-q", "q.fna", "-dna", "-tis"],                               # no database file
    ["-db", "in.fna", "-q", "q.fna", "-dna", "-tis"],              # several files need -indexname
    ["-db", "in.fna", "-q", "-dna", "-tis", "-indexname", "x"],    # -q without a file
])
def test_query_option_errors(tmp_path, args):
    compare(tmp_path, {"in.fna": DNA, "q.fna": QUERY}, args, expect_rc=None)


EMBL = (b"ID   AB000001   standard; DNA; FUN; 60 BP.\nXX\nAC   AB000001;\nXX\n"
        b"DE   Made-up spacer sequence one, partial\nDE   second description line.\nXX\n"
        b"SQ   Sequence 60 BP; 15 A; 15 C; 15 G; 15 T; 0 other;\n"
        b"     acgtacgtac gtacgtacgn nacgtacgta cgtacgtacg tacgtacgta cgtacgtacg        60\n//\n"
        b"ID   AB000002   standard; DNA; FUN; 24 BP.\nXX\nDE   Second record\nXX\n"
        b"SQ   Sequence 24 BP;\n     ttttggggcc ccaaaarykm acgt        24\n//\n")
GENBANK = (b"LOCUS       TST00001      48 bp    DNA             PLN       01-JAN-2000\n"
           b"DEFINITION  Synthetic test locus one,\n            continued.\nACCESSION   TST00001\n"
           b"ORIGIN      \n        1 acgtacgtnn acgtacgtac gtacgtacgt acgtacgtac gtacgtac\n//\n"
           b"LOCUS       TST00002      20 bp    DNA             PLN       01-JAN-2000\n"
           b"DEFINITION  Locus two.\nORIGIN\n        1 ggggccccaa aattttacgt\n//\n")
SWISS = (b"ID   TEST_HUMAN     STANDARD;      PRT;    30 AA.\nAC   P00001;\n"
         b"DE   MADE-UP PROTEIN (TEST).\nSQ   SEQUENCE   30 AA;  3000 MW;  ABCDEF CRC32;\n"
         b"     MKVLAAGIVG LLLAQWERTY PLKHGFDSAC\n//\n")


@needs_ref
@pytest.mark.parametrize("name,data,flag", [("e.embl", EMBL, "-dna"), ("g.gb", GENBANK, "-dna"),
                                            ("s.sp", SWISS, "-protein")])
def test_flat_file_formats(tmp_path, name, data, flag):
    """GenBank, EMBL and SwissProt records are read as FASTA records ">ID description" (parsemultiform.c:225-385)."""
    compare(tmp_path, {name: data}, ["-db", name, flag, "-tis", "-ois", "-pl", "-v"])
    (tmp_path / "gz").mkdir()
    compare(tmp_path / "gz", {name + ".gz": gzip.compress(data)}, ["-db", name + ".gz", flag, "-tis", "-ois"])


@needs_ref
@pytest.mark.parametrize("data", [
    EMBL.replace(b"60 BP.\n", b"60 BASES\n", 1),                      # first line does not tell the type
    EMBL.replace(b"\nDE   ", b"\nDX   "),                              # no description key
    EMBL.replace(b"\nSQ   ", b"\nSX   "),                              # no sequence key
    GENBANK.replace(b"\nORIGIN", b"\nORIGEN"),
    GENBANK.replace(b"\n//\n", b"\n/ \n", 1),                         # no second slash
    b"\n \n" + GENBANK, b" " + EMBL, b"\t\n" + GENBANK,              # a flat file has to start with its key
])
def test_flat_file_errors(tmp_path, data):
    compare(tmp_path, {"x.txt": data}, ["-db", "x.txt", "-dna", "-tis"], expect_rc=1)


@needs_ref
@pytest.mark.parametrize("fixture", ["Emblfile", "GenBankfile", "Emblfile.gz", "GenBankfile.gz"])
def test_flat_file_fixtures_of_the_reference(tmp_path, fixture):
    path = os.path.join("/root/reference/src/Mkvtree/DATA", fixture)
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    with open(path, "rb") as f:
        compare(tmp_path, {fixture: f.read()}, ["-db", fixture, "-dna", "-tis", "-ois", "-v"])


TEXT = b"Hello World xyz abcabcabc, said the text; and the text said hello again.\n"


@needs_ref
@pytest.mark.parametrize("data,args", [
    (TEXT, ["-tis", "-ois", "-v"]),                       # no alphabet option, no -pl: the bytes themselves
    (TEXT, ["-tis", "-ois", "-pl", "1", "-v"]),           # bytes numbered in byte order, .al2
    (TEXT, ["-tis", "-pl", "2", "-indexname", "t"]),
    (bytes([65, 66, 255, 65, 254, 66, 0, 1, 65]), ["-tis", "-pl", "1", "-v"]),
    (bytes([65, 66, 255, 65, 254, 66, 0, 1, 65]), ["-tis", "-v"]),
    (TEXT, ["-tis", "-pl", "-v"]),                        # error: -pl needs a number when the alphabet is unknown
    (TEXT, ["-tis", "-pl", "9"]),                         # error: too large for this text and alphabet
    (DNA, ["-tis", "-v"]),                                # error: FASTA needs an alphabet option
    (EMBL, ["-tis"]), (GENBANK, ["-tis"]),
])
def test_without_alphabet_option(tmp_path, data, args):
    """No -dna/-protein/-smap: plain text whose alphabet is the set of its bytes (src/Mkvtree/mkvinput.c:100-170,
    findalphabet alphabet.c:456-487, .al2 mkvprocess.c:554-577)."""
    compare(tmp_path, {"p.txt": data}, ["-db", "p.txt"] + args, expect_rc=None)


@needs_ref
def test_without_alphabet_option_several_files(tmp_path):
    compare(tmp_path, {"a.txt": TEXT, "b.txt": TEXT}, ["-db", "a.txt", "b.txt", "-tis", "-indexname", "x"],
            expect_rc=1)


@needs_ref
def test_host_tables_protein(tmp_path):
    compare(tmp_path, {"in.faa": PROT}, ["-db", "in.faa", "-protein", "-tis", "-ois", "-pl", "1", "-v"])


@needs_ref
def test_crlf_and_leading_whitespace(tmp_path):
    data = b"\n  \n" + DNA.replace(b"\n", b"\r\n")
    compare(tmp_path, {"in.fna": data}, ["-db", "in.fna", "-dna", "-tis", "-ois", "-v"])


@needs_ref
def test_gzipped_input(tmp_path):
    compare(tmp_path, {"in.fna.gz": gzip.compress(DNA)}, ["-db", "in.fna.gz", "-dna", "-tis", "-ois", "-v"])


@needs_ref
def test_multiple_files_and_plain_member(tmp_path):
    compare(tmp_path, {"a.fna": DNA, "b.fna": b">only\nTTTTGGGGCCCCAAAA\n", "c.txt": b"ACGTACGTNNACGT"},
            ["-db", "a.fna", "b.fna", "c.txt", "-dna", "-tis", "-ois", "-indexname", "all", "-v"])


@needs_ref
def test_plain_input(tmp_path):
    compare(tmp_path, {"p.txt": b"ACGTNACGTACCA"}, ["-db", "p.txt", "-dna", "-tis", "-ois", "-v"])


@needs_ref
def test_symbol_map(tmp_path, monkeypatch):
    smap = b"# comment\naA\ncC\ngG\ntTuU\nnsywrkvbdhmNSYWRKVBDHM x\n"
    compare(tmp_path, {"in.fna": DNA, "MyMap": smap}, ["-db", "in.fna", "-smap", "MyMap", "-tis", "-ois", "-pl", "1",
                                                       "-v"])


@needs_ref
def test_existing_index_is_removed(tmp_path):
    compare(tmp_path, {"in.fna": DNA, "in.fna.suf": b"stale", "in.fna.prj": b"stale", "in.fna.llv": b"x"},
            ["-db", "in.fna", "-dna", "-tis", "-v"])


@needs_ref
@pytest.mark.parametrize("args,inputs", [
    ([], {}),
    (["-db"], {}),
    (["-dna", "-tis"], {}),
    (["-db", "in.fna", "-dna", "-tis", "-foo"], {"in.fna": DNA}),
    (["-db", "in.fna", "-dna", "-tis", "-tis"], {"in.fna": DNA}),
    (["-db", "in.fna", "-dna", "-bck"], {"in.fna": DNA}),
    (["-db", "in.fna", "-dna", "-sti1"], {"in.fna": DNA}),
    (["-db", "in.fna", "-dna", "-protein", "-tis"], {"in.fna": DNA}),
    (["-db", "in.fna", "-dna", "-smap", "x", "-tis"], {"in.fna": DNA}),
    (["-db", "in.fna", "in.fna", "-dna", "-tis"], {"in.fna": DNA}),
    (["-db", "in.fna", "-dna", "-indexname", "x"], {"in.fna": DNA}),
    (["-db", "in.fna", "-dna", "-tis", "-indexname"], {"in.fna": DNA}),
    (["-db", "in.fna", "-dna", "-tis", "-pl", "0"], {"in.fna": DNA}),
    (["-db", "in.fna", "-dna", "-tis", "-pl", "1024"], {"in.fna": DNA}),
    (["-db", "in.fna", "-dna", "-tis", "-pl", "9"], {"in.fna": DNA}),
    (["-db", "in.fna", "-dna", "-tis", "stray"], {"in.fna": DNA}),
    (["-db", "in.fna", "-dna", "-skp"], {"in.fna": DNA}),                          # -skp requires -lcp and -suf
    (["-db", "in.fna", "-dna", "-lcp", "-skp"], {"in.fna": DNA}),
    (["-db", "in.fna", "-dna", "-allout"], {"in.fna": DNA}),                       # -allout requires -pl
    (["-db", "in.fna", "-dna", "-pl", "-allout", "-tis"], {"in.fna": DNA}),
    (["-db", "in.fna", "-protein", "-tis"], {"in.fna": DNA}),                      # lower-case letters are illegal
    (["-db", "in.fna", "-dna", "-tis"], {"in.fna": b">a\nACGT\n>b\n\n>c\nAC\n"}),  # empty sequence
    (["-db", "in.fna", "-dna", "-tis"], {"in.fna": b">a\nAC>GT\n"}),               # '>' inside a line
    (["-db", "in.fna", "-dna", "-tis"], {"in.fna": b">a\nACGT\n>b\n"}),            # last sequence empty
    (["-db", "p.txt", "-dna", "-tis"], {"p.txt": b"ACGT\n"}),                      # plain input, illegal newline
    (["-db", "p.txt", "-dna", "-tis"], {"p.txt": b"ACXGT"}),
])
def test_error_behaviour(tmp_path, args, inputs):
    ra, rb = compare(tmp_path, inputs, args, expect_rc=None)
    assert ra[0] != 0


def test_missing_file_message(tmp_path):
    rc, out, err, files = run(OURS, ["-db", "nonexist", "-dna", "-tis"], str(tmp_path))
    assert rc == 1 and 'cannot read file "nonexist": open failed: No such file or directory' in err


@needs_ref
@pytest.mark.parametrize("name", ["kat_survey", "dna_specials_50k", "protein_30k", "human_like_400k"])
def test_host_tables_on_golden_inputs(tmp_path, name):
    tis, sigma = make_input(name)
    kind = "dna" if CASES[name][2] == "-dna" else "protein"
    compare(tmp_path, {"g.fa": synth.to_fasta(tis, kind)}, ["-db", "g.fa", CASES[name][2], "-tis", "-ois", "-v"])


@needs_ref
def test_host_tables_on_reference_fixtures(tmp_path, refdata):
    for rel, flag in (("testdata/at100K1", "-dna"), ("testdata/swiss10K", "-protein"),
                      ("testdata/Grumbach/Wildcards.fna", "-dna")):
        d = tmp_path / os.path.basename(rel)
        d.mkdir()
        with open(os.path.join(refdata, rel), "rb") as f:
            compare(d, {"in": f.read()}, ["-db", "in", flag, "-tis", "-ois", "-pl", "-v"])


# ---------------------------------------------------------------------------------------------------------
# the whole path: needs the GPU
@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("args,data", [
    (["-db", "in.fna", "-dna", "-pl", "-suf", "-lcp", "-bwt", "-tis", "-ois", "-v"], DNA),
    (["-db", "in.fna", "-dna", "-pl", "2", "-suf", "-lcp", "-bwt", "-bck", "-sti1", "-tis", "-ois", "-v"], DNA),
    (["-db", "in.fna", "-dna", "-pl", "1", "-sti1", "-v"], DNA),                   # writes lcp llv bck sti1, no suf
    (["-db", "in.fna", "-dna", "-suf", "-lcp", "-v"], DNA),                        # no -pl: one global sort
    (["-db", "in.fna", "-dna", "-v"], DNA),                                        # sorts, writes nothing
    (["-db", "in.fna", "-dna", "-pl", "-bwt", "-indexname", "x"], DNA),
    (["-db", "in.fna", "-protein", "-pl", "-suf", "-lcp", "-bwt", "-bck", "-sti1", "-tis", "-ois", "-v"], PROT),
    (["-db", "in.fna", "-dna", "-pl", "-suf", "-lcp", "-skp", "-v"], DNA),         # skip table, written last
    (["-db", "in.fna", "-dna", "-suf", "-lcp", "-skp"], DNA),                      # without -pl
    (["-db", "in.fna", "-dna", "-pl", "-allout", "-v"], DNA),
    (["-db", "in.fna", "-protein", "-pl", "1", "-allout", "-v"], PROT),
    (["-db", "in.fna", "-dna", "-pl", "-suf", "-lcp", "-bwt", "-bck", "-sti1", "-tis", "-ois", "-rev", "-v"], DNA),
    (["-db", "in.fna", "-dna", "-pl", "-suf", "-lcp", "-bwt", "-tis", "-rev", "-cpl", "-v"], DNA),   # <name>.rcp.*
    (["-db", "in.fna", "-dna", "-pl", "2", "-allout", "-cpl"], DNA),
    (["-db", "in.fna", "-protein", "-pl", "-suf", "-lcp", "-bwt", "-rev", "-v"], PROT),
])
def test_full_cli_small(tmp_path, args, data):
    compare(tmp_path, {"in.fna": data}, args)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("data,args", [
    (TEXT * 3, ["-pl", "1", "-suf", "-lcp", "-bwt", "-bck", "-sti1", "-tis", "-v"]),
    (TEXT * 60, ["-pl", "2", "-allout", "-v"]),
    (TEXT * 3, ["-suf", "-lcp", "-bwt", "-tis", "-v"]),                                   # raw bytes, one global sort
    (bytes([65, 66, 255, 65, 254, 66, 0, 1, 65, 66, 65]) * 5, ["-pl", "1", "-suf", "-lcp", "-bwt", "-tis"]),
    (bytes([65, 66, 255, 65, 254, 66, 0, 1, 65, 66, 65]) * 5, ["-suf", "-lcp", "-bwt", "-tis"]),  # all bytes are letters here too
])
def test_full_cli_without_alphabet_option(tmp_path, data, args):
    compare(tmp_path, {"p.txt": data}, ["-db", "p.txt"] + args)


@needs_ref
@pytest.mark.gpu
def test_full_cli_query_files(tmp_path):
    compare(tmp_path, {"in.fna": DNA, "q.fna": QUERY},
            ["-db", "in.fna", "-q", "q.fna", "-dna", "-pl", "-suf", "-lcp", "-bwt", "-bck", "-sti1", "-tis", "-ois",
             "-indexname", "dbq", "-v"])


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["dna_specials_50k", "dna_repeat_200k", "dna_twocopies_60k", "protein_200k",
                                  "human_like_400k", "poly_a_5000"])
def test_full_cli_golden_inputs(tmp_path, name):
    tis, sigma = make_input(name)
    flag = CASES[name][2]
    kind = "dna" if flag == "-dna" else "protein"
    compare(tmp_path, {"g.fa": synth.to_fasta(tis, kind)},
            ["-db", "g.fa", flag, "-pl", "-suf", "-lcp", "-bwt", "-bck", "-sti1", "-tis", "-ois", "-v"])
    for fn in os.listdir(tmp_path):
        shutil.rmtree(tmp_path / fn)
    compare(tmp_path, {"g.fa": synth.to_fasta(tis, kind)}, ["-db", "g.fa", flag, "-pl", "-allout", "-v"])


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("ranks", [2, 5])
@pytest.mark.parametrize("name", ["dna_specials_50k", "human_like_400k", "protein_200k", "kat_survey"])
def test_full_cli_several_gpus_one_index(tmp_path, monkeypatch, name, ranks):
    """MKVTREE_GPUS: the driver's ranks build the ranges of ONE index (here all on device 0); the files are the
    reference's byte for byte."""
    monkeypatch.setenv("MKVTREE_GPUS", str(ranks))
    monkeypatch.setenv("MKVTREE_GPUS_ONE_DEVICE", "1")
    tis, sigma = make_input(name)
    flag = CASES[name][2]
    kind = "dna" if flag == "-dna" else "protein"
    compare(tmp_path, {"g.fa": synth.to_fasta(tis, kind)},
            ["-db", "g.fa", flag, "-pl", "-suf", "-lcp", "-bwt", "-bck", "-tis", "-ois", "-v"])


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("ranks", [2, 3])
def test_cli_several_gpus_sti1(tmp_path, monkeypatch, ranks):
    """-sti1 on several ranks: the table is indexed by text position, every rank contributes its slice of the text."""
    monkeypatch.setenv("MKVTREE_GPUS", str(ranks))
    monkeypatch.setenv("MKVTREE_GPUS_ONE_DEVICE", "1")
    tis, sigma = make_input("dna_specials_50k")
    compare(tmp_path, {"g.fa": synth.to_fasta(tis, "dna")}, ["-db", "g.fa", "-dna", "-pl", "-sti1", "-bwt", "-v"])


@pytest.mark.gpu
def test_cli_several_gpus_refuses_skp(tmp_path, monkeypatch):
    monkeypatch.setenv("MKVTREE_GPUS", "2")
    monkeypatch.setenv("MKVTREE_GPUS_ONE_DEVICE", "1")
    (tmp_path / "in.fna").write_bytes(DNA)
    rc, out, err, files = run(OURS, ["-db", "in.fna", "-dna", "-pl", "-suf", "-lcp", "-skp"], str(tmp_path))
    assert rc == 1 and "sharded build" in err


@needs_ref
@pytest.mark.gpu
def test_full_cli_multifile_gz(tmp_path):
    compare(tmp_path, {"a.fna.gz": gzip.compress(DNA), "b.fna": b">only\nTTTTGGGGCCCCAAAANNNNACGT\n"},
            ["-db", "a.fna.gz", "b.fna", "-dna", "-pl", "-suf", "-lcp", "-bwt", "-bck", "-tis", "-indexname", "two",
             "-v"])


@pytest.mark.gpu
def test_at1mb_scale_cli_against_reference(tmp_path):
    """Config 1 of BASELINE.json at its real size: a 1 Mbp multi-FASTA with ~2000 records and ~2 % specials."""
    if not ou.have_ref():
        pytest.skip("oracle/_ref/mkvtree.x did not travel")
    tis = synth.dna_with_specials(800000, seed=17, nfrac=0.02, mean_run=8, records=1900)
    compare(tmp_path, {"at.fna": synth.to_fasta(tis, "dna")},
            ["-db", "at.fna", "-dna", "-pl", "-suf", "-lcp", "-bwt", "-tis", "-ois", "-bck", "-sti1", "-v"])


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json configs[0] on the reference's OWN fixture: src/testdata/at1MB (committed gzipped under tests/golden/,
# with Grumbach/Wildcards.fna), command line of SURVEY.md 8d.  The sha256 prefixes are the ones SURVEY.md 8c recorded
# from the reference build; the live reference binary is compared on top when it travelled.
AT1MB_SHA = {"suf": "d35617c3307679b0", "lcp": "010a6436643f0908", "llv": "5bea62168cc44547", "bwt": "8ce08b07227a8a93",
             "bck": "3d3d5e340324cebc", "sti1": "1b8f0b5f567ebebc", "tis": "0d91a7b28cfa0b7a", "ois": "02747886eee026fd",
             "des": "8924f1da9e6ede03", "sds": "d245f9f21ed02b3f", "ssp": "8b7e1d182c7e701a", "al1": "07fca85c244e5b6a"}
AT1MB_PRJ = {"totallength": 772376, "numofsequences": 1952, "specialcharacters": 18923, "specialranges": 18010,
             "longest": 137851, "largelcpvalues": 2317, "maxbranchdepth": 517, "prefixlength": 7}
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def at1mb_bytes():
    with gzip.open(os.path.join(GOLDEN_DIR, "at1MB.gz"), "rb") as f:
        return f.read()


def test_at1mb_fixture_is_the_reference_file():
    import hashlib
    data = at1mb_bytes()
    assert len(data) == 999815 and hashlib.sha256(data).hexdigest()[:16] == "8cb0a609e8c559b6"


@pytest.mark.gpu
def test_config1_at1mb_survey_hashes(tmp_path):
    """`mkvtree -db at1MB -dna -pl -suf -lcp -bwt -tis -ois` (+ -bck; .sti1 in a second run: -sti1 suppresses .suf)."""
    import hashlib
    (tmp_path / "at1MB").write_bytes(at1mb_bytes())
    rc, out, err, files = run(OURS, ["-db", "at1MB", "-dna", "-pl", "-suf", "-lcp", "-bwt", "-tis", "-ois", "-bck"],
                              str(tmp_path))
    assert rc == 0, err
    rc, out, err, files2 = run(OURS, ["-db", "at1MB", "-indexname", "two", "-dna", "-pl", "-sti1"], str(tmp_path))
    assert rc == 0, err
    files["at1MB.sti1"] = files2["two.sti1"]
    for suffix, want in AT1MB_SHA.items():
        got = hashlib.sha256(files["at1MB." + suffix]).hexdigest()[:16]
        assert got == want, f".{suffix}: {got} != {want} (SURVEY.md 8c)"
    prj = ou.parse_prj(files["at1MB.prj"].decode())
    for k, v in AT1MB_PRJ.items():
        assert prj[k] == v, (k, prj[k], v)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("fixture,args", [
    ("at1MB", ["-dna", "-pl", "-suf", "-lcp", "-bwt", "-tis", "-ois", "-bck", "-v"]),
    ("at1MB", ["-dna", "-pl", "-sti1", "-skp", "-suf", "-lcp"]),
    ("Wildcards.fna", ["-dna", "-pl", "-suf", "-lcp", "-bwt", "-tis", "-ois", "-bck", "-sti1", "-v"]),
    ("Wildcards.fna", ["-dna", "-suf", "-lcp", "-bwt"])])
def test_reference_fixtures_cli_against_live_reference(tmp_path, fixture, args):
    data = at1mb_bytes() if fixture == "at1MB" else open(os.path.join(GOLDEN_DIR, fixture), "rb").read()
    compare(tmp_path, {fixture: data}, ["-db", fixture] + args, expect_rc=None)


@needs_ref
@pytest.mark.gpu
def test_c2_cli_against_live_reference_at_full_size(tmp_path_factory):
    """Process seam at BASELINE.json configs[1]: both binaries on the same 100 Mbp FASTA file, every file compared
    (cmp), on tmpfs.  About 20 s of reference time."""
    import tempfile
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    tis = synth.random_dna(100_000_000, seed=42)
    args = ["-db", "c2.fna", "-dna", "-pl", "-tis", "-ois", "-suf", "-lcp", "-bwt", "-bck"]
    with tempfile.TemporaryDirectory(dir=shm) as d:
        fasta = synth.to_fasta(tis, "dna", width=1 << 20)
        del tis
        for sub in ("ref", "ours"):
            os.mkdir(os.path.join(d, sub))
            with open(os.path.join(d, sub, "c2.fna"), "wb") as f:
                f.write(fasta)
        del fasta
        pa = subprocess.run([ou.REF_MKVTREE] + args, cwd=os.path.join(d, "ref"), capture_output=True)
        pb = subprocess.run([OURS] + args, cwd=os.path.join(d, "ours"), capture_output=True)
        assert pa.returncode == 0 and pb.returncode == 0, (pa.stderr, pb.stderr)
        names = sorted(fn for fn in os.listdir(os.path.join(d, "ref")) if fn != "c2.fna")
        assert names == sorted(fn for fn in os.listdir(os.path.join(d, "ours")) if fn != "c2.fna")
        assert {"c2.fna.suf", "c2.fna.lcp", "c2.fna.llv", "c2.fna.bwt", "c2.fna.bck", "c2.fna.prj"} <= set(names)
        for fn in names:
            pr = subprocess.run(["cmp", os.path.join(d, "ref", fn), os.path.join(d, "ours", fn)], capture_output=True)
            assert pr.returncode == 0, f"{fn}: {pr.stdout.decode()}"
