"""Differential test of the mkvtree host driver against the UNMODIFIED reference binary on seeded random inputs:
FASTA with ragged lines, CR/LF, blanks and tabs inside lines, empty lines, lower case, wildcards, several files,
plain members, gzip, query files, -rev/-cpl, -pl, -indexname -- everything that does not need the GPU (-tis/-ois).
Exit code, -v transcript, error text and every file are compared.  Inputs on which the reference itself breaks down
(it crashes on all-empty input and loses the file borders when a file ends inside a description) are skipped."""
import gzip
import os
import random
import subprocess

import pytest

import oracle_util as ou

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "vstree_b200", "bin", "mkvtree")
EXTS = ("prj", "al1", "al2", "tis", "ois", "des", "sds", "ssp")

pytestmark = pytest.mark.skipif(not ou.have_ref(), reason="oracle/_ref/mkvtree.x not built")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(OURS):
        import __graft_entry__
        __graft_entry__.build()


def run(binary, args, cwd, env=None):
    pr = subprocess.run([binary] + args, cwd=cwd, capture_output=True, env=dict(os.environ, **env) if env else None)
    files = {}
    for fn in sorted(os.listdir(cwd)):
        if "." in fn and fn.split(".")[-1] in EXTS:
            with open(os.path.join(cwd, fn), "rb") as f:
                files[fn] = f.read()
    out = [l for l in pr.stdout.decode(errors="replace").splitlines() if not l.startswith("overall space peak:")]
    return pr.returncode, out, pr.stderr.decode(errors="replace").replace(binary, "PROG"), files


def rand_seq(rng, kind, n):
    alpha = "ACGT" * 6 + "acgt" * 3 + "NnRYrykmsw" if kind == "dna" else "LVIFKREDAGSTNQYWPHMC" * 4 + "XUBZ*"
    return "".join(rng.choice(alpha) for _ in range(n))


def rand_fasta(rng, kind):
    out = []
    if rng.random() < 0.15:
        out.append(rng.choice(["\n", " \n", "\n\n", "\t\n"]))
    for _ in range(rng.choice([1, 1, 2, 3, 5, 12])):
        nl = rng.choice(["\n", "\n", "\n", "\r\n"])
        out.append(">" + "".join(rng.choice("abc XYZ123|_.-") for _ in range(rng.randint(0, 25))) + nl)
        n = rng.choice([0, 1, 2, 5, 30, 61, 200]) if rng.random() < 0.2 else rng.randint(1, 150)
        s = rand_seq(rng, kind, n)
        w = rng.choice([10, 60, 70, 1000])
        for ln in [s[i:i + w] for i in range(0, len(s), w)] or [""]:
            if rng.random() < 0.05:
                ln = ln[:len(ln) // 2] + " " + ln[len(ln) // 2:]
            if rng.random() < 0.03:
                ln += "\t"
            out.append(ln + nl)
        if rng.random() < 0.1:
            out.append(nl)
    txt = "".join(out)
    if rng.random() < 0.1:
        txt = txt.rstrip("\r\n")
    if rng.random() < 0.03:
        txt = txt.replace("A", "!", 1)                                   # an illegal character
    if rng.random() < 0.03 and ">" in txt[1:]:
        txt = txt[:len(txt) // 2] + ">" + txt[len(txt) // 2:]            # '>' inside a line
    return txt.encode()


def make_case(seed):
    rng = random.Random(seed)
    kind = rng.choice(["dna", "dna", "protein"])
    files, names = {}, []
    for i in range(rng.choice([1, 1, 1, 2, 3])):
        name, data = f"f{i}.fa", rand_fasta(rng, kind)
        if rng.random() < 0.1:
            name, data = f"f{i}.txt", rand_seq(rng, kind, rng.randint(1, 80)).encode()   # a plain member
        if rng.random() < 0.1:
            name, data = name + ".gz", gzip.compress(data)
        files[name] = data
        names.append(name)
    args = ["-db"] + names
    if rng.random() < 0.2:
        files["q.fa"] = rand_fasta(rng, kind)
        args += ["-q", "q.fa"]
    args += ["-dna" if kind == "dna" else "-protein"]
    args += rng.choice([["-tis"], ["-tis", "-ois"], ["-ois"], ["-tis", "-ois"]])
    if rng.random() < 0.5:
        args += ["-pl"] + ([str(rng.randint(1, 4))] if rng.random() < 0.5 else [])
    if rng.random() < 0.2:
        args += ["-rev"]
    if rng.random() < 0.2:
        args += ["-cpl"]
    if len(names) + ("-q" in args) > 1 or rng.random() < 0.3:
        args += ["-indexname", "idx"]
    if rng.random() < 0.7:
        args += ["-v"]
    return files, args


@pytest.mark.parametrize("block", range(6))
def test_host_driver_equals_reference_on_random_inputs(tmp_path, block):
    compared = 0
    for seed in range(block * 40, block * 40 + 40):
        files, args = make_case(seed)
        res = []
        # "threaded": the FASTA parser cut into pieces whatever the size of the input (vstree_b200/host/mkvtree.c:
        # parsefastathreaded); it must agree with the sequential one and with the reference on every case
        threaded = {"MKVTREE_PARSE_THREADS": str(2 + seed % 4), "MKVTREE_PARSE_MIN": "1"}
        for which, binary, env in (("ref", ou.REF_MKVTREE, None), ("ours", OURS, {"MKVTREE_PARSE_THREADS": "1"}),
                                   ("threaded", OURS, threaded)):
            d = tmp_path / f"{seed}_{which}"
            d.mkdir()
            for name, content in files.items():
                (d / name).write_bytes(content)
            r = run(binary, args, str(d), env)
            res.append((r[0], r[1], r[2], {k: v for k, v in r[3].items() if k not in files}))
        ref, ours, piecewise = res
        assert piecewise == ours, f"seed {seed} {args}: the threaded parser differs from the sequential one"
        prj = b"".join(v for k, v in ref[3].items() if k.endswith(".prj"))
        if ref[0] < 0 or "cannot find position" in ref[2] or b"1844674407370955" in prj:
            continue  # the reference crashed, or lost the file borders (a file that ends inside a description)
        compared += 1
        assert ours[0] == ref[0], f"seed {seed} {args}: exit code {ours[0]} vs {ref[0]}: {ours[2]} | {ref[2]}"
        assert ours[1] == ref[1], f"seed {seed} {args}: transcript"
        assert ours[2] == ref[2], f"seed {seed} {args}: error text"
        assert sorted(ours[3]) == sorted(ref[3]), f"seed {seed} {args}: files {sorted(ours[3])} vs {sorted(ref[3])}"
        for k in ref[3]:
            assert ours[3][k] == ref[3][k], f"seed {seed} {args}: {k} differs"
    assert compared >= 30


def big_fasta(rng, records, kind="dna"):
    """many records with everything that crosses a piece border: N runs at the ends of sequences (special ranges that
    continue into the separator), empty lines, CR/LF, '>' inside descriptions, ragged lines"""
    out = []
    for r in range(records):
        nl = "\r\n" if rng.random() < 0.1 else "\n"
        out.append(">" + "".join(rng.choice("abc >XYZ123|_.-") for _ in range(rng.randint(0, 30))) + nl)
        body = rand_seq(rng, kind, rng.randint(1, 400))
        if rng.random() < 0.3:
            body += ("N" if kind == "dna" else "X") * rng.randint(1, 5)
        if rng.random() < 0.3:
            body = ("n" if kind == "dna" else "*") * rng.randint(1, 5) + body
        w = rng.choice([7, 60, 61, 1000])
        for ln in [body[i:i + w] for i in range(0, len(body), w)]:
            out.append(ln + nl)
        if rng.random() < 0.1:
            out.append(nl)
    return "".join(out).encode()


@pytest.mark.parametrize("seed,kind", [(1, "dna"), (2, "dna"), (3, "protein")])
def test_threaded_fasta_parser_on_many_records(tmp_path, seed, kind):
    rng = random.Random(1000 + seed)
    data = big_fasta(rng, 1500, kind)
    args = ["-db", "in.fa", "-dna" if kind == "dna" else "-protein", "-tis", "-ois", "-v"]
    outs = {}
    for which, binary, env in [("ref", ou.REF_MKVTREE, None), ("seq", OURS, {"MKVTREE_PARSE_THREADS": "1"})] + [
            (f"t{t}", OURS, {"MKVTREE_PARSE_THREADS": str(t), "MKVTREE_PARSE_MIN": "1"}) for t in (2, 3, 7, 16, 64)]:
        d = tmp_path / which
        d.mkdir()
        (d / "in.fa").write_bytes(data)
        outs[which] = run(binary, args, str(d), env)
    assert outs["ref"][0] == 0 and len(outs["ref"][3]) >= 6
    for which in outs:
        assert outs[which] == outs["ref"], f"{which} differs from the reference"
    # and the pieces really ran
    d = tmp_path / "report"
    d.mkdir()
    (d / "in.fa").write_bytes(data)
    r = run(OURS, args, str(d), {"MKVTREE_PARSE_THREADS": "7", "MKVTREE_PARSE_MIN": "1", "MKVTREE_PARSE_REPORT": "1"})
    assert "parsefasta: 7 pieces" in r[2]


def test_threaded_fasta_parser_falls_back_on_errors(tmp_path):
    """an illegal character or a '>' inside a line in ANY piece: the message (file, line number) of the sequential parser"""
    rng = random.Random(5)
    good = big_fasta(rng, 300).decode()
    pos = good.index("\n", 2 * len(good) // 3)
    for bad in (good[:pos] + "!" + good[pos:], good[:pos] + ">x" + good[pos:]):
        res = []
        for which, binary, env in (("ref", ou.REF_MKVTREE, None),
                                   ("threaded", OURS, {"MKVTREE_PARSE_THREADS": "4", "MKVTREE_PARSE_MIN": "1"})):
            d = tmp_path / f"{which}{len(res)}{len(bad)}"
            d.mkdir()
            (d / "in.fa").write_bytes(bad.encode())
            res.append(run(binary, ["-db", "in.fa", "-dna", "-tis"], str(d)) if env is None else
                       run(binary, ["-db", "in.fa", "-dna", "-tis"], str(d), env))
        assert res[0][0] != 0 and res[1][:3] == res[0][:3], (res[0][2], res[1][2])
