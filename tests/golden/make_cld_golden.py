"""Regenerates tests/golden/cld_golden.json: the child tables (.cld, .cld1) the UNMODIFIED reference tools write for
the seeded inputs of tests/golden_inputs.py.  Run here (needs oracle/_ref built from /root/reference):
    python tests/golden/make_cld_golden.py                   # the seeded small cases
    python tests/golden/make_cld_golden.py --full c2 c5m c3  # BASELINE-size cases (minutes of reference time each)
Per case: oracle/_ref/mkvtree.x -suf -lcp (+ -tis, mkcld maps SUFTAB | LCPTAB), then oracle/_ref/mkcld.x on the index
(src/Mkvtree/mkcld.mn.c).  Stores a 64-bit sha256 prefix of both files, the sha of the .lcp/.llv they were made from, and
for tiny cases the tables themselves.

The reference never writes the last byte of .cld1 (nor the one before it when lcp[n-1] > lcp[n]: compresscldtab's loops
stop below totallength, src/kurtz/mkcld.c:230-283) and leaves there what malloc returned; `cld1_unwritten` lists those
indices and the golden bytes at them are forced to 0 before hashing, which is what this repository writes."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_util as ou  # noqa: E402
from golden_inputs import CASES, make_input  # noqa: E402
from vstree_b200 import synth  # noqa: E402

REF_MKCLD = os.path.join(ou.ORACLE_DIR, "_ref", "mkcld.x")


def sha(b):
    return hashlib.sha256(b).hexdigest()[:16]


def reference_childtab(tis, flag, pl, workdir):
    """(lcp bytes, llv bytes, cld bytes, cld1 bytes) as the reference tools write them."""
    path = os.path.join(workdir, "in.fna")
    with open(path, "wb") as f:
        f.write(synth.to_fasta(tis, "dna" if flag == "-dna" else "protein"))
    args = [flag, "-pl"] + ([str(pl)] if pl else []) + ["-tis", "-suf", "-lcp"]
    rc, out, err, files = ou.run_mkvtree(ou.REF_MKVTREE, path, args, keepdir=workdir)
    assert rc == 0, err
    pr = subprocess.run([REF_MKCLD, "idx"], cwd=workdir, capture_output=True)
    assert pr.returncode == 0, pr.stderr.decode(errors="replace")
    out = {}
    for t in ("lcp", "llv", "cld", "cld1"):
        with open(os.path.join(workdir, "idx." + t), "rb") as f:
            out[t] = f.read()
    return out


def unwritten_cld1(lcp, llv, cld):
    """Indices of .cld1 that none of compresscldtab's three loops writes."""
    n = len(lcp) - 1
    L = lcp.astype(np.uint32)
    pairs = llv.reshape(-1, 2)
    L[pairs[:, 0].astype(np.int64)] = pairs[:, 1].astype(np.uint32)
    c = cld.reshape(-1, 3)
    up, down, nextl = c[:, 0], c[:, 1], c[:, 2]
    w = np.zeros(n + 1, bool)
    w[:n] |= (nextl > 0)[:n]                       # loop 1: nextlIndex defined
    w[:n] |= ((nextl == 0) & (down > 0))[:n]       # loop 2: no nextlIndex, but a down field
    tail = np.arange(max(0, n - 255), n)
    w[tail[tail + nextl[tail] == n]] = True        #         or nextlIndex == totallength
    if n >= 2:
        w[:n - 1] |= (L[:n - 1] > L[1:n]) & ((up[1:n] != 0) | (down[1:n] != 0))   # loop 3
    return np.nonzero(~w)[0]


def make_full(short):
    """BASELINE-size cases (the generators and names of tests/golden/make_bench_golden.py); full sha256."""
    import make_bench_golden as mb
    name = mb.SHORT.get(short, short)
    gen, n, kw, sigma, flag = mb.CASES[name]
    tis = getattr(synth, gen)(n, **kw)
    k = int(ou.oracle_lib().oracle_recommendedprefixlength(sigma, n))
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=shm, prefix="cg_" + name) as d:
        with open(os.path.join(d, "in.fna"), "wb") as f:
            f.write(synth.to_fasta(tis, "dna" if flag == "-dna" else "protein", width=1 << 20))
        del tis
        pr = subprocess.run([ou.REF_MKVTREE, "-db", "in.fna", "-indexname", "idx", flag, "-pl", str(k), "-tis", "-suf",
                             "-lcp"], cwd=d, capture_output=True)
        assert pr.returncode == 0, pr.stderr.decode(errors="replace")
        pr = subprocess.run([REF_MKCLD, "idx"], cwd=d, capture_output=True)
        assert pr.returncode == 0, pr.stderr.decode(errors="replace")
        lcp = np.fromfile(os.path.join(d, "idx.lcp"), np.uint8)
        llv = np.fromfile(os.path.join(d, "idx.llv"), np.uint64)
        cld = np.fromfile(os.path.join(d, "idx.cld"), np.uint8)
        cld1 = np.fromfile(os.path.join(d, "idx.cld1"), np.uint8)
    un = unwritten_cld1(lcp, llv, cld)
    cld1[un] = 0
    full = lambda a: hashlib.sha256(a.tobytes()).hexdigest()
    return name, {"n": n, "prefixlength": k, "sha256": {"lcp": full(lcp), "llv": full(llv), "cld": full(cld),
                                                        "cld1": full(cld1)},
                  "cld1_unwritten": [int(x) for x in un[:16]], "cld1_unwritten_count": int(len(un)),
                  "large_child_values": int(np.count_nonzero(cld == 255))}


def main():
    path = os.path.join(HERE, "cld_golden.json")
    if len(sys.argv) > 2 and sys.argv[1] == "--full":
        with open(path) as f:
            data = json.load(f)
        for short in sys.argv[2:]:
            name, entry = make_full(short)
            data.setdefault("full", {})[name] = entry
            print(name, entry)
            with open(path, "w") as f:
                json.dump(data, f, indent=1)
        return
    full = {}
    if os.path.exists(path):
        with open(path) as f:
            full = json.load(f).get("full", {})
    cases = []
    for name, (_, sigma, flag, pl) in CASES.items():
        tis, sigma = make_input(name)
        with tempfile.TemporaryDirectory() as d:
            f = reference_childtab(tis, flag, pl, d)
        lcp = np.frombuffer(f["lcp"], np.uint8)
        llv = np.frombuffer(f["llv"], np.uint64)
        cld = np.frombuffer(f["cld"], np.uint8)
        cld1 = np.frombuffer(f["cld1"], np.uint8).copy()
        un = unwritten_cld1(lcp, llv, cld)
        garbage = int(np.count_nonzero(cld1[un]))
        cld1[un] = 0
        entry = {"name": name, "n": int(len(tis)), "sha": {"lcp": sha(f["lcp"]), "llv": sha(f["llv"]),
                                                          "cld": sha(f["cld"]), "cld1": sha(cld1.tobytes())},
                 "cld1_unwritten": [int(x) for x in un[:16]], "cld1_unwritten_count": int(len(un)),
                 "cld1_unwritten_nonzero_in_reference_file": garbage}
        if len(tis) <= 64:
            entry["tables"] = {"lcp": lcp.tolist(), "llv": llv.tolist(), "cld": cld.tolist(), "cld1": cld1.tolist()}
        cases.append(entry)
        print(name, entry["n"], entry["sha"], "unwritten", len(un), "garbage", garbage)
    with open(path, "w") as f:
        json.dump({"generator": "oracle/_ref/mkvtree.x + oracle/_ref/mkcld.x (reference vstree 2.3.1, unmodified), "
                                "tests/golden/make_cld_golden.py", "cases": cases, "full": full}, f, indent=1)


if __name__ == "__main__":
    main()
