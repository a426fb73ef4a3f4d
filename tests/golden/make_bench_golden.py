"""Full-size goldens of the BASELINE.json configurations, made by the UNMODIFIED reference binary
(oracle/_ref/mkvtree.x, built from /root/reference by oracle/build_ref.sh).  Run here, once per case:

    python tests/golden/make_bench_golden.py c2 c4 c3 c5s [c5m]

Each case: the seeded generator of vstree_b200/synth.py (same call as bench.py's workload), rendered as FASTA on tmpfs,
`mkvtree -db in.fna -indexname idx <flag> -pl <k> -suf -lcp -bwt -bck`; the sha256 of every table file and the .prj
facts are merged into tests/golden/bench_golden.json.  The reference is single-threaded: c2 takes ~10 s, c4 minutes,
c3 tens of minutes, c5s (300 Mbp human-like, exact satellite arrays with lcp values beyond 300 000) about an hour.
The full 3 Gbp C5 is out of the reference's reach in this container (28 GB RSS, 33 GB of output and a sum of lcp
values beyond 1e13 character steps): it is covered by 8-GPU == 1-GPU table hashes plus c5s/c5m against the reference.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle_util as ou  # noqa: E402
from vstree_b200 import synth  # noqa: E402

# name -> (generator, n, kwargs, alphabet size, flag)
CASES = {
    "c2_100mbp": ("random_dna", 100_000_000, {"seed": 42}, 4, "-dna"),
    "c3_250mbp": ("repeat_rich_dna", 250_000_000, {"seed": 7}, 4, "-dna"),
    "c4_500mres": ("random_protein", 500_000_000, {"seed": 11}, 20, "-protein"),
    "c5s_300mbp": ("human_like_dna", 300_000_000, {"seed": 2026, "records": 24}, 4, "-dna"),
    "c5m_30mbp": ("human_like_dna", 30_000_000, {"seed": 2026, "records": 24}, 4, "-dna"),
    "c2n_100mbp": ("dna_with_specials", 100_000_000, {"seed": 43, "nfrac": 0.001, "mean_run": 50, "records": 1000}, 4,
                   "-dna"),
}
SHORT = {"c2": "c2_100mbp", "c3": "c3_250mbp", "c4": "c4_500mres", "c5s": "c5s_300mbp", "c5m": "c5m_30mbp",
         "c2n": "c2n_100mbp"}


def file_sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def make(name):
    gen, n, kw, sigma, flag = CASES[name]
    tis = getattr(synth, gen)(n, **kw)
    k = int(ou.oracle_lib().oracle_recommendedprefixlength(sigma, n))
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=shm, prefix="bg_" + name) as d:
        with open(os.path.join(d, "in.fna"), "wb") as f:
            f.write(synth.to_fasta(tis, "dna" if flag == "-dna" else "protein", width=1 << 20))
        tis_sha = hashlib.sha256(tis.tobytes()).hexdigest()
        del tis
        t0 = time.time()
        pr = subprocess.run([ou.REF_MKVTREE, "-db", "in.fna", "-indexname", "idx", flag, "-pl", str(k), "-suf", "-lcp",
                             "-bwt", "-bck"], cwd=d, capture_output=True)
        secs = time.time() - t0
        if pr.returncode != 0:
            raise SystemExit(f"{name}: reference mkvtree failed: {pr.stderr.decode(errors='replace')}")
        with open(os.path.join(d, "idx.prj")) as f:
            prj = ou.parse_prj(f.read())
        sha = {t: file_sha(os.path.join(d, "idx." + t)) for t in ("suf", "lcp", "llv", "bwt", "bck")}
    return {"generator": gen, "n": n, "kwargs": kw, "sigma": sigma, "flag": flag, "prefixlength": k,
            "reference_seconds": round(secs, 1), "tis_sha256": tis_sha, "sha256": sha,
            "sha": {t: v[:16] for t, v in sha.items()},
            "prj": {f: prj[f] for f in ("totallength", "specialcharacters", "specialranges", "numofsequences", "longest",
                                        "prefixlength", "largelcpvalues", "maxbranchdepth")}}


def main():
    path = os.path.join(HERE, "bench_golden.json")
    for a in sys.argv[1:]:
        name = SHORT.get(a, a)
        entry = make(name)
        # merge under a lock-free read-modify-write: cases are made by separate processes that finish minutes apart
        data = {"generator": "oracle/_ref/mkvtree.x (reference vstree 2.3.1, unmodified), tests/golden/make_bench_golden.py"}
        if os.path.exists(path):
            with open(path) as f:
                data = json.load(f)
        data[name] = entry
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
        os.replace(tmp, path)
        print(name, json.dumps(entry), flush=True)


if __name__ == "__main__":
    main()
