"""Regenerates tests/golden/golden.json by running the UNMODIFIED reference binary (oracle/_ref/mkvtree.x) on the
seeded inputs of tests/golden_inputs.py.  Run here (needs oracle/_ref built from /root/reference):
    python tests/golden/make_golden.py
Stores a 64-bit sha256 prefix of every table plus the .prj facts; tiny cases also keep the full tables."""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_util as ou  # noqa: E402
from golden_inputs import CASES, make_input  # noqa: E402
from vstree_b200 import synth  # noqa: E402


def sha(b):
    return hashlib.sha256(b).hexdigest()[:16]


def main():
    cases = []
    for name, (_, sigma, flag, pl) in CASES.items():
        tis, sigma = make_input(name)
        fasta = synth.to_fasta(tis, "dna" if flag == "-dna" else "protein")
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "in.fna")
            with open(path, "wb") as f:
                f.write(fasta)
            args = [flag, "-pl"] + ([str(pl)] if pl else []) + ["-tis", "-ois", "-suf", "-lcp", "-bwt", "-bck", "-sti1"]
            rc, out, err, files = ou.run_mkvtree(ou.REF_MKVTREE, path, args)
            # the skip table in a second run (-sti1 suppresses .suf, -skp needs it: mkvtree.c:376,474-475)
            args2 = [flag, "-pl"] + ([str(pl)] if pl else []) + ["-suf", "-lcp", "-skp"]
            rc2, out2, err2, files2 = ou.run_mkvtree(ou.REF_MKVTREE, path, args2)
        assert rc == 0, (name, err)
        assert rc2 == 0, (name, err2)
        files["skp"] = files2["skp"]
        assert files["tis"] == tis.tobytes(), f"{name}: FASTA round trip does not reproduce tis"
        prj = ou.parse_prj(files["prj"].decode())
        entry = {"name": name, "sigma": sigma, "flag": flag, "n": int(len(tis)), "prefixlength": prj["prefixlength"],
                 "longest": prj["longest"], "maxbranchdepth": prj["maxbranchdepth"],
                 "largelcpvalues": prj["largelcpvalues"], "specialcharacters": prj["specialcharacters"],
                 "specialranges": prj["specialranges"], "numofsequences": prj["numofsequences"],
                 "sha": {k: sha(files[k]) for k in ("tis", "ois", "suf", "lcp", "llv", "bwt", "bck", "sti1", "des", "sds", "skp")
                         if k in files}}
        if "ssp" in files:
            entry["sha"]["ssp"] = sha(files["ssp"])
        if len(tis) <= 64:
            entry["tables"] = {k: np.frombuffer(files[k], ou.TABLE_DTYPES[k]).tolist()
                               for k in ("suf", "lcp", "llv", "bwt", "bck", "sti1", "skp")}
        cases.append(entry)
        print(name, entry["n"], "k=%d" % entry["prefixlength"], "maxlcp=%d" % entry["maxbranchdepth"],
              "llv=%d" % entry["largelcpvalues"])
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump({"generator": "oracle/_ref/mkvtree.x (reference vstree 2.3.1, unmodified)", "cases": cases}, f,
                  indent=1)


if __name__ == "__main__":
    main()
