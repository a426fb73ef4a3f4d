"""End-to-end drop-in check at the process seam: the reference mkvtree and this repository's
mkvtree on the same FASTA file, same command line, output on tmpfs; wall time from process start to exit, and every
index file compared byte for byte.
    python tools/cli_compare.py [workload=c2] [runs of the product=3]"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vstree_b200 import synth  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
desc, n, sigma, flag, gen, kw = bench.WORKLOADS[workload]
tis = bench.make_tis(workload)[0]
base = "/dev/shm" if os.path.isdir("/dev/shm") else None
work = tempfile.mkdtemp(prefix="clicmp_", dir=base)
fasta = os.path.join(work, "in.fna")
with open(fasta, "wb") as f:
    f.write(synth.to_fasta(tis, "dna" if flag == "-dna" else "protein"))
opts = [flag, "-pl", "-tis", "-ois", "-suf", "-lcp", "-bwt"]
out = {"workload": desc, "n": int(n), "options": " ".join(opts), "filesystem": "tmpfs" if base else "tmp"}


def run(binary, name):
    d = os.path.join(work, name)
    os.makedirs(d, exist_ok=True)
    t0 = time.perf_counter()
    pr = subprocess.run([binary, "-db", fasta] + opts + ["-indexname", "idx"], cwd=d, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert pr.returncode == 0, pr.stderr[-2000:]
    return dt, d


def hashes(d):
    h = {}
    for fn in sorted(os.listdir(d)):
        with open(os.path.join(d, fn), "rb") as f:
            m = hashlib.sha256()
            while True:
                b = f.read(1 << 24)
                if not b:
                    break
                m.update(b)
        h[fn] = m.hexdigest()[:16]
    return h


prod = os.path.join(ROOT, "vstree_b200", "bin", "mkvtree")
times = []
for i in range(runs):
    dt, dprod = run(prod, "prod")
    times.append(round(dt, 3))
out["product_wall_s"] = times
out["product_Mbp_per_s"] = round(n / min(times) / 1e6, 1)
ref = os.path.join(ROOT, "oracle", "_ref", "mkvtree.x")
if os.path.exists(ref):
    dt, dref = run(ref, "ref")
    out["reference_wall_s"] = round(dt, 3)
    out["reference_Mbp_per_s"] = round(n / dt / 1e6, 1)
    hp, hr = hashes(dprod), hashes(dref)
    out["files"] = sorted(hr)
    out["identical"] = hp == hr
    if hp != hr:
        out["differ"] = [k for k in sorted(set(hp) | set(hr)) if hp.get(k) != hr.get(k)]
    out["speedup_best"] = round(dt / min(times), 2)
shutil.rmtree(work, ignore_errors=True)
print(json.dumps(out))
