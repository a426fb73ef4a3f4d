"""Builds one BASELINE.json configuration at full size on one GPU (or sharded over the visible GPUs, one thread per
device), checks the tables with the device-side property checks and prints one JSON line with timings.
    python tools/run_config.py c3|c4|c5|c2 [ngpus] [n override] [devices: "same" = all ranks on device 0]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vstree_b200 import Esa, load_library  # noqa: E402
from vstree_b200.capi import EsaGroup  # noqa: E402

workload = sys.argv[1]
ngpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
desc, n, sigma, flag, gen, kw = bench.WORKLOADS[workload]
if len(sys.argv) > 3 and int(sys.argv[3]) > 0:
    bench.WORKLOADS[workload] = (desc, int(sys.argv[3]), sigma, flag, gen, kw)
same_device = len(sys.argv) > 4 and sys.argv[4] == "same"
t0 = time.time()
tis = np.ascontiguousarray(bench.make_tis(workload)[0])
n = len(tis)
t_gen = time.time() - t0
k = int(load_library().esa_recommended_prefixlength(sigma, n))
out = {"workload": desc, "n": n, "prefixlength": k, "ngpus": ngpus, "generate_s": round(t_gen, 1),
       "specials": int((tis >= 254).sum())}
N = n + 1
if ngpus == 1:
    with Esa(0) as esa:
        esa.upload(tis, sigma)
        for rep in range(2):
            st = esa.construct(k, bench.TABLES)
        if os.environ.get("RUN_CONFIG_NO_CHECK") is None:
            esa.check()
            out["check"] = "esa_check passed (permutation, order, exact lcp / llv, bwt on the device)"
        t1 = time.time()
        tabs = esa.download()
        out["download_s"] = round(time.time() - t1, 2)
        stats = [st]
else:
    with EsaGroup([0] * ngpus if same_device else list(range(ngpus))) as grp:
        grp.upload(tis, sigma)
        for rep in range(2):
            stats = grp.construct(k, bench.TABLES)
        t1 = time.time()
        tabs = grp.download()
        out["download_s"] = round(time.time() - t1, 2)
suf = tabs["suf"]
assert len(suf) == N and int(suf.sum(dtype=np.uint64)) % (1 << 56) == (N * (N - 1) // 2) % (1 << 56), "not a permutation"
assert suf[N - 1] == n and suf[tabs["longest"]] == 0
llv = tabs["llv"].reshape(-1, 2)
big = np.flatnonzero(tabs["lcp"] == 255)
assert np.array_equal(llv[:, 0], big.astype(np.uint64)) and (len(llv) == 0 or llv[:, 1].min() >= 255)
out["maxbranchdepth"], out["largelcpvalues"] = tabs["maxbranchdepth"], tabs["largelcpvalues"]
import hashlib
out["sha"] = {t: hashlib.sha256(np.ascontiguousarray(tabs[t]).tobytes()).hexdigest()[:16] for t in ("suf", "lcp", "llv", "bwt")}
out["checksums_suf_lcp_bwt_llv_nllv"] = (lambda c: ["%016x" % c[t] for t in ("suf", "lcp", "bwt", "llv")] + [
    tabs["largelcpvalues"]])(bench.table_checksum(0, suf, tabs["lcp"], tabs["bwt"], tabs["llv"]))
out["longest"] = tabs["longest"]
ms = max(s["ms_total"] for s in stats)
out["ms_build_device"] = round(ms, 2)
out["Msym_per_s"] = round(n / ms / 1e3, 1)
s0 = stats[0]
out["rounds"], out["lcp_path"] = s0["rounds"], s0["lcp_path"]
out["active_first"], out["active_sum"] = [s["active_first"] for s in stats], [s["active_sum"] for s in stats]
out["shard_count"] = [s["shard_count"] for s in stats]
out["device_bytes_peak"] = max(s["device_bytes_peak"] for s in stats)
out["stages_ms_rank0"] = {kk: round(v["ms"], 2) for kk, v in s0["kernel"].items() if v["launches"]}
print(json.dumps(out), flush=True)
