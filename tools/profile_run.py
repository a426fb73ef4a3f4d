"""Small driver for ncu: upload one workload, run `reps` device-resident builds.  Usage:
    python tools/profile_run.py [workload] [reps] [tables,comma,separated]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vstree_b200 import Esa, load_library  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tables = tuple(sys.argv[3].split(",")) if len(sys.argv) > 3 else bench.TABLES
desc, n, sigma, flag, gen, kw = bench.WORKLOADS[workload]
tis = bench.make_tis(workload)[0]
k = int(load_library().esa_recommended_prefixlength(sigma, n))
with Esa(0) as esa:
    esa.upload(tis, sigma)
    for i in range(reps):
        st = esa.construct(k, tables)
        print("rep", i, "ms_total %.3f" % st["ms_total"], "launches", st["launches"], "rounds", st["rounds"],
              " ".join("%s=%.3f" % (k, v["ms"]) for k, v in st["kernel"].items() if v["launches"]))
    if os.environ.get("RUN_CONFIG_NO_CHECK") is None:
        esa.check()
