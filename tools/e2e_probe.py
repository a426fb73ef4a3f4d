"""Per-step breakdown of esa_build end to end (host buffers in, pinned tables out) on the C2 workload.
    python tools/e2e_probe.py [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from vstree_b200 import Esa, load_library  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
desc, n, sigma, flag, gen, kw = bench.WORKLOADS["c2"]
tis = bench.make_tis("c2")[0]
pinned = torch.from_numpy(tis).pin_memory().numpy()
k = int(load_library().esa_recommended_prefixlength(sigma, n))
with Esa(0) as esa:
    mode = os.environ.get("PROBE_PRE", "")
    if mode:
        # what bench.py does in front of its end-to-end leg: staged builds, optionally under the NVML clock sampler
        esa.upload(pinned, sigma)
        sampler = None
        if "nvml" in mode:
            sampler = bench.ClockSampler(0)
            sampler.start()
        for _ in range(8):
            esa.construct(k, bench.TABLES)
        if sampler is not None:
            print("sampler", sampler.stop(), flush=True)
        if "check" in mode:
            esa.check()
    for i in range(steps):
        t0 = time.perf_counter()
        tabs, st = esa.build(pinned, sigma, k, bench.TABLES, want_stats=True, copy=False)
        wall = (time.perf_counter() - t0) * 1e3
        esa.free(tabs)
        print("step %2d wall %7.2f h2d %6.2f device %6.2f (finalize %5.2f) d2h_tail %6.2f threads %d" % (
            i, wall, st["ms_h2d"], st["ms_total"], st["kernel"]["finalize"]["ms"], st["ms_d2h"],
            st["host_widen_threads"]), flush=True)
