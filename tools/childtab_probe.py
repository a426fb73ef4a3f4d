"""Device times of the child-table kernels at BASELINE sizes (ESA_CHILDTAB_TIMING=1 makes the library print them):
    python tools/childtab_probe.py c2 c3 > gpurun_out/childtab_times.txt 2>&1
Builds the index (suf + lcp), then calls esa_childtab on the resident lcp table three times."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ESA_CHILDTAB_TIMING"] = "1"
from vstree_b200 import Esa, synth  # noqa: E402

CASES = {"c2": ("random_dna", 100_000_000, {"seed": 42}, 4, 11),
         "c3": ("repeat_rich_dna", 250_000_000, {"seed": 7}, 4, 11),
         "c4": ("random_protein", 500_000_000, {"seed": 11}, 20, 5)}
with Esa(0) as esa:
    for name in sys.argv[1:] or ["c2"]:
        gen, n, kw, sigma, k = CASES[name]
        tis = getattr(synth, gen)(n, **kw)
        tabs, st = esa.build(tis, sigma, k, tables=("suf", "lcp"), copy=False)
        print(name, "n", n, "build ms", round(st["ms_total"], 3), "largelcpvalues", tabs["largelcpvalues"], flush=True)
        esa.free(tabs)
        for _ in range(3):
            t0 = time.perf_counter()
            esa.childtab(n=n)
            print(name, "esa_childtab wall ms (incl. 4 B per entry to pageable host memory)",
                  round((time.perf_counter() - t0) * 1e3, 1), flush=True)
