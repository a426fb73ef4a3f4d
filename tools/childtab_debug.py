"""Where a child table differs from the oracle's (debugging aid): builds C2, runs esa_childtab on the resident lcp table
with and without the bounded scan, prints the first differing records with the lcp values around them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_util as ou  # noqa: E402
from vstree_b200 import Esa, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
tis = synth.random_dna(n, seed=42)
with Esa(0) as esa:
    tabs, st = esa.build(tis, 4, 11, tables=("suf", "lcp"))
    lcp, llv = tabs["lcp"], tabs["llv"]
    want, want1 = ou.oracle_childtab(lcp, llv)
    L = lcp.astype(np.int64)
    for mode in ("default", "no_near", "host_tables"):
        if mode == "no_near":
            os.environ["ESA_CHILDTAB_NO_NEAR"] = "1"
        else:
            os.environ.pop("ESA_CHILDTAB_NO_NEAR", None)
        got, got1 = esa.childtab(lcp, llv) if mode == "host_tables" else esa.childtab(n=n)
        bad = np.flatnonzero(got != want)
        print(mode, "differing bytes", len(bad), "cld1", int((got1 != want1).sum()), flush=True)
        for b in bad[:6]:
            i = int(b // 3)
            print("  entry", i, "field", int(b % 3), "got", got[3 * i:3 * i + 3], "want", want[3 * i:3 * i + 3],
                  "block offset", i % 256, "L", L[max(0, i - 6):i + 12].tolist())
        if len(bad):
            e = np.unique(bad // 3)
            print("  entries", len(e), "offsets in block histogram (16 bins)", np.histogram(e % 256, bins=16)[0].tolist(),
                  "first", e[:5].tolist(), "last", e[-5:].tolist())
