"""Repeats esa_childtab on C2's resident lcp table and counts records that differ from the oracle's (debugging aid)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_util as ou  # noqa: E402
from vstree_b200 import Esa, synth  # noqa: E402

n = 100_000_000
tis = synth.random_dna(n, seed=42)
with Esa(0) as esa:
    tabs, st = esa.build(tis, 4, 11, tables=("suf", "lcp"))
    lcp, llv = tabs["lcp"], tabs["llv"]
    want, want1 = ou.oracle_childtab(lcp, llv)
    for rep in range(4):
        got, got1 = esa.childtab(n=n)
        bad = np.flatnonzero(got != want)
        print(os.environ.get("VSTREE_B200_LIB", "default"), "rep", rep, "differing bytes", len(bad),
              [(int(b // 3), int(b % 3), int(got[b]), int(want[b])) for b in bad[:4]], flush=True)
