"""Builds a fixed set of inputs with whatever sorter variant the ESA_* environment variables select and compares every
table with the oracle.  Run by tests/test_gpu_variants.py in a subprocess per variant (the library reads the
variables once per process)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_util as ou  # noqa: E402
from golden_inputs import make_input  # noqa: E402
from vstree_b200 import Esa, synth  # noqa: E402

TABLES = ("suf", "lcp", "llv", "bwt", "bck", "sti1")


def identical_records(sigma, reclen, copies, seed, wildcard_at=()):
    """Many copies of one record: suffixes at the same offset of different copies agree up to the separator and are
    ordered by the POSITION of that separator (bese.c:26-55) -- equal flagged text keys."""
    rng = np.random.default_rng(seed)
    rec = rng.integers(0, sigma, size=reclen, dtype=np.uint8)
    for w in wildcard_at:
        rec[w] = 254
    return np.concatenate([np.append(rec, np.uint8(255)) for _ in range(copies)])[:-1]


def inputs():
    yield "identical_dna_40x60", identical_records(4, 40, 60, 1), 4, 3
    yield "identical_dna_wild_70x33", identical_records(4, 70, 33, 2, wildcard_at=(20, 21, 50)), 4, 4
    yield "identical_dna_17x500", identical_records(4, 17, 500, 3), 4, 2
    yield "identical_protein_25x80", identical_records(20, 25, 80, 4, wildcard_at=(9,)), 20, 2
    yield "identical_dna_long_3000x7", identical_records(4, 3000, 7, 5), 4, 5
    for name in ("dna_specials_50k", "dna_specials_dense_8k", "dna_repeat_200k", "protein_30k", "human_like_400k",
                 "poly_a_5000", "period3_x1000", "lead_n_records"):
        tis, sigma = make_input(name)
        yield name, tis, sigma, None
    yield "dna_1m_specials", synth.dna_with_specials(1_000_000, seed=61, nfrac=0.004, mean_run=9, records=200), 4, 8
    yield "dna_5m_repeat", synth.repeat_rich_dna(5_000_000, seed=62, copies=25, divergence=0.01, tandem_frac=0.02), 4, 9


def main():
    lib = ou.oracle_lib()
    with Esa(0) as esa:
        for name, tis, sigma, k in inputs():
            if k is None:
                k = max(1, int(lib.oracle_maximalprefixlength(sigma, len(tis))) - 1)
            got, st = esa.build(tis, sigma, k, tables=("suf", "lcp", "bwt", "bck", "sti1"))
            want = ou.oracle_build(tis, sigma, k)
            for t in TABLES:
                assert np.array_equal(got[t], np.asarray(want[t])), f"{name}: {t} differs from the oracle"
            for sname in ("longest", "maxbranchdepth", "largelcpvalues"):
                assert got[sname] == want[sname], f"{name}: {sname}"
            esa.check()
    ranks = int(os.environ.get("VARIANT_RANKS", "0"))
    if ranks > 1:
        # the same variant on a sharded context (ranks share device 0 through the in-process transport)
        from vstree_b200.capi import EsaGroup
        with EsaGroup([0] * ranks) as grp:
            for name, tis, sigma, k in inputs():
                if k is None:
                    k = max(1, int(lib.oracle_maximalprefixlength(sigma, len(tis))) - 1)
                got, st = grp.build(tis, sigma, k, tables=("suf", "lcp", "bwt", "bck"))
                want = ou.oracle_build(tis, sigma, k)
                for t in ("suf", "lcp", "llv", "bwt", "bck"):
                    assert np.array_equal(got[t], np.asarray(want[t])), f"{name} on {ranks} ranks: {t} differs from the oracle"
                for sname in ("longest", "maxbranchdepth", "largelcpvalues"):
                    assert got[sname] == want[sname], f"{name} on {ranks} ranks: {sname}"
    print("VARIANT-OK")


if __name__ == "__main__":
    main()
