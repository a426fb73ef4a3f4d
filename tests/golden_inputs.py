"""Named, seeded inputs shared by the golden generator (tests/golden/make_golden.py), the oracle tests and the
GPU parity tests.  Each entry: name -> (tis uint8 array, alphabet size, alphabet flag)."""
import numpy as np

from vstree_b200 import synth

W, S = synth.WILDCARD, synth.SEPARATOR


def _tiny(codes):
    return np.array(codes, np.uint8)


CASES = {
    # name: (builder, sigma, flag, prefixlength or None for the reference's automatic choice)
    "kat_survey": (lambda: _tiny([0, 1, 2, 3, W, 0, 1, 2, S, 0, 1, 2, 3, S, W, W, 0, 1]), 4, "-dna", 1),
    "single_base": (lambda: _tiny([2]), 4, "-dna", 1),
    "two_equal": (lambda: _tiny([1, 1]), 4, "-dna", 1),
    "poly_a_300": (lambda: np.zeros(300, np.uint8), 4, "-dna", 2),
    "poly_a_5000": (lambda: np.zeros(5000, np.uint8), 4, "-dna", 3),
    "all_wildcards": (lambda: np.full(40, W, np.uint8), 4, "-dna", 1),
    "period3_x1000": (lambda: np.tile(_tiny([0, 1, 2]), 1000), 4, "-dna", 4),
    "acgt_with_tail_n": (lambda: _tiny([0, 1, 2, 3] * 10 + [W]), 4, "-dna", 1),
    "lead_n_records": (lambda: _tiny([W, W, 0, 1, S, 3, 3, 3, S, W, 2, W, 2, S, 0]), 4, "-dna", 1),
    "dna_20k": (lambda: synth.random_dna(20000, seed=1), 4, "-dna", None),
    "dna_specials_50k": (lambda: synth.dna_with_specials(50000, seed=2, nfrac=0.01, mean_run=5, records=40), 4, "-dna",
                         None),
    "dna_specials_dense_8k": (lambda: synth.dna_with_specials(8000, seed=3, nfrac=0.2, mean_run=2, records=300), 4,
                              "-dna", 5),
    "dna_repeat_200k": (lambda: synth.repeat_rich_dna(200000, seed=7, copies=20, divergence=0.005, tandem_frac=0.02,
                                                      tandem_len=1500), 4, "-dna", None),
    "dna_twocopies_60k": (lambda: np.tile(synth.random_dna(30000, seed=9), 2), 4, "-dna", None),
    "dna_300k_pl8": (lambda: synth.dna_with_specials(300000, seed=4, nfrac=0.002, mean_run=30, records=12), 4, "-dna",
                     8),
    "protein_30k": (lambda: synth.random_protein(30000, seed=11, records=60, xfrac=0.002), 20, "-protein", None),
    "protein_lowcomplex_20k": (lambda: (synth.random_protein(20000, seed=12, records=10, xfrac=0.0) % 3).astype(
        np.uint8), 20, "-protein", 2),
    "protein_200k": (lambda: synth.random_protein(200000, seed=13, records=400, xfrac=0.0005), 20, "-protein", None),
    "human_like_400k": (lambda: synth.human_like_dna(400000, seed=2026, records=5), 4, "-dna", None),
}


def make_input(name):
    build, sigma, flag, pl = CASES[name]
    tis = build()
    return np.ascontiguousarray(tis, np.uint8), sigma
