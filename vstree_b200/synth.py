"""Deterministic synthetic sequence workloads (SURVEY.md 8d / BASELINE.json configs) for tests and bench.py.

Every generator returns the *transformed* sequence `tis` (uint8 codes 0..sigma-1, 254 = wildcard, 255 = separator)
exactly as mkvtree's alphabet transformation would produce it (reference: src/kurtz-basic/alphabet.c:329-351 for
-dna, :384-400 for -protein; src/include/chardef.h:19-43), so that it can be fed straight into the C-ABI.
`to_fasta` turns such an array back into a multi-FASTA text for CLI-level tests.
Not part of the index-construction path itself.
"""
import numpy as np

WILDCARD = 254
SEPARATOR = 255
DNA_LETTERS = b"acgt"
PROTEIN_LETTERS = b"LVIFKREDAGSTNQYWPHMC"


def random_dna(n, seed=42):
    """C2: n i.i.d. uniform bases, one record."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 4, size=n, dtype=np.uint8)


def _place_runs(tis, rng, total, mean_run, value):
    """Overwrite about `total` positions with `value`, in runs of geometric length (mean mean_run)."""
    n = len(tis)
    if total <= 0 or n == 0:
        return
    nruns = max(1, int(total / mean_run))
    starts = rng.integers(0, n, size=nruns)
    lens = rng.geometric(1.0 / mean_run, size=nruns)
    for s, l in zip(starts.tolist(), lens.tolist()):
        tis[s:s + l] = value


def _place_separators(tis, rng, records):
    """Turn records-1 random positions into separators, never adjacent and never first/last (no empty record)."""
    n = len(tis)
    if records <= 1 or n < 3 * records:
        return
    pos = np.unique(rng.integers(1, n - 1, size=records - 1))
    keep = np.ones(len(pos), bool)
    keep[1:] = np.diff(pos) > 1
    tis[pos[keep]] = SEPARATOR


def dna_with_specials(n, seed=43, nfrac=0.001, mean_run=50, records=1000):
    """C2 twin: uniform DNA with 0.1 % N in geometric runs and ~1000 records."""
    rng = np.random.default_rng(seed)
    tis = rng.integers(0, 4, size=n, dtype=np.uint8)
    _place_runs(tis, rng, int(n * nfrac), mean_run, WILDCARD)
    _place_separators(tis, rng, records)
    return tis


def repeat_rich_dna(n, seed=7, copies=50, divergence=0.005, tandem_frac=0.01, max_period=50, tandem_len=2000):
    """C3: a uniform base of n/copies bases repeated `copies` times with i.i.d. substitutions, plus exact tandem
    blocks (period 2..max_period) covering about tandem_frac of the positions to force very deep LCPs."""
    rng = np.random.default_rng(seed)
    base_len = max(1, n // copies)
    base = rng.integers(0, 4, size=base_len, dtype=np.uint8)
    tis = np.tile(base, copies + 1)[:n].copy()
    nsub = int(n * divergence)
    if nsub:
        pos = rng.integers(0, n, size=nsub)
        tis[pos] = (tis[pos] + rng.integers(1, 4, size=nsub, dtype=np.uint8)) & 3
    nblocks = int(n * tandem_frac / tandem_len) if tandem_len else 0
    for _ in range(nblocks):
        period = int(rng.integers(2, max_period + 1))
        length = int(rng.integers(tandem_len // 2, tandem_len * 3 // 2 + 1))
        start = int(rng.integers(0, max(1, n - length)))
        unit = rng.integers(0, 4, size=period, dtype=np.uint8)
        reps = -(-length // period)
        tis[start:start + length] = np.tile(unit, reps)[:min(length, n - start)]
    return tis


def random_protein(n, seed=11, records=100000, xfrac=0.0005):
    """C4: i.i.d. uniform over the 20 amino acids, `records` records, a few X wildcards."""
    rng = np.random.default_rng(seed)
    tis = rng.integers(0, 20, size=n, dtype=np.uint8)
    nx = int(n * xfrac)
    if nx:
        tis[rng.integers(0, n, size=nx)] = WILDCARD
    _place_separators(tis, rng, records)
    return tis


def human_like_dna(n, seed=2026, records=24):
    """C5: 50 % unique uniform sequence, 45 % interspersed repeats (Alu-like 300 bp family at 10-15 % divergence,
    L1-like 6 kb family at 5-20 %, exact satellite arrays of a 171 bp unit), 5 % N in long runs, `records` records."""
    rng = np.random.default_rng(seed)
    tis = rng.integers(0, 4, size=n, dtype=np.uint8)
    alu = rng.integers(0, 4, size=300, dtype=np.uint8)
    l1 = rng.integers(0, 4, size=6000, dtype=np.uint8)
    sat = rng.integers(0, 4, size=171, dtype=np.uint8)

    def plant(family, total, dlo, dhi):
        flen = len(family)
        ncopies = int(total // flen)
        if ncopies == 0 or n <= flen:
            return
        starts = rng.integers(0, n - flen, size=ncopies)
        chunk = 1 << 16
        for c0 in range(0, ncopies, chunk):
            st = starts[c0:c0 + chunk]
            k = len(st)
            block = np.broadcast_to(family, (k, flen)).copy()
            div = rng.uniform(dlo, dhi, size=(k, 1))
            mut = rng.random((k, flen)) < div
            block[mut] = (block[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
            idx = st[:, None] + np.arange(flen)[None, :]
            tis[idx.ravel()] = block.ravel()

    plant(alu, 0.25 * n, 0.10, 0.15)
    plant(l1, 0.17 * n, 0.05, 0.20)
    # satellites: exact tandem arrays, 3 % of the sequence, arrays of 20..2000 units
    remaining = int(0.03 * n)
    while remaining > 171 * 20 and n > 171 * 2000:
        units = int(rng.integers(20, 2001))
        length = min(units * 171, remaining)
        start = int(rng.integers(0, n - length))
        tis[start:start + length] = np.tile(sat, -(-length // 171))[:length]
        remaining -= length
    _place_runs(tis, rng, int(0.05 * n), max(1000, n // 2000), WILDCARD)
    _place_separators(tis, rng, records)
    return tis


def to_fasta(tis, alphabet="dna", width=70, wildcard=None, name="seq"):
    """Multi-FASTA bytes whose mkvtree transformation gives back `tis` (one record per separator-delimited piece)."""
    letters = DNA_LETTERS if alphabet == "dna" else PROTEIN_LETTERS
    if wildcard is None:
        wildcard = b"n" if alphabet == "dna" else b"X"
    lut = np.zeros(256, np.uint8)
    lut[:len(letters)] = np.frombuffer(letters, np.uint8)
    lut[WILDCARD] = wildcard[0]
    tis = np.asarray(tis, np.uint8)
    seps = np.flatnonzero(tis == SEPARATOR)
    bounds = np.concatenate([[-1], seps, [len(tis)]])
    out = []
    for r in range(len(bounds) - 1):
        piece = lut[tis[bounds[r] + 1:bounds[r + 1]]].tobytes()
        out.append(b">%s%d\n" % (name.encode(), r + 1))
        for o in range(0, len(piece), width):
            out.append(piece[o:o + width] + b"\n")
    return b"".join(out)
