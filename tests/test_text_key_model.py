"""The order argument behind the text round (csrc/suffix_sort.cu text_key_kernel), checked by
brute force on a Python restatement of the key: among suffixes that agree on their first h symbols, sorting by the text
key of the next L symbols (stable, equal flagged keys kept in text order) gives the reference's order
(src/Mkvtree/bese.c:26-55: letters by code, a special at position p compares as p + 256), and equal unflagged keys
agree on h + L letters.  CPU only; the CUDA kernel itself is covered by the GPU parity tests."""
import numpy as np
import pytest


def text_key(tis, q, L, sigma, flag_bits):
    """(code << flag_bits) | flag for the window of L symbols at q; positions at or behind len(tis) are special."""
    code, flag = 0, 0
    n = len(tis)
    for i in range(L):
        a = tis[q + i] if q + i < n else 255
        if a >= 254:
            for _ in range(i, L):
                code = code * sigma + (sigma - 1)
            flag = L - i
            break
        code = code * sigma + int(a)
    return (code << flag_bits) | flag


def suffix_sort_key(tis, s):
    n = len(tis)
    return [int(tis[p]) if tis[p] < 254 else 256 + p for p in range(s, n)] + [256 + n]


@pytest.mark.parametrize("seed", range(40))
def test_text_key_orders_like_the_reference(seed):
    rng = np.random.default_rng(seed)
    sigma = int(rng.choice([2, 3, 4, 20]))
    n = int(rng.integers(30, 400))
    tis = rng.integers(0, sigma, size=n, dtype=np.uint8)
    sp = rng.random(n) < rng.choice([0.0, 0.05, 0.3])
    tis[sp] = rng.choice([254, 255], size=int(sp.sum()))
    if rng.random() < 0.7:                      # plant repeats so that groups are not trivial
        w = n // 3
        tis[n - w:] = tis[:w]
    h = int(rng.integers(1, 6))
    L = int(rng.integers(1, 9))
    flag_bits = L.bit_length()
    # groups: suffixes without a special among their first h symbols, by those h symbols
    groups = {}
    for s in range(n):
        if s + h <= n and all(tis[s + i] < 254 for i in range(h)):
            groups.setdefault(tuple(int(x) for x in tis[s:s + h]), []).append(s)   # text order inside a group
    for members in groups.values():
        keyed = sorted(members, key=lambda s: text_key(tis, s + h, L, sigma, flag_bits))   # stable
        want = sorted(members, key=lambda s: suffix_sort_key(tis, s))
        # suffixes with equal unflagged keys are not ordered yet (they stay one group), everything else must be final
        i = 0
        while i < len(keyed):
            k = text_key(tis, keyed[i] + h, L, sigma, flag_bits)
            j = i
            while j < len(keyed) and text_key(tis, keyed[j] + h, L, sigma, flag_bits) == k:
                j += 1
            block, flagged = keyed[i:j], (k & ((1 << flag_bits) - 1)) != 0
            assert sorted(block) == sorted(want[i:j]), "a key block is not a contiguous run of the true order"
            if flagged:
                assert block == want[i:j], "equal flagged keys must already stand in their final (text) order"
            elif len(block) > 1:
                first = [int(x) for x in tis[block[0]:block[0] + h + L]]
                assert len(first) == h + L and all(x < 254 for x in first)
                assert all([int(x) for x in tis[s:s + h + L]] == first for s in block), "group must agree on h + L"
            i = j
