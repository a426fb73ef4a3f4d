"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/vstree_b200.h declares,
fails loudly without a GPU, and its host scalar math equals the oracle's (which restates the reference)."""
import ctypes
import os
import re

import pytest

import oracle_util as ou
import vstree_b200
from vstree_b200 import Esa, EsaError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(vstree_b200.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    return vstree_b200.load_library()


def declared_functions():
    with open(os.path.join(ROOT, "include", "vstree_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(esa_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions()
    assert {"esa_init", "esa_build", "esa_free", "esa_upload", "esa_construct", "esa_download", "esa_check",
            "esa_shutdown", "esa_numofcodes", "esa_recommended_prefixlength", "esa_maximal_prefixlength",
            "esa_version"} <= set(names)
    raw = ctypes.CDLL(vstree_b200.library_path())
    for n in names:
        assert hasattr(raw, n), f"{n} is declared in include/vstree_b200.h but not exported"


def test_version(lib):
    assert b"sm_100a" in lib.esa_version()


def test_scalar_math_matches_oracle(lib):
    o = ou.oracle_lib()
    for sigma in (2, 4, 5, 12, 20, 31):
        for n in (1, 10, 64, 1000, 772376, 20_000_000, 100_000_000, 250_000_000, 500_000_000, 3_000_000_000):
            assert lib.esa_recommended_prefixlength(sigma, n) == o.oracle_recommendedprefixlength(sigma, n)
            assert lib.esa_maximal_prefixlength(sigma, n) == o.oracle_maximalprefixlength(sigma, n)
        for k in range(0, 8):
            assert lib.esa_numofcodes(sigma, k) == o.oracle_numofcodes(sigma, k) == sigma ** k
    # BASELINE.json config sizes: k = 11, 11, 5, 13
    assert lib.esa_recommended_prefixlength(4, 100_000_000) == 11
    assert lib.esa_recommended_prefixlength(4, 250_000_000) == 11
    assert lib.esa_recommended_prefixlength(20, 500_000_000) == 5
    assert lib.esa_recommended_prefixlength(4, 3_000_000_000) == 13


def test_no_gpu_means_loud_failure(lib):
    """No CPU fallback: without a device esa_init must fail with ESA_ERR_NODEVICE (-2)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(EsaError) as ei:
        Esa(0)
    assert ei.value.code == -2 and "no CPU fallback" in ei.value.message


def test_product_does_not_reference_the_oracle():
    """Nothing under vstree_b200/ or include/ may import, link or execute oracle/."""
    for base in ("vstree_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp")):
                    with open(os.path.join(dirpath, fn), errors="replace") as f:
                        text = f.read()
                    assert "oracle_util" not in text and "esa_oracle" not in text and "oracle/" not in text, fn
    with open(os.path.join(ROOT, "Makefile")) as f:
        assert "oracle" not in f.read()
