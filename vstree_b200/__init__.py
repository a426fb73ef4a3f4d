"""vstree_b200 -- B200-native enhanced-suffix-array construction (the mkvtree hot path of genometools/vstree).

The product is the shared library vstree_b200/lib/libvstree_b200.so (hand-written sm_100a CUDA kernels behind the
C ABI of include/vstree_b200.h) and the `mkvtree` host driver in vstree_b200/bin/.  This Python package only holds
the ctypes binding that tests/ and bench.py use to call the C ABI, and the synthetic workload generators.
"""
from .capi import (DEMAND_BITS, Esa, EsaError, KERNEL_FAMILIES, library_path, load_library)  # noqa: F401

__all__ = ["Esa", "EsaError", "DEMAND_BITS", "KERNEL_FAMILIES", "library_path", "load_library"]
