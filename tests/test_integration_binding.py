"""The reference-side binding shown in INTEGRATION.md section B (integration/gpusuffixsort.c) compiles against the
REFERENCE's own headers -- Virtualtree, Input, PairUint, the ERROR macros -- and references exactly the C-ABI entry
points the library exports.  CPU only; needs the reference tree (skipped on the GPU box)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFSRC = "/root/reference/src"


def test_binding_in_integration_md_is_the_compiled_one():
    with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
        snippet = re.search(r"```c\n(#include \"vstree_b200.h\".*?)```", f.read(), re.S).group(1)
    with open(os.path.join(ROOT, "integration", "gpusuffixsort.c")) as f:
        assert snippet in f.read(), "INTEGRATION.md section B and integration/gpusuffixsort.c drifted apart"


def test_binding_compiles_against_the_reference_headers(tmp_path):
    if not os.path.isdir(os.path.join(REFSRC, "include")):
        pytest.skip("reference tree not mounted")
    obj = str(tmp_path / "gpusuffixsort.o")
    cmd = ["gcc", "-c", "-Wall", "-Werror=incompatible-pointer-types", "-Werror=int-conversion",
           "-Werror=implicit-function-declaration", "-Wno-unused-function", "-D_GNU_SOURCE",
           "-I" + os.path.join(REFSRC, "include"), "-I" + os.path.join(REFSRC, "Mkvtree"),
           "-I" + os.path.join(ROOT, "oracle", "stub"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "integration", "gpusuffixsort.c"), "-o", obj]
    pr = subprocess.run(cmd, capture_output=True, text=True)
    assert pr.returncode == 0, pr.stderr[-3000:]
    syms = subprocess.run(["nm", obj], capture_output=True, text=True).stdout
    undefined = {l.split()[-1] for l in syms.splitlines() if " U " in l}
    assert {"esa_init", "esa_build"} <= undefined
    import ctypes
    import vstree_b200
    lib = ctypes.CDLL(vstree_b200.library_path())
    for name in undefined:
        if name.startswith("esa_"):
            assert hasattr(lib, name), f"the binding calls {name}, which the library does not export"


def test_sorter_replacement_compiles_against_the_reference_headers(tmp_path):
    """integration/besegpu.c: besesuffixsort / besesuffixsortspecial with the reference's own signatures
    (src/Mkvtree/bese.c:855), on top of esa_build -- what oracle/build_ref.sh links into oracle/_ref/gpu/*."""
    if not os.path.isdir(os.path.join(REFSRC, "include")):
        pytest.skip("reference tree not mounted")
    obj = str(tmp_path / "besegpu.o")
    cmd = ["gcc", "-c", "-Wall", "-Werror=incompatible-pointer-types", "-Werror=int-conversion",
           "-Werror=implicit-function-declaration", "-D_GNU_SOURCE", "-DWITHLCP",
           "-I" + os.path.join(REFSRC, "include"), "-I" + os.path.join(REFSRC, "Mkvtree"),
           "-I" + os.path.join(ROOT, "oracle", "stub"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "integration", "besegpu.c"), "-o", obj]
    pr = subprocess.run(cmd, capture_output=True, text=True)
    assert pr.returncode == 0, pr.stderr[-3000:]
    syms = subprocess.run(["nm", obj], capture_output=True, text=True).stdout
    defined = {l.split()[-1] for l in syms.splitlines() if " T " in l}
    undefined = {l.split()[-1] for l in syms.splitlines() if " U " in l}
    assert {"besesuffixsort", "besesuffixsortspecial"} <= defined
    assert {"esa_init", "esa_build", "esa_free"} <= undefined
    # the prototypes are the reference's: compile a caller that includes the reference's own prototype of the two
    proto = tmp_path / "proto.c"
    proto.write_text('#include "types.h"\n#include "arraydef.h"\n#include "virtualdef.h"\n#include "mkvaux.h"\n'
                     '#include "bese.pr"\n#include "besespecial.pr"\n')
    if os.path.exists(os.path.join(REFSRC, "Mkvtree", "bese.pr")):
        pr = subprocess.run(["gcc", "-c", "-DWITHLCP", "-I" + os.path.join(REFSRC, "include"),
                             "-I" + os.path.join(REFSRC, "Mkvtree"), "-I" + os.path.join(ROOT, "oracle", "stub"),
                             "-include", str(proto), "-I" + os.path.join(ROOT, "include"),
                             os.path.join(ROOT, "integration", "besegpu.c"), "-o", str(tmp_path / "b2.o")],
                            capture_output=True, text=True)
        assert pr.returncode == 0, pr.stderr[-3000:]


def test_child_table_replacement_compiles_against_the_reference_headers(tmp_path):
    """integration/mkcldgpu.c: mkcldtabnextlIndex / mkcldtabupdown / compresscldtab with the reference's signatures
    (src/kurtz/mkcld.c:41,110,225; prototypes in src/include/mkcld.pr), on top of esa_childtab."""
    if not os.path.isdir(os.path.join(REFSRC, "include")):
        pytest.skip("reference tree not mounted")
    obj = str(tmp_path / "mkcldgpu.o")
    base = ["gcc", "-c", "-Wall", "-Werror=incompatible-pointer-types", "-Werror=int-conversion",
            "-Werror=implicit-function-declaration", "-D_GNU_SOURCE", "-I" + os.path.join(REFSRC, "include"),
            "-I" + os.path.join(REFSRC, "kurtz"), "-I" + os.path.join(ROOT, "oracle", "stub"),
            "-I" + os.path.join(ROOT, "include")]
    pr = subprocess.run(base + [os.path.join(ROOT, "integration", "mkcldgpu.c"), "-o", obj], capture_output=True, text=True)
    assert pr.returncode == 0, pr.stderr[-3000:]
    syms = subprocess.run(["nm", obj], capture_output=True, text=True).stdout
    defined = {l.split()[-1] for l in syms.splitlines() if " T " in l}
    undefined = {l.split()[-1] for l in syms.splitlines() if " U " in l}
    assert {"mkcldtabnextlIndex", "mkcldtabupdown", "compresscldtab"} <= defined
    assert {"esa_init", "esa_childtab"} <= undefined
    if os.path.exists(os.path.join(REFSRC, "include", "mkcld.pr")):  # the reference's own prototypes in front of the file
        proto = tmp_path / "proto.c"
        proto.write_text('#include "types.h"\n#include "arraydef.h"\n#include "virtualdef.h"\n#include "mkcld.pr"\n')
        pr = subprocess.run(base + ["-include", str(proto), os.path.join(ROOT, "integration", "mkcldgpu.c"), "-o",
                                    str(tmp_path / "m2.o")], capture_output=True, text=True)
        assert pr.returncode == 0, pr.stderr[-3000:]


def test_family_binaries_are_built_where_the_reference_is_present():
    if not os.path.isdir(os.path.join(REFSRC, "include")):
        pytest.skip("reference tree not mounted")
    for t in ("mkvtree.x", "mkvram.x", "mkrcidx.x", "mkdna6idx.x", "vmatch.x", "mkcld.x"):
        assert os.access(os.path.join(ROOT, "oracle", "_ref", "gpu", t), os.X_OK), t
    for t in ("mkvram.x", "mkrcidx.x", "mkdna6idx.x", "mkcld.x"):
        assert os.access(os.path.join(ROOT, "oracle", "_ref", t), os.X_OK), t
