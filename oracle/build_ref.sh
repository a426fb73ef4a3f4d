#!/bin/bash
# Build the UNMODIFIED reference mkvtree (and its -DDEBUG self-checking twin) from the sources
# where they lie under /root/reference/src into oracle/_ref/ (git-ignored, travels to the GPU box).
# Test infrastructure only: the product never links or runs anything from here.
# Recipe follows SURVEY.md 8c (reference: src/Mkvtree/Makefile:3,15-22,175-186; src/Makedef-linux-gcc-64:24).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
S="${VSTREE_REF_SRC:-/root/reference/src}"
OUT="$HERE/_ref"
if [ ! -d "$S/Mkvtree" ]; then
  echo "build_ref.sh: reference sources not found at $S (fine on the GPU box: prebuilt oracle/_ref is used)" >&2
  exit 0
fi
# The UNMODIFIED vmatch, the consumer of the index (downstream acceptance test: tests/test_zz_vmatch_acceptance.py).
# Recipe: src/Vmatch/Makefile:38-84 (LIBVMATCHOBJ + vmatch.mn.c), src/Vmengine/Makefile:40-56 and Vmatdef2.mf
# (vmatfind.c / vmatcount.c once per alphabet size, plus the dynamic and the streaming variant), libmkvtree.
build_vmatch() {
  local O="$OUT/objvm"
  mkdir -p "$O"
  local CV="$CF -I$S/Vmengine -I$S/Vmatch"
  local LIST="vmatch parsevm optstring procargs runself runquery mlendistrib mapdemand procmatch queryindex verifysub lrseq echomatch substpos markmat nomatch showmasked vmcluster assigndig explain detmatch opensel vmdbfunc parsedbms chncallparse chainvm chainof initpost matchlenparm xmlfunc vplugin-open vmotif-start cpridx-start parsepp parsequery parsedbcl keepflags parsemcl cluedist matchclust clpos allmclust threadchain procfinal mokay vmatch.mn"
  ( for f in "$S"/kurtz-basic/*.c; do echo "gcc $CF -c $f -o $O/kb_$(basename "$f" .c).o"; done
    for f in "$S"/kurtz/*.c;       do echo "gcc $CF -c $f -o $O/k_$(basename "$f" .c).o"; done
    for f in mkvtree mkvprocess bese ppsort mkvinput cprsuf remainsort; do
      echo "gcc $CF -DWITHLCP -c $S/Mkvtree/$f.c -o $O/m_$f.o"; done
    echo "gcc $CF -DWITHLCP -DSPECIAL -c $S/Mkvtree/bese.c -o $O/m_besespecial.o"
    for f in "$S"/Vmengine/*.c; do b=$(basename "$f" .c)
      case $b in vmatfind|vmatcount|vmatgen) ;; *) echo "gcc $CV -c $f -o $O/e_$b.o";; esac; done
    for a in 4 12 21 31 42 53 66; do
      echo "gcc $CV -DALPHABETSIZE=$a -c $S/Vmengine/vmatfind.c -o $O/e_vmatfind$a.o"
      echo "gcc $CV -DALPHABETSIZE=$a -c $S/Vmengine/vmatcount.c -o $O/e_vmatcount$a.o"; done
    echo "gcc $CV -DDYNAMICALPHABET -c $S/Vmengine/vmatfind.c -o $O/e_vmatfind-dyn.o"
    echo "gcc $CV -DDYNAMICALPHABET -DESASTREAMACCESS -c $S/Vmengine/vmatfind.c -o $O/e_vmatfind-strm.o"
    for b in $LIST; do echo "gcc $CV -c $S/Vmatch/$b.c -o $O/v_$b.o"; done
    echo "gcc $CF -c $HERE/stub/bzstub.c -o $O/bzstub.o"
  ) | xargs -P "$J" -I{} sh -c "{}"
  ar rcs "$O/libkb.a" "$O"/kb_*.o
  ar rcs "$O/libk.a" "$O"/k_*.o
  ar rcs "$O/libm.a" "$O"/m_*.o
  ar rcs "$O/libe.a" "$O"/e_*.o
  mv "$O/v_vmatch.mn.o" "$O/main_vmatch.o"
  ar rcs "$O/libv.a" "$O"/v_*.o
  local L="$O/libv.a $O/libe.a $O/libm.a $O/libk.a $O/libkb.a"
  gcc -o "$OUT/vmatch.x" "$O/main_vmatch.o" $L $L "$O/bzstub.o" -lz -lm -ldl -lpthread
  rm -rf "$O"
}
# The reference's OTHER callers of the sorter (mkrcidx, mkdna6idx, mkvram: src/Mkvtree/Makefile:144-153), unmodified, and
# the "gpu family" next to them (oracle/_ref/gpu/): the very same reference objects, except that bese.o, besespecial.o
# and ppsort.o -- the suffix sorter -- are replaced by integration/besegpu.c, which defines besesuffixsort[special] on
# top of libvstree_b200 (esa_build).  mkvtree, mkvram, mkrcidx, mkdna6idx (callmkvtree[generic]) and vmatch's on-the-fly
# query index (completevirtualtree) then build on the GPU: SURVEY.md 8f row 3, tests/test_gpu_refseam.py.
build_families() {
  local O="$OUT/objfam" ROOT
  ROOT="$(cd "$HERE/.." && pwd)"
  [ -f "$ROOT/vstree_b200/lib/libvstree_b200.so" ] || { echo "build_ref.sh: libvstree_b200.so not built yet: gpu family skipped" >&2; return 0; }
  mkdir -p "$O" "$OUT/gpu"
  local CV="$CF -I$S/Vmengine -I$S/Vmatch"
  local LIST="vmatch parsevm optstring procargs runself runquery mlendistrib mapdemand procmatch queryindex verifysub lrseq echomatch substpos markmat nomatch showmasked vmcluster assigndig explain detmatch opensel vmdbfunc parsedbms chncallparse chainvm chainof initpost matchlenparm xmlfunc vplugin-open vmotif-start cpridx-start parsepp parsequery parsedbcl keepflags parsemcl cluedist matchclust clpos allmclust threadchain procfinal mokay vmatch.mn"
  ( for f in "$S"/kurtz-basic/*.c; do echo "gcc $CF -c $f -o $O/kb_$(basename "$f" .c).o"; done
    for f in "$S"/kurtz/*.c;       do echo "gcc $CF -c $f -o $O/k_$(basename "$f" .c).o"; done
    for f in mkvtree mkvprocess mkvinput cprsuf remainsort; do
      echo "gcc $CF -DWITHLCP -c $S/Mkvtree/$f.c -o $O/m_$f.o"; done
    for f in bese ppsort; do echo "gcc $CF -DWITHLCP -c $S/Mkvtree/$f.c -o $O/s_$f.o"; done
    echo "gcc $CF -DWITHLCP -DSPECIAL -c $S/Mkvtree/bese.c -o $O/s_besespecial.o"
    for f in mkvfile mkvram mkrcidx mkdna6idx; do echo "gcc $CF -DWITHLCP -c $S/Mkvtree/$f.c -o $O/main_$f.o"; done
    echo "gcc $CF -DWITHLCP -I$S/Mkvtree -I$ROOT/include -Wall -c $ROOT/integration/besegpu.c -o $O/g_besegpu.o"
    for f in "$S"/Vmengine/*.c; do b=$(basename "$f" .c)
      case $b in vmatfind|vmatcount|vmatgen) ;; *) echo "gcc $CV -c $f -o $O/e_$b.o";; esac; done
    for a in 4 12 21 31 42 53 66; do
      echo "gcc $CV -DALPHABETSIZE=$a -c $S/Vmengine/vmatfind.c -o $O/e_vmatfind$a.o"
      echo "gcc $CV -DALPHABETSIZE=$a -c $S/Vmengine/vmatcount.c -o $O/e_vmatcount$a.o"; done
    echo "gcc $CV -DDYNAMICALPHABET -c $S/Vmengine/vmatfind.c -o $O/e_vmatfind-dyn.o"
    echo "gcc $CV -DDYNAMICALPHABET -DESASTREAMACCESS -c $S/Vmengine/vmatfind.c -o $O/e_vmatfind-strm.o"
    for b in $LIST; do echo "gcc $CV -c $S/Vmatch/$b.c -o $O/v_$b.o"; done
    echo "gcc $CF -c $HERE/stub/bzstub.c -o $O/bzstub.o"
  ) | xargs -P "$J" -I{} sh -c "{}"
  ar rcs "$O/libkb.a" "$O"/kb_*.o
  ar rcs "$O/libk.a" "$O"/k_*.o
  ar rcs "$O/libm.a" "$O"/m_*.o          # libmkvtree WITHOUT a sorter
  ar rcs "$O/libsort_ref.a" "$O"/s_*.o   # the reference's sorter
  ar rcs "$O/libe.a" "$O"/e_*.o
  mv "$O/v_vmatch.mn.o" "$O/main_vmatch.o"
  ar rcs "$O/libv.a" "$O"/v_*.o
  local KL="$O/libk.a $O/libkb.a $O/libk.a $O/libkb.a $O/bzstub.o -lz -lm -ldl -lpthread"
  local GPU="$O/g_besegpu.o -L$ROOT/vstree_b200/lib -lvstree_b200"
  local RPATH='-Wl,-rpath,$ORIGIN/../../../vstree_b200/lib'   # oracle/_ref/gpu/ -> vstree_b200/lib, wherever the tree lies
  for t in mkvram mkrcidx mkdna6idx; do
    gcc -o "$OUT/$t.x" "$O/main_$t.o" "$O/libm.a" "$O/libsort_ref.a" "$O/libm.a" $KL
    gcc -o "$OUT/gpu/$t.x" "$O/main_$t.o" "$O/libm.a" $GPU "$RPATH" "$O/libm.a" $KL
  done
  gcc -o "$OUT/gpu/mkvtree.x" "$O/main_mkvfile.o" "$O/libm.a" $GPU "$RPATH" "$O/libm.a" $KL
  local L="$O/libv.a $O/libe.a $O/libm.a"
  gcc -o "$OUT/gpu/vmatch.x" "$O/main_vmatch.o" $L $L $GPU "$RPATH" "$O/libm.a" $KL
  rm -rf "$O"
}
# The reference's child-table tool (src/Mkvtree/Makefile:132-134: mkcld.mn.o + libkurtz + libkurtz-basic), unmodified:
# reads .lcp/.llv of an index, writes .cld and .cld1.  Golden maker and live comparison of SURVEY.md 8f row 4
# (tests/golden/make_cld_golden.py, tests/test_gpu_childtab.py).
build_tools() {
  local O="$OUT/objtools"
  mkdir -p "$O"
  ( for f in "$S"/kurtz-basic/*.c; do echo "gcc $CF -c $f -o $O/kb_$(basename "$f" .c).o"; done
    for f in "$S"/kurtz/*.c;       do echo "gcc $CF -c $f -o $O/k_$(basename "$f" .c).o"; done
    echo "gcc $CF -c $S/Mkvtree/mkcld.mn.c -o $O/main_mkcld.o"
    echo "gcc $CF -c $HERE/stub/bzstub.c -o $O/bzstub.o"
  ) | xargs -P "$J" -I{} sh -c "{}"
  ar rcs "$O/libkb.a" "$O"/kb_*.o
  ar rcs "$O/libk.a" "$O"/k_*.o
  gcc -o "$OUT/mkcld.x" "$O/main_mkcld.o" "$O/libk.a" "$O/libkb.a" "$O/libk.a" "$O/libkb.a" "$O/bzstub.o" -lz -lm -ldl -lpthread
  # the same main with the three child-table functions of libkurtz (mkcld.o) replaced by integration/mkcldgpu.c
  local ROOT
  ROOT="$(cd "$HERE/.." && pwd)"
  if [ -f "$ROOT/vstree_b200/lib/libvstree_b200.so" ]; then
    mkdir -p "$OUT/gpu"
    gcc $CF -I"$ROOT/include" -Wall -c "$ROOT/integration/mkcldgpu.c" -o "$O/g_mkcldgpu.o"
    rm -f "$O/k_mkcld.o" "$O/libk_nocld.a"
    ar rcs "$O/libk_nocld.a" "$O"/k_*.o
    gcc -o "$OUT/gpu/mkcld.x" "$O/main_mkcld.o" "$O/g_mkcldgpu.o" "$O/libk_nocld.a" "$O/libkb.a" "$O/libk_nocld.a" "$O/libkb.a" \
        "$O/bzstub.o" -L"$ROOT/vstree_b200/lib" -lvstree_b200 '-Wl,-rpath,$ORIGIN/../../../vstree_b200/lib' -lz -lm -ldl -lpthread
  fi
  rm -rf "$O"
}
CF="-D_LARGEFILE_SOURCE -D_FILE_OFFSET_BITS=64 -O3 -m64 -fno-stack-protector -U_FORTIFY_SOURCE -D_GNU_SOURCE -DWITHSYSCONF -I$HERE/stub -I$S/include -w"
J=$(nproc)
ROOT_="$(cd "$HERE/.." && pwd)"
families_stale() {  # rebuild when a binary is missing or integration/besegpu.c is newer
  [ ! -x "$OUT/gpu/vmatch.x" ] || [ ! -x "$OUT/gpu/mkvtree.x" ] || [ ! -x "$OUT/mkrcidx.x" ] || [ "$ROOT_/integration/besegpu.c" -nt "$OUT/gpu/mkvtree.x" ]
}
if [ -x "$OUT/mkvtree.x" ] && [ -x "$OUT/mkvtree.dbg.x" ] && [ "${1:-}" != "--force" ]; then
  if [ ! -x "$OUT/vmatch.x" ]; then
    mkdir -p "$OUT"
    build_vmatch || echo "build_ref.sh: vmatch.x could not be built (the acceptance test will be skipped)" >&2
  fi
  if families_stale; then
    build_families || echo "build_ref.sh: the gpu family could not be built (tests/test_gpu_refseam.py will be skipped)" >&2
  fi
  if [ ! -x "$OUT/mkcld.x" ] || [ ! -x "$OUT/gpu/mkcld.x" ] || [ "$ROOT_/integration/mkcldgpu.c" -nt "$OUT/gpu/mkcld.x" ]; then
    build_tools || echo "build_ref.sh: mkcld.x could not be built (the live child-table comparison will be skipped)" >&2
  fi
  exit 0
fi
mkdir -p "$OUT/obj" "$OUT/objdbg"
CF="-D_LARGEFILE_SOURCE -D_FILE_OFFSET_BITS=64 -O3 -m64 -fno-stack-protector -U_FORTIFY_SOURCE -D_GNU_SOURCE -DWITHSYSCONF -I$HERE/stub -I$S/include -w"
J=$(nproc)
build_variant() {   # $1 = objdir, $2 = extra flags, $3 = suffix of binaries
  local O="$1" X="$2" SUF="$3"
  ( for f in "$S"/kurtz-basic/*.c; do echo "gcc $CF $X -c $f -o $O/kb_$(basename "$f" .c).o"; done
    for f in "$S"/kurtz/*.c;       do echo "gcc $CF $X -c $f -o $O/k_$(basename "$f" .c).o"; done
    for f in mkvtree mkvprocess bese ppsort mkvinput cprsuf remainsort mkvfile mkvcmp; do
      echo "gcc $CF $X -DWITHLCP -c $S/Mkvtree/$f.c -o $O/m_$f.o"; done
    echo "gcc $CF $X -DWITHLCP -DSPECIAL -c $S/Mkvtree/bese.c -o $O/m_besespecial.o"
    echo "gcc $CF -c $HERE/stub/bzstub.c -o $O/bzstub.o"
  ) | xargs -P "$J" -I{} sh -c "{}"
  rm -f "$O/libkb.a" "$O/libk.a"
  ar rcs "$O/libkb.a" "$O"/kb_*.o
  ar rcs "$O/libk.a" "$O"/k_*.o
  local MK="$O/m_mkvtree.o $O/m_mkvprocess.o $O/m_bese.o $O/m_besespecial.o $O/m_ppsort.o $O/m_mkvinput.o $O/m_cprsuf.o $O/m_remainsort.o"
  local L="$O/libk.a $O/libkb.a $O/libk.a $O/libkb.a $O/bzstub.o -lz -lm -ldl -lpthread"
  gcc -o "$OUT/mkvtree$SUF" $MK "$O/m_mkvfile.o" $L
  if [ -n "$X" ]; then gcc -o "$OUT/mkvcmp$SUF" $MK "$O/m_mkvcmp.o" $L; fi   # compareVirtualtree exists only under -DDEBUG
}
build_variant "$OUT/obj" "" ".x"
build_variant "$OUT/objdbg" "-DDEBUG" ".dbg.x"
rm -rf "$OUT/obj" "$OUT/objdbg"
build_vmatch || echo "build_ref.sh: vmatch.x could not be built (the acceptance test will be skipped)" >&2
build_families || echo "build_ref.sh: the gpu family could not be built (tests/test_gpu_refseam.py will be skipped)" >&2
build_tools || echo "build_ref.sh: mkcld.x could not be built (the live child-table comparison will be skipped)" >&2
echo "built: $(ls "$OUT")"
