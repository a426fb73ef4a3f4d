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
if [ -x "$OUT/mkvtree.x" ] && [ -x "$OUT/mkvtree.dbg.x" ] && [ "${1:-}" != "--force" ]; then
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
echo "built: $(ls "$OUT")"
