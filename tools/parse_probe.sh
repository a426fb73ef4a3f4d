#!/bin/bash
# Host ingest by thread count (no GPU work: -tis only): the "parse" phase of MKVTREE_TIMING on a 100 Mbp FASTA on tmpfs.
cd "$(dirname "$0")/.."
D=$(mktemp -d /dev/shm/pp.XXXX)
python - "$D" <<'P'
import sys
sys.path.insert(0, ".")
from vstree_b200 import synth
tis = synth.dna_with_specials(100_000_000, seed=43, nfrac=0.001, mean_run=50, records=1000)
open(sys.argv[1] + "/in.fna", "wb").write(synth.to_fasta(tis, "dna"))
P
for t in 1 2 4 8 default; do
  for r in 1 2 3; do
    if [ $t = default ]; then unset MKVTREE_PARSE_THREADS; else export MKVTREE_PARSE_THREADS=$t; fi
    ( cd $D && MKVTREE_TIMING=1 $OLDPWD/vstree_b200/bin/mkvtree -db in.fna -dna -tis -indexname x 2>&1 | grep MKVTREE_TIMING | sed "s/, wait.*//; s/^/threads $t: /" )
  done
done
nproc
rm -rf "$D"
