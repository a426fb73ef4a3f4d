#!/bin/bash
# tools/build_variant.sh NAME "-DESA_SORT_THREADS=256 ..."  -> build/variants/NAME/libvstree_b200.so (tuning experiments)
set -e
cd "$(dirname "$0")/.."
NAME=$1; DEFS=$2
OUT=build/variants/$NAME
mkdir -p $OUT
for f in vstree_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $DEFS -c $f -o $OUT/$(basename $f .cu).o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libvstree_b200.so $OUT/*.o
echo built $OUT/libvstree_b200.so
