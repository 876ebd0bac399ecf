#!/bin/bash
# build_variant.sh NAME "-DFLAG=.. ..." : libmgcn variant with extra nvcc defines -> variants/libmgcn_NAME.so (MGCN_LIB=... to use)
set -e
name=$1; shift
d=variants/obj_$name; mkdir -p $d
for f in meta_gcn_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --fmad=true -cudart static $@ -I include -I meta_gcn_b200/csrc -c $f -o $d/$(basename $f .cu).o &
done
wait
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -o variants/libmgcn_$name.so $d/*.o
rm -rf $d
