#!/bin/bash
# Build a variant of libapnetg.so with extra nvcc flags into ab/<name>.so (A/B timing; select with AP_NETG_LIB).
#   bash tools/build_variant.sh name "-DAP_UMMA_DIAG"
set -e
NAME=$1; EXTRA=$2
mkdir -p ab/obj_$NAME
for f in netg conv_umma conv_halo conv_stem conv_out landmark conv_simt elementwise compose conditioning flownet; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $EXTRA \
    -c animateportrait_b200/csrc/$f.cu -o ab/obj_$NAME/$f.o &
done
wait
nvcc -shared -o ab/$NAME.so ab/obj_$NAME/*.o -gencode arch=compute_100a,code=sm_100a
ls -la ab/$NAME.so
