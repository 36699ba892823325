#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=1 ..." : compile the CTA-per-member kernels (tu_run, tu_init, tu_adaptive) with extra
# flags and link them with the other objects of build/ into tools/variants/NAME.so (select it with PNMOL_B200_LIB=...).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/var_$name tools/variants
for tu in tu_run tu_init tu_adaptive tu_small tu_large_run; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -c $@ \
    -o build/var_$name/$tu.o pnmol-experiments_b200/csrc/$tu.cu &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o tools/variants/$name.so \
  build/var_$name/tu_run.o build/var_$name/tu_init.o build/var_$name/tu_adaptive.o build/var_$name/tu_small.o build/var_$name/tu_large_run.o build/api.o \
  build/tu_large_init.o build/tu_large_adaptive.o build/tu_misc.o
echo built tools/variants/$name.so
