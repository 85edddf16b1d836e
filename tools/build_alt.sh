#!/bin/bash
# Build an alternative libcloudsc_b200.so with extra nvcc flags: tools/build_alt.sh <tag> <flags...>
# -> dwarf_p_cloudsc_b200/lib_alt/<tag>/libcloudsc_b200.so ; use with CLOUDSC_B200_LIB=<that path> python bench.py ...
set -e
TAG=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/dwarf_p_cloudsc_b200/lib_alt/$TAG
mkdir -p $OUT
CS=$ROOT/dwarf_p_cloudsc_b200/csrc
for f in cloudsc_abi cloudsc_kernels; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ -I$ROOT/include -I$CS "$@" -c $CS/$f.cu -o $OUT/$f.o &
done
wait
/usr/bin/gcc -std=c11 -O2 -fPIC -I$ROOT/include -c $ROOT/dwarf_p_cloudsc_b200/host/cloudsc_globals.c -o $OUT/cloudsc_globals.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o $OUT/libcloudsc_b200.so $OUT/*.o -lcudart
echo $OUT/libcloudsc_b200.so
