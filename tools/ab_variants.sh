#!/bin/bash
# Build variant libraries that differ only in compile-time switches of ONE source file, for same-box A/B timing:
#   tools/ab_variants.sh attention_sm100.cu name1 "-DX=1" name2 "-DX=2 -DY=1" ...
#   -> multimodalsum_b200/libmmsum_b200_<name>.so ; select with MMSUM_LIB_PATH.  (The other objects come from the last normal build.)
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
src=$1; shift
csrc=$root/multimodalsum_b200/csrc
others=$(ls $csrc/*.o | grep -v "${src%.cu}.o" | grep -v "_var_")
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  (
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 --expt-relaxed-constexpr $flags \
         -c $csrc/$src -o $csrc/_var_$name.o
    nvcc -shared -o $root/multimodalsum_b200/libmmsum_b200_$name.so $csrc/_var_$name.o $others -gencode arch=compute_100a,code=sm_100a
    rm -f $csrc/_var_$name.o
    echo built $name "($flags)"
  ) &
done
wait
