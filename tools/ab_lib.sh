#!/bin/bash
# Build a second copy of the library from the csrc/ of another git revision (default HEAD) for same-box A/B timing:
#   tools/ab_lib.sh [rev]  ->  multimodalsum_b200/libmmsum_b200_base.so ; select it with MMSUM_LIB_PATH=...
set -e
rev=${1:-HEAD}
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p $tmp/multimodalsum_b200/csrc $tmp/include
for f in $(git -C $root ls-tree --name-only $rev multimodalsum_b200/csrc/); do git -C $root show $rev:$f > $tmp/$f; done
git -C $root show $rev:include/mmsum_b200.h > $tmp/include/mmsum_b200.h
objs=""
for s in $tmp/multimodalsum_b200/csrc/*.cu; do
  o=${s%.cu}.o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 --expt-relaxed-constexpr -c $s -o $o &
  objs="$objs $o"
done
wait
nvcc -shared -o $root/multimodalsum_b200/libmmsum_b200_base.so $objs -gencode arch=compute_100a,code=sm_100a
rm -rf $tmp
echo $root/multimodalsum_b200/libmmsum_b200_base.so
