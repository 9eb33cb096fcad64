#!/bin/bash
# Build one A/B variant of libmcd.so into build/ab/<name>.so (tuning tool):
#   scripts/build_variant.sh <name> [-DMCD_TK_CHUNK=8 ...]
# Only mcd_traj.cu and mcd_api.cu see the extra flags; the other objects come from the normal build.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/mcdiscrete_b200/csrc
out=$root/build/ab; tmp=$root/build/ab/obj_$name
mkdir -p $tmp
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v -I$src"
make -C $src -j4 > /dev/null
for f in mcd_traj mcd_api; do
  $NV "$@" -c $src/$f.cu -o $tmp/$f.o 2> $tmp/$f.ptxas.log &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/$name.so $tmp/mcd_traj.o $tmp/mcd_api.o $src/mcd_tables.o $src/mcd_unit.o $src/mcd_bench.o
grep -A1 "traj_kernel_poolILi3ELb0" $tmp/mcd_traj.ptxas.log | grep -o "Used [0-9]* registers.*" | head -1
grep -A2 "traj_kernel_poolILi3ELb0" $tmp/mcd_traj.ptxas.log | grep -o "[0-9]* bytes spill stores" | head -1
echo "built $out/$name.so"
