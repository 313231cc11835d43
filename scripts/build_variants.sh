#!/bin/bash
# Development aid: builds libhalmd_b200 variants of force.cu, integrate.cu and neighbour.cu with extra -D flags into
# scripts/_build/<name>.so for A/B timing on one box (scripts/ab_variants.sh).
# usage: scripts/build_variants.sh name1 "-DSHELL_MINB=7" name2 "-DNL_THREADS_N=128" ...   (select with HALMD_B200_LIB=...)
set -e
cd "$(dirname "$0")/../halmd_b200/csrc"
make -s >/dev/null
B=../../scripts/_build
mkdir -p $B
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O3 -Xptxas -v"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  for src in force integrate neighbour; do
    $NV $flags -c $src.cu -o $B/${src}_$name.o 2> $B/${src}_$name.log
  done
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $B/$name.so \
      core.o binning.o order.o $B/neighbour_$name.o $B/force_$name.o external.o $B/integrate_$name.o observables.o dd.o modules.o -cudart static
  echo "$name: $(grep -A2 'k_force_shellILi2ELi2ELi4ELb0' $B/force_$name.log | grep -o 'Used [0-9]* registers') (k_force_shell)," \
       "$(grep -A2 'k_nlist_buildILb1' $B/neighbour_$name.log | grep -o 'Used [0-9]* registers') (k_nlist_build)"
done
