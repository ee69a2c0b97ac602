#!/bin/bash
# builds a variant of libsplendor_b200.so with extra -D flags for some translation units (experiments; select it with SPL_B200_LIB=<path>)
#   profiles/tools/build_variant.sh NAME "unit1 unit2" -DMW=2 -DDESC_MINB=10 ...        (units without .cu, e.g. "spl_mcts")
set -e
cd "$(dirname "$0")/../.."
name=$1; units=$2; shift; shift
P=alphazero-general-ori_b200
mkdir -p $P/build/variants
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
objs=""
for u in $(cd $P/csrc; ls *.cu | sed 's/\.cu$//'); do
  if echo " $units " | grep -q " $u "; then
    nvcc $F "$@" -Xptxas -v -c -o $P/build/variants/${u}_$name.o $P/csrc/$u.cu 2> $P/build/variants/${u}_$name.ptxas.log &
    objs="$objs $P/build/variants/${u}_$name.o"
  else
    objs="$objs $P/build/$u.o"
  fi
done
wait
nvcc $F -shared -o $P/build/variants/libsplendor_b200_$name.so $objs
echo $P/build/variants/libsplendor_b200_$name.so
