#!/bin/bash
# Build an experimental variant of libgstex_b200.so: recompile ONE source with extra -D flags and relink with the
# production objects.   Usage: tools/build_variant.sh <name> <source.cu> <flags...>   -> experiments/_variants/lib_<name>.so
set -e
cd /root/repo
name=$1; src=$2; shift 2
python -m gstex_cuda_b200.build > /dev/null
B=gstex_cuda_b200/csrc/_build
obj=/tmp/variant_${name}.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr \
  "$@" -c gstex_cuda_b200/csrc/$src -o $obj 2> /tmp/variant_${name}.ptxas
grep -A2 "ILb1ELb0" /tmp/variant_${name}.ptxas | grep "Used\|spill" | tr '\n' ' '; echo
objs=$(ls $B/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -o experiments/_variants/lib_${name}.so $objs $obj -gencode arch=compute_100a,code=sm_100a -lcudart
echo built experiments/_variants/lib_${name}.so
