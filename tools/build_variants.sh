#!/bin/bash
# Experimental builds of the library with other compile-time settings -> kalibr_b200/_exp/<name>.so (git-ignored; they travel to the
# GPU box).  tools/la_timing.py times each of them.   usage: tools/build_variants.sh name "-DKB_LA_VARIANT=1" [name flags ...]
cd "$(dirname "$0")/../kalibr_b200/csrc"
mkdir -p ../_exp
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared $flags \
    -Xptxas -v -o ../_exp/$name.so kb_kernels.cu kb_init.cu kb_host.cpp -ldl 2>&1 | grep -A3 "linearise_assemble[a-z_]*kernelILi0ELb1ELb0" | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores, [0-9]* bytes spill loads" | tr '\n' ' '
  echo " <- $name ($flags)" ) &
done
wait
