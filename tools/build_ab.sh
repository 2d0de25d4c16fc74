#!/bin/bash
# A/B builds: tools/build_ab.sh NAME [-DFLAG ...]  ->  ab/NAME.so (git-ignored; travels to the GPU box).
# Select it with SM_LIB_PATH=ab/NAME.so (read by slime_mold_b200/_lib.py, A/B runs only).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p ab
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-O2,-ffp-contract=off,-fno-fast-math -Xptxas -v"
for s in engine gauss exchange; do
  nvcc $F "$@" -c -o ab/$name.$s.o slime_mold_b200/csrc/$s.cu > ab/$name.$s.log 2>&1 &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ab/$name.so ab/$name.engine.o ab/$name.gauss.o ab/$name.exchange.o -ldl
rm -f ab/$name.*.o
grep -A2 "k_trail_rowsILi2ELi1ELi4ELb0E" ab/$name.engine.log | grep -E "spill|Used" | head -2
