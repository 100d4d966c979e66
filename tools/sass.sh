#!/bin/bash
# usage: tools/sass.sh <file.cu> <mangled-function-substring>  -> compiles one source, prints ptxas info, dumps SASS to /tmp/<name>.sass
set -e
cd /root/repo/ditreeonlineplanner_b200
src=$1; base=$(basename $src .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v -c csrc/$base.cu -o build/$base.o 2>&1 | grep -E "Compiling entry|Used|stack|error|warning" || true
cuobjdump -sass build/$base.o > /tmp/$base.sass
