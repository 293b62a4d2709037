#!/bin/bash
# tools_build_exp.sh NAME "-DFLAG=.. ..." : experiment build of the streamed kernel only (other objects reused)
set -e
cd "$(dirname "$0")/qcrypto-ldpc_b200/csrc"
mkdir -p build_exp
nvcc -gencode arch=compute_100a,code=sm_100a $2 -O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC -Xptxas -v -c layered_i8s.cu -o build_exp/$1.o 2> build_exp/$1.log
grep -E "Used" build_exp/$1.log | sort | uniq -c
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../exp_$1.so build_exp/$1.o build/layered_i8.o build/layered_generic.o build/flooding.o build/flooding_qc.o build/bitops.o build/postproc.o build/api.o build/code.o -lcudart_static -ldl -lrt -lpthread
