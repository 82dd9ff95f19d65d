#!/bin/bash
# usage: tools/build_variant.sh NAME -DFOO=1 ...   -> build/variants/librhfq_NAME.so
name=$1; shift
cd /root/repo
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr \
  -Xcompiler -fPIC,-fopenmp,-ffp-contract=off -shared -lgomp -ccbin /usr/bin/g++ "$@" \
  -I include -o build/variants/librhfq_$name.so reheatfunq_b200/csrc/rhfq_engine.cu 2>&1 | grep -i "error"
cuobjdump -res-usage build/variants/librhfq_$name.so 2>/dev/null | grep -A1 "rhfq_saq_step_bli" | grep -o "REG:[0-9]*"
