#!/bin/bash
# build_variant.sh <name> <nvcc -D flags...>: libhpfw_b200_<name>.so = the current objects with cqt.cu recompiled under the flags
# (A/B experiments on the GPU box: HPFW_B200_LIB=hpfw_b200/libhpfw_b200_<name>.so python scripts/cqt_tune.py 48)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr "$@" \
  -c hpfw_b200/csrc/cqt.cu -o hpfw_b200/_build/cqt_$name.o
objs=$(ls hpfw_b200/_build/*.o | grep -v "/cqt" )
nvcc -shared -o hpfw_b200/libhpfw_b200_$name.so $objs hpfw_b200/_build/cqt_$name.o -gencode arch=compute_100a,code=sm_100a -lcudart_static
echo built hpfw_b200/libhpfw_b200_$name.so
