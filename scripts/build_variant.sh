#!/bin/bash
# build_variant.sh <name> <files: comma-separated .cu basenames> <nvcc -D flags...>
# libhpfw_b200_<name>.so = the current objects with the named sources recompiled under the flags (A/B experiments on the GPU
# box: HPFW_B200_LIB=hpfw_b200/libhpfw_b200_<name>.so python scripts/cqt_tune.py 48)
set -e
name=$1; files=$2; shift; shift
cd "$(dirname "$0")/.."
objs=""
skip=""
for f in ${files//,/ }; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr "$@" \
    -c hpfw_b200/csrc/$f.cu -o hpfw_b200/_build/var_${name}_$f.o
  objs="$objs hpfw_b200/_build/var_${name}_$f.o"
  skip="$skip|/$f.o"
done
rest=$(ls hpfw_b200/_build/*.o | grep -v "/var_" | grep -v -E "${skip:1}")
nvcc -shared -o hpfw_b200/libhpfw_b200_$name.so $rest $objs -gencode arch=compute_100a,code=sm_100a -lcudart_static
echo built hpfw_b200/libhpfw_b200_$name.so
