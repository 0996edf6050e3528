#!/bin/bash
# Build a variant of the library with extra -D flags into build/var_<name>/ (tuning experiments only).
# usage: tools/build_variant.sh <name> <nvcc flags...>;  run a tool against it with LD_LIBRARY_PATH=build/var_<name>
set -e
name=$1; shift
out=build/var_$name
mkdir -p $out/obj
for f in radvlm_b200/csrc/*.cu radvlm_b200/csrc/*.cpp; do
  b=$(basename $f)
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -x cu "$@" -c $f -o $out/obj/$b.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libradvlm_b200.so $out/obj/*.o
echo built $out/libradvlm_b200.so
