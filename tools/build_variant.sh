#!/bin/bash
# build_variant.sh NAME "EXTRA_NVCC_FLAGS" SRC...  -- recompile the given .cu files with extra flags and link a
# variant library csrc/_variants/lib_NAME.so from them plus the regular objects (tuning experiments only).
set -e
cd "$(dirname "$0")/../online_convex_optimization_b200/csrc"
name=$1; extra=$2; shift 2
mkdir -p _variants/_obj_$name
objs=""
for o in _obj/*.o; do
  b=$(basename $o .o); skip=0
  for s in "$@"; do [ "$b" = "$(basename $s .cu)" ] && skip=1; done
  [ $skip = 0 ] && objs="$objs $o"
done
pids=""
for s in "$@"; do
  b=$(basename $s .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       -Xptxas -v $extra -c $s -o _variants/_obj_$name/$b.o > _variants/_obj_$name/$b.log 2>&1 &
  pids="$pids $!"
  objs="$objs _variants/_obj_$name/$b.o"
done
for p in $pids; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o _variants/lib_$name.so $objs
echo built _variants/lib_$name.so
