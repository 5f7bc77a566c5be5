#!/bin/bash
# tools/build_exp.sh NAME "-DFLAG ..." file.cu [file.cu ...]  — experiment build of libhuygens_cuda.so into
# huygens_b200/lib/exp/lib_NAME.so: the listed sources recompiled with the extra flags, everything else from build/.
# Select it with HG_LIB_PATH.
set -e
cd "$(dirname "$0")/../huygens_b200/csrc"
name=$1; flags=$2; shift 2
make -s >/dev/null
mkdir -p ../lib/exp
objs=""
for src in common comm filterbank filterbank_ru filterbank_ws_f32 filterbank_ws_f64 filterbank_capi filterbank_scan oscbank rmsbank cascade; do
  o=build/$src.o
  for f in "$@"; do
    if [ "$f" = "$src.cu" ]; then
      o=build/exp_${name}_$src.o
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags -c $src.cu -o $o 2>/dev/null
    fi
  done
  objs="$objs $o"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/exp/lib_$name.so $objs -lcudart -ldl
echo built lib_$name.so
