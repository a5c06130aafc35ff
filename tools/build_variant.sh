#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags...]  -> gpurun_variants/libccb200_<name>.so (same ABI; select with CCB_LIB_PATH)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/chunk-compaction-in-vectorized-execution-simd_b200/csrc
out=$root/gpurun_variants; mkdir -p $out/obj_$name
for f in runtime tables probe_chunk probe_batch compactor chain_fused partition pjoin; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I$root/include -I$src "$@" -c $src/$f.cu -o $out/obj_$name/$f.o &
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -I$root/include -I$src -c $src/tuner.cpp -o $out/obj_$name/tuner.o &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libccb200_$name.so $out/obj_$name/*.o -lcudart_static -lpthread -ldl -lrt
rm -rf $out/obj_$name
echo built $out/libccb200_$name.so
