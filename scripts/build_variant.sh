#!/bin/bash
# build_variant.sh NAME "-DFLAG ..." : compile a tuning variant of the library into build_variants/NAME.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants/obj_$1
objs=""
for f in vp_context vp_splat vp_mesh vp_rle vp_nodes vp_edit vp_worldfile vp_worldgen_dev vp_multi; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --cudart static $2 -c voxplat_b200/csrc/$f.cu -o build_variants/obj_$1/$f.o &
  objs="$objs build_variants/obj_$1/$f.o"
done
wait
nvcc -shared --cudart static -Wno-deprecated-gpu-targets -o build_variants/$1.so $objs
echo build_variants/$1.so
