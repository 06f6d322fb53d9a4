#!/bin/bash
# Builds a variant of the whole library for kernel A/B measurements: tools/build_variant.sh NAME -DMACRO=... [...]
# -> tools/variants/libhdmoe_NAME.so (use with HDMOE_B200_LIB=tools/variants/libhdmoe_NAME.so python tools/perf_shapes.py)
set -e
name=$1; shift
PKG=heterogeneous-moe-for-diffusion-models_b200
mkdir -p tools/variants build/var_$name
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Iinclude -Wno-deprecated-gpu-targets"
objs=""
for f in $PKG/csrc/*.cu; do
  b=$(basename $f .cu)
  case $b in
    gconv2|gwgrad2) nvcc $FLAGS "$@" -c $f -o build/var_$name/$b.o; objs="$objs build/var_$name/$b.o";;
    *) objs="$objs build/$b.o";;
  esac
done
nvcc -shared -o tools/variants/libhdmoe_$name.so $objs -lcudart
echo built tools/variants/libhdmoe_$name.so
