#!/bin/bash
# usage: tools/build_variant.sh <name> <unit.cu> <extra nvcc flags...>   -> luminary_b200/liblumb200_<name>.so
# Rebuilds ONE translation unit with extra flags and links it with the stock objects (tuning experiments).
set -e
name=$1; unit=$2; shift 2
cd "$(dirname "$0")/../luminary_b200/csrc"
flags="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -I ../../include"
case $unit in shade.cu|sky.cu) flags="$flags --use_fast_math";; trace.cu|bvh_build.cu) flags="$flags -fmad=false -prec-div=true -prec-sqrt=true";; esac
/usr/local/cuda/bin/nvcc $flags "$@" -c $unit -o /tmp/variant_$name.o
objs=""; for o in bvh_build.o trace.o shade.o sky.o device_api.o comm.o host/light_tree.o; do [ "$o" = "${unit%.cu}.o" ] && objs="$objs /tmp/variant_$name.o" || objs="$objs $o"; done
/usr/local/cuda/bin/nvcc -shared -o ../liblumb200_$name.so $objs -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -ldl
echo ../liblumb200_$name.so
