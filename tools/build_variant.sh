#!/bin/bash
# Build an A/B variant of libtfswa_b200.so: tools/build_variant.sh <name> <file.cu> <extra nvcc flags...>
# -> tfswa-unet_b200/lib/libtfswa_b200_<name>.so ; run with TFSWA_B200_LIB=<that path>
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
P=$ROOT/tfswa-unet_b200
name=$1; src=$2; shift 2
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC "$@" \
  -c $P/csrc/$src -o $P/build/${src%.cu}_$name.o
objs=""
for f in $P/csrc/*.cu; do b=$(basename ${f%.cu}); [ "$b" != "${src%.cu}" ] && objs="$objs $P/build/$b.o"; done
nvcc -shared -o $P/lib/libtfswa_b200_$name.so $objs $P/build/${src%.cu}_$name.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo $P/lib/libtfswa_b200_$name.so
