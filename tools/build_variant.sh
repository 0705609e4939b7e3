#!/bin/bash
# tools/build_variant.sh NAME [nvcc -D switches ...]   A/B build of librelem with other compile-time switches:
#   rnaelem_b200/variants/librelem_NAME.so   (git-ignored; select it with RELEM_LIBRARY=<path> for bench.py / the probes)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
CS=$ROOT/rnaelem_b200/csrc
OUT=$ROOT/rnaelem_b200/variants
B=$OUT/_build_$NAME
mkdir -p $B
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
COMMON="-std=c++17 -O3 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a"
$NVCC $COMMON -fmad=false "$@" -c $CS/relem_api.cu -o $B/relem_api.o &
$NVCC $COMMON "$@" -c $CS/relem_lin.cu -o $B/relem_lin.o &
$NVCC $COMMON -fmad=false "$@" -c $CS/host_model.cpp -o $B/host_model.o &
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/librelem_$NAME.so $B/relem_api.o $B/relem_lin.o $B/host_model.o -ldl
rm -rf $B
echo $OUT/librelem_$NAME.so
