#!/bin/bash
# Builds libnfc_b200.so (sm_100a only) in-tree: two translation units compiled side by side, then linked.  Usage: ./build.sh [-v]
set -e
cd "$(dirname "$0")"
PKG="neuralfreeconvection.jl_b200"
EXTRA=""
[ "$1" = "-v" ] && EXTRA="-Xptxas -v"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -diag-suppress 128 $EXTRA -Xcompiler -fPIC"
mkdir -p build
nvcc $FLAGS -c -o build/nfc_api.o "$PKG/csrc/nfc_api.cu" &
P1=$!
nvcc $FLAGS -c -o build/nfc_adjoint_tc.o "$PKG/csrc/nfc_adjoint_tc.cu" &
P2=$!
wait $P1
wait $P2
nvcc -gencode arch=compute_100a,code=sm_100a --shared -o "$PKG/libnfc_b200.so" build/nfc_api.o build/nfc_adjoint_tc.o
