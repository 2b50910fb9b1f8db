#!/bin/bash
# Builds libnfc_b200.so (sm_100a only) in-tree.  Usage: ./build.sh [-v]
set -e
cd "$(dirname "$0")"
PKG="neuralfreeconvection.jl_b200"
EXTRA=""
[ "$1" = "-v" ] && EXTRA="-Xptxas -v"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -diag-suppress 128 $EXTRA --shared -Xcompiler -fPIC \
     -o "$PKG/libnfc_b200.so" "$PKG/csrc/nfc_api.cu"
