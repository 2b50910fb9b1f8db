#!/bin/bash
# Builds libnfc_b200.so (sm_100a only) in-tree: two translation units compiled side by side, then linked.
# Usage: ./build.sh [-v] [-f]     -v: ptxas resource usage, -f: rebuild everything (default: only objects older than their sources)
set -e
cd "$(dirname "$0")"
PKG="neuralfreeconvection.jl_b200"
EXTRA=""; FORCE=0
for a in "$@"; do [ "$a" = "-v" ] && EXTRA="-Xptxas -v"; [ "$a" = "-f" ] && FORCE=1; done
FLAGS="$NFC_EXTRA_FLAGS -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -diag-suppress 128 $EXTRA -Xcompiler -fPIC"
BUILD=${NFC_BUILD_DIR:-build}; SO=${NFC_SO:-libnfc_b200.so}     # debug builds: NFC_EXTRA_FLAGS=-DNFC_DEBUG_TIMERS NFC_BUILD_DIR=build_dbg NFC_SO=libnfc_b200_dbg.so
mkdir -p $BUILD
stale() {   # $1 object, rest: sources
    local o="$1"; shift
    [ "$FORCE" = 1 ] || [ ! -f "$o" ] && return 0
    for s in "$@"; do [ "$s" -nt "$o" ] && return 0; done
    return 1
}
COMMON=$(ls $PKG/csrc/*.cuh $PKG/csrc/*.h include/nfc.h | grep -v nfc_adjoint_tc)
PIDS=""
if stale $BUILD/nfc_api.o $PKG/csrc/nfc_api.cu $COMMON; then nvcc $FLAGS -c -o $BUILD/nfc_api.o "$PKG/csrc/nfc_api.cu" & PIDS="$PIDS $!"; fi
if stale $BUILD/nfc_adjoint_tc.o $PKG/csrc/nfc_adjoint_tc.cu $PKG/csrc/nfc_adjoint_tc.cuh $COMMON; then nvcc $FLAGS -c -o $BUILD/nfc_adjoint_tc.o "$PKG/csrc/nfc_adjoint_tc.cu" & PIDS="$PIDS $!"; fi
for p in $PIDS; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a --shared -ldl -o "$PKG/$SO" $BUILD/nfc_api.o $BUILD/nfc_adjoint_tc.o
