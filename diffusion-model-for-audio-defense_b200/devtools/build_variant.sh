#!/bin/bash
# Development aid: build devtools/ab/lib_<name>.so from the current sources with extra -D flags on ap_wavenet_tc.cu
# (the other objects come from the last regular build).  Usage: build_variant.sh <name> [-DFLAG ...]
set -e
PKG="$(cd "$(dirname "$0")/.." && pwd)"
name="$1"; shift
mkdir -p "$PKG/devtools/ab"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC "$@" \
  -c "$PKG/csrc/ap_wavenet_tc.cu" -o "$PKG/devtools/ab/tc_$name.o"
objs=""
for o in ap_update ap_query ap_wavenet ap_mel ap_classifier ap_conv_tc ap_unet; do objs="$objs $PKG/build/$o.o"; done
nvcc -shared -o "$PKG/devtools/ab/lib_$name.so" $objs "$PKG/devtools/ab/tc_$name.o" -cudart static
rm -f "$PKG/devtools/ab/tc_$name.o"
echo "$PKG/devtools/ab/lib_$name.so"
