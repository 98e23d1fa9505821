#!/bin/bash
# tools/build_variant.sh NAME FILE.cu [-DFLAGS...]: builds tools/sweep/libjn_NAME.so with one translation
# unit recompiled with extra flags (kernel tuning sweeps; select with JN_ELAS_LIB=tools/sweep/libjn_NAME.so)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; SRC=$2; shift 2
PKG="$ROOT/jackal-navigation_b200"
mkdir -p "$ROOT/tools/sweep" /tmp/jnv_$NAME
for f in api descriptor support delaunay planes_grid dense post scan rectify navigate calib jpeg; do cp "$PKG/_obj/$f.o" /tmp/jnv_$NAME/; done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC "$@" \
  -c "$PKG/csrc/$SRC.cu" -o /tmp/jnv_$NAME/$SRC.o 2>&1 | grep -v "deprecated" || true
nvcc -shared -o "$ROOT/tools/sweep/libjn_$NAME.so" /tmp/jnv_$NAME/*.o -cudart static -lnvjpeg_static -lculibos 2>&1 | grep -v "deprecated" || true
echo "$ROOT/tools/sweep/libjn_$NAME.so"
