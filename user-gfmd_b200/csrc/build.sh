#!/bin/sh
# Builds libgfmd_b200.so in-tree for sm_100a (no other architecture, no fallback).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=${GFMD_OUT:-"$HERE/../libgfmd_b200.so"}
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared ${GFMD_NVCC_EXTRA} \
  -o "$OUT" "$HERE/gfmd_b200.cu" -ldl
echo "built $OUT"
