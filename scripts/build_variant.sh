#!/bin/bash
# Developer helper: build a named variant of the engine with extra -D flags (experiments only).
# usage: scripts/build_variant.sh NAME "-DGW_DEBUG -DWALK_DEBUG"
set -e
cd "$(dirname "$0")/.."
mkdir -p summersph_b200/variants
nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared --expt-relaxed-constexpr $2 \
  -o summersph_b200/variants/libsph_$1.so summersph_b200/csrc/sph_engine.cu -ldl
echo built summersph_b200/variants/libsph_$1.so
