#!/bin/bash
# Tuning builds of libharpb200.so with different compile-time knobs, selected at run time with HARP_LIB=<path>.
#   bash tools/build_variants.sh && HARP_LIB=harp_b200/variants/libharpb200_A.so python bench.py --no-e2e --no-cpu-baseline
set -e
cd "$(dirname "$0")/.."
mkdir -p harp_b200/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared"
build() { /usr/local/cuda/bin/nvcc $FLAGS $2 harp_b200/csrc/*.cu -o harp_b200/variants/libharpb200_$1.so; }
build c24m12 "-DHARP_PV_CHUNK=24 -DHARP_PV_MINCTAS=12" &   # tile rasteriser: 16.7 KB of shared memory -> 12 resident CTAs (registers) instead of 11
build c20m14 "-DHARP_PV_CHUNK=20 -DHARP_PV_MINCTAS=14" &   # 15.4 KB, <= 73 registers -> 14 CTAs / SM
build c16m16 "-DHARP_PV_CHUNK=16 -DHARP_PV_MINCTAS=16" &   # 13.9 KB, <= 64 registers -> 16 CTAs / SM
wait
ls -la harp_b200/variants
