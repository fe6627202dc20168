#!/bin/bash
# Tuning builds of libharpb200.so with different compile-time knobs, selected at run time with HARP_LIB=<path>.
#   bash tools/build_variants.sh && HARP_LIB=harp_b200/variants/libharpb200_A.so python bench.py --no-e2e --no-cpu-baseline
set -e
cd "$(dirname "$0")/.."
mkdir -p harp_b200/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared"
build() { /usr/local/cuda/bin/nvcc $FLAGS $2 harp_b200/csrc/*.cu -o harp_b200/variants/libharpb200_$1.so; }
build q192m15 "-DHARP_PV_QCAP=192 -DHARP_PV_MINCTAS=15" &   # tile rasteriser: 14.3 KB, <= 68 registers -> 15 CTAs / SM
build q192m16 "-DHARP_PV_QCAP=192 -DHARP_PV_MINCTAS=16 -DHARP_PV_CHUNK=16" &   # 13.6 KB, <= 64 registers -> 16 CTAs / SM
wait
ls -la harp_b200/variants
