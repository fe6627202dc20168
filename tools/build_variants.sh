#!/bin/bash
# Tuning builds of libharpb200.so with different compile-time knobs, selected at run time with HARP_LIB=<path>.
#   bash tools/build_variants.sh && HARP_LIB=harp_b200/variants/libharpb200_A.so python bench.py --no-e2e --no-cpu-baseline
set -e
cd "$(dirname "$0")/.."
mkdir -p harp_b200/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared"
build() { /usr/local/cuda/bin/nvcc $FLAGS $2 harp_b200/csrc/*.cu -o harp_b200/variants/libharpb200_$1.so; }
build A "-DHARP_PV_MINCTAS=20 -DHARP_PV_CHUNK=48 -DHARP_PV_QCAP=160" &   # pixel_vis at 20 CTAs / SM
build B "-DHARP_SHADE_MIN_CTAS=5" &                                        # shade at 5 CTAs / SM (<= 102 registers)
build C "-DHARP_SHADE_GRID=192" &                                          # shade: 192 CTAs per frame
build D "-DHARP_PV_CHUNK=32 -DHARP_PV_QCAP=352 -DHARP_PV_MINCTAS=16" &     # pixel_vis: smaller staging chunks
wait
ls -la harp_b200/variants
