#!/bin/bash
# tuning builds of the 3-D sweep kernels: scripts/build3d_variant.sh NAME -DW5_3D_NT=... -> build/libweno5mhd_NAME.so
# (run with WENO5_LIB=build/libweno5mhd_NAME.so python scripts/bench3d.py)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
    -o build/libweno5mhd_$name.so julia_b200/csrc/weno5_kernels.cu julia_b200/csrc/weno5_line1d.cu \
    julia_b200/csrc/weno5_api.cu julia_b200/csrc/weno5_3d.cu -ldl
