#!/bin/bash
# builds the measurement probes (not part of the library): tools/csrc/imma_probe.cu (tensor-core product in isolation; modes
# 3 / 4 = no epilogue / no operand loads: the MMA stream alone) and tools/csrc/int_peaks.cu (integer-pipe ceilings)
set -e
cd "$(dirname "$0")/.."
mkdir -p pvw-rs_b200/build
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"
nvcc $F -o pvw-rs_b200/build/imma_probe tools/csrc/imma_probe.cu pvw-rs_b200/csrc/imma.cu
nvcc $F -o pvw-rs_b200/build/int_peaks tools/csrc/int_peaks.cu
