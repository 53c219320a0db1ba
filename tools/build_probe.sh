#!/bin/bash
# Builds the developer probe (tools/gemm_probe) for sm_100a.
set -e
cd "$(dirname "$0")/.."
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xptxas -v \
    -o tools/gemm_probe tools/gemm_probe.cu hermipy_b200/csrc/hb_gemm.cu -lcublas 2>&1 | grep -E "error|warning|registers|spill|Compiling entry" | head -60
