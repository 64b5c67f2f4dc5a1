#!/usr/bin/env bash
# Builds the C-ABI CUDA library in-tree (sm_100a only) and the oracle helper objects.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=diffusion_conductor_b200/libdc_b200.so
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared ${DC_NVCC_EXTRA:-} \
  -o "$OUT" diffusion_conductor_b200/csrc/dc_api.cu
echo "built $OUT"
