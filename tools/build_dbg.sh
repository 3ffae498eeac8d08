#!/bin/bash
# debug build of the C-ABI library with in-kernel clock64 stamps (-DSSDBOX_PHASE_TIMING) -> tools/libssdbox_dbg.so
set -e
cd "$(dirname "$0")/.."
S=object-detection-pytorch_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC \
  -Xcompiler -fvisibility=hidden -cudart static -DSSDBOX_PHASE_TIMING "$@" -I include -I $S \
  $S/abi.cu $S/boxops.cu $S/match.cu $S/loss.cu $S/detect.cu $S/evalpost.cu $S/heads.cu $S/voceval.cu -o tools/libssdbox_dbg.so
