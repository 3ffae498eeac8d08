#!/bin/bash
# experiment build of the C-ABI library (-DSSDBOX_EXPERIMENTS: the SSDBOX_RING_* / SSDBOX_CARVEOUT / SSDBOX_BWD_TWO_PASS
# environment knobs are compiled in; the release library reads no environment) -> tools/libssdbox_exp.so
set -e
cd "$(dirname "$0")/.."
S=object-detection-pytorch_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC \
  -Xcompiler -fvisibility=hidden -cudart static -DSSDBOX_EXPERIMENTS "$@" -I include -I $S \
  $S/abi.cu $S/boxops.cu $S/match.cu $S/loss.cu $S/detect.cu $S/evalpost.cu $S/heads.cu $S/voceval.cu -o tools/libssdbox_exp.so
