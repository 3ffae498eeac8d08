#!/bin/bash
# experiment: number of match warps in loss_stream_kernel (kMatchWarps; 4 warps also leave room for a 5th ring stage)
S=object-detection-pytorch_b200/csrc
for w in "$@"; do
  sed -i "s/^constexpr int kMatchWarps = [0-9]*;/constexpr int kMatchWarps = $w;/" $S/loss.cu
  python object-detection-pytorch_b200/build.py > /dev/null 2>&1 || { echo "build failed for $w"; continue; }
  echo "kMatchWarps=$w"
  bash tools/exp_ring_grid.sh "0 0 0" "0 0 0"
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "loss" 2>&1 | tail -1
done
sed -i "s/^constexpr int kMatchWarps = [0-9]*;/constexpr int kMatchWarps = 8;/" $S/loss.cu
python object-detection-pytorch_b200/build.py > /dev/null 2>&1
