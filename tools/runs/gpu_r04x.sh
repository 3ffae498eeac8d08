#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "backward or bwd or grad or refine or full_size or full_batch or loss" 2>&1 | tail -3
export SSDBOX_EXP=1
{
for i in 1 2 3; do
python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_NO_LOC_WARP=1 python tools/run_bwd.py 2>&1 | tail -1
done
} | tee $O/r04x_bwd.log
