#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "backward or bwd or grad or refine or full_size or full_batch or loss" 2>&1 | tail -3
export SSDBOX_EXP=1
{
for i in 1 2; do
python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_ABLATE=128 python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_ABLATE=176 python tools/run_bwd.py 2>&1 | tail -1
done
SSDBOX_BWD_ABLATE=8 python tools/run_bwd.py 2>&1 | tail -1
} | tee $O/r04q_bwd.log
