#!/bin/bash
O=gpurun_out; mkdir -p $O
export SSDBOX_EXP=1
{
for a in 0 1 2 4 5 6 7 8 15; do SSDBOX_BWD_ABLATE=$a python tools/run_bwd.py 2>&1 | tail -1; done
} | tee $O/r04e_bwd_ablate.log
