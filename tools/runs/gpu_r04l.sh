#!/bin/bash
O=gpurun_out; mkdir -p $O
export SSDBOX_EXP=1
{
python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_ABLATE=32 python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_ABLATE=16 python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_ABLATE=48 python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_ABLATE=8 python tools/run_bwd.py 2>&1 | tail -1
python tools/run_bwd.py 2>&1 | tail -1
} | tee $O/r04l_bwd_prefetch.log
