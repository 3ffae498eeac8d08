#!/bin/bash
O=gpurun_out; mkdir -p $O
export SSDBOX_EXP=1
{
SSDBOX_BWD_GS=1 SSDBOX_BWD_STAMPS=1 python tools/run_bwd.py 2>&1 | tail -4
SSDBOX_BWD_GS=1 SSDBOX_BWD_STAMPS=1 SSDBOX_BWD_ABLATE=8 python tools/run_bwd.py 2>&1 | tail -3
} | tee $O/r04j_bwd_gs.log
