#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "backward or bwd or grad or refine or full_size or full_batch" 2>&1 | tail -3
{
python tools/run_bwd.py 2>&1 | tail -1
export SSDBOX_EXP=1
SSDBOX_BWD_OLD=1 python tools/run_bwd.py 2>&1 | tail -1
python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_ABLATE=1 python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_ABLATE=8 python tools/run_bwd.py 2>&1 | tail -1
} | tee $O/r04g_bwd_gs.log
