#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "backward or bwd or grad" 2>&1 | tail -3
{
python tools/run_bwd.py 2>&1 | tail -1
for n in 1 2 4; do SSDBOX_EXP=libssdbox_exp_split$n.so python tools/run_bwd.py 2>&1 | tail -1; done
SSDBOX_BWD_ABLATE=2 SSDBOX_EXP=libssdbox_exp_split4.so python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_ABLATE=15 SSDBOX_EXP=libssdbox_exp_split4.so python tools/run_bwd.py 2>&1 | tail -1
} | tee $O/r04f_bwd_split.log
