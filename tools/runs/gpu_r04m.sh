#!/bin/bash
O=gpurun_out; mkdir -p $O
export SSDBOX_EXP=1
{
python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_ABLATE=64 python tools/run_bwd.py 2>&1 | tail -1
} | tee $O/r04m_bwd.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_bwd_stream -s 3 -c 1 -f -o $O/r04m_bwd python tools/run_bwd.py > $O/r04m_ncu.log 2>&1; echo "ncu exit $?"
