#!/bin/bash
O=gpurun_out
mkdir -p $O
python tools/run_bwd.py 0 > $O/r02i_bwd_plain.log 2>&1; echo "plain exit $?"; tail -2 $O/r02i_bwd_plain.log
ncu --set full --clock-control none --import-source on -k regex:loss_bwd -s 2 -c 1 -f -o $O/r02i_bwd_fill python tools/run_bwd.py 0 > $O/r02i_ncu_fill.log 2>&1; echo "ncu fill exit $?"
ncu --set full --clock-control none --import-source on -k regex:loss_bwd -s 2 -c 1 -f -o $O/r02i_bwd_tma python tools/run_bwd.py 32 > $O/r02i_ncu_tma.log 2>&1; echo "ncu tma exit $?"
