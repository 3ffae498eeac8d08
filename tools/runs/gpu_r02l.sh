#!/bin/bash
O=gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"bwd|fill" -c 12 --csv --log-file $O/r02l_bwd_list.csv python tools/run_bwd.py 0 > $O/r02l_ncu.log 2>&1; echo "exit $?"
grep -v "^==" $O/r02l_bwd_list.csv | cut -d, -f5,13- | tail -40
