#!/bin/bash
O=gpurun_out
python tools/run_dense.py > $O/r02r_dense_plain.log 2>&1; tail -2 $O/r02r_dense_plain.log
ncu --set full --clock-control none --import-source on -k regex:"detect_overflow_chunk|detect_segment_kernel" -s 2 -c 2 -f -o $O/r02r_dense python tools/run_dense.py > $O/r02r_ncu.log 2>&1; echo "ncu exit $?"
