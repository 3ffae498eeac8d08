#!/bin/bash
O=gpurun_out
TAG=r02a
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-graph --no-side-phases"
$CMD > $O/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${TAG}_ncu_launches.csv $CMD > $O/${TAG}_ncu_list.log 2>&1
echo "ncu list exit $?"
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-graph --no-voc-eval"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${TAG}_ncu_launches_with_side_phases.csv $CMD2 > $O/${TAG}_ncu_list2.log 2>&1
echo "ncu list2 exit $?"
