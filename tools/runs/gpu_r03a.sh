#!/bin/bash
O=gpurun_out
X="--no-side-phases --no-cpu-baseline --e2e-steps 1 --steps 200"
for M in "" "--serial"; do
  T=two; [ -n "$M" ] && T=serial
  timeout 200 python bench.py $X $M > $O/r03a_TD_$T.json 2> $O/r03a_TD_$T.err; echo "TD $T exit $?"
  for W in ssd300_voc fssd300_coco rfb300_voc; do
    timeout 200 python bench.py $X $M --workload $W > $O/r03a_${W}_$T.json 2> $O/r03a_${W}_$T.err; echo "$W $T exit $?"
  done
done
