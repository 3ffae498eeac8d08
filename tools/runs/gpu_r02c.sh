#!/bin/bash
# round-2 GPU call C: where do the two halves of the step stand in graph replay, and what do the ring knobs buy?
O=gpurun_out
mkdir -p $O
X="--no-side-phases --no-cpu-baseline --e2e-steps 1 --steps 200"
run() { # tag, env..., -- args
  tag=$1; shift
  env "$@" > /dev/null 2>&1
}
for ONLY in T D; do
  for CFG in "I0_S0_G0" "I148_S0_G0" "I0_S4_G0" "I148_S4_G0" "I0_S0_G132" "I148_S0_G132" "I148_S4_G132" "I8_S0_G0" "I37_S0_G0"; do
    I=$(echo $CFG | sed 's/I\([0-9]*\)_.*/\1/'); S=$(echo $CFG | sed 's/.*_S\([0-9]*\)_.*/\1/'); G=$(echo $CFG | sed 's/.*_G\([0-9]*\)/\1/')
    unset SSDBOX_RING_INTERLEAVE SSDBOX_RING_MAX_STAGES SSDBOX_RING_GRID
    [ $I != 0 ] && export SSDBOX_RING_INTERLEAVE=$I
    [ $S != 0 ] && export SSDBOX_RING_MAX_STAGES=$S
    [ $G != 0 ] && export SSDBOX_RING_GRID=$G
    timeout 200 python tools/exp_bench.py $X --only $ONLY > $O/r02c_${ONLY}_${CFG}.json 2> $O/r02c_${ONLY}_${CFG}.err; echo "$ONLY $CFG exit $?"
  done
done
unset SSDBOX_RING_INTERLEAVE SSDBOX_RING_MAX_STAGES SSDBOX_RING_GRID
for F in 2 4; do
  timeout 200 python tools/exp_bench.py $X --only T --loss-flags $F > $O/r02c_T_flags$F.json 2> $O/r02c_T_flags$F.err; echo "T flags $F exit $?"
done
for W in ssd300_voc rfb300_voc; do for ONLY in T D; do
  timeout 200 python tools/exp_bench.py $X --only $ONLY --workload $W > $O/r02c_${W}_${ONLY}.json 2> $O/r02c_${W}_${ONLY}.err; echo "$W $ONLY exit $?"
done; done
