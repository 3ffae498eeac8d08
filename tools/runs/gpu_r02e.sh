#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 300 python tools/phase_timing.py > $O/r02e_phase_timing.log 2>&1; echo "phase timing exit $?"
timeout 600 python -m pytest tests -m gpu -x -q -k "peer or multi or graph" > $O/r02e_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02e_pytest.log
X="--no-side-phases --no-cpu-baseline --e2e-steps 1 --steps 200"
for F in 0 4; do
timeout 200 python bench.py $X --workload ssd300_voc --only T --loss-flags $F > $O/r02e_ssd300_T_f$F.json 2> $O/r02e_ssd300_T_f$F.err; echo "ssd300 T flags $F exit $?"
done
