#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "backward or full_batch or smoke or random_shapes or edge" > $O/r02k_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02k_pytest.log
python tools/run_bwd.py 0 2>&1 | tail -2
X="--no-cpu-baseline --e2e-steps 1 --steps 100 --no-voc-eval"
for F in 0 32; do
timeout 400 python bench.py $X --loss-flags $F > $O/r02k_bench_f$F.json 2> $O/r02k_bench_f$F.err; echo "bench flags $F exit $?"
done
