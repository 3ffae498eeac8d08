#!/bin/bash
# round-2 GPU call A: parity suite, bench line, carve-out experiment, small workloads, mining phase stamps
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02a_smi.txt
nproc > $O/r02a_nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02a_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02a_pytest.log
timeout 600 python bench.py > $O/r02a_bench.json 2> $O/r02a_bench.err; echo "bench exit $?"
X="--no-side-phases --no-cpu-baseline --e2e-steps 1 --steps 100"
for W in ssd512_coco ssd300_voc fssd300_coco rfb300_voc; do
  for CV in none 100; do
    if [ $CV = none ]; then unset SSDBOX_CARVEOUT; else export SSDBOX_CARVEOUT=$CV; fi
    timeout 300 python tools/exp_bench.py $X --workload $W > $O/r02a_exp_${W}_carve${CV}.json 2> $O/r02a_exp_${W}_carve${CV}.err; echo "exp $W carve=$CV exit $?"
  done
done
unset SSDBOX_CARVEOUT
timeout 300 python bench.py $X --workload refinedet320_voc > $O/r02a_bench_refinedet320_voc.json 2> $O/r02a_bench_refinedet320_voc.err; echo "refinedet exit $?"
timeout 300 python tools/phase_timing.py > $O/r02a_phase_timing.log 2>&1; echo "phase timing exit $?"
ls -la $O | tail -20
