#!/bin/bash
O=gpurun_out; mkdir -p $O
X="--no-side-phases --no-cpu-baseline --e2e-steps 1 --steps 200 --workload refinedet320_voc"
for P in 8192 4096; do
for i in 1 2; do
SSDBOX_PAIR_MIN_P=$P python tools/exp_bench.py $X 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('refinedet pair_min_p=$P %.2f us' % (1e3*d['ms_per_step']), 'mine %.1f' % d['phases']['kernels_us']['mine_reduce'], d['sanity']['loss_c'])"
done
done 2>&1 | tee $O/r04_pairp.log
