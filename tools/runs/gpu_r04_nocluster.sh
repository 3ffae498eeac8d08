#!/bin/bash
O=gpurun_out; mkdir -p $O
X="--no-side-phases --no-cpu-baseline --e2e-steps 1 --steps 200"
for W in ssd300_voc fssd300_coco rfb300_voc; do
for F in 0 4; do
python bench.py $X --workload $W --loss-flags $F 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$W flags=$F %.2f us' % (1e3*d['ms_per_step']), 'mine %.1f' % d['phases']['kernels_us']['mine_reduce'], d['sanity']['loss_c'])"
done
done 2>&1 | tee $O/r04_nocluster.log
