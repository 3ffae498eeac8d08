#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "multibox_loss or match or refine or full_batch or loss" 2>&1 | tail -2
X="--no-side-phases --no-cpu-baseline --e2e-steps 1 --steps 200"
for W in rfb300_voc ssd300_voc refinedet320_voc ssd512_coco; do
python bench.py $X --workload $W 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$W %.2f us' % (1e3*d['ms_per_step']), 'loss_stream %.1f' % d['phases']['kernels_us']['loss_stream'], d['sanity']['loss_c'])"
done 2>&1 | tee $O/r04_mw10.log
