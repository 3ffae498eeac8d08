#!/bin/bash
O=gpurun_out; mkdir -p $O
X="--no-side-phases --no-cpu-baseline --e2e-steps 1 --steps 200"
for W in ssd512_coco ssd300_voc; do
for i in 1 2; do
python tools/exp_bench.py $X --workload $W 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$W base %.2f us' % (1e3*d['ms_per_step']), d['sanity']['loss_c'])"
SSDBOX_PDL=1 python tools/exp_bench.py $X --workload $W 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$W PDL  %.2f us' % (1e3*d['ms_per_step']), d['sanity']['loss_c'])"
done
done 2>&1 | tee $O/r04zz_pdl.log
