#!/bin/bash
O=gpurun_out; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
X="--no-side-phases --no-cpu-baseline --steps 200 --workload refinedet320_voc"
timeout 200 python bench.py $X > $O/r04r_refine_1gpu.json 2> $O/r04r_refine_1gpu.err; echo "N=1 exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29931 bench.py --gpus $N $X > $O/r04r_refine_${N}gpu.json 2> $O/r04r_refine_${N}gpu.err; echo "N=$N exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29932 bench.py --gpus $N $X --serial > $O/r04r_refine_${N}gpu_serial.json 2> $O/r04r_refine_${N}gpu_serial.err; echo "N=$N serial exit $?"
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r04r_refine_*gpu*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('/')[-1], d['n_gpus'], '%.1f us' % (1e3 * d['ms_per_step']), '%.0f images/s' % d['value'], (d.get('sanity') or {}).get('mgpu'))
    except Exception as e:
        print(f, 'FAILED', e)
P
tail -3 $O/r04r_refine_${N}gpu.err
