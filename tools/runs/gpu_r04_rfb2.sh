#!/bin/bash
O=gpurun_out; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
X="--no-side-phases --no-cpu-baseline --steps 200 --workload rfb300_voc"
timeout 200 python bench.py $X > $O/r04_rfb_1gpu.json 2> $O/r04_rfb_1gpu.err; echo "N=1 exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29951 bench.py --gpus $N $X > $O/r04_rfb_${N}gpu.json 2> $O/r04_rfb_${N}gpu.err; echo "N=$N exit $?"
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r04_rfb_*gpu.json')):
    d = json.loads([l for l in open(f) if l.startswith('{')][-1])
    print(f.split('/')[-1], d['n_gpus'], '%.1f us' % (1e3 * d['ms_per_step']), '%.0f images/s' % d['value'], (d.get('sanity') or {}).get('mgpu'))
P
