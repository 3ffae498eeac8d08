#!/bin/bash
# gpurun --gpus N: every BASELINE.json config at N = 1 and at N = all GPUs of the box (step only, graph replay)
O=gpurun_out; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
X="--no-side-phases --no-cpu-baseline --steps 200"
P=29920
for W in ssd300_voc ssd512_coco rfb300_voc fssd300_coco refinedet320_voc; do
  timeout 300 python bench.py --workload $W $X > $O/cfg_${W}_1gpu.json 2> $O/cfg_${W}_1gpu.err; echo "$W N=1 exit $?"
  P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload $W $X > $O/cfg_${W}_${N}gpu.json 2> $O/cfg_${W}_${N}gpu.err; echo "$W N=$N exit $?"
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/cfg_*gpu.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('/')[-1], d['n_gpus'], '%.1f us' % (1e3 * d['ms_per_step']), '%.0f images/s' % d['value'], (d.get('sanity') or {}).get('mgpu', {}).get('bit_equal'))
    except Exception as e:
        print(f, 'FAILED', e)
P
