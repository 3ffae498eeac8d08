#!/bin/bash
O=gpurun_out; mkdir -p $O
X="--no-side-phases --no-cpu-baseline --steps 200"
for W in ssd300_voc ssd512_coco rfb300_voc fssd300_coco refinedet320_voc; do
  timeout 200 python bench.py --workload $W $X > $O/r04z_cfg_${W}_1gpu.json 2> $O/r04z_cfg_${W}_1gpu.err; echo "$W exit $?"
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r04z_cfg_*_1gpu.json')):
    d = json.loads([l for l in open(f) if l.startswith('{')][-1])
    print(f.split('/')[-1], '%.1f us' % (1e3 * d['ms_per_step']), '%.0f images/s' % d['value'], {k: round(v, 1) for k, v in d['phases']['kernels_us'].items()})
P
