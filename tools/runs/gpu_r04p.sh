#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > $O/r04p_bench.json 2> $O/r04p_bench.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r04p_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['phases']['step_hbm_frac'])
print('bwd', d['phases']['train_bwd']['kernel_us_sum'], d['phases']['train_bwd']['hbm_frac'])
print('heads', d['phases']['head_layout'])
P
