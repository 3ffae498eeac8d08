#!/bin/bash
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/final_bench_reference.json 2> $O/final_bench_reference.err; echo "reference exit $?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/final_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['phases']['step_hbm_frac'], d['gpu_launches'], d['e2e']['value'], d['clocks'])
print('bwd', d['phases']['train_bwd']['kernel_us_sum'], 'heads', d['phases']['head_layout']['us'], 'dense', d['phases']['dense']['us'])
r=json.loads([l for l in open('gpurun_out/final_bench_reference.json') if l.startswith('{')][-1])
print('reference', r['value'], r['unit'], r['cpu_baseline']['kind'])
P
