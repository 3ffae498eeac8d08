#!/bin/bash
O=gpurun_out
timeout 600 python bench.py > $O/r03c_bench.json 2> $O/r03c_bench.err; echo "bench exit $?"
X="--no-side-phases --no-cpu-baseline --e2e-steps 1 --steps 200"
for W in ssd300_voc fssd300_coco rfb300_voc refinedet320_voc; do
  timeout 200 python bench.py $X --workload $W > $O/r03c_${W}.json 2> $O/r03c_${W}.err; echo "$W exit $?"
done
timeout 200 python bench.py $X --workload refinedet320_voc --serial > $O/r03c_refinedet320_voc_serial.json 2> $O/r03c_refinedet320_voc_serial.err; echo "refine serial exit $?"
timeout 600 python -m pytest tests -m gpu -x -q > $O/r03c_pytest.log 2>&1; tail -2 $O/r03c_pytest.log
