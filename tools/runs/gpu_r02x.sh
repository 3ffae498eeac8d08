#!/bin/bash
O=gpurun_out
TAG=${1:-r02x}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 $O/${TAG}_pytest.log
X="--no-side-phases --no-cpu-baseline --e2e-steps 2 --steps 200"
timeout 200 python bench.py $X > $O/${TAG}_TD.json 2> $O/${TAG}_TD.err; echo "TD exit $?"
for W in ssd300_voc refinedet320_voc; do
  timeout 200 python bench.py $X --workload $W > $O/${TAG}_${W}.json 2> $O/${TAG}_${W}.err; echo "$W exit $?"
done
