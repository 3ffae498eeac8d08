#!/bin/bash
O=gpurun_out
TAG=${1:-r02n}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/${TAG}_pytest.log
X="--no-side-phases --no-cpu-baseline --e2e-steps 1 --steps 100"
timeout 300 python bench.py $X --dense > $O/${TAG}_dense.json 2> $O/${TAG}_dense.err; echo "dense exit $?"
timeout 300 python bench.py $X --only D > $O/${TAG}_D.json 2> $O/${TAG}_D.err; echo "D exit $?"
timeout 300 python bench.py $X > $O/${TAG}_TD.json 2> $O/${TAG}_TD.err; echo "TD exit $?"
for W in ssd300_voc rfb300_voc; do
timeout 300 python bench.py $X --dense --only D --workload $W > $O/${TAG}_dense_$W.json 2> $O/${TAG}_dense_$W.err; echo "dense $W exit $?"
done
