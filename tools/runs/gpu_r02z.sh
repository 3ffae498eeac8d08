#!/bin/bash
O=gpurun_out
TAG=${1:-r02z}
timeout 600 python -m pytest tests -m gpu -x -q -k "dense or density or fused_softmax or uniform or clean or graph_replay or refinedet_fused" > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $O/${TAG}_pytest.log
X="--no-side-phases --no-cpu-baseline --e2e-steps 1 --steps 100"
timeout 300 python bench.py $X --dense --only D > $O/${TAG}_dense.json 2> $O/${TAG}_dense.err; echo "dense exit $?"
