#!/bin/bash
# round-2 GPU call D: parity + timing after the packed-fp32x2 row scan and the fine-bin / gather-rank mining select
O=gpurun_out
TAG=${1:-r02d}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/${TAG}_pytest.log
X="--no-side-phases --no-cpu-baseline --e2e-steps 1 --steps 200"
for ONLY in T D; do
  timeout 200 python bench.py $X --only $ONLY > $O/${TAG}_${ONLY}.json 2> $O/${TAG}_${ONLY}.err; echo "$ONLY exit $?"
done
timeout 200 python bench.py $X > $O/${TAG}_TD.json 2> $O/${TAG}_TD.err; echo "TD exit $?"
for W in ssd300_voc fssd300_coco rfb300_voc refinedet320_voc; do
  timeout 200 python bench.py $X --workload $W > $O/${TAG}_${W}.json 2> $O/${TAG}_${W}.err; echo "$W exit $?"
done
timeout 200 python bench.py $X --workload ssd300_voc --only T > $O/${TAG}_ssd300_voc_T.json 2> $O/${TAG}_ssd300_voc_T.err; echo "ssd300 T exit $?"
