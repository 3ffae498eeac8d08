#!/bin/bash
O=gpurun_out
mkdir -p $O
X="--no-cpu-baseline --e2e-steps 1 --steps 100 --no-voc-eval"
for F in 0 32; do
timeout 400 python bench.py $X --loss-flags $F > $O/r02h_bench_f$F.json 2> $O/r02h_bench_f$F.err; echo "bench flags $F exit $?"
done
for W in ssd300_voc fssd300_coco rfb300_voc; do
timeout 400 python bench.py $X --workload $W > $O/r02h_bench_$W.json 2> $O/r02h_bench_$W.err; echo "bench $W exit $?"
done
