#!/bin/bash
O=gpurun_out
X="--no-side-phases --no-cpu-baseline --steps 200"
timeout 300 python bench.py $X > $O/r03d_bench_1gpu.json 2> $O/r03d_bench_1gpu.err; echo "bench 1 exit $?"
P=29900
for N in 8 4 2; do
  P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N $X > $O/r03d_bench_${N}gpu.json 2> $O/r03d_bench_${N}gpu.err; echo "bench $N exit $?"
done
