#!/bin/bash
O=gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests -m gpu -x -q -k "two_stream or sharded or deferred" > $O/r03b_pytest_${N}gpu.log 2>&1; echo "pytest exit $?"; tail -2 $O/r03b_pytest_${N}gpu.log
X="--no-side-phases --no-cpu-baseline --steps 200"
timeout 300 python bench.py $X > $O/r03b_bench_1gpu.json 2> $O/r03b_bench_1gpu.err; echo "bench 1 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus $N $X > $O/r03b_bench_${N}gpu.json 2> $O/r03b_bench_${N}gpu.err; echo "bench $N exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus $N $X --serial > $O/r03b_bench_${N}gpu_serial.json 2> $O/r03b_bench_${N}gpu_serial.err; echo "bench $N serial exit $?"
