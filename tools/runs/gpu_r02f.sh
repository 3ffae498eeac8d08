#!/bin/bash
# round-2 multi-GPU call (gpurun --gpus N): N>1 parity test + bench lines at 1 and N GPUs with sanity.mgpu
O=gpurun_out
N=${1:-2}
TAG=${2:-r02f}
mkdir -p $O
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_${N}gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/${TAG}_pytest_${N}gpu.log
X="--no-side-phases --no-cpu-baseline --steps 200"
timeout 300 python bench.py $X > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench_1gpu.err; echo "bench 1 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N $X > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err; echo "bench $N exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N $X --no-numa-bind > $O/${TAG}_bench_${N}gpu_nobind.json 2> $O/${TAG}_bench_${N}gpu_nobind.err; echo "bench $N nobind exit $?"
tail -2 $O/${TAG}_bench_${N}gpu.err
