#!/bin/bash
# gpurun --gpus 2 (or more): the whole GPU suite incl. tests/test_multi_gpu.py, then the bench at N = all GPUs
O=gpurun_out; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/mgpu_check_pytest_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus $N --no-side-phases --no-cpu-baseline --steps 200 > $O/mgpu_check_bench_${N}gpu.json 2> $O/mgpu_check_bench_${N}gpu.err; echo "bench $N exit $?"
tail -c 1500 $O/mgpu_check_bench_${N}gpu.json
