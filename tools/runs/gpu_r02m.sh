#!/bin/bash
# round-2 8-GPU call: N>1 parity test at 8 ranks, bench lines at 8 (NUMA-bound / unbound) and 4 GPUs with sanity.mgpu and e2e probes
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r02m_topo.txt 2>&1
lscpu | head -30 > $O/r02m_lscpu.txt 2>&1
cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c > $O/r02m_numa_nodes.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "sharded" > $O/r02m_pytest_8gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02m_pytest_8gpu.log
X="--no-side-phases --no-cpu-baseline --steps 200"
P=29700
for N in 8 4; do
  P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N $X > $O/r02m_bench_${N}gpu.json 2> $O/r02m_bench_${N}gpu.err; echo "bench $N exit $?"
done
P=$((P+1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 $X --no-numa-bind > $O/r02m_bench_8gpu_nobind.json 2> $O/r02m_bench_8gpu_nobind.err; echo "bench 8 nobind exit $?"
timeout 300 python bench.py $X > $O/r02m_bench_1gpu.json 2> $O/r02m_bench_1gpu.err; echo "bench 1 exit $?"
