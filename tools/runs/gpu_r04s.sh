#!/bin/bash
O=gpurun_out; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
X="--no-side-phases --no-cpu-baseline --steps 200 --workload refinedet320_voc"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29941 bench.py --gpus $N $X > $O/r04s_refine_${N}gpu.json 2> $O/r04s_refine_${N}gpu.err; echo "N=$N exit $?"
tail -c 900 $O/r04s_refine_${N}gpu.json
