#!/bin/bash
# One gpurun call: full bench line, reference arm, ncu launch list and one ncu --set full capture.
#   gpurun --timeout 1500 -- 'bash tools/profile_run.sh r01b'
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "reference exit $?"
# the step only (the side phases -- backward, fused softmax, eval post, head layout, VOC eval, dense Detect -- would mix
# their launches of the same kernels into the per-kernel means)
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-graph --no-side-phases"
$CMD > $O/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${TAG}_ncu_launches.csv $CMD > $O/${TAG}_ncu_list.log 2>&1
echo "ncu list exit $?"
$CMD > $O/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:loss_stream|loss_bwd_stream|detect_stream|mine_reduce|detect_segment|detect_overflow' -s 10 -c 14 -f -o $O/${TAG}_prof $CMD > $O/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la $O | tail -12
