#!/bin/bash
O=gpurun_out; mkdir -p $O
export SSDBOX_EXP=1
{
for i in 1 2; do
python tools/run_heads.py ssd512_coco 10 | tail -2
SSDBOX_HEADS_WARPS=16 python tools/run_heads.py ssd512_coco 10 | tail -2
done
SSDBOX_HEADS_WARPS=16 python tools/run_heads.py refinedet320_voc 10 | tail -2
SSDBOX_HEADS_WARPS=16 python tools/run_heads.py fssd300_coco 10 | tail -2
} 2>&1 | tee $O/r04t_heads.log
