#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "head" 2>&1 | tail -3
{
python tools/run_heads.py ssd512_coco 10 | tail -2
export SSDBOX_EXP=1
SSDBOX_HEADS_ALIGN=0 python tools/run_heads.py ssd512_coco 10 | tail -1
SSDBOX_HEADS_MODE=2 python tools/run_heads.py ssd512_coco 10 | tail -1
SSDBOX_HEADS_ALIGN=0 SSDBOX_HEADS_MODE=2 python tools/run_heads.py ssd512_coco 10 | tail -1
SSDBOX_HEADS_STAGES=2 python tools/run_heads.py ssd512_coco 10 | tail -1
} 2>&1 | tee $O/r04d_heads.log
