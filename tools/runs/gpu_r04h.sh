#!/bin/bash
O=gpurun_out; mkdir -p $O
./tools/micro/fill_patterns > $O/r04_fill_patterns3.txt 2>&1; tail -12 $O/r04_fill_patterns3.txt
python -m pytest tests -m gpu -x -q -k "head" 2>&1 | tail -3
{
python tools/run_heads.py ssd512_coco 10 | tail -2
export SSDBOX_EXP=1
SSDBOX_HEADS_STG=1 python tools/run_heads.py ssd512_coco 10 | tail -1
SSDBOX_HEADS_STAGES=2 python tools/run_heads.py ssd512_coco 10 | tail -1
python tools/run_heads.py refinedet320_voc 10 | tail -1
python tools/run_heads.py fssd300_coco 10 | tail -1
} 2>&1 | tee $O/r04h_heads.log
