#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "detect or nms or dense or refine or two_stream or graph" 2>&1 | tail -3
python tools/run_dense.py 2>&1 | tail -6 | tee $O/r04y_dense.log
