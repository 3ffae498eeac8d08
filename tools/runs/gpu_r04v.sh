#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 800 python -m pytest tests -m gpu -x -q -k "any_top_k or large_path or nms or detect" 2>&1 | tail -6
