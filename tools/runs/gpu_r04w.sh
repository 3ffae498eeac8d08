#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 800 python -m pytest tests -m gpu -x -q -k "global_max or lse_global or multibox_loss or refine" 2>&1 | tail -8
