#!/bin/bash
# experiment: interleaved tile assignment (SSDBOX_RING_INTERLEAVE) x grid size; read the detect_stream column
export SSDBOX_RING_INTERLEAVE=1
bash tools/exp_ring_grid.sh "0 0 0" "140 0 0" "132 0 0" "124 0 0" "0 4 0" "132 4 0"
export EXTRA="--dense"; bash tools/exp_ring_grid.sh "0 0 0"
export EXTRA="--workload rfb300_voc"; bash tools/exp_ring_grid.sh "0 0 0" "132 0 0"
