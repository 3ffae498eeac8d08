#!/bin/bash
# experiment: grouped interleave (SSDBOX_RING_INTERLEAVE=g: g consecutive CTAs pool their tile runs); read the
# detect_stream column in the sparse and the dense regime
for g in 0 2 4 8 16 148; do
  if [ "$g" = "0" ]; then unset SSDBOX_RING_INTERLEAVE; else export SSDBOX_RING_INTERLEAVE=$g; fi
  echo "interleave group = $g"
  export EXTRA=""; bash tools/exp_ring_grid.sh "0 0 0"
  export EXTRA="--dense"; bash tools/exp_ring_grid.sh "0 0 0"
done
