#!/bin/bash
# experiments on the two streaming kernels: SSDBOX_RING_GRID (CTAs), SSDBOX_RING_MAX_STAGES, loss flags.
#   bash tools/exp_ring_grid.sh "148 0 0" "112 4 0" ...      each argument = "grid max_stages loss_flags" (0 = default)
for cfg in "$@"; do
  set -- $cfg
  g=$1; ns=$2; fl=$3
  env=""
  [ "$g" != "0" ] && export SSDBOX_RING_GRID=$g || unset SSDBOX_RING_GRID
  [ "$ns" != "0" ] && export SSDBOX_RING_MAX_STAGES=$ns || unset SSDBOX_RING_MAX_STAGES
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-voc-eval --loss-flags $fl $EXTRA 2>/dev/null > /tmp/exp.json
  python - "$cfg" <<'PY'
import json, sys
d = json.loads(open("/tmp/exp.json").read().strip().splitlines()[-1])
k = d["phases"]["kernels_us"]
print("grid/stages/flags", sys.argv[1], "| step_us %.1f" % (d["ms_per_step"] * 1e3), "loss_stream %.1f" % k["loss_stream"], "match %.1f" % k.get("match", 0.0),
      "detect_stream %.1f" % k["detect_stream"], "mine %.1f" % k["mine_reduce"])
PY
done
