"""Prints the SASS instructions with the most warp-stall samples for one kernel of an .ncu-rep
(ncu -i REP --page source --csv --kernel-name regex:NAME) with a little context."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several launches may be present: keep the first block
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
    elif cur is not None:
        cur.append(r)
blk = blocks[0]
hdr = blk[0]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in blk[1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("kernel", kern, "instructions", len(data), "samples", tot)
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    print("%5d  %6s smp %5.1f%%  exec %-9s  %s" % (i, r[ix["# Samples"]], 100.0 * int(r[ix["# Samples"]] or 0) / max(tot, 1),
                                                  r[ix["Instructions Executed"]], r[ix["Source"]].strip()))
