"""Turns one tools/profile_run.sh capture (gpurun_out/<tag>_*) into the committed evidence under
profiles/: bench lines, ncu launch list, raw ncu CSV of the --set full capture, a short summary
and profiles/traffic.json (dram bytes per launch of the two streaming kernels, read by bench.py).

    python tools/summarize_profile.py r01b
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
Pdir = os.path.join(ROOT, "profiles")
os.makedirs(Pdir, exist_ok=True)

for suffix in ("bench.json", "bench_reference.json", "ncu_launches.csv", "smi.txt"):
    src = os.path.join(G, "%s_%s" % (tag, suffix))
    if os.path.isfile(src):
        shutil.copy(src, os.path.join(Pdir, "%s_%s" % (tag, suffix)))

rep = os.path.join(G, tag + "_prof.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(Pdir, tag + "_ncu_full_raw.csv"), "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct")]


def tobytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


per = collections.OrderedDict()
for r in data:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    d = per.setdefault(name, collections.defaultdict(list))
    for m, short in want:
        if m in ix:
            v, u = r[ix[m]], units[ix[m]]
            try:
                d[short].append(tobytes(v, u) if short in ("dram_rd", "dram_wr") else float(v.replace(",", "")))
            except ValueError:
                pass

lines = ["# %s: ncu --set full --clock-control none (per launch, mean over the captured launches)" % tag, "",
         "| kernel | launches | time us | dram read MB | dram write MB | dram % of ncu peak | warps active % | issue active % | fma pipe % | regs | grid x block |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
mean = lambda v: sum(v) / len(v) if v else float("nan")
for name, d in per.items():
    lines.append("| %s | %d | %.1f | %.2f | %.2f | %.1f | %.1f | %.1f | %.1f | %d | %d x %d |" % (
        name, len(d["time"]), mean(d["time"]), mean(d["dram_rd"]) / 1e6, mean(d["dram_wr"]) / 1e6, mean(d["dram_pct"]),
        mean(d["warps_active_pct"]), mean(d["issue_active_pct"]), mean(d["fma_pipe_pct"]), mean(d["regs"]), mean(d["grid"]), mean(d["block"])))
    for key in ("loss_stream", "detect_stream"):
        if key in name:
            traffic[key] = mean(d["dram_rd"]) + mean(d["dram_wr"])
traffic["source"] = "profiles/%s_ncu_full_raw.csv (dram__bytes_read.sum + dram__bytes_write.sum per launch)" % tag
json.dump(traffic, open(os.path.join(Pdir, "traffic.json"), "w"), indent=1)

# launch list: share of the step per kernel
ll = os.path.join(G, tag + "_ncu_launches.csv")
if os.path.isfile(ll):
    agg = collections.OrderedDict()
    for r in csv.reader(open(ll)):
        if len(r) > 14 and r[0].isdigit():
            nm = r[4].split("(")[0].replace("void ", "")
            agg.setdefault(nm, []).append(float(r[-1]) / 1e3)
    ours = {k: v for k, v in agg.items() if k.startswith("ssdbox::")}
    fwd = [k for k in ours if not any(s in k for s in ("zero_fill", "loss_bwd", "priorbox", ", 1>", "compact_", "peer_finish", "heads_to_rows", "voc_", "radix_", "crop_", "init_kernel"))]
    tot = sum(mean(ours[k]) for k in fwd)      # (init_kernel runs once per workspace: not part of a step any more)
    lines += ["", "## launch list (ncu --metrics gpu__time_duration.sum, cold cache, serialised): share of one step (T fwd + D)", "",
              "| kernel | launches | mean us | share of step |", "|---|---|---|---|"]
    for k in ours:
        m = mean(ours[k])
        mult = 1
        share = "%.1f%%" % (100 * m * mult / tot) if k in fwd else "(first-call init / backward / setup / side phases of bench.py, not in the step)"
        lines.append("| %s%s | %d | %.1f | %s |" % (k, " (x2 per step)" if mult == 2 else "", len(ours[k]), m, share))
    lines.append("| sum over one step | | %.1f | |" % tot)

bj = os.path.join(G, tag + "_bench.json")
if os.path.isfile(bj):
    b = json.load(open(bj))
    lines += ["", "## bench line (CUDA events, graph replay)", "",
              "value %.0f images/s, %.1f us/step; dominant kernel %s %.1f us = %.0f GB/s = %.3f of measured peak %.1f GB/s; e2e %.0f images/s; cpu_baseline %s" % (
                  b["value"], 1e3 * b["ms_per_step"], b["roofline"]["kernel"], b["roofline"]["avg_launch_us"], b["roofline"]["achieved"],
                  b["roofline"]["frac"], b["roofline"]["peak"], b["e2e"]["value"], json.dumps(b.get("cpu_baseline"))),
              "", "kernels_us (CUDA events inside the library, eager): " + json.dumps(b["phases"]["kernels_us"])]
open(os.path.join(Pdir, tag + "_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
