#!/usr/bin/env python
"""bench.py -- SSD512-COCO box hot path on B200: match + MultiBoxLoss forward AND Detect/NMS.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ssdbox|reference]

One "step" = one pass of the hot path over one batch of synthetic input of BASELINE.json
configs[1] (SSD512 VGG16 COCO: P=24564 priors, C=81 classes, B=64 images PER GPU):
    T  match + MultiBoxLoss forward  (loss_stream with the matching on dedicated warps, mine_reduce)
    D  DetectOut: threshold, top-200, NMS (detect_stream, detect_segments, overflow select + rewritten lists)
`value` = images/s with inputs resident in HBM (whole job: N * B / max-over-ranks step time; every
image passes through both T and D), timed with CUDA events around K CUDA-graph replays.  T and D are independent
ops on the same batch: by default D is submitted to a side stream first (ssdbox.TwoStreamStep), so that its
latency-bound tail kernels run under T's HBM-bound conf pass; `--serial` runs the six kernels back to back.
`e2e` = same metric through the public modules (MultiBoxLoss.forward / DetectOut.__call__) with
pinned HOST inputs: H2D of loc/conf/targets/scores and D2H of the losses and the detection tensor
inside the timed region.  `--impl reference` times the reference's own CPU path on the host cores: the
unmodified copy of its `lib/` package that oracle/build_ref.py places under oracle/_ref (git-ignored, travels
to the GPU box), or the in-repo oracle port when that copy is absent.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "object-detection-pytorch_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "SSD512 COCO images/s: match+MultiBoxLoss & Detect/NMS, 1/2/4/8 B200"
WORKLOAD = "ssd512_coco"
VAR = [0.1, 0.2]
# kernels launched per step (the init launches are gone from the second call on: self-cleaning workspaces).
# plain: T = loss_stream, mine_reduce; D = detect_stream, segments, overflow select, rewritten lists.
# refine: ARM loss (2) + ODM loss (2) + RefineDetectOut (4): the ARM outputs are consumed inside the kernels (fused)
LAUNCHES = {"plain": 6, "refine": 8}
# BASELINE.json configs; `--workload` selects one (the headline metric is quoted on ssd512_coco)
DESCR = {
    "ssd300_voc": "SSD300 VGG16 VOC",
    "ssd512_coco": "SSD512 VGG16 COCO",
    "rfb300_voc": "RFBNet300 VGG16 VOC",
    "fssd300_coco": "FSSD300 VGG COCO",
    "refinedet320_voc": "RefineDet320 VOC two-step",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ssdbox", choices=["ssdbox", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: the config's batch)")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--loss-flags", type=int, default=0, help="experiments: ssdbox_loss_cfg.flags (1 = matching as its own kernel)")
    ap.add_argument("--no-voc-eval", action="store_true", help="skip the VOC evaluation side phase")
    ap.add_argument("--dense", action="store_true", help="detect scores with background bias 4 (worst case)")
    ap.add_argument("--only", default="", choices=["", "T", "D"], help="experiments: time only the loss half (T) or the Detect half (D) of the step")
    ap.add_argument("--serial", action="store_true", help="run T and D back to back on ONE stream (default: D on a second stream, its tail overlaps T's conf pass)")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: do not pin the rank to the CPUs / memory next to its GPU")
    ap.add_argument("--no-side-phases", action="store_true", help="skip backward / fused softmax / eval post / head layout / VOC eval / dense phases")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# bytes model (SURVEY.md section 8d / BASELINE.md section 3), per image
# ----------------------------------------------------------------------------------------------
def bytes_T(P, C, G, B):
    return P * (4 * C + 16) + 20 * G + 16.0 * P / B


def bytes_D(P, C, top_k):
    return P * (4 * C + 16) + 20 * C * top_k


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """dram bytes per launch of the dominant kernels from the committed ncu capture (or None)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU side (oracle port of the reference algorithm) -- the only place bench.py executes oracle/
# ----------------------------------------------------------------------------------------------
def cpu_impl(name):
    """(kind, loss_fn, detect_fn): the reference's OWN functions (box_utils.match / MultiBoxLoss.forward /
    DetectOut.forward / nms, unmodified copy under oracle/_ref placed there by oracle/build_ref.py, run through
    oracle/ref_loader.py's torch >= 1.x shims) when they are present, else the in-repo restatement
    (oracle/ssd_oracle.py, proved bit-identical to them by tests/test_oracle_vs_reference.py).  RefineDet has
    no reference code: always the restatement."""
    from oracle import ssd_oracle as O
    if name != "refinedet320_voc":
        try:
            from oracle import ref_loader as RL
            if RL.use_vendored():
                R = RL.load()
                return ("reference",
                        lambda loc, conf, pri, tg, C: R.multibox_loss(C, (loc, conf, pri), tg),
                        lambda loc, sc, pri, C: R.detect(C, loc, sc, pri))
        except Exception as e:      # pragma: no cover - e.g. a missing import of the reference's chain
            sys.stderr.write("bench: reference copy under oracle/_ref unusable (%r); timing the oracle port\n" % (e,))
    return ("port", lambda loc, conf, pri, tg, C: O.multibox_loss(loc, conf, pri, tg, C),
            lambda loc, sc, pri, C: O.detect(loc, sc, pri, C))


def cpu_reference_pass(name, C, P, priors_cpu, bt, bd, seed, dense, impl=None):
    """times ONE pass: T on `bt` images + D on `bd` images; returns (t_T, t_D) seconds."""
    from oracle import ssd_oracle as O
    from ssdbox import configs, synth
    _, c = configs.get(name)
    kind, loss_fn, det_fn = impl or cpu_impl(name)
    tg = synth.gen_targets(bt, C, c["gt_max"], seed)
    loc = synth.gen_loc(max(bt, bd), P, seed)
    conf = synth.gen_train_logits(bt, P, C, seed)
    sc = synth.gen_detect_scores(bd, P, C, seed, bkg_bias=4.0 if dense else 10.0)
    if name == "refinedet320_voc":          # two-step path: ARM loss + ODM loss + refined Detect (own restatement)
        arm_loc, arm_conf = synth.gen_arm_outputs(max(bt, bd), P, seed)
        t0 = time.perf_counter()
        O.refine_multibox_loss(arm_loc[:bt], arm_conf[:bt], loc[:bt], conf, priors_cpu, tg, 2, use_arm=False)
        O.refine_multibox_loss(arm_loc[:bt], arm_conf[:bt], loc[:bt], conf, priors_cpu, tg, C, use_arm=True)
        t1 = time.perf_counter()
        O.refine_detect(arm_loc[:bd], arm_conf[:bd], loc[:bd], sc, priors_cpu, C)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1
    t0 = time.perf_counter()
    loss_fn(loc[:bt], conf, priors_cpu, tg, C)
    t1 = time.perf_counter()
    det_fn(loc[:bd], sc, priors_cpu, C)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def run_reference(args):
    """--impl reference: the reference algorithm on the host cores (oracle port, all threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import ssd_oracle as O
    from ssdbox import configs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, c = configs.get(args.workload)
    C = cfg.MODEL.NUM_CLASSES
    pri = O.prior_boxes(cfg.MODEL, c["layer_dims"])
    P = pri.size(0)
    bt, bd = 4, 2
    impl = cpu_impl(args.workload)
    for i in range(args.warmup):
        cpu_reference_pass(args.workload, C, P, pri, bt, bd, i, args.dense, impl)
    tt = td = 0.0
    t_start = time.perf_counter()
    for i in range(args.steps):
        a, b = cpu_reference_pass(args.workload, C, P, pri, bt, bd, 100 + i, args.dense, impl)
        tt += a
        td += b
    wall = time.perf_counter() - t_start
    per_img = tt / (args.steps * bt) + td / (args.steps * bd)
    value = 1.0 / per_img
    what = ("the reference's own box_utils.match + MultiBoxLoss.forward / DetectOut.forward + nms (unmodified copy in oracle/_ref)"
            if impl[0] == "reference" else "the oracle port of the reference algorithm (oracle/ssd_oracle.py)")
    sample = "per step: match+MultiBoxLoss fwd on %d images + Detect on %d images (%s scores), %s, torch CPU, %d threads" % (
        bt, bd, "dense" if args.dense else "sparse", what, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (tt + td) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %s box path, P=%d C=%d, CPU sample of the B=%d step" % (args.workload, DESCR.get(args.workload, args.workload), P, C, c["batch"]),
                   "inputs": "synthetic seeded (ssdbox.synth), same generator as the GPU arm"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": impl[0], "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "phases": {"train_fwd_images_per_s": args.steps * bt / tt, "detect_images_per_s": args.steps * bd / td},
        "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


def side_phases(args, ssdbox, _abi, synth, crit, det, det_out, cfg, c, loc, conf, sc, priors, gt, offs, gmax, B, P, C, top_k, dev, n_gpus):
    """phases reported beside the headline (not part of the step): forward + backward, DetectOut on raw logits,
    eval post-processing, head-output layout, VOC evaluation."""
    # forward + backward through the autograd module
    loc_g = loc.clone().requires_grad_(True)
    conf_g = conf.clone().requires_grad_(True)
    for i in range(4):
        if i == 1:
            _abi.timers_enable(True)
        a_, b_ = crit.forward_packed(loc_g, conf_g, priors, gt, offs, gmax)
        (a_ + b_).backward()
        loc_g.grad = None
        conf_g.grad = None
    torch.cuda.synchronize()
    kb = _abi.timers_read()
    _abi.timers_enable(False)
    bwd_us = 1e3 * kb["loss_bwd"][0] / max(kb["loss_bwd"][1], 1)
    del loc_g, conf_g

    # SURVEY 8f rank 2 (reported beside the headline): DetectOut on raw logits with the softmax fused
    # into the candidate pass, against torch.softmax + the plain DetectOut of the step
    det_lg = ssdbox.DetectOut(C, 0, top_k, 0.01, 0.45, VAR, conf_is_logits=True)
    lg = torch.log(sc.clamp_min(1e-30))          # logits whose softmax is the step's score tensor
    with torch.no_grad():
        det_lg.forward(loc, lg, priors, out=det_out)
        torch.cuda.synchronize()
        _abi.timers_enable(True)
        for _ in range(5):
            det_lg.forward(loc, lg, priors, out=det_out)
        torch.cuda.synchronize()
        kl = _abi.timers_read()
        _abi.timers_enable(False)
        fused_us = sum(1e3 * kl[k][0] / kl[k][1] for k in ("init", "detect_stream", "detect_segment", "detect_segment_big", "detect_overflow") if kl[k][1])
        fused_stream_us = 1e3 * kl["detect_stream"][0] / max(kl["detect_stream"][1], 1)
        fused_dets = int((det_out[..., 0] > 0).sum())
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        sm_out = torch.empty_like(lg)
        torch.softmax(lg, -1, out=sm_out)
        ev[0].record()
        for _ in range(5):
            torch.softmax(lg, -1, out=sm_out)
        ev[1].record()
        torch.cuda.synchronize()
        softmax_us = 1e3 * ev[0].elapsed_time(ev[1]) / 5
    del lg, sm_out

    # SURVEY 8f rank 1 (reported beside the headline): detections -> flat [n,7] result rows
    from ssdbox import evaluate_utils as EU
    extra_hw = torch.tensor([[480.0, 640.0]] * B, device=dev)
    img_ids = torch.arange(B, dtype=torch.float32, device=dev)
    with torch.no_grad():
        det.forward(loc, sc, priors, out=det_out)
        EU.coco_result_rows(det_out, extra_hw, img_ids, sync=False)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(10):
            rows_buf, rows_total, _ = EU.coco_result_rows(det_out, extra_hw, img_ids, sync=False)
        ev[1].record()
        torch.cuda.synchronize()
        evalpost_us = 1e3 * ev[0].elapsed_time(ev[1]) / 10
        evalpost_rows = int(rows_total)
    del rows_buf

    # SURVEY 8f rank 3 (reported beside the headline): multibox head outputs (NCHW) -> conf [B,P,C]
    from ssdbox import heads as HD
    per_cell = ssdbox.PriorBoxSSD(cfg).num_priors     # anchors per cell of every source layer (prior_box.py:46-50)
    with torch.no_grad():
        head_outs = [torch.randn(B, a_ * C, h_, w_, device=dev) for a_, (h_, w_) in zip(per_cell, c["layer_dims"])]
        rows_out = torch.empty(B, P, C, device=dev)
        HD.heads_to_rows(head_outs, C, out=rows_out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ref_rows = torch.empty(B, P * C, device=dev)
        ref_rows.zero_()          # (not timed) keeps the GPU busy while the first launches are queued: the region times kernels, not the host
        ev[0].record()
        for _ in range(5):
            HD.heads_to_rows(head_outs, C, out=rows_out)
        ev[1].record()
        ref_rows = torch.cat([o.permute(0, 2, 3, 1).contiguous().view(B, -1) for o in head_outs], 1)
        ev[2].record()
        for _ in range(3):
            ref_rows = torch.cat([o.permute(0, 2, 3, 1).contiguous().view(B, -1) for o in head_outs], 1)
        ev[3].record()
        torch.cuda.synchronize()
        heads_us = 1e3 * ev[0].elapsed_time(ev[1]) / 5
        heads_torch_us = 1e3 * ev[2].elapsed_time(ev[3]) / 3
        heads_equal = bool(torch.equal(ref_rows.view(B, P, C), rows_out))
    del head_outs, rows_out, ref_rows

    # SURVEY 8f rank 4 (reported beside the headline): PASCAL VOC evaluation of an accumulated result set
    voc_phase = None
    if n_gpus == 1 and not args.no_voc_eval:      # single-GPU side phase: a rank-asymmetric pause would trip the peers' bounded wait
        from ssdbox import voc_eval as VE
        n_img = 4952                                   # VOC2007 test
        case = synth.gen_voc_eval_case(n_img, 21, 11, fp_max=12)
        vgt = VE.VOCGroundTruth(case["gt_boxes"], case["gt_labels"], case["gt_difficult"], case["gt_offsets"], dev)
        vrows, vseg = torch.as_tensor(case["rows"]).to(dev), torch.as_tensor(case["seg"]).to(dev)
        for _ in range(3):                              # warm-up: workspace, allocator blocks of these sizes
            res = VE.voc_eval(vrows, vseg, vgt, 21)
        torch.cuda.synchronize()
        voc_us = []
        for _ in range(9):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            res = VE.voc_eval(vrows, vseg, vgt, 21)
            ev[1].record()
            torch.cuda.synchronize()
            voc_us.append(1e3 * ev[0].elapsed_time(ev[1]))
        voc_us.sort()
        voc_phase = {"note": "ssdbox_voc_eval: write_voc_results_file + voc_eval + voc_ap (lib/datasets/voc_eval.py:58-242) for all 20 classes of a "
                             "synthetic VOC2007-test sized result set; median per call incl. the host mirror's allocations and its D2H read",
                     "images": n_img, "detections": int(vrows.size(0)), "truths": int(case["gt_boxes"].shape[0]),
                     "us": voc_us[len(voc_us) // 2], "us_min_max": [voc_us[0], voc_us[-1]], "calls": len(voc_us), "mean_ap": res.mean_ap}
        if not args.no_cpu_baseline:
            from oracle import voc_oracle as _V    # CPU leg only: the reference algorithm's port timed beside the GPU call
            sub = synth.gen_voc_eval_case(300, 21, 12, fp_max=12)
            t0 = time.perf_counter()
            _V.voc_eval_rows(sub["rows"], sub["seg"], 300, 21, sub["gt_boxes"], sub["gt_labels"], sub["gt_difficult"], sub["gt_offsets"])
            dt = time.perf_counter() - t0
            voc_phase["cpu_port"] = {"sample": "300 images, %d detections, numpy oracle, 1 thread" % sub["rows"].shape[0],
                                     "detections_per_s": sub["rows"].shape[0] / dt,
                                     "extrapolated_s_for_this_set": int(vrows.size(0)) * dt / sub["rows"].shape[0]}
        del vrows, vseg, vgt, res

    return (bwd_us, fused_us, fused_stream_us, fused_dets, softmax_us, evalpost_us, evalpost_rows, heads_us, heads_torch_us,
            heads_equal, voc_phase)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def main():
    T0 = time.perf_counter()
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import ssdbox
    from ssdbox import _abi, configs, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the box path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from ssdbox import dist as sdist
    numa = sdist.bind_to_local_numa(local_rank) if (world > 1 and not args.no_numa_bind) else {"bound": False, "note": "single process: not bound"}
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL announces its version on stdout at communicator creation: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    n_gpus = world

    def log(msg):
        if rank == 0:
            sys.stderr.write("bench[%.1fs]: %s\n" % (time.perf_counter() - T0, msg))
            sys.stderr.flush()

    cfg, c = configs.get(args.workload)
    C = cfg.MODEL.NUM_CLASSES
    B = args.batch or c["batch"]
    top_k = 200
    priors = ssdbox.PriorBoxSSD(cfg).forward(c["layer_dims"], keep_on_device=True)
    P = priors.size(0)

    # ---- synthetic inputs: drawn on the CPU (seeded per rank), kept in pinned memory for e2e ----
    seed = 1000 * rank
    tg = synth.gen_targets(B, C, c["gt_max"], seed)
    gt_h, offs_h = synth.pack_targets(tg)
    gmax = max(int(t.size(0)) for t in tg)
    g_avg = float(gt_h.size(0)) / B
    loc_h = synth.gen_loc(B, P, seed).pin_memory()
    conf_h = synth.gen_train_logits(B, P, C, seed).pin_memory()
    sc_h = synth.gen_detect_scores(B, P, C, seed, bkg_bias=4.0 if args.dense else 10.0).pin_memory()
    gt_h, offs_h = gt_h.pin_memory(), offs_h.pin_memory()
    loc, conf, sc = loc_h.to(dev), conf_h.to(dev), sc_h.to(dev)
    gt, offs = gt_h.to(dev), offs_h.to(dev)

    refine = args.workload == "refinedet320_voc"
    crit = ssdbox.MultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, distributed=(world > 1))
    crit.abi_flags |= args.loss_flags
    det = ssdbox.DetectOut(C, 0, top_k, 0.01, 0.45, VAR)
    det_out = torch.empty(B, C, top_k, 5, dtype=torch.float32, device=dev)
    if refine:
        # cfg5 is measured as RefineDet: ARM loss (C=2, raw priors) + ODM loss (per-image refined anchors,
        # negative-anchor filtering) + RefineDetectOut -- SURVEY.md 8a-R, own restatement (parity unpinned)
        arm_loc_h, arm_conf_h = synth.gen_arm_outputs(B, P, seed)
        arm_loc_h, arm_conf_h = arm_loc_h.pin_memory(), arm_conf_h.pin_memory()
        arm_loc, arm_conf = arm_loc_h.to(dev), arm_conf_h.to(dev)
        arm_crit = ssdbox.RefineMultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, use_ARM=False, distributed=(world > 1))
        odm_crit = ssdbox.RefineMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, use_ARM=True, distributed=(world > 1))
        rdet = ssdbox.RefineDetectOut(C, 0, top_k, 0.01, 0.45, VAR, theta=0.01)
        crit = odm_crit

        three_streams = ssdbox.MultiStreamStep(3, dev)

        def step():
            with torch.no_grad():
                if args.serial:
                    al, ac = arm_crit.forward_packed_two_step(arm_loc, arm_conf, loc, conf, priors, gt, offs, gmax)
                    ol, oc = odm_crit.forward_packed_two_step(arm_loc, arm_conf, loc, conf, priors, gt, offs, gmax)
                    out = rdet.forward(arm_loc, arm_conf, loc, sc, priors, out=det_out)
                else:      # the three ops are independent: RefineDetectOut and the ODM loss on side streams, the ARM loss here
                    (al, ac), out, (ol, oc) = three_streams([
                        lambda: arm_crit.forward_packed_two_step(arm_loc, arm_conf, loc, conf, priors, gt, offs, gmax),
                        lambda: rdet.forward(arm_loc, arm_conf, loc, sc, priors, out=det_out),
                        lambda: odm_crit.forward_packed_two_step(arm_loc, arm_conf, loc, conf, priors, gt, offs, gmax)])
            return al + ol, ac + oc, out
    else:
        zero = torch.zeros((), device=dev)
        two_streams = ssdbox.TwoStreamStep(dev)

        def step():
            with torch.no_grad():
                if args.only == "D":
                    return zero, zero, det.forward(loc, sc, priors, out=det_out)
                if args.only == "T":
                    return crit.forward_packed_deferred(loc, conf, priors, gt, offs, gmax).wait() + (det_out,)
                if args.serial:
                    # T, then D on one stream; N > 1: the wait for the other ranks' sums rides on a Detect kernel
                    pending = crit.forward_packed_deferred(loc, conf, priors, gt, offs, gmax)
                    out = det.forward(loc, sc, priors, out=det_out, pending=pending)
                    ll, lc = pending.wait()
                    return ll, lc, out
                # T and D are independent ops on the same batch: D goes to a second stream (submitted first).  Its
                # HBM-bound candidate pass and T's conf pass still run one after the other (each fills every SM), but
                # D's latency-bound tail (segments + the overflow kernels, ~11 us) now runs UNDER T's conf pass.
                # (N > 1: the mining kernel is the last kernel of the step here, so its last CTA posts AND collects the
                # ranks' sums itself -- nothing is left to overlap a deferred wait with)
                (ll, lc), out = two_streams(lambda: crit.forward_packed(loc, conf, priors, gt, offs, gmax),
                                            lambda: det.forward(loc, sc, priors, out=det_out))
            return ll, lc, out

    log("inputs ready")
    # eager warm-up (lazy init, workspace allocation), also the functional sanity of the step
    for _ in range(2):
        ll, lc, out = step()
    torch.cuda.synchronize()
    sanity = {"loss_l": float(ll), "loss_c": float(lc), "detections": int((out[..., 0] > 0).sum()), "mgpu": None}

    # ---- per-kernel device times (eager, CUDA events inside the library on the launch stream) ---
    # measured with the ops back to back on ONE stream: in the two-stream step the kernels of T and D overlap, and
    # an event pair around one of them would also time its neighbours
    was_serial = args.serial
    args.serial = True
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    _abi.timers_enable(True)
    n_prof = max(3, min(args.steps, 20))
    for _ in range(n_prof):
        step()
    torch.cuda.synchronize()
    args.serial = was_serial
    kt = _abi.timers_read()
    _abi.timers_enable(False)
    kernels_us = {k: (1e3 * v[0] / v[1]) for k, v in kt.items() if v[1]}

    side = not (args.no_side_phases or refine)
    bwd_us = fused_us = fused_stream_us = softmax_us = evalpost_us = heads_us = heads_torch_us = 0.0
    fused_dets = evalpost_rows = 0
    heads_equal = None
    voc_phase = None
    if side:
        (bwd_us, fused_us, fused_stream_us, fused_dets, softmax_us, evalpost_us, evalpost_rows, heads_us, heads_torch_us,
         heads_equal, voc_phase) = side_phases(args, ssdbox, _abi, synth, crit, det, det_out, cfg, c, loc, conf, sc, priors,
                                               gt, offs, gmax, B, P, C, top_k, dev, n_gpus)
    log("per-kernel timers done")
    # ---- the timed region: K replays of the captured step (or eager launches) -------------------
    use_graph = not args.no_graph
    graph = None
    if use_graph:
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                step()
                step()
            torch.cuda.current_stream().wait_stream(s)
            # captured on the warm-up stream: the modules' per-stream workspaces already exist there and are in their
            # clean state, so the captured step carries no init launch (SSDBOX_*_WS_CLEAN)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):
                step()
        except Exception as e:  # e.g. NCCL capture unsupported
            sys.stderr.write("bench: CUDA-graph capture failed (%r); timing eager launches\n" % (e,))
            graph = None
            torch.cuda.synchronize()

    def run_once():
        if graph is not None:
            graph.replay()
        else:
            step()

    log("graph captured" if graph is not None else "eager mode")
    for _ in range(max(args.warmup, 3)):
        run_once()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        # ranks leave the host-side barrier up to ~0.5 ms apart; without this the rank that leaves first is
        # charged the others' delay inside its timed region (its first peer wait).  One more untimed step:
        # its peer exchange lines the GPUs up, and the start event is recorded on the stream right behind it.
        run_once()
    e0.record()
    for _ in range(args.steps):
        run_once()
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    elapsed_ms = e0.elapsed_time(e1)
    rank_ms = [elapsed_ms / args.steps]
    if dist is not None:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_ms = [float(x.item()) / args.steps for x in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    # keep the GPUs busy a little longer so the clock sampler sees load even for short K; the step
    # holds a collective when N > 1, so every rank runs the SAME number of extra steps
    n_extra = int(min(20000, max(1, 0.6 / max(elapsed_ms / args.steps * 1e-3, 1e-6))))
    for _ in range(n_extra):
        run_once()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = elapsed_ms / args.steps
    value = n_gpus * B / (ms_per_step * 1e-3)

    log("timed region done")
    # ---- e2e: public modules, pinned host inputs, H2D + D2H inside the timed region -------------
    loc_d, conf_d, sc_d = torch.empty_like(loc), torch.empty_like(conf), torch.empty_like(sc)
    out_h = torch.empty(B, C, top_k, 5, dtype=torch.float32).pin_memory()
    loss_h = torch.empty(2, dtype=torch.float32).pin_memory()
    tg_h = [t.pin_memory() for t in tg]

    copy_stream = torch.cuda.Stream()
    if refine:
        arm_loc_d, arm_conf_d = torch.empty_like(arm_loc), torch.empty_like(arm_conf)

    def e2e_step():
        # H2D on a copy stream, kernels on the current stream: T runs while the scores are still in
        # flight, the D2H of the detections leaves as soon as D is done (PCIe is full duplex)
        cur = torch.cuda.current_stream()
        with torch.no_grad():
            copy_stream.wait_stream(cur)               # the previous step's kernels are done with the buffers
            with torch.cuda.stream(copy_stream):
                loc_d.copy_(loc_h, non_blocking=True)
                conf_d.copy_(conf_h, non_blocking=True)
                if refine:
                    arm_loc_d.copy_(arm_loc_h, non_blocking=True)
                    arm_conf_d.copy_(arm_conf_h, non_blocking=True)
                tgd = tg_h                                   # host targets: packed into ONE pinned staging copy by pack_targets
                ev_t = torch.cuda.Event()
                ev_t.record(copy_stream)
                sc_d.copy_(sc_h, non_blocking=True)
                ev_d = torch.cuda.Event()
                ev_d.record(copy_stream)
            cur.wait_event(ev_t)
            if refine:
                al, ac = arm_crit((arm_loc_d, arm_conf_d, loc_d, conf_d, priors), tgd)
                ol, oc = odm_crit((arm_loc_d, arm_conf_d, loc_d, conf_d, priors), tgd)
                ll, lc = al + ol, ac + oc
            else:
                ll, lc = crit((loc_d, conf_d, priors), tgd)
            loss_h.copy_(torch.stack([ll, lc]), non_blocking=True)
            cur.wait_event(ev_d)
            o = rdet(arm_loc_d, arm_conf_d, loc_d, sc_d, priors) if refine else det(loc_d, sc_d, priors)
            out_h.copy_(o, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    h2d = loc_h.numel() * 4 + conf_h.numel() * 4 + sc_h.numel() * 4 + gt_h.numel() * 4 + offs_h.numel() * 4
    if refine:
        h2d += arm_loc_h.numel() * 4 + arm_conf_h.numel() * 4
    d2h = out_h.numel() * 4 + 8
    e2e_rank_s = [e2e_s]
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        e2e_rank_s = [float(x.item()) for x in allt]
        e2e_s = max(e2e_rank_s)
    # a bare pinned H2D copy of the same size on this rank, alone on the link (what PCIe gives one GPU) ...
    probe = torch.empty_like(conf_h, device=dev)
    probe.copy_(conf_h, non_blocking=True)
    torch.cuda.synchronize()
    solo_gbps = None
    for r in range(world):                           # the ranks take turns: each copy has the host side to itself
        if dist is not None:
            dist.barrier()
        if r == rank:
            t0 = time.perf_counter()
            probe.copy_(conf_h, non_blocking=True)
            torch.cuda.synchronize()
            solo_gbps = conf_h.numel() * 4 / (time.perf_counter() - t0) / 1e9
    all_gbps = None
    solo_all = None
    if dist is not None:
        t = torch.tensor([solo_gbps], dtype=torch.float64, device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        solo_all = [float(x.item()) for x in allt]
        # ... and with every rank copying at once (what the host's memory / PCIe root complexes give N GPUs)
        dist.barrier()
        t0 = time.perf_counter()
        probe.copy_(conf_h, non_blocking=True)
        torch.cuda.synchronize()
        mine = conf_h.numel() * 4 / (time.perf_counter() - t0) / 1e9
        t = torch.tensor([mine], dtype=torch.float64, device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        all_gbps = [float(x.item()) for x in allt]
    del probe
    e2e = {"value": n_gpus * B / e2e_s, "unit": "images/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s,
           "rank_ms_per_step": [1e3 * x for x in e2e_rank_s],
           "h2d_gbps_per_rank": [h2d / x / 1e9 for x in e2e_rank_s], "numa": numa,
           "h2d_probe_gbps": {"one_rank_alone_this_rank": solo_gbps, "one_rank_alone_every_rank": solo_all, "all_ranks_at_once": all_gbps,
                              "note": "bare pinned cudaMemcpyAsync of conf (%.0f MB): the PCIe / host-memory ceiling of the e2e number" % (conf_h.numel() * 4 / 1e6)},
           "api": ("ssdbox.RefineMultiBoxLoss.forward x2 + ssdbox.RefineDetectOut.__call__" if refine else
                   "ssdbox.MultiBoxLoss.forward + ssdbox.DetectOut.__call__") +
                  " from pinned host tensors (H2D on a copy stream, kernels overlap the next copy)"}

    log("e2e done")

    # ---- N > 1: is the peer-reduced loss the loss of the global batch? (printed in sanity.mgpu) -------------
    mgpu = None
    if dist is not None and refine:
        # both losses of the two-step path: the peer-exchanged sums against the rank-ordered fp64 sum of the LOCAL sums
        with torch.no_grad():
            step()
            torch.cuda.synchronize()
            mgpu = {"reduce": odm_crit.reduce_used, "bit_equal": True, "identical_on_all_ranks": True,
                    "note": "ARM and ODM loss: peer-exchanged {sum_l, sum_c, N} == rank-ordered fp64 sum of the NCCL-gathered local sums, on every rank"}
            for name, dcrit, ncls, use_arm in (("arm", arm_crit, 2, False), ("odm", odm_crit, C, True)):
                lcrit = ssdbox.RefineMultiBoxLoss(ncls, 0.5, True, 0, True, 3, 0.5, False, use_ARM=use_arm, distributed=False)
                lcrit.forward_packed_two_step(arm_loc, arm_conf, loc, conf, priors, gt, offs, gmax)
                local_sums = lcrit._last[0].clone()
                global_sums = dcrit._last[0].clone()
                gathered = [torch.empty_like(local_sums) for _ in range(world)]
                dist.all_gather(gathered, local_sums)
                tot = [0.0, 0.0, 0.0]
                for r in range(world):
                    v = gathered[r].cpu().tolist()
                    for k in range(3):
                        tot[k] += v[k]
                mine = global_sums.cpu().tolist()
                ok = torch.tensor([1 if all(tot[k] == mine[k] for k in range(3)) else 0], dtype=torch.int32, device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                same = [torch.empty_like(global_sums) for _ in range(world)]
                dist.all_gather(same, global_sums)
                mgpu["bit_equal"] = mgpu["bit_equal"] and bool(int(ok.item()))
                mgpu["identical_on_all_ranks"] = mgpu["identical_on_all_ranks"] and all(torch.equal(v, same[0]) for v in same)
                mgpu[name + "_global_sums"] = mine
        dist.barrier()
    if dist is not None and not refine:
        with torch.no_grad():
            local_crit = ssdbox.MultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, distributed=False)
            local_crit.forward_packed(loc, conf, priors, gt, offs, gmax)
            local_sums = local_crit._last[0].clone()
            gl_, gc_, _ = step()
            global_sums = crit._last[0].clone()         # after PendingLoss.wait(): the GLOBAL {sum_l, sum_c, N}
            global_losses = torch.stack([gl_, gc_]).double()
        gathered = [torch.empty_like(local_sums) for _ in range(world)]
        dist.all_gather(gathered, local_sums)            # NCCL: the ranks' LOCAL sums
        tot = [0.0, 0.0, 0.0]
        for r in range(world):                           # rank-ordered fp64 adds on the host, like the kernel does
            v = gathered[r].cpu().tolist()
            for k in range(3):
                tot[k] += v[k]
        mine = global_sums.cpu().tolist()
        ok = torch.tensor([1 if all(tot[k] == mine[k] for k in range(3)) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = [torch.empty_like(global_sums) for _ in range(world)]
        dist.all_gather(same, global_sums)
        mgpu = {"reduce": crit.reduce_used, "bit_equal": bool(int(ok.item())),
                "note": "peer-exchanged {sum_l, sum_c, N} == rank-ordered fp64 sum of the NCCL-gathered local sums, on every rank",
                "identical_on_all_ranks": all(torch.equal(v, same[0]) for v in same),
                "global_sums": mine, "n_pos_global": int(mine[2])}
        if world <= 2:
            # rank 0 recomputes the criterion over the concatenated global batch on its own GPU (the reference
            # computes it on one device over the gathered batch, train.py:137-142)
            rel = None
            if rank == 0:
                locs, confs, tgs = [loc], [conf], list(tg)
                for r in range(1, world):
                    tr = synth.gen_targets(B, C, c["gt_max"], 1000 * r)
                    tgs += tr
                    locs.append(synth.gen_loc(B, P, 1000 * r).to(dev))
                    confs.append(synth.gen_train_logits(B, P, C, 1000 * r).to(dev))
                g_all, o_all = synth.pack_targets(tgs)
                with torch.no_grad():
                    one = ssdbox.MultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, distributed=False)
                    wl, wc = one.forward_packed(torch.cat(locs), torch.cat(confs), priors, g_all.to(dev), o_all.to(dev),
                                                max(int(t.size(0)) for t in tgs))
                want = torch.stack([wl, wc]).double()
                rel = float(((global_losses - want).abs() / want.abs()).max())
                mgpu["single_device_losses"] = want.cpu().tolist()
                mgpu["n_pos_single_device"] = int(one._last[0][2])
                del locs, confs
            mgpu["global_batch_on_one_gpu_rel_err"] = rel
            mgpu["global_batch_ok"] = (rel is not None and rel <= 1e-6 and mgpu["n_pos_single_device"] == mgpu["n_pos_global"]) if rank == 0 else None
        mgpu["losses"] = global_losses.cpu().tolist()
        dist.barrier()

    # ---- dense-score regime of Detect (SURVEY 8d "worst case": background bias 4) as a second phase ----------
    dense_phase = None
    if side and n_gpus == 1 and not args.dense:
        sc_dense = synth.gen_detect_scores(B, P, C, seed, bkg_bias=4.0).to(dev)
        with torch.no_grad():
            for _ in range(2):
                det.forward(loc, sc_dense, priors, out=det_out)
            torch.cuda.synchronize()
            _abi.timers_enable(True)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(5):
                det.forward(loc, sc_dense, priors, out=det_out)
            ev[1].record()
            torch.cuda.synchronize()
            kd = _abi.timers_read()
            _abi.timers_enable(False)
        dense_us = 1e3 * ev[0].elapsed_time(ev[1]) / 5
        dense_phase = {"note": "DetectOut alone on dense scores (bkg bias 4: ~50 % of the (prior, class) pairs pass 0.01, every class saturates top_k)",
                       "us": dense_us, "images_per_s": B / (dense_us * 1e-6), "detections": int((det_out[..., 0] > 0).sum()),
                       "pass_fraction": float((sc_dense[:2, :, 1:] > 0.01).float().mean()),
                       "kernels_us": {k: 1e3 * v[0] / v[1] for k, v in kd.items() if v[1]},
                       "hbm_frac": bytes_D(P, C, top_k) * B / (dense_us * 1e-6) / 1e9 / load_peak()[0]}
        del sc_dense

    def teardown():
        # a process group whose collectives were captured in a CUDA graph can block in
        # destroy_process_group(): drop the graph, drain, rendezvous, then leave without it
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        teardown()
        return

    # ---- roofline of the dominant kernel + per-phase fractions ----------------------------------
    peak, peak_src = load_peak()
    traffic = load_traffic()
    t_us = sum(kernels_us.get(k, 0.0) for k in ("init", "match", "loss_stream", "mine_reduce"))
    d_us = sum(kernels_us.get(k, 0.0) for k in ("init", "detect_stream", "detect_segment", "detect_segment_big", "detect_overflow"))
    dom = max(("loss_stream", "detect_stream"), key=lambda k: kernels_us.get(k, 0.0))
    if refine:
        dom = "detect_stream"        # the library's loss_stream timer averages the ARM (C=2) and ODM (C=21) launches
    bT, bD = bytes_T(P, C, g_avg, B), bytes_D(P, C, top_k)
    if refine:
        # + ARM loss P*(4*2+16) + 20G, + the re-read of arm_loc / arm_conf (24 B per anchor) by the ODM loss and by Detect
        bT += P * (4 * 2 + 16) + 20 * g_avg + 24 * P
        bD += 24 * P
    dom_us = kernels_us[dom]
    dom_bytes = float(B) * P * 4 * C       # the compulsory read of conf / scores [B,P,C] fp32
    achieved = dom_bytes / (dom_us * 1e-6) / 1e9
    roofline = {"bound": "hbm", "kernel": dom + "_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic.get(dom) if args.workload == WORKLOAD else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (profiles/traffic.json), not measured in this run",
                "avg_launch_us": dom_us,
                "avg_launch_us_source": "CUDA events recorded by the library around each launch on the launch stream, in an eager pass with the ops back to back on ONE stream (in the two-stream timed region the kernels of T and D overlap)",
                "algorithmic_bytes_per_launch": dom_bytes, "peak_source": peak_src}
    other = "detect_stream" if dom == "loss_stream" else "loss_stream"
    phases = {
        "step_bytes_model": "T: P*(4C+16)+20G+16P/B per image; D: P*(4C+16)+20*C*top_k per image (SURVEY.md 8d)" +
                            ("; two-step: T += P*(4*2+16)+20G (ARM loss) + 24P (arm_loc/arm_conf re-read by the ODM loss), D += 24P" if refine else ""),
        "train_fwd": {"kernel_us_sum": t_us, "bytes_per_image": bT,
                      "hbm_frac_of_kernel_sum": bT * B / (t_us * 1e-6) / 1e9 / peak if t_us and not refine else None},
        "train_bwd": {"kernel_us_sum": bwd_us, "note": "loss_bwd_stream_kernel: one pass, TMA bulk stores of zero tiles carrying the selected rows; grad_conf/grad_loc fully written",
                      "bytes_written_per_image": P * (4 * C + 16),
                      "hbm_frac": P * (4 * C + 16) * B / (bwd_us * 1e-6) / 1e9 / peak if bwd_us else None},
        "detect": {"kernel_us_sum": d_us, "bytes_per_image": bD,
                   "hbm_frac_of_kernel_sum": bD * B / (d_us * 1e-6) / 1e9 / peak if d_us else None},
        "detect_fused_softmax": {"note": "DetectOut(conf_is_logits=True): softmax of ssd_v3.py:123-124 fused into the candidate pass (SURVEY 8f rank 2), same detections",
                                 "kernel_us_sum": fused_us, "detect_stream_us": fused_stream_us, "detections": fused_dets,
                                 "unfused_us": softmax_us + d_us, "torch_softmax_us": softmax_us,
                                 "hbm_frac_of_kernel_sum": bytes_D(P, C, top_k) * B / (fused_us * 1e-6) / 1e9 / peak if fused_us else None},
        "eval_post": {"note": "ssdbox_detections_compact after DetectOut: rescale + convert_ssd_result + COCO post_proc (evaluate_utils.py:63-70,175-203), 2 launches incl. host launch overhead",
                      "us": evalpost_us, "rows": evalpost_rows, "bytes_read": B * C * top_k * 5 * 4},
        "head_layout": {"note": "ssdbox_heads_to_rows: conf head outputs NCHW -> [B,P,C] in one launch (ssd_v3.py:114-121) against torch permute().contiguous() + cat",
                        "us": heads_us, "torch_us": heads_torch_us, "bytes_moved": 2 * B * P * C * 4, "identical_to_torch": heads_equal,
                        "hbm_frac": 2 * B * P * C * 4 / (heads_us * 1e-6) / 1e9 / peak if heads_us else None},
        "voc_eval": voc_phase,
        "dense": dense_phase,
        "step_hbm_frac": (bT + bD) * B / (ms_per_step * 1e-3) / 1e9 / peak,
        other + "_kernel": {"avg_launch_us": kernels_us.get(other), "achieved_GBps": dom_bytes / (kernels_us[other] * 1e-6) / 1e9 if other in kernels_us else None,
                            "traffic": traffic.get(other)},
        "kernels_us": kernels_us,
    }

    # ---- CPU baseline: the oracle port on this box's host cores (bounded sample) ----------------
    cpu_baseline = None
    if not args.no_cpu_baseline and n_gpus == 1:
        import warnings
        warnings.filterwarnings("ignore")
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        pri_cpu = priors.cpu()
        bt, bd, reps = 16, 8, 10          # a bounded sample: ~10 s of CPU work incl. the generation of its inputs
        impl = cpu_impl(args.workload)
        cpu_reference_pass(args.workload, C, P, pri_cpu, bt, bd, 7, args.dense, impl)
        tt = td = 0.0
        for i in range(reps):
            a, b = cpu_reference_pass(args.workload, C, P, pri_cpu, bt, bd, 50 + i, args.dense, impl)
            tt += a
            td += b
        per_img = tt / (reps * bt) + td / (reps * bd)
        cpu_baseline = {"value": 1.0 / per_img, "unit": "images/s", "cores": cores, "kind": impl[0],
                        "sample": "%d reps of (match+MultiBoxLoss fwd on %d images + Detect on %d images) of the same workload, %s, torch CPU, %d threads"
                                  % (reps, bt, bd, "the reference's own functions (unmodified copy in oracle/_ref)" if impl[0] == "reference"
                                     else "the oracle port (oracle/ssd_oracle.py)", cores),
                        "train_fwd_images_per_s": reps * bt / tt, "detect_images_per_s": reps * bd / td}

    launches_per_step = LAUNCHES["refine" if refine else "plain"]
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %s box path, P=%d priors, C=%d classes, B=%d images per GPU, step = %s" % (
                       args.workload, DESCR.get(args.workload, args.workload), P, C, B,
                       "ARM MultiBoxLoss (C=2) + ODM RefineMultiBoxLoss (refined anchors, negative-anchor filtering) fwd + "
                       "RefineDetectOut(top_k=200, conf 0.01, nms 0.45, theta 0.01)" if refine else
                       "match+MultiBoxLoss fwd + DetectOut(top_k=200, conf 0.01, nms 0.45)"),
                   "only": args.only or None, "global_batch": n_gpus * B, "detect_scores": "dense (bkg bias 4)" if args.dense else "sparse/realistic (bkg bias 10)",
                   "l2": "inputs larger than L2 (conf and scores are %.0f MB each vs 126 MB L2)" % (conf.numel() * 4 / 1e6),
                   "launch": ("CUDA graph replay" if graph is not None else "eager launches") +
                             ("" if (args.serial or args.only) else
                              ("; three streams inside the step (RefineDetectOut, ODM loss, ARM loss)" if refine else
                               "; two streams inside the step: DetectOut on a side stream (submitted first), its tail kernels overlap MultiBoxLoss's conf pass")),
                   "parallelism": ("images sharded by rank; {sum_l, sum_c, N_pos} reduced per step: %s" % (
                       ("posted over NVLink peer memory by the mining kernel (six self-validating 8-byte words per peer, no fence) and collected by " +
                        ("one warp of a Detect kernel" if args.serial else "the same kernel's last CTA") + " (no NCCL launch, no extra kernel)") if crit.reduce_used == "p2p"
                       else "one NCCL all-reduce")) if n_gpus > 1 else "single GPU"},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps, "rank_ms_per_step": rank_ms, "clocks": clocks, "phases": phases, "sanity": sanity,
    }
    sanity["mgpu"] = mgpu
    print(json.dumps(line), flush=True)
    teardown()


if __name__ == "__main__":
    main()
