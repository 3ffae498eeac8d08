#!/usr/bin/env python
"""bench.py -- SSD512-COCO box hot path on B200: match + MultiBoxLoss forward AND Detect/NMS.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ssdbox|reference]

One "step" = one pass of the hot path over one batch of synthetic input of BASELINE.json
configs[1] (SSD512 VGG16 COCO: P=24564 priors, C=81 classes, B=64 images PER GPU):
    T  match + MultiBoxLoss forward  (init, match, loss_stream, mine_reduce kernels)
    D  DetectOut: threshold, top-200, NMS (init, detect_stream, detect_segment small/big, detect_overflow)
`value` = images/s with inputs resident in HBM (whole job: N * B / max-over-ranks step time; every
image passes through both T and D), timed with CUDA events around K CUDA-graph replays.
`e2e` = same metric through the public modules (MultiBoxLoss.forward / DetectOut.__call__) with
pinned HOST inputs: H2D of loc/conf/targets/scores and D2H of the losses and the detection tensor
inside the timed region.  `--impl reference` times the reference algorithm's CPU path (the oracle
port: /root/reference itself is Python and not present on the GPU box) on the host cores.
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
def cpu_reference_pass(name, C, P, priors_cpu, bt, bd, seed, dense):
    """times ONE pass: T on `bt` images + D on `bd` images; returns (t_T, t_D) seconds."""
    from oracle import ssd_oracle as O
    from ssdbox import configs, synth
    _, c = configs.get(name)
    tg = synth.gen_targets(bt, C, c["gt_max"], seed)
    loc = synth.gen_loc(max(bt, bd), P, seed)
    conf = synth.gen_train_logits(bt, P, C, seed)
    sc = synth.gen_detect_scores(bd, P, C, seed, bkg_bias=4.0 if dense else 10.0)
    t0 = time.perf_counter()
    O.multibox_loss(loc[:bt], conf, priors_cpu, tg, C)
    t1 = time.perf_counter()
    O.detect(loc[:bd], sc, priors_cpu, C)
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
    for i in range(args.warmup):
        cpu_reference_pass(args.workload, C, P, pri, bt, bd, i, args.dense)
    tt = td = 0.0
    t_start = time.perf_counter()
    for i in range(args.steps):
        a, b = cpu_reference_pass(args.workload, C, P, pri, bt, bd, 100 + i, args.dense)
        tt += a
        td += b
    wall = time.perf_counter() - t_start
    per_img = tt / (args.steps * bt) + td / (args.steps * bd)
    value = 1.0 / per_img
    sample = "per step: match+MultiBoxLoss fwd on %d images + Detect on %d images (sparse scores), torch CPU, %d threads" % (bt, bd, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (tt + td) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: P=%d C=%d, CPU sample of the B=%d step" % (args.workload, P, C, c["batch"]),
                   "inputs": "synthetic seeded (ssdbox.synth), same generator as the GPU arm"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "phases": {"train_fwd_images_per_s": args.steps * bt / tt, "detect_images_per_s": args.steps * bd / td},
        "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


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

    crit = ssdbox.MultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, distributed=(world > 1))
    crit.abi_flags |= args.loss_flags
    det = ssdbox.DetectOut(C, 0, top_k, 0.01, 0.45, VAR)
    det_out = torch.empty(B, C, top_k, 5, dtype=torch.float32, device=dev)

    def step():
        # T, then D, then the (multi-GPU) wait for the other ranks' loss sums: D overlaps that wait
        with torch.no_grad():
            pending = crit.forward_packed_deferred(loc, conf, priors, gt, offs, gmax)
            out = det.forward(loc, sc, priors, out=det_out)
            ll, lc = pending.wait()
        return ll, lc, out

    log("inputs ready")
    # eager warm-up (lazy init, workspace allocation), also the functional sanity of the step
    for _ in range(2):
        ll, lc, out = step()
    torch.cuda.synchronize()
    sanity = {"loss_l": float(ll), "loss_c": float(lc), "detections": int((out[..., 0] > 0).sum())}

    # ---- per-kernel device times (eager, CUDA events inside the library on the launch stream) ---
    _abi.timers_enable(True)
    n_prof = max(3, min(args.steps, 20))
    for _ in range(n_prof):
        step()
    torch.cuda.synchronize()
    kt = _abi.timers_read()
    _abi.timers_enable(False)
    kernels_us = {k: (1e3 * v[0] / v[1]) for k, v in kt.items() if v[1]}

    # forward + backward through the autograd module (reported beside the headline, not part of it)
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
            torch.cuda.current_stream().wait_stream(s)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
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

    def e2e_step():
        # H2D on a copy stream, kernels on the current stream: T runs while the scores are still in
        # flight, the D2H of the detections leaves as soon as D is done (PCIe is full duplex)
        cur = torch.cuda.current_stream()
        with torch.no_grad():
            copy_stream.wait_stream(cur)               # the previous step's kernels are done with the buffers
            with torch.cuda.stream(copy_stream):
                loc_d.copy_(loc_h, non_blocking=True)
                conf_d.copy_(conf_h, non_blocking=True)
                tgd = [t.to(dev, non_blocking=True) for t in tg_h]
                ev_t = torch.cuda.Event()
                ev_t.record(copy_stream)
                sc_d.copy_(sc_h, non_blocking=True)
                ev_d = torch.cuda.Event()
                ev_d.record(copy_stream)
            cur.wait_event(ev_t)
            ll, lc = crit((loc_d, conf_d, priors), tgd)
            loss_h.copy_(torch.stack([ll, lc]), non_blocking=True)
            cur.wait_event(ev_d)
            o = det(loc_d, sc_d, priors)
            out_h.copy_(o, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = loc_h.numel() * 4 + conf_h.numel() * 4 + sc_h.numel() * 4 + gt_h.numel() * 4
    d2h = out_h.numel() * 4 + 8
    e2e = {"value": n_gpus * B / e2e_s, "unit": "images/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s,
           "api": "ssdbox.MultiBoxLoss.forward + ssdbox.DetectOut.__call__ from pinned host tensors (H2D on a copy stream, kernels overlap the next copy)"}

    log("e2e done")

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
    dom_us = kernels_us[dom]
    dom_bytes = float(B) * P * 4 * C       # the compulsory read of conf / scores [B,P,C] fp32
    achieved = dom_bytes / (dom_us * 1e-6) / 1e9
    roofline = {"bound": "hbm", "kernel": dom + "_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic.get(dom), "avg_launch_us": dom_us,
                "algorithmic_bytes_per_launch": dom_bytes, "peak_source": peak_src}
    other = "detect_stream" if dom == "loss_stream" else "loss_stream"
    phases = {
        "step_bytes_model": "T: P*(4C+16)+20G+16P/B per image; D: P*(4C+16)+20*C*top_k per image (SURVEY.md 8d)",
        "train_fwd": {"kernel_us_sum": t_us, "bytes_per_image": bytes_T(P, C, g_avg, B),
                      "hbm_frac_of_kernel_sum": bytes_T(P, C, g_avg, B) * B / (t_us * 1e-6) / 1e9 / peak if t_us else None},
        "train_bwd": {"kernel_us_sum": bwd_us, "note": "loss_bwd_stream_kernel: one pass, TMA bulk stores of zero tiles carrying the selected rows; grad_conf/grad_loc fully written",
                      "bytes_written_per_image": P * (4 * C + 16),
                      "hbm_frac": P * (4 * C + 16) * B / (bwd_us * 1e-6) / 1e9 / peak if bwd_us else None},
        "detect": {"kernel_us_sum": d_us, "bytes_per_image": bytes_D(P, C, top_k),
                   "hbm_frac_of_kernel_sum": bytes_D(P, C, top_k) * B / (d_us * 1e-6) / 1e9 / peak if d_us else None},
        "detect_fused_softmax": {"note": "DetectOut(conf_is_logits=True): softmax of ssd_v3.py:123-124 fused into the candidate pass (SURVEY 8f rank 2), same detections",
                                 "kernel_us_sum": fused_us, "detect_stream_us": fused_stream_us, "detections": fused_dets,
                                 "unfused_us": softmax_us + d_us, "torch_softmax_us": softmax_us,
                                 "hbm_frac_of_kernel_sum": bytes_D(P, C, top_k) * B / (fused_us * 1e-6) / 1e9 / peak if fused_us else None},
        "eval_post": {"note": "ssdbox_detections_compact after DetectOut: rescale + convert_ssd_result + COCO post_proc (evaluate_utils.py:63-70,175-203), 2 launches incl. host launch overhead",
                      "us": evalpost_us, "rows": evalpost_rows, "bytes_read": B * C * top_k * 5 * 4},
        "head_layout": {"note": "ssdbox_heads_to_rows: conf head outputs NCHW -> [B,P,C] in one launch (ssd_v3.py:114-121) against torch permute().contiguous() + cat",
                        "us": heads_us, "torch_us": heads_torch_us, "bytes_moved": 2 * B * P * C * 4, "identical_to_torch": heads_equal,
                        "hbm_frac": 2 * B * P * C * 4 / (heads_us * 1e-6) / 1e9 / peak},
        "voc_eval": voc_phase,
        "step_hbm_frac": (bytes_T(P, C, g_avg, B) + bytes_D(P, C, top_k)) * B / (ms_per_step * 1e-3) / 1e9 / peak,
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
        bt, bd, reps = 4, 2, 3
        cpu_reference_pass(args.workload, C, P, pri_cpu, bt, bd, 7, args.dense)
        tt = td = 0.0
        for i in range(reps):
            a, b = cpu_reference_pass(args.workload, C, P, pri_cpu, bt, bd, 50 + i, args.dense)
            tt += a
            td += b
        per_img = tt / (reps * bt) + td / (reps * bd)
        cpu_baseline = {"value": 1.0 / per_img, "unit": "images/s", "cores": cores, "kind": "port",
                        "sample": "%d reps of (match+MultiBoxLoss fwd on %d images + Detect on %d images) of the same workload, torch CPU oracle, %d threads"
                                  % (reps, bt, bd, cores),
                        "train_fwd_images_per_s": reps * bt / tt, "detect_images_per_s": reps * bd / td}

    # T: init, loss_stream, mine_reduce; D: init, detect_stream, segment (warp / CTA), overflow; N > 1: + the collect kernel
    launches_per_step = 8 + (1 if n_gpus > 1 else 0)
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: SSD512 VGG16 COCO box path, P=%d priors, C=%d classes, B=%d images per GPU, "
                               "step = match+MultiBoxLoss fwd + DetectOut(top_k=200, conf 0.01, nms 0.45)" % (args.workload, P, C, B),
                   "global_batch": n_gpus * B, "detect_scores": "dense (bkg bias 4)" if args.dense else "sparse/realistic (bkg bias 10)",
                   "l2": "inputs larger than L2 (conf and scores are %.0f MB each vs 126 MB L2)" % (conf.numel() * 4 / 1e6),
                   "launch": "CUDA graph replay" if graph is not None else "eager launches",
                   "parallelism": ("images sharded by rank; {sum_l, sum_c, N_pos} reduced per step: %s" % (
                       "posted over NVLink peer memory by the mining kernel, collected by a 1-warp kernel after DetectOut (no NCCL launch)" if crit.reduce_used == "p2p"
                       else "one NCCL all-reduce")) if n_gpus > 1 else "single GPU"},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps, "rank_ms_per_step": rank_ms, "clocks": clocks, "phases": phases, "sanity": sanity,
    }
    print(json.dumps(line), flush=True)
    teardown()


if __name__ == "__main__":
    main()
