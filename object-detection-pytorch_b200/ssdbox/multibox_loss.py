"""MultiBoxLoss -- drop-in for lib/layers/modules/multibox_loss.py:10-117 (and the
empty-target variant multibox_loss_v1.py:70-71).  Same constructor and
forward(predictions, targets) -> (loss_l, loss_c); both outputs are differentiable w.r.t.
loc_data / conf_data (train.py:142-144).  All arithmetic runs in libssdbox.so."""
import ctypes as C

import torch
import torch.nn as nn

from . import _abi


def pack_targets(targets, device):
    """list of B tensors [G_b,5] (or the 1-element sentinel of multibox_loss_v1.py:70) ->
    (gt [sum G,5] f32 on device, offsets int32 [B+1] on device, gmax).  Only shapes are read on
    the host (no device sync).  Host targets (the reference's data loader yields CPU tensors,
    train.py:131-135) travel as ONE pinned staging buffer -- offsets and boxes in a single H2D copy;
    device targets are gathered by one concatenation."""
    offs = [0]
    rows = []
    gmax = 0
    for t in targets:
        g = int(t.size(0)) if (t.dim() == 2 and t.size(-1) == 5) else 0
        offs.append(offs[-1] + g)
        gmax = max(gmax, g)
        if g:
            rows.append(t)
    B1 = len(offs)
    total = offs[-1]
    if all(not t.is_cuda for t in rows):
        # staging layout (4-byte words): [offsets int32 (B+1) | pad to 4 words | gt f32 (total*5)]
        head = (B1 + 3) // 4 * 4
        stage = torch.empty(head + max(total, 1) * 5, dtype=torch.float32, pin_memory=torch.cuda.is_available())
        stage[:B1].view(torch.int32).copy_(torch.tensor(offs, dtype=torch.int32))
        if rows:
            torch.cat([r.to(torch.float32) for r in rows], 0, out=stage[head:head + total * 5].view(total, 5))
        else:
            stage[head:].zero_()
        dstage = stage.to(device, non_blocking=True)
        return dstage[head:].view(-1, 5), dstage[:B1].view(torch.int32), gmax
    rows = [r if r.device == device else r.to(device, non_blocking=True) for r in rows]
    gt = torch.cat(rows, 0).to(torch.float32).contiguous()
    offsets = torch.tensor(offs, dtype=torch.int32, pin_memory=True).to(device, non_blocking=True)
    return gt, offsets, gmax


class _State(object):
    """Per-module reusable buffers, one set per (batch shape, device, stream): stable pointers (the forward can
    be captured in a CUDA graph) and no sharing between two streams that drive the same module."""

    def __init__(self):
        self.ws = _abi.Workspace()
        self.slots = {}
        self.scratch = None
        self.sel = self.tidx = self.sums = self.losses = None

    def ensure(self, B, P, device):
        key = (B, P, device, torch.cuda.current_stream(device).cuda_stream)
        slot = self.slots.get(key)
        if slot is None:
            slot = (torch.empty(B, P, dtype=torch.int16, device=device),
                    torch.empty(B, P, dtype=torch.int16, device=device),
                    torch.zeros(3, dtype=torch.float64, device=device),
                    torch.zeros(2, dtype=torch.float32, device=device))
            self.slots[key] = slot
        self.sel, self.tidx, self.sums, self.losses = slot


def make_refine(refine):
    """(arm_loc [B,P,4], arm_conf [B,P,2] | None, theta) -> (ctypes ssdbox_refine, the tensors it points to)"""
    arm_loc, arm_conf, theta = refine
    arm_loc = _abi.as_f32(arm_loc.detach())
    arm_conf = _abi.as_f32(arm_conf.detach()) if arm_conf is not None else None
    r = _abi.Refine(_abi.ptr(arm_loc, torch.float32, "arm_loc"), _abi.ptr(arm_conf, torch.float32, "arm_conf", True), float(theta), 0)
    return r, (arm_loc, arm_conf)


def loss_forward_raw(state, loc, conf, priors, gt, offsets, gmax, num_classes, threshold, negpos_ratio,
                     variance, anchors_xyxy=None, pool=None, binarize=False, finalize=True, debug=None,
                     fresh=False, flags=0, peers=None, refine=None, lse_group=False):
    """One call of ssdbox_multibox_loss_fwd on validated CUDA tensors.  Returns
    (cfg, sums[3] f64, losses[2] f32, sel[B,P] i16, tidx[B,P] i16).  With fresh=False the outputs are
    the module's persistent buffers (overwritten by the next call); fresh=True allocates new ones
    (what autograd saves for backward)."""
    B, P = loc.size(0), loc.size(1)
    dev = loc.device
    if fresh:
        sel = torch.empty(B, P, dtype=torch.int16, device=dev)
        tidx = torch.empty(B, P, dtype=torch.int16, device=dev)
        sums = torch.empty(3, dtype=torch.float64, device=dev)
        losses = torch.empty(2, dtype=torch.float32, device=dev)
    else:
        state.ensure(B, P, dev)
        sel, tidx, sums, losses = state.sel, state.tidx, state.sums, state.losses
    if priors.dim() == 3 and priors.size(0) == 1:          # the reference docstring's [1, num_priors, 4]
        priors = priors[0]
    per_image = priors.dim() == 3
    if per_image and priors.size(0) != B:
        raise ValueError("ssdbox: per-image priors have batch %d, loc_data has %d" % (priors.size(0), B))
    if priors.size(-2) != P or priors.size(-1) != 4:
        raise ValueError("ssdbox: priors must be [%d,4] or [%d,%d,4], got %s" % (P, B, P, tuple(priors.shape)))
    # gmax is an upper bound of the truths per image: rounded up to a power of two (>= 8) the workspace layout --
    # and with it the "clean" state the previous call left behind -- stays the same from batch to batch
    gmax = int(gmax)
    gq = 8
    while gq < gmax:
        gq *= 2
    gmax = gq if gmax > 0 else 0
    ws, n, clean = state.ws.acquire(_abi.workspace_bytes(_abi.OP_LOSS_FWD, B, P, num_classes, gmax), dev,
                                    ("loss", B, P, int(num_classes), gmax))
    if int(flags) & _abi.LOSS_LSE_SHIFT:
        # fidelity mode (box_utils.py:272-273): the log-sum-exp subtracts the maximum of the whole batch_conf -- of the
        # GLOBAL batch when it is sharded over ranks -- handed to the kernels in sums[0]
        if state.scratch is None or state.scratch.device != dev:
            state.scratch = torch.empty(256, dtype=torch.uint8, device=dev)
        _abi.check(_abi.lib().ssdbox_global_max(_abi.ptr(conf, torch.float32, "conf_data"), conf.numel(), _abi.ptr(sums),
                                                _abi.ptr(state.scratch), 256, _abi.stream_ptr(dev)))
        if lse_group is not False and torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(sums[:1], op=torch.distributed.ReduceOp.MAX, group=lse_group)
    cfg = _abi.LossCfg(B, P, int(num_classes), gmax, float(threshold), int(negpos_ratio),
                       float(variance[0]), float(variance[1]), 1 if binarize else 0, 1 if finalize else 0,
                       4 * P if per_image else 0, int(flags) | (_abi.LOSS_WS_CLEAN if clean else 0), 0)
    dbg = debug or {}
    if refine is not None:       # RefineDet fused: the ARM outputs go to the kernels as they are
        rf, keepalive = make_refine(refine)
        _abi.check(_abi.lib().ssdbox_multibox_loss_fwd_refine(
            C.byref(cfg), _abi.ptr(loc, torch.float32, "loc_data"), _abi.ptr(conf, torch.float32, "conf_data"),
            _abi.ptr(priors, torch.float32, "priors"), C.byref(rf), _abi.ptr(gt, torch.float32, "gt"),
            _abi.ptr(offsets, torch.int32, "gt_offsets"), _abi.ptr(sums), _abi.ptr(losses),
            _abi.ptr(sel), _abi.ptr(tidx), _abi.ptr(dbg.get("conf_t"), torch.int64, "conf_t", True),
            _abi.ptr(dbg.get("loc_t"), torch.float32, "loc_t", True), _abi.ptr(dbg.get("neg"), torch.uint8, "neg", True),
            _abi.ptr(dbg.get("keys"), torch.float32, "keys", True),
            C.byref(peers) if peers is not None else None, ws, n, _abi.stream_ptr(dev)))
        state.ws.commit()
        return cfg, sums, losses, sel, tidx
    _abi.check(_abi.lib().ssdbox_multibox_loss_fwd_peers(
        C.byref(cfg), _abi.ptr(loc, torch.float32, "loc_data"), _abi.ptr(conf, torch.float32, "conf_data"),
        _abi.ptr(priors, torch.float32, "priors"), _abi.ptr(anchors_xyxy, torch.float32, "anchors", True),
        _abi.ptr(pool, torch.uint8, "pool", True), _abi.ptr(gt, torch.float32, "gt"),
        _abi.ptr(offsets, torch.int32, "gt_offsets"), _abi.ptr(sums), _abi.ptr(losses),
        _abi.ptr(sel), _abi.ptr(tidx), _abi.ptr(dbg.get("conf_t"), torch.int64, "conf_t", True),
        _abi.ptr(dbg.get("loc_t"), torch.float32, "loc_t", True), _abi.ptr(dbg.get("neg"), torch.uint8, "neg", True),
        _abi.ptr(dbg.get("keys"), torch.float32, "keys", True),
        C.byref(peers) if peers is not None else None, ws, n, _abi.stream_ptr(dev)))
    state.ws.commit()
    return cfg, sums, losses, sel, tidx


class _MultiBoxLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loc, conf, priors, gt, offsets, anchors_xyxy, pool, mod, gmax, refine=None):
        st = mod._state
        distributed = mod._is_distributed()
        # the only collective of the path: {sum smooth-L1, sum CE, N_pos} summed over ranks.  Preferred:
        # inside the mining kernel over NVLink peer memory; otherwise one NCCL all-reduce + finalize.
        if getattr(mod, "_force_peers", False):
            peers = mod._peers.group
        else:
            peers = mod._peer_group(loc.device) if distributed else None
        if peers is not None:
            mod._check_no_pending()
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        cfg, sums, losses, sel, tidx = loss_forward_raw(
            st, loc, conf, priors, gt, offsets, gmax, mod.num_classes, mod.threshold, mod.negpos_ratio,
            mod.variance, anchors_xyxy, pool, mod.binarize_labels, finalize=(not distributed) or peers is not None,
            debug=mod._debug, fresh=need_grad, flags=mod.abi_flags | (_abi.LOSS_LSE_SHIFT if mod.lse_global_max else 0), peers=peers,
            refine=refine, lse_group=(mod.process_group if distributed else False))
        if distributed and peers is None:
            import torch.distributed as dist
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=mod.process_group)
            _abi.check(_abi.lib().ssdbox_multibox_loss_finalize(_abi.ptr(sums), _abi.ptr(losses),
                                                                _abi.stream_ptr(loc.device)))
        mod._last = (sums, sel, tidx)
        ctx.grad_mul = 1.0
        if distributed and mod.ddp_average:
            import torch.distributed as dist
            ctx.grad_mul = float(dist.get_world_size(mod.process_group))
        if need_grad:
            ctx.cfg = cfg
            ctx.refine = refine
            ctx.save_for_backward(loc, conf, priors, gt, offsets, sel, tidx, sums)
            return losses[0], losses[1]
        out = losses.clone()
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_l, g_c):
        loc, conf, priors, gt, offsets, sel, tidx, sums = ctx.saved_tensors
        dev = loc.device
        gout = torch.stack([g_l.reshape(()).to(torch.float32), g_c.reshape(()).to(torch.float32)]).contiguous()
        if ctx.grad_mul != 1.0:
            gout = gout * ctx.grad_mul
        grad_loc = torch.empty_like(loc)
        grad_conf = torch.empty_like(conf)
        if ctx.refine is not None:
            rf, keepalive = make_refine(ctx.refine)
            _abi.check(_abi.lib().ssdbox_multibox_loss_bwd_refine(
                C.byref(ctx.cfg), _abi.ptr(loc), _abi.ptr(conf), _abi.ptr(priors), C.byref(rf), _abi.ptr(gt), _abi.ptr(offsets),
                _abi.ptr(sel), _abi.ptr(tidx), _abi.ptr(sums), _abi.ptr(gout), _abi.ptr(grad_loc), _abi.ptr(grad_conf),
                _abi.stream_ptr(dev)))
            return grad_loc, grad_conf, None, None, None, None, None, None, None, None
        _abi.check(_abi.lib().ssdbox_multibox_loss_bwd(
            C.byref(ctx.cfg), _abi.ptr(loc), _abi.ptr(conf), _abi.ptr(priors), _abi.ptr(gt), _abi.ptr(offsets),
            _abi.ptr(sel), _abi.ptr(tidx), _abi.ptr(sums), _abi.ptr(gout), _abi.ptr(grad_loc), _abi.ptr(grad_conf),
            _abi.stream_ptr(dev)))
        return grad_loc, grad_conf, None, None, None, None, None, None, None, None


class PendingLoss(object):
    """Result of MultiBoxLoss.forward_packed_deferred: wait() enqueues ssdbox_multibox_loss_peer_finish
    on the current stream and returns (loss_l, loss_c) of the GLOBAL batch."""

    def __init__(self, peers, sums, losses, ready, owner=None):
        self._peers, self._sums, self._losses, self._ready = peers, sums, losses, ready
        self._owner = owner
        self._collected = False

    def _detect_tail_args(self):
        """(peer group, sums, losses) for ssdbox_detect_peers, or None when there is nothing left to collect."""
        if self._ready is not None or self._collected:
            return None
        return self._peers, self._sums, self._losses

    def _completed_by_detect(self):
        self._collected = True
        if self._owner is not None:
            self._owner._pending_unwaited = False

    def wait(self):
        if self._ready is None and self._collected:
            out = self._losses.clone()
            self._ready = (out[0], out[1])
        if self._ready is None:
            if self._owner is not None:
                self._owner._pending_unwaited = False
            _abi.check(_abi.lib().ssdbox_multibox_loss_peer_finish(
                C.byref(self._peers), _abi.ptr(self._sums), _abi.ptr(self._losses),
                _abi.stream_ptr(self._losses.device)))
            out = self._losses.clone()
            self._ready = (out[0], out[1])
        return self._ready


class MultiBoxLoss(nn.Module):
    """SSD weighted loss (multibox_loss.py:10-46).  Arguments as in the reference; as there,
    prior_for_matching / bkg_label / neg_mining / neg_overlap / encode_target are stored but
    unused.  `variance` replaces the reference's read of the global cfg (multibox_loss.py:46);
    `distributed=True` (or an initialised default process group with world size > 1) sums the loss
    numerators and the positive count across ranks (images are sharded by rank).  `reduce`:
    "p2p" = inside the mining kernel over NVLink peer memory (ssdbox.dist.PeerExchange), "nccl" = one
    all-reduce after it, "auto" = p2p when symmetric memory is available, else nccl.  `ddp_average`: see
    the comment in __init__ (set it when the network is wrapped in DistributedDataParallel)."""

    def __init__(self, num_classes, overlap_thresh, prior_for_matching, bkg_label, neg_mining, neg_pos,
                 neg_overlap, encode_target, use_gpu=True, variance=(0.1, 0.2), distributed=None,
                 process_group=None, reduce="auto", ddp_average=False):
        super(MultiBoxLoss, self).__init__()
        self.use_gpu = use_gpu
        self.num_classes = num_classes
        self.threshold = overlap_thresh
        self.background_label = bkg_label
        self.encode_target = encode_target
        self.use_prior_for_matching = prior_for_matching
        self.do_neg_mining = neg_mining
        self.negpos_ratio = neg_pos
        self.neg_overlap = neg_overlap
        self.variance = list(variance)
        self.binarize_labels = False
        self.distributed = distributed
        self.process_group = process_group
        if reduce not in ("auto", "p2p", "nccl"):
            raise ValueError("reduce must be 'auto', 'p2p' or 'nccl'")
        self.reduce = reduce
        self.reduce_used = None     # "p2p" / "nccl" once the first distributed forward has run
        # Distributed losses are normalised by the GLOBAL N, so a rank's backward yields its shard's part of the
        # single-device gradient: parameter gradients must be SUMMED over ranks.  DistributedDataParallel
        # averages them -- ddp_average=True multiplies the backward by the world size so that DDP's mean
        # equals the reference's single-device gradient (train.py:137-144); the returned loss values stay global.
        self.ddp_average = bool(ddp_average)
        self._pending_unwaited = False
        self._peers = None
        self.abi_flags = 0          # _abi.LOSS_SEPARATE_MATCH: matching as its own kernel
        # fidelity switch (default off): log_sum_exp with ONE maximum over the whole (global) batch like box_utils.py:272-273
        # instead of each row's own maximum -- same values up to fp32 rounding, but also the reference's underflow for rows
        # far below the batch maximum; costs a pass over conf and the generic streaming kernel
        self.lse_global_max = False
        self._state = _State()
        self._debug = None
        self._last = None

    def _is_distributed(self):
        if self.distributed is False:
            return False
        import torch.distributed as dist
        ok = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1
        if self.distributed and not ok:
            raise RuntimeError("ssdbox: distributed=True but no process group with world size > 1 is initialised")
        return ok

    def _peer_group(self, device):
        """ctypes ssdbox_peer_group for the peer-memory reduction, or None (use NCCL).  Built lazily at
        the first distributed forward (collective: all ranks get here together)."""
        if self.reduce == "nccl":
            self.reduce_used = "nccl"
            return None
        if self._peers is None:
            from . import dist as sdist
            try:
                self._peers = sdist.PeerExchange(device, self.process_group)
            except Exception as e:
                self._peers = False
                self._peers_error = repr(e)
            # all ranks must take the same route: fall back together if any rank failed
            import torch.distributed as dist
            ok = torch.tensor([1 if self._peers else 0], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.process_group)
            if int(ok.item()) == 0:
                if self.reduce == "p2p":
                    raise RuntimeError("ssdbox: reduce='p2p' but the peer exchange could not be set up: %s"
                                       % getattr(self, "_peers_error", "failed on another rank"))
                self._peers = False
        self.reduce_used = "p2p" if self._peers else "nccl"
        return self._peers.group if self._peers else None

    def _check_no_pending(self):
        """The deferred protocol posts in call k and collects in PendingLoss.wait(); a second peer-reduced
        forward before that wait would advance the epoch and overwrite the slot bank the un-collected call
        still has to read (two banks by epoch parity).  A dropped PendingLoss therefore raises here instead
        of silently desynchronising the ranks."""
        if self._pending_unwaited:
            raise RuntimeError("ssdbox: the previous forward_packed_deferred() has not been completed -- call "
                               "PendingLoss.wait() before the next peer-reduced forward")

    def use_local_peer_exchange(self, device):
        """tests: run the peer-exchange path with a world of one rank (no process group needed)."""
        from . import dist as sdist
        self._peers = sdist.LocalPeerExchange(device)
        self._force_peers = True

    def forward(self, predictions, targets):
        loc_data, conf_data, priors = predictions
        if not loc_data.is_cuda:
            raise RuntimeError("ssdbox: MultiBoxLoss runs on CUDA tensors only (no CPU path)")
        dev = loc_data.device
        loc = _abi.as_f32(loc_data)
        conf = _abi.as_f32(conf_data)
        if conf.numel() != loc.size(0) * loc.size(1) * self.num_classes:
            raise ValueError("conf_data has %d elements, MultiBoxLoss(num_classes=%d) expects %d x %d x %d"
                             % (conf.numel(), self.num_classes, loc.size(0), loc.size(1), self.num_classes))
        conf = conf.view(loc.size(0), loc.size(1), self.num_classes)       # (an empty shard has no -1 to infer)
        pri = _abi.as_f32(priors, dev)[:loc.size(1), :].contiguous()        # multibox_loss.py:62
        gt, offsets, gmax = pack_targets(targets, dev)
        return self.forward_packed(loc, conf, pri, gt, offsets, gmax)

    def forward_packed(self, loc, conf, priors, gt, offsets, gmax, anchors_xyxy=None, pool=None, refine=None):
        """Same as forward with the targets already in the C-ABI layout (no host work: capturable).
        `refine` = (arm_loc, arm_conf | None, theta): RefineDet fused (anchors refined and filtered inside the kernels)."""
        return _MultiBoxLossFn.apply(loc, conf, priors, gt, offsets, anchors_xyxy, pool, self, int(gmax), refine)

    def forward_packed_deferred(self, loc, conf, priors, gt, offsets, gmax):
        """Inference-style (no autograd) forward whose cross-rank wait is deferred: returns a
        PendingLoss; kernels enqueued before `pending.wait()` (e.g. DetectOut of the same step) overlap
        the wait for the other ranks' sums, so rank skew does not stall the step.  Without the
        peer-memory reduction (single GPU, reduce='nccl') it simply runs the normal forward."""
        with torch.no_grad():
            distributed = self._is_distributed()
            if getattr(self, "_force_peers", False):
                peers = self._peers.group
            else:
                peers = self._peer_group(loc.device) if distributed else None
            if peers is None:
                ll, lc = self.forward_packed(loc, conf, priors, gt, offsets, gmax)
                return PendingLoss(None, None, None, (ll, lc))
            self._check_no_pending()
            cfg, sums, losses, sel, tidx = loss_forward_raw(
                self._state, loc, conf, priors, gt, offsets, gmax, self.num_classes, self.threshold,
                self.negpos_ratio, self.variance, None, None, self.binarize_labels, finalize=True,
                debug=None, fresh=False, flags=self.abi_flags | _abi.LOSS_DEFER_PEER_WAIT | (_abi.LOSS_LSE_SHIFT if self.lse_global_max else 0),
                peers=peers, lse_group=self.process_group)
            self._last = (sums, sel, tidx)
            self._pending_unwaited = True
            return PendingLoss(peers, sums, losses, None, owner=self)

    def intermediates(self, predictions, targets):
        """Runs the forward and also materialises the reference's intermediates
        (conf_t, loc_t, neg mask, mining keys) -- used by the parity tests."""
        loc_data = predictions[0]            # (loc, conf, priors) or RefineDet's (arm_loc, arm_conf, odm_loc, odm_conf, priors)
        dev = loc_data.device
        B, P = loc_data.size(0), loc_data.size(1)
        self._debug = dict(conf_t=torch.empty(B, P, dtype=torch.int64, device=dev),
                           loc_t=torch.empty(B, P, 4, dtype=torch.float32, device=dev),
                           neg=torch.empty(B, P, dtype=torch.uint8, device=dev),
                           keys=torch.empty(B, P, dtype=torch.float32, device=dev))
        try:
            ll, lc = self.forward(predictions, targets)
            d = dict(self._debug)
        finally:
            self._debug = None
        sums, sel, tidx = self._last
        d.update(loss_l=ll, loss_c=lc, sums=sums.clone(), sel=sel.clone(), tidx=tidx.clone())
        return d
