"""torchrun worker of tests/test_multi_gpu.py: one process per GPU, NCCL.  Every rank computes the
loss of its image shard with (a) the NVLink peer-memory reduction inside the mining kernel and
(b) the NCCL all-reduce; both must equal the single-GPU loss over the whole batch (computed on rank
0's GPU by the same library, itself pinned to the oracle by test_gpu_parity.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch_b200"))
import torch
import torch.distributed as dist


def main():
    import ssdbox
    from ssdbox import dist as sdist
    from tests import _util as U
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B = 2 * world + 1                                       # uneven shards
    x = U.seeded_inputs("ssd300_voc", B, 21)
    loc, conf, tg = sdist.shard_batch(x["loc"], x["conf"], x["targets"], rank, world)
    pri = x["priors"].to(dev)
    tgd = [t.to(dev) for t in tg]
    res = {}
    for mode in ("p2p", "nccl"):
        crit = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False, distributed=True, reduce=mode)
        for it in range(3):                                 # several calls: both epoch banks of the exchange
            lo = loc.to(dev).requires_grad_(True)
            co = conf.to(dev).requires_grad_(True)
            ll, lc = crit((lo, co, pri), tgd)
            (ll + lc).backward()
        assert crit.reduce_used == mode, (crit.reduce_used, mode)
        res[mode] = (float(ll), float(lc), crit._last[0].clone(), lo.grad.clone(), co.grad.clone())
    # single-GPU reference over the whole batch
    full = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False, distributed=False)
    lo = x["loc"].to(dev).requires_grad_(True)
    co = x["conf"].to(dev).requires_grad_(True)
    fl, fc = full((lo, co, pri), [t.to(dev) for t in x["targets"]])
    (fl + fc).backward()
    b, e = sdist.shard_range(B, rank, world)
    for mode in ("p2p", "nccl"):
        ll, lc, sums, gl, gc = res[mode]
        assert int(sums[2]) == int(full._last[0][2]), (mode, sums, full._last[0])
        assert abs(ll - float(fl)) <= 1e-6 * abs(float(fl)) and abs(lc - float(fc)) <= 1e-6 * abs(float(fc)), (mode, ll, lc, float(fl), float(fc))
        U.assert_close_rel(gl, lo.grad[b:e], 1e-5, 1e-8, mode + " grad_loc shard")
        U.assert_close_rel(gc, co.grad[b:e], 1e-5, 1e-8, mode + " grad_conf shard")
    # deferred wait: post in the mining kernel, other work, then collect -- same global losses
    from ssdbox import synth
    crit = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False, distributed=True, reduce="p2p")
    gt, offs = synth.pack_targets(tg)
    gmax = max(int(t.size(0)) for t in tg)
    det = ssdbox.DetectOut(x["C"], 0, 200, 0.01, 0.45, [0.1, 0.2])
    sc_d = x["scores"][b:e].to(dev)
    for it in range(4):
        pend = crit.forward_packed_deferred(loc.to(dev), conf.to(dev), pri, gt.to(dev), offs.to(dev), gmax)
        _ = torch.zeros(1 << 20, device=dev).sum()          # unrelated work between post and collect
        if it % 2 and e > b:                                # collect rides on the last Detect kernel (ssdbox_detect_peers)
            det.forward(loc.to(dev), sc_d, pri, pending=pend)
        dl, dc = pend.wait()
    assert abs(float(dl) - res["p2p"][0]) <= 1e-7 * abs(res["p2p"][0]) and abs(float(dc) - res["p2p"][1]) <= 1e-7 * abs(res["p2p"][1])
    assert torch.equal(crit._last[0], res["p2p"][2])
    # (a) global batch smaller than the world: the last rank's shard is EMPTY, it still takes part (posts zeros)
    Bs = world - 1
    xs = U.seeded_inputs("ssd300_voc", max(Bs, 1), 33)
    sl, sc_, st = sdist.shard_batch(xs["loc"][:Bs], xs["conf"][:Bs], xs["targets"][:Bs], rank, world)
    fulls = ssdbox.MultiBoxLoss(xs["C"], 0.5, True, 0, True, 3, 0.5, False, distributed=False)
    with torch.no_grad():
        wl, wc = fulls((xs["loc"][:Bs].to(dev), xs["conf"][:Bs].to(dev), pri), [t.to(dev) for t in xs["targets"][:Bs]])
    for mode in ("p2p", "nccl"):
        crit = ssdbox.MultiBoxLoss(xs["C"], 0.5, True, 0, True, 3, 0.5, False, distributed=True, reduce=mode)
        for it in range(2):
            with torch.no_grad():
                el, ec = crit((sl.to(dev), sc_.to(dev), pri), [t.to(dev) for t in st])
        assert abs(float(el) - float(wl)) <= 1e-6 * abs(float(wl)) and abs(float(ec) - float(wc)) <= 1e-6 * abs(float(wc)), \
            ("empty shard", mode, rank, float(el), float(ec), float(wl), float(wc))
        if mode == "p2p":
            assert crit._peers.timeouts() == 0
    # (b) DistributedDataParallel averages parameter gradients; with ddp_average=True the result equals the
    # single-device gradient of the reference's criterion over the gathered batch (train.py:137-144)
    from torch.nn.parallel import DistributedDataParallel as DDP
    F_ = 6

    class Heads(torch.nn.Module):
        def __init__(self, C):
            super().__init__()
            self.loc = torch.nn.Linear(F_, 4)
            self.conf = torch.nn.Linear(F_, C)

        def forward(self, f):
            return self.loc(f), self.conf(f)

    torch.manual_seed(7)
    heads = Heads(x["C"]).to(dev)
    single = Heads(x["C"]).to(dev)
    single.load_state_dict(heads.state_dict())
    feat = torch.randn(B, x["P"], F_, generator=torch.Generator().manual_seed(9))
    ddp = DDP(heads, device_ids=[local])
    crit = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False, distributed=True, ddp_average=True)
    lo, co = ddp(feat[b:e].to(dev))
    ll, lc = crit((lo, co, pri), tgd)
    (ll + lc).backward()
    lo, co = single(feat.to(dev))
    fl2, fc2 = full((lo, co, pri), [t.to(dev) for t in x["targets"]])
    (fl2 + fc2).backward()
    assert abs(float(ll) - float(fl2)) <= 1e-6 * abs(float(fl2))
    for (n1, p1), (n2, p2) in zip(heads.named_parameters(), single.named_parameters()):
        U.assert_close_rel(p1.grad, p2.grad, 1e-4, 1e-7, "DDP grad " + n1)
    # every rank holds bit-identical global sums (rank-ordered fp64 adds)
    mine = res["p2p"][2]
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    for v in allv:
        assert torch.equal(v, allv[0])
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("MGPU_OK world=%d p2p=(%.6f, %.6f) nccl=(%.6f, %.6f)" % (world, res["p2p"][0], res["p2p"][1], res["nccl"][0], res["nccl"][1]), flush=True)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
