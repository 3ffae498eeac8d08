"""CPU, world_size 2 over gloo: the multi-GPU host logic.  Images are sharded by rank; each rank
produces {sum smooth-L1, sum CE, N_pos} for its shard (here the oracle stands in for the CUDA
kernels, which cannot run without a GPU); one all-reduce(SUM) + divide must reproduce the
single-process loss over the whole batch."""
import os
import socket
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "object-detection-pytorch_b200"))
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import ssd_oracle as O
    from ssdbox import dist as sdist
    from tests import _util as U
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    x = U.seeded_inputs("refinedet320_voc", 5, 3)          # 5 images over 2 ranks: shards of 3 and 2
    loc, conf, tg = sdist.shard_batch(x["loc"], x["conf"], x["targets"], rank, world)
    d = O.multibox_loss(loc, conf, x["priors"], tg, x["C"], detail=True)
    sums = torch.tensor([float(d["sum_l"]), float(d["sum_c"]), float(d["n"])], dtype=torch.float64)
    sdist.allreduce_loss_sums(sums)
    ll, lc = sdist.finalize_losses(sums)
    torch.save({"ll": ll, "lc": lc, "n": sums[2], "shard": sdist.shard_range(5, rank, world)},
               os.path.join(outdir, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_loss_equals_full_batch():
    from oracle import ssd_oracle as O
    from tests import _util as U
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, _free_port(), d), nprocs=2, join=True)
        r0 = torch.load(os.path.join(d, "r0.pt"))
        r1 = torch.load(os.path.join(d, "r1.pt"))
    assert r0["shard"] == (0, 3) and r1["shard"] == (3, 5)
    x = U.seeded_inputs("refinedet320_voc", 5, 3)
    full = O.multibox_loss(x["loc"], x["conf"], x["priors"], x["targets"], x["C"], detail=True)
    for r in (r0, r1):
        assert int(r["n"]) == int(full["n"])
        assert abs(float(r["ll"]) - float(full["loss_l"])) <= 1e-5 * abs(float(full["loss_l"]))
        assert abs(float(r["lc"]) - float(full["loss_c"])) <= 1e-5 * abs(float(full["loss_c"]))
    assert float(r0["ll"]) == float(r1["ll"]) and float(r0["lc"]) == float(r1["lc"])


def _gather_worker(rank, world, port, outdir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "object-detection-pytorch_b200"))
    from ssdbox import dist as sdist
    from ssdbox import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = synth.gen_voc_eval_case(7, 5, 21, gt_max=3, fp_max=3)
    C = 5
    b, e = sdist.shard_range(7, rank, world)                 # 4 + 3 images
    seg = torch.as_tensor(case["seg"])
    r0, r1 = int(seg[b * C]), int(seg[e * C])
    rows = torch.as_tensor(case["rows"])[r0:r1]
    my_seg = (seg[b * C:e * C + 1] - r0).to(torch.int32)
    full_rows, full_seg = sdist.gather_detections(rows, my_seg)
    if rank == 1:
        rows, my_seg = rows[:0], torch.zeros_like(my_seg)    # a rank without any detection
    g_rows, g_seg = sdist.gather_detections(rows, my_seg)
    torch.save({"rows": g_rows, "seg": g_seg, "full_rows": full_rows, "full_seg": full_seg}, os.path.join(outdir, "g%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_detections_rebuilds_the_single_process_layout():
    """world_size 2 over gloo: the rows / segment offsets of two image shards (one of them empty) gathered in
    rank order equal what one process would have accumulated."""
    import numpy as np
    from ssdbox import synth
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_gather_worker, args=(2, _free_port(), d), nprocs=2, join=True)
        g0 = torch.load(os.path.join(d, "g0.pt"))
        g1 = torch.load(os.path.join(d, "g1.pt"))
    case = synth.gen_voc_eval_case(7, 5, 21, gt_max=3, fp_max=3)
    n0 = int(case["seg"][4 * 5])                             # rows of rank 0's four images
    want_seg = np.concatenate([case["seg"][:4 * 5], np.full(3 * 5 + 1, n0, np.int32)])
    for g in (g0, g1):
        assert np.array_equal(g["full_rows"].numpy(), case["rows"]) and np.array_equal(g["full_seg"].numpy(), case["seg"])
        assert np.array_equal(g["rows"].numpy(), case["rows"][:n0])
        assert np.array_equal(g["seg"].numpy(), want_seg)


def test_shard_range_partitions():
    from ssdbox import dist as sdist
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [sdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bench_reference_arm_prints_contract_line():
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.check_output([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference",
                                   "--steps", "1", "--warmup", "0"], stderr=subprocess.DEVNULL, timeout=600)
    line = json.loads(out.decode().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    # "reference" = the unmodified copy of the reference's package under oracle/_ref (oracle/build_ref.py), else the port
    have_ref = os.path.isfile(os.path.join(root, "oracle", "_ref", "lib", "layers", "box_utils.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
