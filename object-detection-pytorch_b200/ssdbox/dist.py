"""Multi-GPU plumbing of the box path: images are independent units, sharded by rank; the only
cross-image coupling of MultiBoxLoss is the normaliser N = sum_b num_pos and the two loss
numerators (multibox_loss.py:114-116), so ONE all-reduce(SUM) of three fp64 scalars per step is
the whole data-path communication.  Detect needs none."""
import torch


def shard_range(num_images, rank, world):
    """[begin, end) of the images rank `rank` owns (contiguous, sizes differ by at most one)."""
    base, rem = divmod(int(num_images), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(loc, conf, targets, rank, world):
    b, e = shard_range(loc.size(0), rank, world)
    return loc[b:e], conf[b:e], targets[b:e]


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def bind_to_local_numa(device_index):
    """Pins this process to the CPUs next to GPU `device_index` (its PCIe root complex's NUMA node), so that pinned
    host buffers allocated afterwards are first-touched -- and therefore placed -- in the memory of that node.  With
    one process per GPU this keeps every rank's host-to-device copies off the inter-socket link; without it all
    ranks' staging buffers can end up on one node and share its memory controllers and one socket's PCIe uplinks.
    Returns a dict describing what was done (for logs), never raises."""
    import os
    info = {"bound": False}
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        with open(base + "/numa_node") as f:
            info["numa_node"] = int(f.read().strip())
        with open(base + "/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = cpus & allowed if cpus & allowed else set()
        info["pci"] = bdf
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
        info["cpus"] = len(cpus) if cpus else len(allowed)
    except Exception as e:          # no sysfs in the container, exotic topology, ...
        info["error"] = repr(e)
    return info


def allreduce_loss_sums(sums, group=None):
    """sums: fp64 [3] = {sum smooth-L1, sum CE, N_pos} of the local shard -> global sums (in place)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def finalize_losses(sums):
    """(loss_l, loss_c) = sums[0:2] / N with the N == 0 convention of the library (zeros)."""
    n = float(sums[2])
    if n <= 0:
        return torch.zeros((), dtype=torch.float32), torch.zeros((), dtype=torch.float32)
    return (sums[0] / n).to(torch.float32), (sums[1] / n).to(torch.float32)


class PeerExchange(object):
    """Exchange buffers for the NVLink peer-memory reduction of the loss sums (ssdbox_peer_group,
    include/ssdbox.h): one small zero-filled buffer per rank, allocated in symmetric memory
    (torch.distributed._symmetric_memory) so that every rank holds a device pointer to every peer's
    buffer.  Built once per (process group, device); collective: every rank of the group must
    construct it at the same point.  Raises when symmetric memory / peer access is unavailable --
    MultiBoxLoss then keeps the NCCL all-reduce."""

    def __init__(self, device, group=None, wait_timeout_s=30.0):
        """wait_timeout_s: how long a rank waits for its peers' sums before it gives up (the losses of
        that call are NaN and timeouts() counts it; the CUDA context survives).  Size it for the longest
        stall a rank may see between two steps (data loading, checkpointing)."""
        import ctypes as C
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import _abi
        grp = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(grp)
        self.rank = dist.get_rank(grp)
        if self.world > _abi.MAX_PEERS:
            raise RuntimeError("ssdbox: peer reduction supports at most %d ranks" % _abi.MAX_PEERS)
        nbytes = int(_abi.lib().ssdbox_peer_buffer_bytes())
        gname = grp.group_name if hasattr(grp, "group_name") else grp
        enable = getattr(symm_mem, "enable_symm_mem_for_group", None)
        if enable is not None:           # needed by older torch releases, a no-op / deprecated later
            try:
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    enable(gname)
            except Exception:
                pass
        self.buf = symm_mem.empty((nbytes + 7) // 8, dtype=torch.int64, device=device)
        self.handle = symm_mem.rendezvous(self.buf, gname)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=grp)          # every buffer is zero before any rank stores into a peer
        ptrs = [int(x) for x in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or int(self.handle.rank) != self.rank:
            raise RuntimeError("ssdbox: symmetric-memory rendezvous returned an unexpected layout")
        self.group = _abi.PeerGroup()
        self.group.rank = self.rank
        self.group.world = self.world
        self.group.wait_timeout_ms = int(wait_timeout_s * 1000)
        for r, ptr in enumerate(ptrs):
            self.group.bufs[r] = C.c_void_p(ptr)


    def timeouts(self):
        """number of calls in which this rank gave up waiting for a peer (word 1 of its exchange buffer;
        synchronises the device).  Non-zero: the ranks are out of step -- build a new PeerExchange."""
        return int(self.buf[1].item())


class LocalPeerExchange(object):
    """World-size-1 peer group on ordinary device memory (tests: drives the exchange code path of the
    mining kernel on a single GPU; the rank stores into and reads from its own buffer)."""

    def __init__(self, device):
        import ctypes as C
        from . import _abi
        nbytes = int(_abi.lib().ssdbox_peer_buffer_bytes())
        self.buf = torch.zeros((nbytes + 7) // 8, dtype=torch.int64, device=device)
        self.world, self.rank = 1, 0
        self.group = _abi.PeerGroup()
        self.group.rank = 0
        self.group.world = 1
        self.group.bufs[0] = C.c_void_p(self.buf.data_ptr())


def gather_detections(rows, seg, group=None):
    """Evaluation shards the image set by rank (contiguous ranges, shard_range) and needs ONE exchange before
    the metric: every rank's result rows.  rows [n_r, ncol] + seg int32 [I_r*C + 1] of this rank's images
    (ssdbox.voc_eval.VOCDetections.flat()) -> the rows / seg of the whole image set in rank order, on every
    rank: two all-gathers (row counts and segment counts, then the padded payloads).  The result is what a
    single process would have accumulated, so ssdbox_voc_eval sees the same file order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rows, seg
    world = dist.get_world_size(group)
    dev = rows.device
    seg = seg.to(dev, torch.int32)
    sizes = torch.tensor([rows.size(0), seg.numel() - 1], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    n_rows = [int(s[0]) for s in all_sizes]
    n_segs = [int(s[1]) for s in all_sizes]
    ncol = rows.size(1)
    pad_rows = torch.zeros(max(max(n_rows), 1), ncol, dtype=rows.dtype, device=dev)
    pad_rows[:rows.size(0)] = rows
    pad_seg = torch.zeros(max(max(n_segs), 1), dtype=torch.int32, device=dev)
    pad_seg[:seg.numel() - 1] = seg[:-1]
    got_rows = [torch.empty_like(pad_rows) for _ in range(world)]
    got_seg = [torch.empty_like(pad_seg) for _ in range(world)]
    dist.all_gather(got_rows, pad_rows, group=group)
    dist.all_gather(got_seg, pad_seg, group=group)
    out_rows, out_seg, base = [], [], 0
    for r in range(world):
        out_rows.append(got_rows[r][:n_rows[r]])
        out_seg.append(got_seg[r][:n_segs[r]] + base)
        base += n_rows[r]
    out_seg.append(torch.tensor([base], dtype=torch.int32, device=dev))
    return torch.cat(out_rows, 0).contiguous(), torch.cat(out_seg).contiguous()
