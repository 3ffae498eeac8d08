"""Two ops of one step on two streams.

MultiBoxLoss (training targets + loss) and DetectOut (inference post-processing) of the same batch are
independent.  Each starts with an HBM-bound pass that fills every SM (one persistent CTA per SM), so the two
passes cannot share the GPU -- but DetectOut's tail (the per-list sort / NMS kernels: latency-bound, a few warps
per SM, 18 KB of shared memory) fits beside MultiBoxLoss's streaming CTAs.  Submitting DetectOut to a second
stream FIRST and the loss to the current stream lets the hardware run

    detect_stream | loss_stream + [segments, overflow, rewritten lists] | mine_reduce

instead of six kernels back to back: at SSD512-COCO B=64 the step drops from 195 to 180 us (0.85 -> 0.93 of the HBM
roofline), at SSD300-VOC B=32 from 47 to 37 us.  Results are identical (same kernels, same inputs; each module
keeps one workspace per stream).  Capturable in a CUDA graph (the side stream forks from and joins the capturing
stream)."""
import torch


class TwoStreamStep(object):
    """step = TwoStreamStep(); ll, lc, out = step(loss_call, detect_call)

    `detect_call()` is submitted to the side stream, `loss_call()` to the current stream; both return whatever
    the wrapped module returns.  The current stream waits for the side stream before the call returns, so the
    caller may use both results in stream order as usual."""

    def __init__(self, device=None):
        self.side = torch.cuda.Stream(device=device)

    def __call__(self, loss_call, detect_call):
        cur = torch.cuda.current_stream(self.side.device)
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            det = detect_call()
        loss = loss_call()
        cur.wait_stream(self.side)
        return loss, det


class MultiStreamStep(object):
    """The same idea for any number of independent ops (e.g. RefineDet's ARM loss, ODM loss and RefineDetectOut of one
    batch): calls[1:] go to side streams in the order given (first submitted, first served by the SMs), calls[0] to
    the current stream, which then waits for all of them.  Returns the list of results."""

    def __init__(self, n, device=None):
        self.sides = [torch.cuda.Stream(device=device) for _ in range(max(n - 1, 0))]

    def __call__(self, calls):
        dev = self.sides[0].device if self.sides else None
        cur = torch.cuda.current_stream(dev)
        res = [None] * len(calls)
        for i, s in enumerate(self.sides[:len(calls) - 1], start=1):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                res[i] = calls[i]()
        res[0] = calls[0]()
        for s in self.sides[:len(calls) - 1]:
            cur.wait_stream(s)
        return res
