"""ctypes binding of libssdbox.so (include/ssdbox.h) + small torch glue.

The library is the product; this file only turns torch tensors into (pointer, size, stream)
arguments.  There is deliberately NO CPU fallback: a non-CUDA tensor raises, and a missing
library raises at import of the first op.
"""
import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libssdbox.so")

OK, EINVAL, ESHAPE, EALIGN, EWORKSPACE, ECUDA = 0, -1, -2, -3, -4, -5
LOSS_SEPARATE_MATCH = 1
LOSS_GENERIC_MINE = 2
LOSS_NO_CLUSTER = 4
LOSS_DEFER_PEER_WAIT = 8
LOSS_WS_CLEAN = 16
LOSS_LSE_SHIFT = 32
LOSS_MINE_HALF_CTA = 64
OP_MATCH, OP_LOSS_FWD, OP_DETECT, OP_NMS, OP_LSE, OP_MINE, OP_COMPACT, OP_VOC_EVAL = 1, 2, 3, 4, 5, 6, 7, 8
MAX_LAYERS, MAX_MIN_SIZES, MAX_RATIOS = 16, 4, 6

# every symbol include/ssdbox.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "ssdbox_abi_version", "ssdbox_last_error", "ssdbox_workspace_bytes", "ssdbox_priorbox_count",
    "ssdbox_priorbox", "ssdbox_point_form", "ssdbox_center_form", "ssdbox_jaccard", "ssdbox_encode",
    "ssdbox_decode", "ssdbox_log_sum_exp", "ssdbox_global_max", "ssdbox_match_encode", "ssdbox_hard_negative_mine",
    "ssdbox_multibox_loss_fwd", "ssdbox_multibox_loss_fwd_peers", "ssdbox_multibox_loss_peer_finish",
    "ssdbox_peer_buffer_bytes",
    "ssdbox_multibox_loss_finalize", "ssdbox_multibox_loss_bwd",
    "ssdbox_multibox_loss_fwd_refine", "ssdbox_multibox_loss_bwd_refine", "ssdbox_detect_refine",
    "ssdbox_nms", "ssdbox_detect", "ssdbox_detect_peers", "ssdbox_detections_compact", "ssdbox_heads_to_rows", "ssdbox_voc_eval", "ssdbox_crop_overlaps", "ssdbox_arm_filter", "ssdbox_timers_enable", "ssdbox_timers_read",
]

KERNEL_NAMES = ["init", "match", "loss_stream", "mine_reduce", "loss_bwd", "detect_stream", "detect_segment",
                "detect_overflow", "materialize", "detect_segment_big"]


class PriorCfg(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int32), ("clip", C.c_int32), ("flip", C.c_int32), ("has_max", C.c_int32),
        ("image_h", C.c_double), ("image_w", C.c_double),
        ("feat_h", C.c_int32 * MAX_LAYERS), ("feat_w", C.c_int32 * MAX_LAYERS),
        ("step", C.c_double * MAX_LAYERS),
        ("num_min", C.c_int32 * MAX_LAYERS),
        ("min_size", (C.c_double * MAX_MIN_SIZES) * MAX_LAYERS),
        ("max_size", C.c_double * MAX_LAYERS),
        ("num_ratio", C.c_int32 * MAX_LAYERS),
        ("ratio", (C.c_double * MAX_RATIOS) * MAX_LAYERS),
    ]


class LossCfg(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("P", C.c_int32), ("C", C.c_int32), ("gmax", C.c_int32),
        ("threshold", C.c_float), ("negpos_ratio", C.c_int32), ("var0", C.c_float), ("var1", C.c_float),
        ("binarize_labels", C.c_int32), ("finalize", C.c_int32), ("prior_batch_stride", C.c_int64),
        ("flags", C.c_int32), ("reserved", C.c_int32),
    ]


class DetectCfg(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("P", C.c_int32), ("C", C.c_int32), ("top_k", C.c_int32),
        ("conf_thresh", C.c_float), ("nms_thresh", C.c_float), ("var0", C.c_float), ("var1", C.c_float),
        ("prior_batch_stride", C.c_int64), ("flags", C.c_int32), ("reserved", C.c_int32),
    ]


DETECT_LOGITS = 1
DETECT_WS_CLEAN = 2
MAX_PEERS = 16
MAX_HEADS = 16


class HeadsCfg(C.Structure):
    """ssdbox_heads_cfg"""
    _fields_ = [("num_layers", C.c_int32), ("B", C.c_int32), ("channels", C.c_int32 * MAX_HEADS),
                ("hw", C.c_int32 * MAX_HEADS), ("src", C.c_void_p * MAX_HEADS)]



class VocEvalCfg(C.Structure):
    """ssdbox_voc_eval_cfg"""
    _fields_ = [("num_images", C.c_int32), ("num_classes", C.c_int32), ("num_rows", C.c_int32),
                ("row_stride", C.c_int32), ("num_gt", C.c_int32), ("use_07_metric", C.c_int32),
                ("ovthresh", C.c_double)]


class Refine(C.Structure):
    """ssdbox_refine: the ARM head's outputs for the fused RefineDet path."""
    _fields_ = [("arm_loc", C.c_void_p), ("arm_conf", C.c_void_p), ("theta", C.c_float), ("reserved", C.c_int32)]


class PeerGroup(C.Structure):
    """ssdbox_peer_group: every rank's exchange buffer as addressable from this device."""
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("bufs", C.c_void_p * MAX_PEERS),
                ("wait_timeout_ms", C.c_int64)]


_lib = None
_lock = threading.Lock()
P_ = C.c_void_p


def _declare(lib):
    i32, i64, f32, sz = C.c_int32, C.c_int64, C.c_float, C.c_size_t
    lib.ssdbox_abi_version.restype = C.c_int
    lib.ssdbox_abi_version.argtypes = []
    lib.ssdbox_last_error.restype = C.c_int
    lib.ssdbox_last_error.argtypes = [C.c_char_p, sz]
    lib.ssdbox_workspace_bytes.restype = sz
    lib.ssdbox_workspace_bytes.argtypes = [C.c_int] * 6
    lib.ssdbox_peer_buffer_bytes.restype = sz
    lib.ssdbox_peer_buffer_bytes.argtypes = []
    lib.ssdbox_priorbox_count.restype = i64
    lib.ssdbox_priorbox_count.argtypes = [C.POINTER(PriorCfg)]
    sigs = {
        "ssdbox_priorbox": [C.POINTER(PriorCfg), P_, i64, P_],
        "ssdbox_point_form": [P_, i64, P_, P_],
        "ssdbox_center_form": [P_, i64, P_, P_],
        "ssdbox_jaccard": [P_, i32, P_, i32, P_, P_],
        "ssdbox_encode": [P_, P_, i64, f32, f32, P_, P_],
        "ssdbox_decode": [P_, P_, i64, i64, f32, f32, P_, P_, P_],
        "ssdbox_log_sum_exp": [P_, i64, i32, P_, P_, sz, P_],
        "ssdbox_global_max": [P_, i64, P_, P_, sz, P_],
        "ssdbox_match_encode": [P_, P_, i32, P_, i64, P_, i32, i32, f32, f32, f32, i32, P_, P_, P_, P_, P_, sz, P_],
        "ssdbox_hard_negative_mine": [P_, P_, P_, i32, i32, i32, P_, P_, sz, P_],
        "ssdbox_multibox_loss_fwd": [C.POINTER(LossCfg)] + [P_] * 15 + [P_, sz, P_],
        "ssdbox_multibox_loss_fwd_peers": [C.POINTER(LossCfg)] + [P_] * 15 + [C.POINTER(PeerGroup), P_, sz, P_],
        "ssdbox_multibox_loss_fwd_refine": [C.POINTER(LossCfg), P_, P_, P_, C.POINTER(Refine)] + [P_] * 10 + [C.POINTER(PeerGroup), P_, sz, P_],
        "ssdbox_multibox_loss_bwd_refine": [C.POINTER(LossCfg), P_, P_, P_, C.POINTER(Refine)] + [P_] * 8 + [P_],
        "ssdbox_detect_refine": [C.POINTER(DetectCfg), P_, P_, P_, C.POINTER(Refine), P_, P_, P_, sz, P_],
        "ssdbox_multibox_loss_finalize": [P_, P_, P_],
        "ssdbox_multibox_loss_peer_finish": [C.POINTER(PeerGroup), P_, P_, P_],
        "ssdbox_multibox_loss_bwd": [C.POINTER(LossCfg)] + [P_] * 11 + [P_],
        "ssdbox_nms": [P_, P_, i32, f32, i32, P_, P_, P_, sz, P_],
        "ssdbox_detect": [C.POINTER(DetectCfg), P_, P_, P_, P_, P_, P_, P_, sz, P_],
        "ssdbox_detect_peers": [C.POINTER(DetectCfg), P_, P_, P_, P_, P_, P_, C.POINTER(PeerGroup), P_, P_, P_, sz, P_],
        "ssdbox_arm_filter": [P_, i64, f32, P_, P_],
        "ssdbox_detections_compact": [P_, i32, i32, i32, P_, P_, i32, P_, i64, P_, P_, P_, sz, P_],
        "ssdbox_heads_to_rows": [C.POINTER(HeadsCfg), P_, P_],
        "ssdbox_voc_eval": [C.POINTER(VocEvalCfg)] + [P_] * 14 + [P_, sz, P_],
        "ssdbox_crop_overlaps": [P_, P_, P_, i32, i32, P_, P_, P_, P_],
    }
    sigs["ssdbox_timers_enable"] = [C.c_int]
    sigs["ssdbox_timers_read"] = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = args


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.isfile(LIB_PATH):
                    raise RuntimeError(
                        "ssdbox: %s is missing -- build it with `python object-detection-pytorch_b200/build.py` "
                        "(there is no CPU / eager fallback)" % LIB_PATH)
                l = C.CDLL(LIB_PATH)
                _declare(l)
                _lib = l
    return _lib


def last_error():
    buf = C.create_string_buffer(512)
    lib().ssdbox_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc):
    if rc == OK:
        return
    msg = last_error()
    if rc == EINVAL:
        raise ValueError(msg)
    raise RuntimeError("ssdbox error %d: %s" % (rc, msg))


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t, dtype=None, name="tensor", allow_none=False):
    """Device pointer of a contiguous CUDA tensor (validated like torch would)."""
    if t is None:
        if allow_none:
            return None
        raise ValueError("ssdbox: %s is None" % name)
    if not isinstance(t, torch.Tensor):
        raise TypeError("ssdbox: %s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("ssdbox: %s lives on %s -- the box path has no CPU implementation; move it to a CUDA device"
                           % (name, t.device))
    if dtype is not None and t.dtype != dtype:
        raise TypeError("ssdbox: %s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("ssdbox: %s must be contiguous" % name)
    return C.c_void_p(t.data_ptr()) if t.numel() else C.c_void_p(0)


def as_f32(t, device=None):
    """contiguous fp32 CUDA view/copy of t (host tensors are uploaded)."""
    if device is not None and t.device != device:
        t = t.to(device, non_blocking=True)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class Workspace(object):
    """Grow-only scratch buffer, one per (device, stream) that uses the owner: two streams driving the same
    module never share scratch (the C ABI is re-entrant only with distinct workspaces).  The pointer of a
    (device, stream) stays stable once large enough: graph-safe.

    The loss and Detect ops keep a little state in their workspace and hand it back clean after every call
    (SSDBOX_LOSS_WS_CLEAN / SSDBOX_DETECT_WS_CLEAN in include/ssdbox.h).  acquire() / commit() track that: acquire
    reports `clean` when the last COMPLETED call on this buffer carried the same tag (op + shape); the caller
    then sets the flag and the library skips its init launch."""

    def __init__(self):
        self.bufs = {}
        self.tags = {}
        self._pending = None

    def _key(self, device):
        return (device, torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0)

    def get(self, nbytes, device):
        ptr, n, _ = self.acquire(nbytes, device, None)
        return ptr, n

    def acquire(self, nbytes, device, tag):
        nbytes = int(nbytes)
        key = self._key(device)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            self.bufs[key] = buf
            self.tags[key] = None
        clean = tag is not None and self.tags.get(key) == tag
        self.tags[key] = None                 # unknown until the call has been enqueued without an error
        self._pending = (key, tag)
        return C.c_void_p(buf.data_ptr()), buf.numel(), clean

    def commit(self):
        if self._pending is not None:
            key, tag = self._pending
            self.tags[key] = tag
            self._pending = None

    @property
    def buf(self):
        """the most recently created buffer (tests / introspection)"""
        return next(reversed(self.bufs.values())) if self.bufs else None


def workspace_bytes(op, B=0, P=0, Cn=0, gmax=0, top_k=0):
    return int(lib().ssdbox_workspace_bytes(op, int(B), int(P), int(Cn), int(gmax), int(top_k)))


def timers_enable(on):
    check(lib().ssdbox_timers_enable(1 if on else 0))


def timers_read():
    """{kernel name: (total device ms, launches)} accumulated since timers_enable(True)."""
    out = {}
    for i, name in enumerate(KERNEL_NAMES):
        ms, n = C.c_double(0), C.c_int64(0)
        check(lib().ssdbox_timers_read(i, C.byref(ms), C.byref(n)))
        out[name] = (ms.value, n.value)
    return out
