"""CPU: the C-ABI library loads, exports exactly the symbols include/ssdbox.h declares, and its
argument validation answers without touching a GPU.  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from ssdbox import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_abi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _abi.lib()


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ssdbox.h")).read()
    return sorted(set(re.findall(r"SSDBOX_API\s+[\w\s\*]+?\b(ssdbox_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _header_symbols() == sorted(_abi.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", _abi.LIB_PATH]).decode()
    exported = sorted(set(re.findall(r"\bT (ssdbox_\w+)", out)))
    assert exported == _header_symbols()
    for s in _abi.SYMBOLS:
        assert hasattr(lib, s)


def test_library_has_sm100a_code_and_tma():
    sass = subprocess.run(["cuobjdump", "-sass", _abi.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass
    assert "UBLKCP" in sass                      # cp.async.bulk (TMA bulk copy) in the streaming kernels
    assert "SYNCS" in sass                       # mbarrier
    assert "REDUX" in sass                       # warp reductions in match
    assert "UTMALDG" in sass                     # cp.async.bulk.tensor (tensor-map TMA) in the head-layout kernel


def test_version_and_struct_layouts(lib):
    assert lib.ssdbox_abi_version() == 1
    # struct sizes must match the C definitions (checked against a tiny C program)
    prog = r'''
    #include <stdio.h>
    #include "ssdbox.h"
    #include <stddef.h>
    int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(ssdbox_prior_cfg), sizeof(ssdbox_loss_cfg), sizeof(ssdbox_detect_cfg),
                      sizeof(ssdbox_peer_group), offsetof(ssdbox_peer_group, wait_timeout_ms), sizeof(ssdbox_heads_cfg),
                      sizeof(ssdbox_voc_eval_cfg), offsetof(ssdbox_voc_eval_cfg, ovthresh));return 0;}
    '''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(_abi.PriorCfg), C.sizeof(_abi.LossCfg), C.sizeof(_abi.DetectCfg),
                     C.sizeof(_abi.PeerGroup), _abi.PeerGroup.wait_timeout_ms.offset, C.sizeof(_abi.HeadsCfg),
                     C.sizeof(_abi.VocEvalCfg), _abi.VocEvalCfg.ovthresh.offset]


def test_workspace_query(lib):
    assert _abi.workspace_bytes(_abi.OP_LOSS_FWD, 64, 24564, 81, 32) > 64 * 24564 * 10
    assert _abi.workspace_bytes(_abi.OP_DETECT, 64, 24564, 81, 0, 200) >= 64 * 81 * 1024 * 8
    assert _abi.workspace_bytes(99) == 0
    # any top_k: beyond 1024 the lists are sorted in the workspace (keys of every prior / box, a slice per CTA)
    assert _abi.workspace_bytes(_abi.OP_NMS, 0, 30000, 0, 0, 200) < 30000 * 4 + 1024
    assert _abi.workspace_bytes(_abi.OP_NMS, 0, 30000, 0, 0, 5000) >= 30000 * 4 + 32768 * 8 + 5000 * 24
    assert _abi.workspace_bytes(_abi.OP_DETECT, 2, 8732, 21, 0, 1500) >= 148 * (8732 * 4 + 16384 * 8 + 1500 * 24)
    # flags of the C ABI and of the host mirror agree
    hdr = open(os.path.join(ROOT, "include", "ssdbox.h")).read()
    for name in ("LOSS_SEPARATE_MATCH", "LOSS_GENERIC_MINE", "LOSS_NO_CLUSTER", "LOSS_DEFER_PEER_WAIT", "LOSS_WS_CLEAN", "LOSS_LSE_SHIFT",
                 "LOSS_MINE_HALF_CTA"):
        m = re.search(r"#define SSDBOX_%s (\d+)" % name, hdr)
        assert m and int(m.group(1)) == getattr(_abi, name), name


def test_argument_validation_without_gpu(lib):
    cfg = _abi.DetectCfg(1, 4, 3, 200, 0.01, 0.0, 0.1, 0.2, 0)
    assert lib.ssdbox_detect(C.byref(cfg), None, None, None, None, None, None, None, 0, None) == _abi.EINVAL
    assert "nms_threshold must be non negative" in _abi.last_error()
    cfg = _abi.DetectCfg(1, 4, 3, 70000, 0.01, 0.45, 0.1, 0.2, 0)      # beyond the any-top_k path (65536)
    assert lib.ssdbox_detect(C.byref(cfg), None, None, None, None, None, None, None, 0, None) == _abi.ESHAPE
    lc = _abi.LossCfg(2, 8, 3, 9999, 0.5, 3, 0.1, 0.2, 0, 1, 0)
    rc = lib.ssdbox_multibox_loss_fwd(C.byref(lc), *([None] * 15), None, 0, None)
    assert rc == _abi.ESHAPE
    assert lib.ssdbox_nms(None, None, 5, 0.45, 0, None, None, None, 0, None) == _abi.ESHAPE
    vc = _abi.VocEvalCfg(1, 2000, 0, 7, 0, 1, 0.5)
    assert lib.ssdbox_voc_eval(C.byref(vc), *([None] * 14), None, 0, None) == _abi.ESHAPE
    vc = _abi.VocEvalCfg(1, 3, 0, 4, 0, 1, 0.5)
    assert lib.ssdbox_voc_eval(C.byref(vc), *([None] * 14), None, 0, None) == _abi.EINVAL
    assert lib.ssdbox_crop_overlaps(None, None, None, -1, 2, None, None, None, None) == _abi.EINVAL
    assert _abi.workspace_bytes(_abi.OP_VOC_EVAL, 0, 600000, 21, 17000) > 600000 * 24
    pc = _abi.PriorCfg()
    pc.num_layers = 99
    assert lib.ssdbox_priorbox_count(C.byref(pc)) == _abi.ESHAPE


def test_prior_count_matches_configs(lib):
    from ssdbox import PriorBoxSSD, configs
    for name, c in configs.CONFIGS.items():
        cfg, _ = configs.get(name)
        pb = PriorBoxSSD(cfg)
        assert int(lib.ssdbox_priorbox_count(pb._abi_cfg(c["layer_dims"]))) == c["num_priors"]


def test_host_error_behaviour():
    import torch
    from ssdbox import DetectOut, MultiBoxLoss, PriorBoxSSD, configs
    with pytest.raises(ValueError, match="nms_threshold must be non negative"):
        DetectOut(21, 0, 200, 0.01, 0.0, [0.1, 0.2])                # detection.py:19-20
    cfg, _ = configs.get("ssd300_voc")
    bad = configs.AttrDict(MODEL=configs.AttrDict(dict(cfg.MODEL)))
    bad.MODEL.VARIANCE = [0.1, -0.2]
    with pytest.raises(ValueError, match="Variances must be greater than 0"):
        PriorBoxSSD(bad)                                              # prior_box.py:33-35
    bad = configs.AttrDict(MODEL=configs.AttrDict(dict(cfg.MODEL)))
    bad.MODEL.MIN_SIZES = [30, 60]
    with pytest.raises(Exception, match="length does not match"):
        PriorBoxSSD(bad)                                              # prior_box.py:42-43
    crit = MultiBoxLoss(21, 0.5, True, 0, True, 3, 0.5, False)
    with pytest.raises(RuntimeError, match="no CPU path"):           # no silent CPU fallback
        crit((torch.zeros(1, 4, 4), torch.zeros(1, 4, 21), torch.ones(4, 4)), [torch.zeros(1, 5)])
    with pytest.raises(RuntimeError, match="no CPU path"):
        DetectOut(21, 0, 200, 0.01, 0.45, [0.1, 0.2])(torch.zeros(1, 4, 4), torch.zeros(1, 4, 21), torch.ones(4, 4))
