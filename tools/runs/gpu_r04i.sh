#!/bin/bash
O=gpurun_out; mkdir -p $O
export SSDBOX_EXP=1
{
python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_GS=1 python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_GS=1 SSDBOX_BWD_ABLATE=1 python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_GS=1 SSDBOX_BWD_ABLATE=8 python tools/run_bwd.py 2>&1 | tail -1
SSDBOX_BWD_GS=1 SSDBOX_BWD_ABLATE=9 python tools/run_bwd.py 2>&1 | tail -1
} | tee $O/r04i_bwd_gs.log
SSDBOX_BWD_GS=1 SSDBOX_EXP_LIB=libssdbox_exp.so python - <<'P' 2>&1 | tail -5
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "object-detection-pytorch_b200")
from ssdbox import _abi
_abi.LIB_PATH = os.path.join("tools", "libssdbox_exp.so")
import pytest
sys.exit(pytest.main(["tests", "-m", "gpu", "-x", "-q", "-k", "backward or bwd or grad or refine or full_size or full_batch"]))
P
