"""experiments: bench.py on the experiment build (tools/libssdbox_exp.so, see tools/build_exp.sh), e.g.
    SSDBOX_CARVEOUT=100 python tools/exp_bench.py --no-side-phases --no-cpu-baseline --e2e-steps 1"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch_b200"))
from ssdbox import _abi

_abi.LIB_PATH = os.path.join(ROOT, "tools", os.environ.get("SSDBOX_EXP_LIB", "libssdbox_exp.so"))
import bench

bench.main()
