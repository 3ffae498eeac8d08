#!/bin/bash
O=gpurun_out
N=${1:-2}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02g_pytest_${N}gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02g_pytest_${N}gpu.log
ls $O | grep mgpu
