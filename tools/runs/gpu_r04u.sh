#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "head_layout or head_output or multibox_loss_backward or test_full_batch_small_configs" > $O/r04u_memcheck.log 2>&1; echo "memcheck exit $?"
tail -8 $O/r04u_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" $O/r04u_memcheck.log
