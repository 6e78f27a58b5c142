#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02c12_npoly.log
for np in 4 3 5 4; do
  VGPA_ATTN_NPOLY8=$np timeout 300 python bench.py --primary-only --steps 4 --warmup 3 2>/dev/null | tail -1 | sed "s/^/NPOLY8=$np /" >> gpurun_out/r02c12_npoly.log
done
( timeout 300 python -m pytest tests/test_gpu_t5.py -x -q 2>&1 | tail -3 ) > gpurun_out/r02c12_tests.log 2>&1
timeout 300 python tools/dev_t5.py 2>&1 | tail -2 > gpurun_out/r02c12_t5.log
cut -c1-330 gpurun_out/r02c12_npoly.log; cat gpurun_out/r02c12_tests.log gpurun_out/r02c12_t5.log
