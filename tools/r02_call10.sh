#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_t5.py -x -q 2>&1 | tail -6 ) > gpurun_out/r02c10_tests.log 2>&1
for b in 0 1; do VGPA_T5_ATTN_MMA=$b timeout 300 python tools/dev_t5.py 2>&1 | tail -2 | sed "s/^/MMA=$b /"; done > gpurun_out/r02c10_t5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"t5_attention" -s 2 -c 2 python tools/dev_t5.py 2>&1 | grep -E "t5_attention|gpu__time" >> gpurun_out/r02c10_t5.log
cat gpurun_out/r02c10_tests.log gpurun_out/r02c10_t5.log
