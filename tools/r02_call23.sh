#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_dit.py tests/test_gpu_wan.py -x -q 2>&1 | tail -4 ) > gpurun_out/r02c23_tests.log 2>&1
timeout 300 python tools/dev_ln.py > gpurun_out/r02c23_ln.log 2>&1
timeout 300 python bench.py --primary-only --steps 4 --warmup 3 2>/dev/null | tail -1 | cut -c1-330 >> gpurun_out/r02c23_ln.log
cat gpurun_out/r02c23_tests.log gpurun_out/r02c23_ln.log
