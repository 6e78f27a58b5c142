#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_scorer.py -x -q 2>&1 | tail -5 ) > gpurun_out/r02c3_tests.log 2>&1
timeout 300 python tools/dev_scorer_bench.py 2>&1 | grep MVCS > gpurun_out/r02c3_scorer.log
timeout 200 python tools/dev_gemm_l2.py > gpurun_out/r02c3_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mvcs_pairs -s 1 -c 1 -f -o gpurun_out/r02c3_mvcs python tools/dev_profile_kernels.py > gpurun_out/r02c3_mvcs_ncu.log 2>&1
cat gpurun_out/r02c3_tests.log gpurun_out/r02c3_scorer.log gpurun_out/r02c3_gemm.log
