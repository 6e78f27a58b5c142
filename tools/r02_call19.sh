#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dit.py tests/test_gpu_t5.py tests/test_gpu_vae.py tests/test_gpu_wan.py -x -q 2>&1 | tail -8 ) > gpurun_out/r02c19_tests.log 2>&1
for b in 0 1; do VGPA_GEMM_SKINNY_TILES=$b timeout 300 python tools/dev_t5.py 2>&1 | tail -2 | sed "s/^/SKINNY_TILES=$b /"; done > gpurun_out/r02c19_t5.log 2>&1
cat gpurun_out/r02c19_tests.log gpurun_out/r02c19_t5.log
