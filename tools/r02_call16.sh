#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_vae.py "tests/test_gpu_parity_full.py::test_vae_decode_full_size_tiled" -q 2>&1 | tail -12 ) > gpurun_out/r02c16_tests.log 2>&1
VAE_STREAMS=9 VAE_GRAPH=0 timeout 300 python tools/dev_vae.py 2>&1 | tail -2 > gpurun_out/r02c16_vae.log
cat gpurun_out/r02c16_tests.log gpurun_out/r02c16_vae.log
