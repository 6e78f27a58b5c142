#!/bin/bash
# Last validation of round 2 with the final library: GPU tests, smoke, bench (N = 1), ncu --set full of the scorer / T5 kernels.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/r02g_gputests.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r02g_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"mvcs_pairs" -s 1 -c 1 -f -o gpurun_out/r02g_mvcs python tools/dev_scorer_ncu.py > gpurun_out/r02g_ncu.log 2>&1
tail -3 gpurun_out/r02g_gputests.log; cat gpurun_out/r02g_smoke.log; head -c 400 gpurun_out/r02g_bench.json; echo; tail -2 gpurun_out/r02g_bench.err
