#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02c7_gputests.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r02c7_smoke.log 2>&1
timeout 300 python tools/dev_scorer_bench.py > gpurun_out/r02c7_scorer.log 2>&1
timeout 900 python bench.py > gpurun_out/r02c7_bench.json 2> gpurun_out/r02c7_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"mvcs_pairs|dpo_partial|dpo_finalize" -s 3 -c 3 -f -o gpurun_out/r02c7_scorer_kernels python tools/dev_scorer_ncu.py > gpurun_out/r02c7_ncu.log 2>&1
tail -6 gpurun_out/r02c7_gputests.log; cat gpurun_out/r02c7_smoke.log; grep -E "MVCS" gpurun_out/r02c7_scorer.log; head -c 900 gpurun_out/r02c7_bench.json; tail -3 gpurun_out/r02c7_bench.err
