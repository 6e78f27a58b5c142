#!/bin/bash
# End-of-round sequence: GPU tests, smoke, bench (N = 1), reference arm, launch list of a short bench, sanitizer.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r02f_gputests.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r02f_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/r02f_ref.json 2> gpurun_out/r02f_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 480 --csv --log-file gpurun_out/r02f_launches.csv \
    python bench.py --layers 4 --steps 1 --warmup 1 --primary-only > gpurun_out/r02f_launches_bench.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python tools/dev_sanitize.py > gpurun_out/r02f_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/dev_sanitize.py > gpurun_out/r02f_racecheck.log 2>&1
tail -4 gpurun_out/r02f_gputests.log; cat gpurun_out/r02f_smoke.log; head -c 700 gpurun_out/r02f_bench.json; echo; tail -2 gpurun_out/r02f_bench.err
head -c 500 gpurun_out/r02f_ref.json; echo; tail -2 gpurun_out/r02f_memcheck.log; tail -2 gpurun_out/r02f_racecheck.log
