#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/dev_sanitize.py > gpurun_out/r02c15_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/dev_sanitize.py > gpurun_out/r02c15_racecheck.log 2>&1
tail -4 gpurun_out/r02c15_memcheck.log; tail -4 gpurun_out/r02c15_racecheck.log
