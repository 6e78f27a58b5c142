#!/bin/bash
mkdir -p gpurun_out
for m in 4 3; do VGPA_MVCS_MINB=$m timeout 300 python tools/dev_scorer_bench.py 2>&1 | grep "MVCS batched" | sed "s/^/MINB=$m /"; done > gpurun_out/r02c4_mvcs.log 2>&1
: > gpurun_out/r02c4_vae.log
for v in "4 0" "4 1" "9 1" "2 1" "1 1"; do
  set -- $v
  echo "== VAE_STREAMS=$1 VAE_GRAPH=$2" >> gpurun_out/r02c4_vae.log
  VAE_STREAMS=$1 VAE_GRAPH=$2 timeout 300 python tools/dev_vae.py 2>&1 | tail -2 >> gpurun_out/r02c4_vae.log
done
VAE_STREAMS=1 VAE_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r02c4_vae_launches.csv python tools/dev_vae_one.py > gpurun_out/r02c4_vae_ncu.log 2>&1
python tools/dev_vae_launches.py gpurun_out/r02c4_vae_launches.csv > gpurun_out/r02c4_vae_summary.txt 2>&1
cat gpurun_out/r02c4_mvcs.log gpurun_out/r02c4_vae.log; head -40 gpurun_out/r02c4_vae_summary.txt
