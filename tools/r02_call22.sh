#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02c22_ab.log
for f in 1 0 1; do
  VGPA_ATTN_FAST=$f timeout 300 python bench.py --primary-only --steps 4 --warmup 3 2>/dev/null | tail -1 | sed "s/^/ATTN_FAST=$f /" >> gpurun_out/r02c22_ab.log
done
cut -c1-360 gpurun_out/r02c22_ab.log
