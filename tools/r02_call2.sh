#!/bin/bash
# Round-2 call 2: MVCS three-tier kernel (tests + bench), GEMM L2 knob sweep (time + dram traffic)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_scorer.py tests/test_gpu_dit.py -x -q 2>&1 | tail -8 ) > gpurun_out/r02c2_tests.log 2>&1
timeout 300 python tools/dev_scorer_bench.py > gpurun_out/r02c2_scorer.log 2>&1
: > gpurun_out/r02c2_gemm.log
for v in "16 0 0" "32 0 0" "32 1 0" "32 1 1" "48 1 1" "24 1 1" "16 1 1" "32 1 5" "64 1 1"; do
  set -- $v
  export VGPA_GEMM_GROUP_M=$1 VGPA_GEMM_STREAM_OUT=$2 VGPA_GEMM_HINTS=$3
  timeout 200 python tools/dev_gemm_l2.py >> gpurun_out/r02c2_gemm.log 2>&1
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 -s 2 -c 1 \
     python tools/dev_gemm_l2.py ff1 2>&1 | grep -E "dram__bytes|gpu__time" | tr '\n' ' ' >> gpurun_out/r02c2_gemm.log
  echo " <- ncu ff1 G=$1 S=$2 H=$3" >> gpurun_out/r02c2_gemm.log
done
unset VGPA_GEMM_GROUP_M VGPA_GEMM_STREAM_OUT VGPA_GEMM_HINTS
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mvcs_pairs -s 1 -c 1 -f -o gpurun_out/r02c2_mvcs python tools/dev_profile_kernels.py > gpurun_out/r02c2_mvcs_ncu.log 2>&1
cat gpurun_out/r02c2_tests.log; cat gpurun_out/r02c2_scorer.log | tail -12; cat gpurun_out/r02c2_gemm.log
